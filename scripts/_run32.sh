run() { VG_OPTIONS=$1 VG_BENCH_DEVICE_BUILD=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c4 2>/dev/null | python scripts/bench_brief.py /dev/stdin "$1" | grep -v headline; }
run primary_per_lane_motion=0
run primary_per_lane_motion=1
