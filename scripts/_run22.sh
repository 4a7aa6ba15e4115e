mkdir -p gpurun_out
run() { VG_OPTIONS="$1" VG_BENCH_DEVICE_BUILD=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin "$1"; }
run "shadow_per_lane=1"
