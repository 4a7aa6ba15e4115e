run() { VG_OPTIONS=$1 VG_BENCH_DEVICE_BUILD=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c4 2>/dev/null | python scripts/bench_brief.py /dev/stdin "$1" | grep "c4"; }
run shadow_level0_per_lane=0
run shadow_level0_per_lane=1
run shadow_level0_per_lane=2
timeout 300 python -m pytest tests/test_gpu_round2.py tests/test_gpu_render.py -m gpu -q -x -k "level0 or motion" 2>&1 | tail -3
