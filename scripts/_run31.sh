run() { if [ -n "$1" ]; then export VG_SO_PATH=$1; else unset VG_SO_PATH; fi; VG_OPTIONS=$3 VG_BENCH_DEVICE_BUILD=0 timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3,c4 2>/dev/null | python scripts/bench_brief.py /dev/stdin "$2" | grep -v headline; }
run "" base shadow_level0_per_lane=0
run $PWD/build_variants/lib_spec.so spec shadow_level0_per_lane=0
VG_SO_PATH=$PWD/build_variants/lib_spec.so timeout 300 python -m pytest tests/test_gpu_render.py tests/test_gpu_baseline_configs.py -m gpu -q -x 2>&1 | tail -3
