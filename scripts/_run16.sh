set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -8
VG_BENCH_DEVICE_BUILD=0 python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin base
for v in plwide uni16 uni24 uni28; do
  VG_SO_PATH=$PWD/build_variants/lib_$v.so python -m pytest tests/test_gpu_trace.py tests/test_gpu_render.py -m gpu -q -x 2>&1 | tail -1
  VG_SO_PATH=$PWD/build_variants/lib_$v.so python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin $v
done
