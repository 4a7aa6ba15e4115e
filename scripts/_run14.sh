set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -8
VG_SO_PATH=$PWD/build_variants/lib_stackstats.so python scripts/stack_depth.py 2>&1 | grep "max traversal" | tee gpurun_out/stack_depth.txt
VG_BENCH_DEVICE_BUILD=0 python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin base
for v in nm2 nm8 nm16 ra8 ra16 ra32 rb16 rb30; do
  VG_SO_PATH=$PWD/build_variants/lib_$v.so python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin $v
done
