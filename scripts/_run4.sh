set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -15
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02b.json 2> gpurun_out/bench_r02b.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_r02b.err
python scripts/bench_brief.py gpurun_out/bench_r02b.json main
for v in nof2 f2tm6; do
  VG_SO_PATH=$PWD/build_variants/lib_$v.so python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin $v
done
