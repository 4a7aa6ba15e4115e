set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests -m gpu -x -q 2>&1 | tail -15
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02a.json 2> gpurun_out/bench_r02a.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_r02a.err
python scripts/bench_brief.py gpurun_out/bench_r02a.json main
( time python bench.py --impl reference --steps 5 --warmup 2 > gpurun_out/bench_r02a_ref.json ) 2>&1 | tail -4
for v in tm6 tm5 sh7 sh6; do
  VG_SO_PATH=$PWD/build_variants/lib_$v.so python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --configs c2 2>/dev/null | python scripts/bench_brief.py /dev/stdin $v
done
