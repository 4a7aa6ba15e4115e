set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -25
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02e.json 2> gpurun_out/bench_r02e.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_r02e.err
python scripts/bench_brief.py gpurun_out/bench_r02e.json main
