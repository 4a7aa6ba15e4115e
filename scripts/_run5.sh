set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -80 > gpurun_out/pytest_r02c.txt
tail -5 gpurun_out/pytest_r02c.txt
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02c.json 2> gpurun_out/bench_r02c.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_r02c.err
python scripts/bench_brief.py gpurun_out/bench_r02c.json main
