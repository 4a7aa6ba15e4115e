set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -12
python scripts/leafmax_scan.py 2>&1 | grep leaf_max | tee gpurun_out/leafmax_scan.txt
( time python bench.py --steps 5 --warmup 3 > gpurun_out/bench_r02d.json 2> gpurun_out/bench_r02d.err ) 2>&1 | tail -4
tail -3 gpurun_out/bench_r02d.err
python scripts/bench_brief.py gpurun_out/bench_r02d.json main
