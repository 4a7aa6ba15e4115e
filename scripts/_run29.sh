mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time python bench.py > gpurun_out/bench_r02q.json 2> gpurun_out/bench_r02q.err ) 2>&1 | tail -3
python scripts/bench_brief.py gpurun_out/bench_r02q.json final
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
