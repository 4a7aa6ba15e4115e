mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_r02n.json 2> gpurun_out/bench_r02n.err ) 2>&1 | tail -3
python scripts/bench_brief.py gpurun_out/bench_r02n.json final
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
