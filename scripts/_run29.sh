mkdir -p gpurun_out
( time python bench.py > gpurun_out/bench_r02m.json 2> gpurun_out/bench_r02m.err ) 2>&1 | tail -3
python scripts/bench_brief.py gpurun_out/bench_r02m.json final
( time python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r02m_ref.json 2> gpurun_out/bench_r02m_ref.err ) 2>&1 | tail -3
cut -c1-600 gpurun_out/bench_r02m_ref.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02d_bench_steps1.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2 > gpurun_out/launches_r02d.log 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
