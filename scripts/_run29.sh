mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2 2>/dev/null | python scripts/bench_brief.py /dev/stdin last
