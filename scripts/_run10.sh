set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -25
VG_BENCH_DEVICE_BUILD=0 python bench.py --steps 3 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2 2>/dev/null | python scripts/bench_brief.py /dev/stdin quick
