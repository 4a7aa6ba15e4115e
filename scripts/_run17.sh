set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -q --tb=short -rf 2>&1 | tail -8
for o in "accumulate_tiled=0" "accumulate_tiled=1"; do
VG_OPTIONS=$o VG_BENCH_DEVICE_BUILD=0 python bench.py --steps 5 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c4,c1 2>/dev/null | python scripts/bench_brief.py /dev/stdin "$o"
done
