mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | tail -6
run() { VG_OPTIONS="$1" VG_BENCH_DEVICE_BUILD=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c4,c3,c1 2>/dev/null | tee gpurun_out/bench_r02k.json | python scripts/bench_brief.py /dev/stdin "$1"; }
run ""
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r02k.json'))
for k,c in d['configs'].items(): print(k, c.get('shadow_level0_kernel'), c.get('precise_trig'))
P
