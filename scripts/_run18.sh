set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -rf -k "pd_ray or compact" 2>&1 | tail -8
python bench.py --steps 5 --warmup 3 --no-wavefront --no-nonparity --configs c2 2>gpurun_out/bench_r02h.err | tee gpurun_out/bench_r02h.json | python scripts/bench_brief.py /dev/stdin "pd"
python - <<'P'
import json
d=json.load(open('gpurun_out/bench_r02h.json'))
print(json.dumps(d["e2e"],indent=1)); print(d.get("host_placement")); print(d["speedup_vs_cpu"])
P
nvidia-smi topo -m 2>&1 | head -20; lscpu | grep -i "numa\|socket\|model name\|^CPU(s)"
