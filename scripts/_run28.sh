mkdir -p gpurun_out
nvidia-smi topo -m 2>&1 | head -14
lscpu | grep -i "numa\|socket\|^CPU(s)"
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r02p_n8.json 2> gpurun_out/bench_r02p_n8.err ) 2>&1 | tail -3
python scripts/bench_brief.py gpurun_out/bench_r02p_n8.json n8
python - <<'P'
import json
d=json.loads([l for l in open('gpurun_out/bench_r02p_n8.json') if l.startswith('{')][-1])
print(d.get('host_placement')); print({k:d['e2e'][k] for k in ('value','ms_per_step','all_modes_ms_rank0')})
for k,c in d['configs'].items(): print(k, c.get('gathered_frame_bit_identical_to_single_gpu'), c['e2e'].get('gather_ms_per_step'))
P
tail -5 gpurun_out/bench_r02p_n8.err
