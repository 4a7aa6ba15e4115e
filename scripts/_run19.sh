set -x
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -rf -k "pd_ray or compact" 2>&1 | tail -12
for c in 16 17 18 19; do
VG_OPTIONS=stream_chunk_log2=$c timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2 2>gpurun_out/bench_r02i.err | tee gpurun_out/bench_r02i_$c.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('chunk_log2', $c, d['e2e']['value'], d['e2e']['all_modes_ms_rank0'])"
done
