set -x
mkdir -p gpurun_out
python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -rf 2>&1 | tail -25
for plain in 1 ""; do
VG_BENCH_E2E_PLAIN=$plain python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 10 --warmup 3 --configs c2,c1 2>/dev/null | python scripts/bench_brief.py /dev/stdin "n2 plain=$plain"
VG_BENCH_E2E_PLAIN=$plain VG_BENCH_DEVICE_BUILD=0 python bench.py --steps 10 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c1 2>/dev/null | python scripts/bench_brief.py /dev/stdin "n1 plain=$plain"
done
