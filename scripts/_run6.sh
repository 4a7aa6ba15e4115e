set -x
mkdir -p gpurun_out
nvidia-smi -L
python -m pytest tests/test_gpu_round2.py -m gpu -q --tb=short -rf -k "nccl or communicator" 2>&1 | tail -20
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --configs c2,c1 > gpurun_out/bench_r02c_n2.json 2> gpurun_out/bench_r02c_n2.err
tail -5 gpurun_out/bench_r02c_n2.err
python scripts/bench_brief.py gpurun_out/bench_r02c_n2.json n2
