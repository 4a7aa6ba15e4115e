set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 5 --warmup 3 > gpurun_out/bench_r02f_n8.json 2> gpurun_out/bench_r02f_n8.err ) 2>&1 | tail -4
tail -5 gpurun_out/bench_r02f_n8.err
python scripts/bench_brief.py gpurun_out/bench_r02f_n8.json n8
python -m pytest tests/test_gpu_round2.py tests/test_gpu_trace.py -m gpu -q --tb=short -rf -k "nccl or nested" 2>&1 | tail -5
