set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --tb=short -rf -x 2>&1 | tail -8
VG_BENCH_DEVICE_BUILD=0 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-nonparity --configs c2,c4,c1,c3 2>gpurun_out/bench_r02j.err | tee gpurun_out/bench_r02j.json | python scripts/bench_brief.py /dev/stdin "stack1reg+publish"
