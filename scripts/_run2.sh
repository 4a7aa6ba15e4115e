set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
scripts/ab_options.sh c2 "iter_group=32,iters_per_batch=32" "iter_group=32,iters_per_batch=64" "iter_group=32,iters_per_batch=32,shadow_per_lane=1" "iter_group=32,iters_per_batch=32,primary_per_lane=0" 2>&1 | tee gpurun_out/ab2_c2.txt
scripts/ab_options.sh c3 "iter_group=32,iters_per_batch=32" 2>&1 | tee gpurun_out/ab2_c3.txt
scripts/ab_options.sh c4 "iter_group=32,iters_per_batch=32" "iter_group=32,iters_per_batch=32,shadow_per_lane=1" 2>&1 | tee gpurun_out/ab2_c4.txt
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_trace_queue|k_shade|k_raygen|k_resolve" -c 5 -o gpurun_out/ncu_r02a -f python scripts/profile_step.py --iters=32 --opt=iters_per_batch=32 > gpurun_out/ncu_r02a.log 2>&1
tail -3 gpurun_out/ncu_r02a.log
ls -la gpurun_out/
