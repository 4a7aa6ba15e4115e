set -x
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none"
cap() {
  name=$1; shift; regex=$1; shift; cnt=$1; shift
  timeout 900 $NCU $NCU_EXTRA -k regex:"$regex" -c $cnt -o /tmp/ncu/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.csv 2>/dev/null
}
NCU_EXTRA="--import-source on" cap ncu_r02d_c2 "k_trace_queue|k_shade|k_raygen|k_resolve|k_accumulate" 12 python scripts/profile_step.py --iters=32 --opt=iters_per_batch=32 --opt=shadow_level0_per_lane=1 --stats-out=gpurun_out/ncu_r02d_c2.stats.json
ncu -i /tmp/ncu/ncu_r02d_c2.ncu-rep --page source --csv --print-source cuda,sass > /tmp/ncu/c2_source.csv 2>/dev/null
for k in "k_trace_queue<(int)0" "k_trace_queue<(int)1" "k_shade"; do python scripts/ncu_lines.py /tmp/ncu/c2_source.csv "$k" 40; echo; done > gpurun_out/ncu_r02d_c2_lines.txt
NCU_EXTRA="-s 5" cap ncu_r02d_inc "k_trace_batch" 1 python scripts/profile_incoherent.py --stats-out=gpurun_out/ncu_r02d_inc.stats.json
NCU_EXTRA="" cap ncu_r02d_c4 "k_trace_queue|k_shade" 14 python scripts/profile_step.py --motion --iters=32 --opt=iters_per_batch=32 --opt=shadow_level0_per_lane=0 --stats-out=gpurun_out/ncu_r02d_c4.stats.json
NCU_EXTRA="" cap ncu_r02d_c3 "k_trace_queue|k_shade" 14 python scripts/profile_step.py --c3 --iters=4 --opt=iters_per_batch=4 --opt=shadow_level0_per_lane=0 --stats-out=gpurun_out/ncu_r02d_c3.stats.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02d_bench_steps1.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2 > gpurun_out/launches_r02d.log 2>&1
du -sh gpurun_out
