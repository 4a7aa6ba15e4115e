set -x
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none"
cap() {  # name, kernel regex, count, extra ncu args..., -- command
  name=$1; shift; regex=$1; shift; cnt=$1; shift
  timeout 900 $NCU $NCU_EXTRA -k regex:"$regex" -c $cnt -o /tmp/ncu/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.csv 2>/dev/null
}
NCU_EXTRA="--import-source on" cap ncu_r02b_c2 "k_trace_queue|k_shade|k_raygen|k_resolve|k_accumulate" 12 python scripts/profile_step.py --iters=32 --opt=iters_per_batch=32 --stats-out=gpurun_out/ncu_r02b_c2.stats.json
ncu -i /tmp/ncu/ncu_r02b_c2.ncu-rep --page source --csv --print-source cuda,sass > gpurun_out/ncu_r02b_c2_source.csv 2>/dev/null
NCU_EXTRA="-s 5" cap ncu_r02b_inc "k_trace_batch" 1 python scripts/profile_incoherent.py --stats-out=gpurun_out/ncu_r02b_inc.stats.json
NCU_EXTRA="" cap ncu_r02b_c4 "k_trace_queue|k_shade" 14 python scripts/profile_step.py --motion --iters=32 --opt=iters_per_batch=32 --stats-out=gpurun_out/ncu_r02b_c4.stats.json
NCU_EXTRA="" cap ncu_r02b_c3 "k_trace_queue|k_shade" 14 python scripts/profile_step.py --c3 --iters=4 --opt=iters_per_batch=4 --stats-out=gpurun_out/ncu_r02b_c3.stats.json
ls -la gpurun_out/ /tmp/ncu
du -sh gpurun_out
