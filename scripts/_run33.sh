set -x
mkdir -p gpurun_out /tmp/ncu
NCU="ncu --set full --clock-control none"
cap() {
  name=$1; shift; regex=$1; shift; cnt=$1; shift
  timeout 900 $NCU $NCU_EXTRA -k regex:"$regex" -c $cnt -o /tmp/ncu/$name -f "$@" > gpurun_out/$name.log 2>&1
  ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > gpurun_out/$name.csv 2>/dev/null
}
NCU_EXTRA="" cap ncu_r02e_c4 "k_trace_queue|k_shade|k_resolve|k_raygen" 14 python scripts/profile_step.py --motion --iters=32 --opt=iters_per_batch=32 --opt=shadow_level0_per_lane=0 --stats-out=gpurun_out/ncu_r02e_c4.stats.json
