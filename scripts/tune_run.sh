#!/bin/bash
# usage: scripts/tune_run.sh <config> <variant...>   — short bench.py per build_variants/lib_<variant>.so; prints value, ms/step and stage times
cfg=$1; shift
for n in "$@"; do
  so=build_variants/lib_$n.so
  VG_BENCH_CONFIG=$cfg VG_SO_PATH=$PWD/$so timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
i=d.get('incoherent_closest_hit') or {}
print('$cfg','$n','opts=${VG_OPTIONS}',round(d['value'],1),'Mrays/s',round(d['ms_per_step'],2),'ms/step',{k:round(v,2) for k,v in d['stage_ms_per_step'].items()},'incoh',round(i.get('value',0),1))"
done
