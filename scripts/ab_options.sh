#!/bin/bash
# usage: scripts/ab_options.sh <config> "<opts1>" "<opts2>" ...   — short bench.py per VG_OPTIONS value; prints value, ms/step and stage times
cfg=$1; shift
for o in "$@"; do
  VG_BENCH_CONFIG=$cfg VG_OPTIONS="$o" VG_BENCH_DEVICE_BUILD=0 timeout 400 python bench.py --steps 3 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
i=d.get('incoherent_closest_hit') or {}
print('$cfg','opts=$o',round(d['value'],1),'Mrays/s',round(d['ms_per_step'],2),'ms/step',{k:round(v,2) for k,v in d['stage_ms_per_step'].items()},'incoh',round(i.get('value',0),1))"
done
