#!/bin/bash
# runs a short bench against every build_variants/lib_*.so (bench.py rebuilds nothing when VG_SO_PATH is set)
for so in build_variants/lib_*.so; do
  n=$(basename $so .so)
  VG_SO_PATH=$PWD/$so timeout 200 python bench.py --steps 2 --warmup 3 --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$n', round(d['value'],1), 'Mrays/s', round(d['ms_per_step'],2), 'ms/step')"
done
