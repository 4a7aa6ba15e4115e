#!/bin/bash
# usage: scripts/tune_variants.sh  — runs perf_trace.py against every build_variants/lib_*.so (and the parity tests on request)
for so in build_variants/lib_*.so; do
  n=$(basename $so .so)
  echo "== $n"
  if [ -n "$VG_TUNE_TEST" ]; then VG_SO_PATH=$PWD/$so timeout 200 python -m pytest tests/test_gpu_trace.py -m gpu -q -x 2>&1 | tail -1; fi
  VG_SO_PATH=$PWD/$so timeout 120 python scripts/perf_trace.py 2>&1 | grep -E "primary|incoherent |shadow|incoh-shuf" | awk '{print "   ", $1, $5, $6}'
done
