#!/bin/bash
# usage: scripts/tune_variants.sh  — runs perf_trace.py (and a short bench) against every build_variants/lib_*.so
for so in build_variants/lib_*.so; do
  n=$(basename $so .so)
  echo "== $n"
  VG_SO_PATH=$PWD/$so timeout 120 python scripts/perf_trace.py 2>&1 | grep -E "primary|incoherent |shadow" | awk '{print "   ", $1, $5, $6}'
done
