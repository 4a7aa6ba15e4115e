#!/bin/bash
# compute-sanitizer passes over the hot path (run on the GPU box: `gpurun -- bash scripts/sanitize.sh`).
# memcheck / racecheck / synccheck / initcheck over smoke() (traversal batch + a small wavefront frame), then memcheck over
# the traversal, build and render parity tests.  Logs go to gpurun_out/sanitize_*.log; the script prints one summary line
# per pass ("ERROR SUMMARY: n errors" as compute-sanitizer reports it).  Each pass is bounded by its own timeout.
set -u
OUT=gpurun_out
mkdir -p $OUT
CS=/usr/local/cuda/bin/compute-sanitizer
SMOKE='import __graft_entry__ as g; g.smoke()'

run() {  # name, timeout, tool options..., -- command
  local name=$1 tmo=$2; shift 2
  local log=$OUT/sanitize_$name.log
  timeout "$tmo" $CS "$@" > "$log" 2>&1
  local rc=$?
  echo "$name rc=$rc $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY' "$log" | tail -1) $(grep -E 'smoke ok|passed|failed' "$log" | tail -1)"
}

run memcheck_smoke 300 --tool memcheck --leak-check no --error-exitcode 1 python -c "$SMOKE"
run racecheck_smoke 400 --tool racecheck --racecheck-report all --error-exitcode 1 python -c "$SMOKE"
run synccheck_smoke 300 --tool synccheck --error-exitcode 1 python -c "$SMOKE"
run initcheck_smoke 300 --tool initcheck --error-exitcode 1 python -c "$SMOKE"
# round 2: the per-lane occlusion kernel (published shared blocks, one-register stack) and the {P, D} ray records
run racecheck_round2 600 --tool racecheck --racecheck-report all --error-exitcode 1 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "level0 or pd_ray"
run memcheck_round2 600 --tool memcheck --leak-check no --error-exitcode 1 python -m pytest tests/test_gpu_round2.py -m gpu -x -q -k "level0 or pd_ray or path_order or stack_overflow"
run memcheck_trace 600 --tool memcheck --leak-check no --error-exitcode 1 python -m pytest tests/test_gpu_trace.py tests/test_gpu_build.py -m gpu -x -q
run memcheck_render 600 --tool memcheck --leak-check no --error-exitcode 1 python -m pytest tests/test_gpu_render.py tests/test_gpu_texture.py -m gpu -x -q

# negative control: the same tool invocations must flag a deliberately broken kernel loaded the same way (ctypes)
nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -shared -Xcompiler -fPIC -o $OUT/libsanitize_control.so scripts/sanitize_control.cu
run control_memcheck 120 --tool memcheck --leak-check no --error-exitcode 1 python -c "import ctypes; ctypes.CDLL('$OUT/libsanitize_control.so').control_oob()"
run control_racecheck 120 --tool racecheck --racecheck-report all --error-exitcode 1 python -c "import ctypes; ctypes.CDLL('$OUT/libsanitize_control.so').control_race()"
rm -f $OUT/libsanitize_control.so
