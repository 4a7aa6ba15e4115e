set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
scripts/ab_options.sh c2 "iter_group=1" "iter_group=2" "iter_group=4" "iter_group=8" "iter_group=16" "iter_group=4,iters_per_batch=32" "iter_group=32,iters_per_batch=32" 2>&1 | tee gpurun_out/ab_itergroup_c2.txt
scripts/ab_options.sh c3 "iter_group=1" "iter_group=4" "iter_group=16" 2>&1 | tee gpurun_out/ab_itergroup_c3.txt
scripts/ab_options.sh c4 "iter_group=1" "iter_group=4" "iter_group=16" 2>&1 | tee gpurun_out/ab_itergroup_c4.txt
python - <<'PY' 2>&1 | tee gpurun_out/peaks.txt
from vermeer_b200.host import Device
d=Device(0)
print(d.measure_peaks())
print(d.measure_peaks())
PY
