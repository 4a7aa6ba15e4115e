mkdir -p gpurun_out
VG_OPTIONS=shadow_level0_per_lane=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r02d_bench_steps1.csv python bench.py --steps 1 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2 > gpurun_out/launches_r02d.log 2>&1
tail -c 600 gpurun_out/launches_r02d.log
