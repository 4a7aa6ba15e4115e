run() { if [ -n "$1" ]; then export VG_SO_PATH=$1; else unset VG_SO_PATH; fi; VG_BENCH_DEVICE_BUILD=0 timeout 300 python bench.py --steps 4 --warmup 3 --no-cpu --no-wavefront --no-nonparity --configs c2,c3 2>/dev/null | python scripts/bench_brief.py /dev/stdin "$2"; }
run "" base
for v in lp0 lp2; do run $PWD/build_variants/lib_$v.so $v; done
