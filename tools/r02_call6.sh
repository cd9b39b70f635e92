#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call6.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call6.log; tail -n 8 $O/$name.log | tee -a $O/call6.log; return $rc; }
run c6_oz_tests 400 python -m pytest tests/test_ozaki.py -x -q
for dbg in 0 1 3 7; do SGDML_B200_OZAKI_DBG=$dbg timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c6_oz_probe2.log; done
for S in 4 5 6; do OZ_S=$S timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c6_oz_probe2.log; done
SGDML_B200_OZAKI_SLICES=7 run c6_oz_solve_s7 400 python tools/solve_check.py --workload aspirin
SGDML_B200_OZAKI_SLICES=6 run c6_oz_solve_s6 400 python tools/solve_check.py --workload aspirin
run c6_dropin 600 python -m pytest tests/test_dropin_cli.py tests/test_iterative.py -q -m gpu
for s in 4 5; do SGDML_B200_OZAKI_PREDICT_SLICES=$s run c6_bench_acala_s$s 400 python bench.py --workload ac-ala3-nhme --steps 3 --warmup 3 --no-cpu-baseline --no-extras; done
OZ_N=4096 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ozaki_gemm -s 3 -c 1 -f -o $O/c6_oz_gemm python tools/ozaki_probe2.py > $O/c6_oz_ncu.log 2>&1
