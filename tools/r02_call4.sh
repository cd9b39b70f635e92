#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call4.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call4.log; tail -n 8 $O/$name.log | tee -a $O/call4.log; return $rc; }
run c4_oz_tests 400 python -m pytest tests/test_ozaki.py -x -q
for dbg in 0 1 3 7; do SGDML_B200_OZAKI_DBG=$dbg timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c4_oz_probe2.log; done
OZ_S=4 timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c4_oz_probe2.log
run c4_oz_probe 200 python tools/ozaki_probe.py
SGDML_B200_OZAKI_SLICES=7 run c4_oz_solve_m1000 400 python tools/solve_check.py --workload aspirin
run c4_iterative 600 python -m pytest tests/test_iterative.py -x -q -m gpu
run c4_gpu_parity 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin_cli.py -x -q -m gpu
run c4_variants 300 python tools/predict_variants.py
run c4_bench 900 python bench.py --steps 5 --warmup 3
OZ_N=4096 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ozaki_gemm -s 3 -c 1 -f -o $O/c4_oz_gemm python tools/ozaki_probe2.py > $O/c4_oz_ncu.log 2>&1
