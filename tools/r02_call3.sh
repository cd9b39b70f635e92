#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call3.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call3.log; tail -n 8 $O/$name.log | tee -a $O/call3.log; return $rc; }
run c3_oz_stage 240 python -m pytest tests/test_ozaki.py -x -q -k "stage1 or stage2"
run c3_oz_rest 400 python -m pytest tests/test_ozaki.py -x -q -k "not stage1 and not stage2"
for dbg in 0 1 3 7; do SGDML_B200_OZAKI_DBG=$dbg timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c3_oz_probe2.log; done
run c3_oz_probe 200 python tools/ozaki_probe.py
SGDML_B200_OZAKI_SLICES=7 run c3_oz_solve_m1000 400 python tools/solve_check.py --workload aspirin
run c3_iterative 600 python -m pytest tests/test_iterative.py -x -q -m gpu
run c3_gpu_tests 600 python -m pytest tests -m gpu -x -q --deselect tests/test_ozaki.py --deselect tests/test_iterative.py
run c3_cg_aspirin 600 python tools/cg_probe.py --workload aspirin --n-train 1000 --max-memory 8
run c3_cg_acala 900 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 500 --max-memory 8 --profile
OZ_N=4096 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ozaki_gemm -s 1 -c 1 -f -o $O/c3_oz_gemm python tools/ozaki_probe2.py > $O/c3_oz_ncu.log 2>&1
