#!/bin/bash
# converged iterative training at (or near) full size on one GPU: BASELINE config 3 (Ac-Ala3-NHMe shape, M = 2000, S = 243)
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call10.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call10.log; tail -n 6 $O/$name.log | tee -a $O/call10.log; return $rc; }
SGDML_B200_OZAKI_PREDICT_SLICES=5 run c10_cg_acala_m2000_int8 1200 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 2000 --max-memory 150 --trace 25
run c10_cg_acala_m1000_fp64 900 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 1000 --max-memory 150 --trace 25
SGDML_B200_OZAKI_PREDICT_SLICES=5 run c10_cg_acala_m500_int8 600 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 500 --max-memory 8
