#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call14.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call14.log; tail -n 12 $O/$name.log | tee -a $O/call14.log; return $rc; }
run c14_gpu_tests 1500 python -m pytest tests -q -m gpu -x
run c14_variants 400 python tools/predict_variants.py
run c14_solve 400 python tools/solve_check.py --workload aspirin --repeat 3
run c14_solve_prof 400 python tools/solve_check.py --workload aspirin --repeat 2 --profile
