#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call23.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call23.log; tail -n 12 $O/$name.log | tee -a $O/call23.log; return $rc; }
run c23_tests 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "variant or graph or predict"
VARIANT_SHAPES=ethanol run c23_variants 300 python tools/predict_variants.py
run c23_bench 900 python bench.py
