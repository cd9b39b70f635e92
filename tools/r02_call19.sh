#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call19.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call19.log; tail -n 14 $O/$name.log | tee -a $O/call19.log; return $rc; }
run c19_asm_tests 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "assemble or c60 or nystroem or train"
run c19_asm 600 python tools/asm_variants.py
