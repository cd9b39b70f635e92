#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call25.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call25.log; tail -n 6 $O/$name.log | tee -a $O/call25.log; return $rc; }
run c25_gpu_tests 1500 python -m pytest tests -q -m gpu
run c25_smoke 300 python __graft_entry__.py smoke
run c25_bench 900 python bench.py --gpus 1 --steps 20 --warmup 5
