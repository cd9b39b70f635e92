#!/bin/bash
# last call of the round: full GPU suite and smoke on the final tree; BASELINE config 3 at full size on ONE B200, defaults
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call27.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call27.log; tail -n 5 $O/$name.log | tee -a $O/call27.log; return $rc; }
run c27_gpu_tests 1500 python -m pytest tests -q -m gpu
run c27_smoke 300 python __graft_entry__.py smoke
run c27_cg_acala_m2000 300 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 2000 --max-memory 170 --trace 50
