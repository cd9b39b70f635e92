#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call9.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call9.log; tail -n 8 $O/$name.log | tee -a $O/call9.log; return $rc; }
run c9_oz_tests 400 python -m pytest tests/test_ozaki.py -x -q
for dbg in 0 1 3 7; do SGDML_B200_OZAKI_DBG=$dbg timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c9_oz_probe2.log; done
for S in 4 5 6; do OZ_S=$S timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/c9_oz_probe2.log; done
run c9_oz_probe 200 python tools/ozaki_probe.py
run c9_solve 400 python tools/solve_check.py --workload aspirin
run c9_solve2 400 python tools/solve_check.py --workload aspirin
run c9_tests 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -k "assemble or potrf or large_outer"
