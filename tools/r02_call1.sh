#!/bin/bash
# Round-2 GPU call 1: baseline parity suite, staged bring-up of the tcgen05 int8 GEMM (every step under its
# own timeout: a barrier that never completes must not hang the box), its throughput, and the solve check.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() {  # run <name> <timeout> <cmd...>
  local name=$1 to=$2; shift 2
  echo "=== $name" | tee -a $O/call1.log
  timeout $to "$@" > $O/$name.log 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a $O/call1.log
  tail -n 6 $O/$name.log | tee -a $O/call1.log
  return $rc
}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $O/call1.log
run gpu_tests 600 python -m pytest tests -m gpu -x -q --deselect tests/test_ozaki.py
run oz_stage1 180 python -m pytest tests/test_ozaki.py -x -q -k "stage1"
run oz_stage2_first 120 python -m pytest tests/test_ozaki.py -x -q -k "stage2 and 128-64-128"
if run oz_stage2 240 python -m pytest tests/test_ozaki.py -x -q -k "stage2"; then
  run oz_rest 400 python -m pytest tests/test_ozaki.py -x -q -k "not stage1 and not stage2 and not potrf and not predictor"
  run oz_potrf 300 python -m pytest tests/test_ozaki.py -x -q -k "potrf"
  run oz_probe 200 python tools/ozaki_probe.py
  SGDML_B200_OZAKI_SLICES=7 run oz_solve_m300 300 python tools/solve_check.py --workload aspirin --n-train 300
  SGDML_B200_OZAKI_SLICES=7 run oz_solve_m1000 400 python tools/solve_check.py --workload aspirin
else
  # keep evidence for the diagnosis: each width separately on the smallest case
  SGDML_B200_OZAKI_BK=128 run oz_diag128 120 python -m pytest tests/test_ozaki.py -x -q -k "stage2 and 128-64-128 and 128"
fi
run solve_m300 300 python tools/solve_check.py --workload aspirin --n-train 300
run solve_m1000 400 python tools/solve_check.py --workload aspirin
run cg_aspirin 600 python tools/cg_probe.py --workload aspirin --n-train 1000 --max-memory 8 --profile
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee -a $O/call1.log
