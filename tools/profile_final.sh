#!/bin/bash
# Final round-1 ncu evidence (run under gpurun, one GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# launch lists (device time per launch; cold-cache, serialised: compare shares, not absolutes)
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01f_launches_bench_aspirin_predict.csv \
  python bench.py --no-train --no-cpu-baseline --steps 2 --warmup 1 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/r01f_launches_train_ethanol.csv \
  python tools/train_probe.py ethanol 200 > /dev/null 2>&1
$NCU --metrics gpu__time_duration.sum --csv -c 1200 --log-file gpurun_out/r01f_launches_train_aspirin250.csv \
  python tools/train_probe.py aspirin 250 > /dev/null 2>&1
# full captures of the top kernels
$NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o gpurun_out/r01f_predict_aspirin \
  python bench.py --no-train --no-cpu-baseline --steps 1 --warmup 1 --batch 24576 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o gpurun_out/r01f_predict_ethanol \
  python bench.py --workload ethanol --no-train --no-cpu-baseline --steps 1 --warmup 1 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_gemm_nt -s 22 -c 1 -f -o gpurun_out/r01f_gemm_trailing \
  python tools/train_probe.py aspirin 500 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_assemble -s 1 -c 1 -f -o gpurun_out/r01f_assemble \
  python tools/train_probe.py aspirin 250 > /dev/null 2>&1
ls -la gpurun_out | grep r01f
