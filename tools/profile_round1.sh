#!/bin/bash
# ncu evidence for round 1 (run under gpurun, one GPU).  Outputs land in gpurun_out/.
set -x
mkdir -p gpurun_out
NCU="ncu --clock-control none"
# 1. launch lists (device time per launch; cold-cache, serialised: compare shares)
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_predict_aspirin.csv \
  python bench.py --no-train --no-cpu-baseline --steps 2 --warmup 1 --batch 65536 > gpurun_out/ncu_bench_predict.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv --log-file gpurun_out/launches_train_ethanol.csv \
  python tools/train_probe.py ethanol 200 > gpurun_out/ncu_train_ethanol.log 2>&1
$NCU --metrics gpu__time_duration.sum --csv -c 800 --log-file gpurun_out/launches_train_aspirin250.csv \
  python tools/train_probe.py aspirin 250 > gpurun_out/ncu_train_aspirin250.log 2>&1
# 2. full captures of the top kernels
$NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o gpurun_out/prof_predict_aspirin \
  python bench.py --no-train --no-cpu-baseline --steps 1 --warmup 1 --batch 24576 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o gpurun_out/prof_predict_ethanol \
  python bench.py --workload ethanol --no-train --no-cpu-baseline --steps 1 --warmup 1 --batch 65536 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_gemm_nt -s 30 -c 1 -f -o gpurun_out/prof_gemm \
  python tools/train_probe.py aspirin 250 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_assemble -s 1 -c 1 -f -o gpurun_out/prof_assemble \
  python tools/train_probe.py aspirin 250 > /dev/null 2>&1
$NCU --set full --import-source on -k regex:k_potf2_tile -s 10 -c 1 -f -o gpurun_out/prof_potf2 \
  python tools/train_probe.py aspirin 125 > /dev/null 2>&1
ls -la gpurun_out
