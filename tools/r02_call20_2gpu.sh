#!/bin/bash
# 2 GPUs: BASELINE configs 5 and 3 at full size again, with the v5 / v4 assembly kernels
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call20.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call20.log; tail -n 4 $O/$name.log | tee -a $O/call20.log; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
export SGDML_B200_OZAKI_PREDICT_SLICES=5
run c20_cg_c60_m3000_n2 360 $TR --master-port 29521 tools/cg_probe.py --workload c60 --n-train 3000 --max-memory 85 --trace 25
run c20_cg_acala_m2000_n2 240 $TR --master-port 29519 tools/cg_probe.py --workload ac-ala3-nhme --n-train 2000 --max-memory 85 --trace 50
