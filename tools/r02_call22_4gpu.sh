#!/bin/bash
# 4 GPUs of one box: the driver's scaling command at N = 4 (replica prediction + sharded legs), then BASELINE config 5 at
# full size over 4 ranks
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call22.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call22.log; tail -n 4 $O/$name.log | tee -a $O/call22.log; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
nvidia-smi -L | tee -a $O/call22.log
run c22_bench_n4 420 $TR --master-port 29517 bench.py --gpus 4 --steps 3 --warmup 3
export SGDML_B200_OZAKI_PREDICT_SLICES=5
run c22_cg_c60_m3000_n4 240 $TR --master-port 29521 tools/cg_probe.py --workload c60 --n-train 3000 --max-memory 43 --trace 25
