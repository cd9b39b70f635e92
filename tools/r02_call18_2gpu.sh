#!/bin/bash
# 2 GPUs: bench at N = 2 with the communicator check in the JSON line; config 4 at full size with the inducing set sized
# against both ranks' memory
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call18.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call18.log; tail -n 6 $O/$name.log | tee -a $O/call18.log; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
run c18_bench_n2 420 $TR --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3
ls gpurun_out/nccl_* 2>/dev/null | head -4 | tee -a $O/call18.log
export SGDML_B200_OZAKI_PREDICT_SLICES=5
run c18_cg_syn100_m5000_n2 500 $TR --master-port 29523 tools/cg_probe.py --workload synthetic100 --n-train 5000 --max-memory 170 --trace 100
