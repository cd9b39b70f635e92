#!/bin/bash
# 2 GPUs of one box: the NCCL test, the bench at N = 2 (replica prediction + sharded legs), a converged sharded CG run
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call8.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call8.log; tail -n 6 $O/$name.log | tee -a $O/call8.log; return $rc; }
nvidia-smi -L | tee -a $O/call8.log
run c8_nccl_test 600 python -m pytest tests/test_iterative.py -q -m gpu -k "two_ranks"
run c8_bench_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3
run c8_cg_c60_n2 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 tools/cg_probe.py --workload c60 --n-train 300 --max-memory 16
run c8_cg_c60_n1 900 python tools/cg_probe.py --workload c60 --n-train 300 --max-memory 16
ls gpurun_out/nccl_* 2>/dev/null | head; grep -h "comm 0x\|nranks\|NVLS\|Channel 00/" gpurun_out/nccl_n2_* 2>/dev/null | head -8 | tee -a $O/call8.log
