#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call11.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call11.log; tail -n 4 $O/$name.log | tee -a $O/call11.log; return $rc; }
run c11_bench_n2 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3
run c11_nccl_test 200 python -m pytest tests/test_iterative.py -q -m gpu -k "two_ranks"
grep -h "nranks\|NVLS\|comm 0x" gpurun_out/nccl_n2_* 2>/dev/null | head -6 | cut -c1-200 | tee -a $O/call11.log
