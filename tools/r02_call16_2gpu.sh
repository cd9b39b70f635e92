#!/bin/bash
# 2 GPUs of one box: NCCL test, short bench at N = 2, then FULL-SIZE converged iterative training of BASELINE configs
# 3, 5 and 4 with the Nystroem factor and K.v row-sharded over the two ranks (NCCL all-gather / all-reduce on device
# buffers inside sgdml_b200_pcg); K.v contractions on 5 int8 slices
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call16.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call16.log; tail -n 6 $O/$name.log | tee -a $O/call16.log; return $rc; }
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
nvidia-smi -L | tee -a $O/call16.log
run c16_nccl_test 300 python -m pytest tests/test_iterative.py -q -m gpu -k "two_ranks"
run c16_bench_n2 600 $TR --master-port 29517 bench.py --gpus 2 --steps 3 --warmup 3 --no-extras
ls gpurun_out/nccl_* 2>/dev/null | head -4 | tee -a $O/call16.log; grep -h "nranks\|NVLS\|Connected all\|Channel 00/" gpurun_out/nccl_n2_* 2>/dev/null | head -6 | tee -a $O/call16.log
export SGDML_B200_OZAKI_PREDICT_SLICES=5
run c16_cg_acala_m2000_n2 420 $TR --master-port 29519 tools/cg_probe.py --workload ac-ala3-nhme --n-train 2000 --max-memory 170 --trace 50
run c16_cg_c60_m3000_n2 600 $TR --master-port 29521 tools/cg_probe.py --workload c60 --n-train 3000 --max-memory 170 --trace 25
run c16_cg_syn100_m5000_n2 360 $TR --master-port 29523 tools/cg_probe.py --workload synthetic100 --n-train 5000 --max-memory 170 --trace 25
