#!/bin/bash
# Round-2 final state: full GPU suite, smoke, both bench arms, ncu capture of the default assembly kernel
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call21.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call21.log; tail -n 8 $O/$name.log | tee -a $O/call21.log; return $rc; }
run c21_gpu_tests 1500 python -m pytest tests -q -m gpu
run c21_smoke 300 python __graft_entry__.py smoke
run c21_bench 900 python bench.py
run c21_bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
NCU="ncu --clock-control none"
timeout 600 $NCU --set full --import-source on -k regex:k_assemble_v4 -s 1 -c 1 -f -o $O/r02b_assemble python tools/train_probe.py aspirin 250 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --csv -c 3000 --log-file $O/r02b_launches_train_aspirin250.csv python tools/train_probe.py aspirin 300 > /dev/null 2>&1
ls -la $O | grep r02b_ | tee -a $O/call21.log
