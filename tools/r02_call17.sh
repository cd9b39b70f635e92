#!/bin/bash
# Round-2 final validation + evidence: full GPU suite, smoke, both bench arms, kernel-variant probes, ncu launch lists and
# full captures of the (new) default predictor kernels and the assembly kernel.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call17.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call17.log; tail -n 8 $O/$name.log | tee -a $O/call17.log; return $rc; }
run c17_gpu_tests 1500 python -m pytest tests -q -m gpu
run c17_smoke 300 python __graft_entry__.py smoke
run c17_asm 400 python tools/asm_variants.py
VARIANT_SHAPES=aspirin,ethanol run c17_variants 400 python tools/predict_variants.py
run c17_bench 900 python bench.py
run c17_bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
NCU="ncu --clock-control none"
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/r02b_launches_bench_aspirin_predict.csv python bench.py --no-train --no-cpu-baseline --no-extras --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o $O/r02b_predict_aspirin python bench.py --no-train --no-cpu-baseline --no-extras --steps 1 --warmup 1 --batch 24576 > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o $O/r02b_predict_ethanol python bench.py --workload ethanol --no-train --no-cpu-baseline --no-extras --steps 1 --warmup 1 > /dev/null 2>&1
ls -la $O | grep r02b_ | tee -a $O/call17.log
