#!/bin/bash
# Round-2 validation + evidence: full GPU suite, default bench line, ncu launch lists and full captures of the top kernels.
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call7.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call7.log; tail -n 6 $O/$name.log | tee -a $O/call7.log; return $rc; }
run c7_gpu_tests 1500 python -m pytest tests -q -m gpu
run c7_smoke 300 python __graft_entry__.py smoke
run c7_bench 900 python bench.py
run c7_bench_ref 600 python bench.py --impl reference --steps 2 --warmup 1
NCU="ncu --clock-control none"
# launch lists (device time per launch; cold-cache, serialised: compare shares, not absolutes)
timeout 600 $NCU --metrics gpu__time_duration.sum --csv --log-file $O/r02_launches_bench_aspirin_predict.csv python bench.py --no-train --no-cpu-baseline --no-extras --steps 2 --warmup 1 > /dev/null 2>&1
timeout 600 $NCU --metrics gpu__time_duration.sum --csv -c 3000 --log-file $O/r02_launches_train_aspirin250.csv python tools/train_probe.py aspirin 300 > /dev/null 2>&1
# full captures of the top kernels
timeout 600 $NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o $O/r02_predict_aspirin python bench.py --no-train --no-cpu-baseline --no-extras --steps 1 --warmup 1 --batch 24576 > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_predict_main -s 3 -c 1 -f -o $O/r02_predict_ethanol python bench.py --workload ethanol --no-train --no-cpu-baseline --no-extras --steps 1 --warmup 1 > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_ozaki_gemm -s 6 -c 1 -f -o $O/r02_ozaki_trailing python tools/train_probe.py aspirin 500 > /dev/null 2>&1
timeout 600 $NCU --set full --import-source on -k regex:k_assemble -s 1 -c 1 -f -o $O/r02_assemble python tools/train_probe.py aspirin 250 > /dev/null 2>&1
ls -la $O | grep r02_
