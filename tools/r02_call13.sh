#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call13.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call13.log; tail -n 6 $O/$name.log | tee -a $O/call13.log; return $rc; }
run c13_gpu_tests 1500 python -m pytest tests -q -m gpu -x
run c13_asm 300 python tools/asm_variants.py
run c13_cg_acala_m1000 900 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 1000 --max-memory 60 --profile
run c13_bench 900 python bench.py
run c13_smoke 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')"
