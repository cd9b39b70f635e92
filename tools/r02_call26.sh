#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call26.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call26.log; tail -n 6 $O/$name.log | tee -a $O/call26.log; return $rc; }
run c26_tests 900 python -m pytest tests/test_ozaki.py tests/test_iterative.py tests/test_dropin_cli.py -q -m gpu -x
run c26_cg_acala_m500 400 python tools/cg_probe.py --workload ac-ala3-nhme --n-train 500 --max-memory 8
