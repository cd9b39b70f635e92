#!/bin/bash
# predict-only lines of BASELINE configs 3-5 (random coefficients) with the final build: FP64 (default) and 5 int8 slices
mkdir -p gpurun_out/r02
O=gpurun_out/r02
run() { local name=$1 to=$2; shift 2; echo "=== $name" | tee -a $O/call28.log; timeout $to "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/call28.log; tail -n 2 $O/$name.log | cut -c1-300 | tee -a $O/call28.log; return $rc; }
for wl in ac-ala3-nhme c60 synthetic100; do
  for s in 0 5; do SGDML_B200_OZAKI_PREDICT_SLICES=$s run c28_bench_${wl}_s$s 200 python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline --no-extras; done
done
