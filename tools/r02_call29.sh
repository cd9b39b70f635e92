#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
timeout 200 ncu --clock-control none --set full --import-source on -k regex:k_ozaki_gemm -s 6 -c 1 -f -o $O/r02b_ozaki_trailing python tools/train_probe.py aspirin 500 > $O/c29_ncu.log 2>&1
ls -la $O | grep r02b_ozaki
