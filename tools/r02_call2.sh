#!/bin/bash
mkdir -p gpurun_out/r02
O=gpurun_out/r02
for dbg in 0 1 2 3 4 5 7; do
  SGDML_B200_OZAKI_DBG=$dbg timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/oz_probe2.log
done
SGDML_B200_OZAKI_BK=128 timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/oz_probe2.log
OZ_S=4 timeout 120 python tools/ozaki_probe2.py 2>&1 | tee -a $O/oz_probe2.log
OZ_N=2048 timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_ozaki_gemm -s 1 -c 1 -f -o $O/oz_gemm python tools/ozaki_probe2.py > $O/oz_ncu.log 2>&1
ls -la $O
