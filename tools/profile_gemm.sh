#!/bin/bash
# ncu capture of the big (k = 512) trailing-update GEMM inside potrf, aspirin M=500 (n = 31500)
ncu --clock-control none --set full --import-source on -k regex:k_gemm_nt -s 22 -c 1 -f -o gpurun_out/prof_gemm_k512 \
  python tools/train_probe.py aspirin 500 > /dev/null 2>&1
