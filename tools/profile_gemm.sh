#!/bin/bash
# ncu capture of the big (k = NBO = 1024) trailing-update GEMM inside potrf, aspirin M=500 (n = 31500).
# Step 1 lists the GEMM launches; step 2 captures the first one with a grid larger than 20000 CTAs.
NCU="ncu --clock-control none"
$NCU --metrics gpu__time_duration.sum --csv -k regex:k_gemm_nt -c 120 --log-file gpurun_out/gemm_list.csv \
  python tools/train_probe.py aspirin 500 > /dev/null 2>&1
IDX=$(python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/gemm_list.csv')))
for i,r in enumerate(rows):
    if r and r[0]=='ID': hdr=r; start=i+1; break
gi=hdr.index('Grid Size')
for n,r in enumerate(rows[start:]):
    g=int(r[gi].strip('()').split(',')[0])
    if g>20000:
        print(n); break
PY
)
echo "first big trailing GEMM is k_gemm_nt* launch #$IDX"
$NCU --set full --import-source on -k regex:k_gemm_nt -s $IDX -c 1 -f -o gpurun_out/r01f_gemm_trailing \
  python tools/train_probe.py aspirin 500 > /dev/null 2>&1
