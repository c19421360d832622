#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do
MFB_AB_NOTAIL=1 MFB_AB_NOENT=1 bash scripts/ab_bench.sh default 2>&1 | sed 's/^/old  /'
MFB_AB_NOTAIL=1 bash scripts/ab_bench.sh default 2>&1 | sed 's/^/ent  /'
bash scripts/ab_bench.sh default 2>&1 | sed 's/^/both /'
done | tee gpurun_out/r2w_ab2.txt
