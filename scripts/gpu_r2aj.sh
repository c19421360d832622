#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kde2d_tc_kernel --launch-skip 3 -c 1 -o gpurun_out/r2aj_k2 -f python scripts/kde2d_ab.py > gpurun_out/r2aj_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2aj_k2.ncu-rep --page source --csv --print-source sass > gpurun_out/r2aj_source_sass.csv 2>/dev/null
rm -f gpurun_out/r2aj_k2.ncu-rep; gzip -f gpurun_out/r2aj_source_sass.csv; ls -la gpurun_out | grep r2aj
