#!/bin/bash
mkdir -p gpurun_out
timeout 400 ncu --set full --clock-control none --import-source on -k regex:kde1d_deposit_kernel --launch-skip 1 -c 1 -o gpurun_out/r2r_dep -f python scripts/prof_step.py > gpurun_out/r2r_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2r_dep.ncu-rep --page source --csv --print-source sass > gpurun_out/r2r_source_sass.csv 2>/dev/null
rm -f gpurun_out/r2r_dep.ncu-rep; gzip -f gpurun_out/r2r_source_sass.csv; ls -la gpurun_out | grep r2r
