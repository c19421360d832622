#!/bin/bash
# source-level profile of one forward flow layer launch (stall samples per SASS instruction)
mkdir -p gpurun_out
timeout 500 ncu --set full --clock-control none --import-source on -k regex:nsf_tc_layer_kernel --launch-skip 12 -c 1 -o gpurun_out/r2d_layer -f python scripts/prof_step.py > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i gpurun_out/r2d_layer.ncu-rep --page raw --csv > gpurun_out/r2d_raw.csv 2>/dev/null
ncu -i gpurun_out/r2d_layer.ncu-rep --page source --csv --print-source sass > gpurun_out/r2d_source_sass.csv 2>/dev/null
ncu -i gpurun_out/r2d_layer.ncu-rep --page source --csv --print-source cuda > gpurun_out/r2d_source_cuda.csv 2>/dev/null
rm -f gpurun_out/r2d_layer.ncu-rep
gzip -f gpurun_out/r2d_source_sass.csv gpurun_out/r2d_source_cuda.csv
ls -la gpurun_out | tail -8
