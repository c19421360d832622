#!/bin/bash
# round 2, session Z: wgrad kernel: loader warps / staging depth A/B (training step), kernel time via bwd_prof under ncu
mkdir -p gpurun_out
for v in default s3b l28 l28s3b; do
  if [ $v = default ]; then unset MENTFLOW_B200_LIB; else export MENTFLOW_B200_LIB=$PWD/variants/lib_$v.so; fi
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'nsf_tc_wgrad_kernel' --launch-skip 5 -c 5 --csv --log-file gpurun_out/r2z_$v.csv python scripts/bwd_prof.py > gpurun_out/r2z_ncu.log 2>&1
  echo $v $(grep wgrad gpurun_out/r2z_$v.csv | awk -F'","' '{gsub(/[",]/,"",$NF); print $NF}' | tr '\n' ' ')
done | tee gpurun_out/r2z_wgrad.txt
