#!/bin/bash
mkdir -p gpurun_out
for v in dense band; do
  if [ $v = dense ]; then export MENTFLOW_B200_LIB=$PWD/variants/lib_dense.so; else unset MENTFLOW_B200_LIB; fi
  timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_membar_per_issue_active.ratio,smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio --clock-control none -k regex:kde2d_tc_kernel --launch-skip 3 -c 1 --csv --log-file gpurun_out/r2ad_$v.csv python scripts/kde2d_ab.py > gpurun_out/r2ad_ncu.log 2>&1
  echo "== $v"; grep kde2d_tc_kernel gpurun_out/r2ad_$v.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"'
done | tee gpurun_out/r2ad.txt
