#!/bin/bash
# round 2, session AK: KDE-1D backward kernel with padded gradient table + factorised taps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kde1d.py tests/test_gpu_entropy_loss.py tests/test_gpu_baseline_sized.py tests/test_gpu_edge_cases.py -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/r2ak_tests.txt
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:kde1d_bwd_kernel -c 2 --csv --log-file gpurun_out/r2ak.csv python scripts/prof_step.py > gpurun_out/r2ak_ncu.log 2>&1
grep kde1d_bwd gpurun_out/r2ak.csv | awk -F'","' '{print $(NF-2), $NF}' | tr -d '"' | tee gpurun_out/r2ak.txt
