#!/bin/bash
# round 2, session U: KDE-1D deposit with packed tap recurrences, warp-blocked bins, unchecked main loop
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kde1d.py tests/test_gpu_baseline_sized.py tests/test_gpu_entropy_loss.py tests/test_gpu_edge_cases.py tests/test_gpu_ment.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2u_tests.txt
cat gpurun_out/r2u_tests.txt
true
