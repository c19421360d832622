#!/bin/bash
# round 2, session S: aligned-pair KDE taps: parity tests, then A/B against the previous deposit kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kde1d.py tests/test_gpu_baseline_sized.py tests/test_gpu_entropy_loss.py tests/test_gpu_edge_cases.py tests/test_gpu_ment.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2s_tests.txt
cat gpurun_out/r2s_tests.txt
bash scripts/ab_bench.sh variants/lib_kdeold.so default variants/lib_kdeold.so default 2>&1 | tee gpurun_out/r2s_ab.txt
