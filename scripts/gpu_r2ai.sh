#!/bin/bash
# round 2, session AI: KDE-2D tensor-core kernel with the A operand in tensor memory
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_kde2d.py tests/test_gpu_baseline_sized.py -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2ai_tests.txt
timeout 200 python scripts/kde2d_ab.py 2>&1 | tail -9 | tee gpurun_out/r2ai_ab.txt
