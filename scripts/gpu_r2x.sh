#!/bin/bash
# round 2, session X: compact dL/dphi rows + ReLU bit masks in the tensor-core backward
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nsf.py tests/test_gpu_entropy_loss.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2x_tests.txt
cat gpurun_out/r2x_tests.txt
bash scripts/ab_bench.sh variants/lib_prev.so default variants/lib_prev.so default 2>&1 | tee gpurun_out/r2x_ab.txt
