#!/bin/bash
# round 2, session T: packed fp32 (FMUL2 / FADD2) in the flow kernel: parity tests, host-spline test, A/B
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nsf.py tests/test_gpu_baseline_sized.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2t_tests.txt
cat gpurun_out/r2t_tests.txt
bash scripts/ab_bench.sh variants/lib_nox2.so default variants/lib_nox2.so default 2>&1 | tee gpurun_out/r2t_ab.txt
