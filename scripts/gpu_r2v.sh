#!/bin/bash
# round 2, session V: warpgroup-major tile dealing and programmatic dependent launch of the flow layers
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_nsf.py tests/test_gpu_entropy_loss.py tests/test_gpu_randn.py -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2v_tests.txt
cat gpurun_out/r2v_tests.txt
bash scripts/ab_bench.sh variants/lib_base.so variants/lib_deal.so default variants/lib_base.so variants/lib_deal.so default 2>&1 | tee gpurun_out/r2v_ab.txt
