#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_nsf.py tests/test_gpu_randn.py tests/test_gpu_entropy_loss.py -q -k "backward_tensor or graph_replay or covariance or benchmark_weights" 2>&1 | grep -v "^  \|^$" | tail -60 > gpurun_out/r2b_tests.txt
for v in comp1 comp2; do
  export MENTFLOW_B200_LIB=$PWD/variants/lib_$v.so
  timeout 600 python scripts/tc_stats.py 2>&1 | grep -E "^golden|^bench|tcgen05" > gpurun_out/r2b_stats_$v.txt
done
unset MENTFLOW_B200_LIB
scripts/ab_bench.sh default variants/lib_comp1.so variants/lib_comp2.so default variants/lib_comp1.so > gpurun_out/r2b_ab.txt 2>&1
cat gpurun_out/r2b_tests.txt | tail -30; cat gpurun_out/r2b_ab.txt
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -3 gpurun_out/r2b_bench.err
