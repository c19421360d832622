#!/bin/bash
# round 2, GPU session A: parity of the re-derived spline (three compensation variants), Philox, A/B timing
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | tail -40 > gpurun_out/r2a_tests.txt
for v in default comp1 comp2; do
  if [ $v = default ]; then unset MENTFLOW_B200_LIB; else export MENTFLOW_B200_LIB=$PWD/variants/lib_$v.so; fi
  timeout 600 python scripts/tc_stats.py > gpurun_out/r2a_stats_$v.txt 2>&1
done
unset MENTFLOW_B200_LIB
scripts/ab_bench.sh default variants/lib_comp1.so variants/lib_comp2.so default > gpurun_out/r2a_ab.txt 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err
tail -5 gpurun_out/r2a_tests.txt; cat gpurun_out/r2a_ab.txt
