#!/bin/bash
mkdir -p gpurun_out
scripts/ab_bench.sh default variants/lib_poly1.so variants/lib_poly2.so default variants/lib_poly1.so variants/lib_poly2.so > gpurun_out/r2q_ab.txt 2>&1
cat gpurun_out/r2q_ab.txt
