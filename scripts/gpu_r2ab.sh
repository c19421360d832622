#!/bin/bash
# round 2, session AB: programmatic dependent launch along the forward chain + leaner host side of a replay
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/r2ab_tests.txt
bash scripts/ab_bench.sh variants/lib_nopdl.so default variants/lib_nopdl.so default 2>&1 | tee gpurun_out/r2ab_ab.txt
