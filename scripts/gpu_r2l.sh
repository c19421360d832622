#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -q 2>&1 | grep -v "^  \|^$" | tail -30 > gpurun_out/r2l_tests.txt
python __graft_entry__.py smoke > gpurun_out/r2l_smoke.txt 2>&1; echo "smoke rc=$?"
tail -14 gpurun_out/r2l_tests.txt; tail -2 gpurun_out/r2l_smoke.txt
