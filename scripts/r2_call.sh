#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1; echo tests $?; tail -3 gpurun_out/c_tests.log
bash scripts/ab_bench.sh "|" "TREEDET_STATS_ROWS=8|" "|"
