#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1; echo tests $?; tail -3 gpurun_out/c_tests.log
bash scripts/ab_bench.sh "|" "|"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c_launches.csv python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --no-files --serial --steps 2 --warmup 3 > gpurun_out/c_ncu.log 2>&1; echo ncu $?
python profiles/summarize_launches.py gpurun_out/c_launches.csv 2>/dev/null | head -10
