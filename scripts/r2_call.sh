#!/bin/bash
# ncu --set full (+ source counters) of the two simplify launches of one serial step
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"simplify_kernel" --launch-skip 6 -c 2 -o gpurun_out/r2_simplify_v4 -f python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --no-files --serial --steps 2 --warmup 3 > gpurun_out/ncu_s4.log 2>&1; echo ncu $?
ls -la gpurun_out/*.ncu-rep
