#!/bin/bash
# GPU tests that cover the changed kernels, a short bench line, ncu --set full (+ source counters) of the two simplify launches and the stats kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1; echo tests $?; tail -3 gpurun_out/c_tests.log
timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-merged --no-files > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo bench $?
python - <<'PY'
import json
for l in open("gpurun_out/c_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"], 1), round(d["ms_per_step"], 3), (d.get("parity") or "NO PARITY")[:40], d["config"].get("stage_ms"))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"simplify_kernel|crown_stats_kernel|decimate_ndvi" --launch-skip 12 -c 5 -o gpurun_out/r2_simplify_v4 -f python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --no-files --serial --steps 2 --warmup 3 > gpurun_out/ncu_s4.log 2>&1; echo ncu $?
ls -la gpurun_out/*.ncu-rep
