#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/c_tests.log 2>&1; echo tests $?; tail -2 gpurun_out/c_tests.log
timeout 200 python bench.py --steps 100 --warmup 5 --no-cpu-baseline --no-merged --no-files > gpurun_out/c_bench.json 2> gpurun_out/c_bench.err; echo bench $?
python - <<'PY'
import json
for l in open("gpurun_out/c_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"], 1), round(d["ms_per_step"], 3), (d.get("parity") or "NO PARITY")[:40], d["config"].get("stage_ms"), d["e2e"]["value"])
PY
