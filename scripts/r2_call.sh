#!/bin/bash
mkdir -p gpurun_out
timeout 400 python bench.py > gpurun_out/v7_bench.json 2> gpurun_out/v7_bench.err; echo bench $?
python - <<'PY'
import json
for l in open("gpurun_out/v7_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"], 1), round(d["ms_per_step"], 3), d["e2e"]["value"], d["combined_path"]["value"], d["path_roofline"]["frac"], (d.get("parity") or "")[:30])
PY
