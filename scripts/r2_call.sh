#!/bin/bash
mkdir -p gpurun_out
( time timeout 900 python bench.py > gpurun_out/v6_bench.json 2> gpurun_out/v6_bench.err ) 2>&1 | grep real; echo bench $?
tail -3 gpurun_out/v6_bench.err
python - <<'PY'
import json
for l in open("gpurun_out/v6_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(round(d["value"], 1), round(d["ms_per_step"], 3), d["e2e"]["value"], d["combined_path"]["value"], d["combined_path"]["ms_per_step"], d["combined_path"].get("exact_size_fallbacks_incl_warmup"), (d.get("parity") or "")[:30])
PY
