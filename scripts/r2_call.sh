#!/bin/bash
# GPU tests, a short bench line, the ncu launch list of two serial steps
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
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/c_launches.csv python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --no-files --serial --steps 2 --warmup 3 > gpurun_out/c_ncu.log 2>&1; echo ncu $?
python profiles/summarize_launches.py gpurun_out/c_launches.csv 2>/dev/null | head -16
