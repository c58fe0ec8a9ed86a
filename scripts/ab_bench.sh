#!/bin/bash
# A/B of bench.py variants on one GPU: prints value / ms_per_step / stage_ms per variant (diagnostic, not a bench value)
# usage: scripts/ab_bench.sh "<ENV=.. ENV=.. | flags>" ...      (the part before '|' is optional)
i=0
for spec in "$@"; do
  i=$((i+1))
  envs=""; flags="$spec"
  if [[ "$spec" == *"|"* ]]; then envs="${spec%%|*}"; flags="${spec#*|}"; fi
  env $envs timeout 250 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-merged --no-files $flags \
      > gpurun_out/ab_$i.json 2> gpurun_out/ab_$i.err || tail -3 gpurun_out/ab_$i.err | cut -c1-300
  python - "$spec" gpurun_out/ab_$i.json <<'PY'
import json, sys
for l in open(sys.argv[2]):
    if l.startswith("{"):
        d = json.loads(l)
        print(sys.argv[1], "=>", round(d["value"], 1), round(d["ms_per_step"], 3), (d.get("parity") or "NO PARITY")[:12],
              "p1_alone", round(d["roofline"].get("ms_per_launch", 0), 3) if "roofline" in d else None,
              {k[:5]: v for k, v in d["config"]["stage_ms"].items()})
PY
done
