#!/bin/bash
# what produced profiles/r02_*_v5 (final build of round 2): the GPU tests, smoke, the default bench line, the reference
# arm, the ncu launch list of whole serial steps and a --set full capture of the heaviest kernels
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/v5_tests.log 2>&1; echo tests $?; tail -2 gpurun_out/v5_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/v5_smoke.log 2>&1; echo smoke $?; tail -1 gpurun_out/v5_smoke.log
( time timeout 900 python bench.py > gpurun_out/v5_bench.json 2> gpurun_out/v5_bench.err ) 2>&1 | grep real; echo bench $?
( time timeout 900 python bench.py --impl reference > gpurun_out/v5_bench_ref.json 2> gpurun_out/v5_bench_ref.err ) 2>&1 | grep real; echo ref $?
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r2_launches_v5.csv python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --no-files --serial --steps 2 --warmup 3 > gpurun_out/ncu_l_v5.log 2>&1; echo ncu1 $?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile_resize_u8_up_warp|trace_walk_kernel|trace_rings_slots|simplify_kernel|paste_pack_kernel|crown_stats_kernel|decimate_ndvi|nms_adjacency" --launch-skip 40 -c 20 -o gpurun_out/r2_step_v5 -f python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --no-files --serial --steps 2 --warmup 3 > gpurun_out/ncu_f_v5.log 2>&1; echo ncu2 $?
python - <<'PY'
import json
for f in ("gpurun_out/v5_bench.json", "gpurun_out/v5_bench_ref.json"):
    for l in open(f):
        if l.startswith("{"):
            d = json.loads(l)
            print(f, d.get("impl"), round(d["value"], 3), d.get("ms_per_step"), d.get("e2e", {}).get("value"), (d.get("parity") or "")[:30])
PY
