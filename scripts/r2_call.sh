#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 300 python -m pytest tests/test_gpu_nccl_strip.py tests/test_gpu_chain.py -x -q > gpurun_out/n2_tests.log 2>&1; echo tests $?; tail -2 gpurun_out/n2_tests.log
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 5 --no-files > gpurun_out/n2_bench.json 2> gpurun_out/n2_bench.err; echo bench $?
python - <<'PY'
import json
for l in open("gpurun_out/n2_bench.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print(d["n_gpus"], round(d["value"], 1), round(d["ms_per_step"], 3), (d.get("parity") or "NO PARITY")[:40], d.get("e2e", {}).get("value"), d["config"].get("strips_per_step"))
PY
