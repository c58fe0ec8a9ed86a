timeout 900 python -m pytest tests/test_gpu_pipeline.py tests/test_gpu_chain_dyn.py tests/test_gpu_configs.py tests/test_gpu_config1.py -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-clocks --no-merged --steps 30 --serial 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'])"
timeout 200 python scripts/kernel_timeline.py 10000 5 2>&1 | grep -v Warn | head -16 | cut -c1-150
