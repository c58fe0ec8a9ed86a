timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python bench.py --no-cpu-baseline --no-clocks --steps 40 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('RES', d['value'], d['ms_per_step'], d['config']['stage_ms'], d['path_roofline']['frac'], d['crowns_merged']['value'])"
