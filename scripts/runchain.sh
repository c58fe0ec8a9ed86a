timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -12
timeout 300 python bench.py --no-cpu-baseline --no-clocks --no-merged --steps 30 --serial 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['config']['chain'][-60:])"
timeout 300 python bench.py --no-cpu-baseline --no-clocks --no-merged --steps 30 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['config']['chain'][-60:])"
