timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
( time timeout 900 python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err ) 2>&1 | grep real
python -c "
import json; d=json.loads(open('gpurun_out/bench_full.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], 'p1frac', d['roofline']['frac'], 'path', d['path_roofline']['frac'], 'e2e', d['e2e'], d.get('cpu_baseline'), d['clocks'], d['gpu_launches'])"
( time timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err ) 2>&1 | grep real
tail -c 1500 gpurun_out/bench_ref.json
