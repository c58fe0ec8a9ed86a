for m in "--p1-after-walk" ""; do
timeout 300 python bench.py --no-cpu-baseline --no-clocks --no-merged --steps 40 $m 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('RES $m', d['value'], d['ms_per_step'], d['config']['stage_ms'])"
done
