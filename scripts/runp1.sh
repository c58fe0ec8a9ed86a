for kb in 60 76 46; do
TREEDET_P1_SMEM_KB=$kb timeout 300 python bench.py --no-cpu-baseline --no-clocks --no-merged --steps 40 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print('RES kb $kb', d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['ms_per_launch'])"
done
