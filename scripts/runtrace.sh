export TREEDET_TRACE_SMEM=27648
for pad in 0 17000 36000 55000; do
echo "pad $pad"; TREEDET_P1_SMEM_PAD=$pad timeout 300 python bench.py --no-cpu-baseline --no-clocks --no-merged --steps 30 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'])"
done
