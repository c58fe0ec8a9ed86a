timeout 600 python -m pytest tests/test_gpu_raster.py tests/test_gpu_api.py -m gpu -x -q 2>&1 | tail -2
for R in 25 20 40; do
echo "rows $R"; TREEDET_P1_ROWS=$R timeout 300 python bench.py --no-cpu-baseline --no-clocks --steps 30 --serial 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['path_roofline']['frac'])"
done
echo overlapped; timeout 300 python bench.py --no-cpu-baseline --no-clocks --steps 30 2>&1 | tail -1 > gpurun_out/b.json; python -c "
import json; d=json.load(open('gpurun_out/b.json')); print(d['value'], d['ms_per_step'], d['config']['stage_ms'], d['roofline']['frac'], d['path_roofline']['frac'])"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tile_resize_u8_up_warp" --launch-skip 7 -c 1 -o gpurun_out/p1_v6d -f python bench.py --no-cpu-baseline --no-clocks --steps 2 --warmup 3 --serial > gpurun_out/ncu_p1v6.log 2>&1; echo $?
