timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --steps 50 > gpurun_out/bench_r01_final.json 2> gpurun_out/bench_r01_final.err; tail -c 300 gpurun_out/bench_r01_final.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_v8.csv python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --serial --steps 2 --warmup 3 > gpurun_out/ncu_l8.log 2>&1; echo ncu1 $?
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"tile_resize_u8_up_warp|trace_walk_kernel|trace_rings_slots|simplify_kernel|paste_pack_kernel|crown_stats_kernel|decimate_kernel|nms_adjacency" --launch-skip 40 -c 20 -o gpurun_out/step_v11 -f python bench.py --no-cpu-baseline --no-clocks --no-merged --no-alone --serial --steps 2 --warmup 3 > gpurun_out/ncu_f8.log 2>&1; echo ncu2 $?
python -c "import __graft_entry__ as g; g.smoke()"
