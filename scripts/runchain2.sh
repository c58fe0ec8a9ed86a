timeout 200 python scripts/kernel_timeline.py 10000 5 2>&1 | grep -v Warn | cut -c1-170 > gpurun_out/timeline_v6.txt; head -4 gpurun_out/timeline_v6.txt
