timeout 900 python -m pytest tests/test_gpu_chain_dyn.py -m gpu -x -q 2>&1 | tail -3
timeout 200 python scripts/kernel_timeline.py 10000 5 2>&1 | grep -v Warn | head -60 | cut -c1-150
