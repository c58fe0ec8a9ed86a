"""Diagnostic: td_ndvi_decimate alone (CUDA events, 50 launches) on a 10 000^2 RGBI raster."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from treedetection_b200 import ops
dev = torch.device("cuda", 0)
rgbi = torch.randint(0, 255, (4, 10000, 10000), dtype=torch.uint8, device=dev)
out = torch.empty((2000, 2000), dtype=torch.float32, device=dev)
for _ in range(5):
    ops.ndvi_decimate(rgbi, 2000, 2000, out=out)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(50):
    ops.ndvi_decimate(rgbi, 2000, 2000, out=out)
b.record()
torch.cuda.synchronize()
print("ndvi_decimate ms", a.elapsed_time(b) / 50, os.environ.get("TREEDET_DECIMATE_GENERIC"))
