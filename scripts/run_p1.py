"""Run only P1 (tile cut + normalise) on the bench workload: for ncu captures and A/B timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from treedetection_b200 import api, ops, synth, tiling

size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
dev = torch.device("cuda:0")
rng = np.random.default_rng(0)
rgbi = torch.from_numpy(rng.integers(0, 256, size=(4, size, size), dtype=np.uint8)).to(dev)
tf = synth.image_transform(synth.ORIGIN_X, synth.ORIGIN_Y + size * 0.2, 0.2)
tiles = tiling.tile_grid("x", tf, size, size, 25832, 50, 50, 20)
win = torch.tensor([m["window"] for m in tiles.values()], dtype=torch.int32)
net = torch.tensor([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in win.tolist()], dtype=torch.int32)
plan = ops.TilePlan(win, net, 1, size, size)
out = torch.empty((plan.total,), dtype=torch.float32, device=dev)
nbytes = int(sum(3 * int(w[2]) * int(w[3]) for w in win.tolist())) + 4 * plan.total
for _ in range(3):
    plan.run(rgbi, out)
torch.cuda.synchronize()
ts = []
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.run(rgbi, out); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ts = np.array(ts)
print(f"P1 {len(tiles)} tiles: median {np.median(ts):.3f} ms  min {ts.min():.3f} ms  {nbytes/1e9:.2f} GB -> "
      f"{nbytes/np.median(ts)/1e6:.0f} GB/s (median)")
