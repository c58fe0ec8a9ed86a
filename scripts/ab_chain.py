"""A/B of the chain's kernel variants on the bench workload (resident inputs, chain alone and next to P1).
Diagnostic only.  usage: ab_chain.py [size]   (variants come from the TREEDET_* environment variables)"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from treedetection_b200 import api, pipeline, synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
dev = torch.device("cuda:0")
sc = synth.make_scene(seed=1234, size_px=size, px=0.2, ndsm_px=0.2, density_per_km2=2500.0)
p = pipeline.PipelineParams()
host = api.HostImage.from_scene(sc, pin=False)
tables = api.TileTables(sc.tiles, dev, 1)
d = {k: getattr(host, k).to(dev) for k in ("rgbi", "ndsm", "boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
p1_out = torch.empty((tables.p1_floats,), dtype=torch.float32, device=dev)
det = {k: d[k] for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
runner = pipeline.ChainRunner(p)
bufs = {}
rasters = lambda: pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p, buffers=bufs)
p1_stream = torch.cuda.Stream(device=dev)
chain_stream = torch.cuda.Stream(device=dev, priority=-1)


def run(overlap, steps):
    last = []
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(steps):
        main = torch.cuda.current_stream()
        if overlap:
            p1_stream.wait_stream(main); chain_stream.wait_stream(main)
            with torch.cuda.stream(p1_stream):
                tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
            with torch.cuda.stream(chain_stream):
                t = runner.submit(det, tables.tile_tf, tables.tile_boxes, rasters)
            main.wait_stream(p1_stream); main.wait_stream(chain_stream)
        else:
            t = runner.submit(det, tables.tile_tf, tables.tile_boxes, rasters)
        if last:
            runner.collect(last.pop())
        last.append(t)
    e.record()
    torch.cuda.synchronize()
    n, f = runner.collect(last.pop())
    return s.elapsed_time(e) / steps, n, len(f)


run(False, 3)
a = run(False, 20)
run(True, 3)
b = run(True, 20)
env = {k: v for k, v in os.environ.items() if k.startswith("TREEDET_")}
print(f"{env}: chain alone {a[0]:.3f} ms, chain + P1 overlapped {b[0]:.3f} ms  (cand {a[1]}, crowns {a[2]}, fallbacks {runner.fallbacks})")
