"""Wall time per stage on a small scene (GPU work negligible): exposes host / launch overhead."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, cProfile, pstats
from treedetection_b200 import api, pipeline, synth, _lib

dev = torch.device("cuda:0")
sc = synth.make_scene(seed=1, size_px=int(sys.argv[1]) if len(sys.argv) > 1 else 1000, px=0.2, ndsm_px=0.2)
p = pipeline.PipelineParams()
host = api.HostImage.from_scene(sc, pin=False)
tables = api.TileTables(sc.tiles, dev, 1)
d = {k: getattr(host, k).to(dev) for k in ("rgbi", "ndsm", "boxes_net", "scores", "probs", "inst_tile", "tile_dims")}

def step(timing=None):
    t = [time.perf_counter()]
    def mark():
        if timing is not None:
            torch.cuda.synchronize(); t.append(time.perf_counter())
    tables.plan(d["rgbi"]).run(d["rgbi"]); mark()
    table = pipeline.predict_stage(d["boxes_net"], d["scores"], d["probs"], d["inst_tile"], d["tile_dims"], tables.tile_tf, tables.tile_boxes, p); mark()
    rasters = pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p); mark()
    feats = pipeline.postprocess_stage(table, rasters, p); mark()
    if timing is not None:
        timing.append([1e3 * (b - a) for a, b in zip(t[:-1], t[1:])])
for _ in range(5): step()
torch.cuda.synchronize()
T = []
for _ in range(20): step(T)
import numpy as np
print("instances", len(host.scores), "stage wall ms (P1, P2-4, P5, P6-9):", np.round(np.median(np.array(T), 0), 3))
l0 = _lib.launch_count; step(); print("own kernel launches per step:", _lib.launch_count - l0)
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(20): step()
torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
