"""In-situ kernel durations of one resident step (CUPTI through torch.profiler): per-kernel
mean time, GPU busy time per step and the idle gaps between kernels.  Diagnostic only -- numbers
taken under a profiler are never bench values."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import collections

import torch
from torch.profiler import ProfilerActivity, profile

from treedetection_b200 import api, pipeline, synth

size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
dev = torch.device("cuda:0")
sc = synth.make_scene(seed=1234, size_px=size, px=0.2, ndsm_px=0.2, density_per_km2=2500.0)
p = pipeline.PipelineParams()
host = api.HostImage.from_scene(sc, pin=False)
tables = api.TileTables(sc.tiles, dev, 1)
d = {k: getattr(host, k).to(dev) for k in ("rgbi", "ndsm", "boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
p1_out = torch.empty((tables.p1_floats,), dtype=torch.float32, device=dev)


runner = pipeline.ChainRunner(p)
det = {k: d[k] for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
last = []


overlap = len(sys.argv) > 3 and sys.argv[3] == "overlap"
p1_stream = torch.cuda.Stream(device=dev)
chain_stream = torch.cuda.Stream(device=dev, priority=-1)


def step():
    if overlap:      # as bench.py's default step: P1 next to the chain
        main = torch.cuda.current_stream()
        p1_stream.wait_stream(main); chain_stream.wait_stream(main)
        with torch.cuda.stream(p1_stream):
            tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
        with torch.cuda.stream(chain_stream):
            t = runner.submit(det, tables.tile_tf, tables.tile_boxes,
                              lambda: pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p))
        main.wait_stream(p1_stream); main.wait_stream(chain_stream)
        if last:
            runner.collect(last.pop())
        last.append(t)
        return
    tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
    t = runner.submit(det, tables.tile_tf, tables.tile_boxes,
                      lambda: pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p))
    if last:
        runner.collect(last.pop())
    last.append(t)


for _ in range(5):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(steps):
        step()
    torch.cuda.synchronize()

evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
agg = collections.OrderedDict()
for e in evs:
    a = agg.setdefault(e.name, [0, 0.0])
    a[0] += 1
    a[1] += e.time_range.elapsed_us()
busy = sum(v[1] for v in agg.values())
span = evs[-1].time_range.end - evs[0].time_range.start
print(f"steps {steps}: span {span / steps / 1e3:.3f} ms/step, GPU busy {busy / steps / 1e3:.3f} ms/step, "
      f"idle {(span - busy) / steps / 1e3:.3f} ms/step, {len(evs) / steps:.0f} device ops/step")
# idle attributed to the op that FOLLOWS the gap
gap = collections.defaultdict(float)
for a, b in zip(evs[:-1], evs[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 0:
        gap[b.name] += g
print("\n   us/step   n/step   us/call   gap-before us/step   kernel")
for name, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / steps:10.1f} {n / steps:8.1f} {us / n:9.1f} {gap[name] / steps:12.1f}   {name[:110]}")
