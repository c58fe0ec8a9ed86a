"""Diagnostic: where the FIRST image of a process spends its time (cold CUDA context, lazy module loading,
allocations) against the second.  Wraps the functions of treedetection_b200.ops / pipeline with a device
synchronisation on both sides; prints per-function seconds for call 1 and call 2 of api.run_image."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

t_start = time.time()
from treedetection_b200 import api, ops, pipeline, synth  # noqa: E402

size = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
sc = synth.make_scene(seed=1234, size_px=size, px=0.2, ndsm_px=0.2, density_per_km2=2500.0, stem="FDOP20_000000_rgbi")
acc, depth = {}, [0]


def wrap(mod, name):
    fn = getattr(mod, name)

    def inner(*a, **k):
        if depth[0] > 0:
            return fn(*a, **k)
        depth[0] += 1
        torch.cuda.synchronize()
        t = time.time()
        try:
            return fn(*a, **k)
        finally:
            torch.cuda.synchronize()
            acc.setdefault(f"{mod.__name__.split('.')[-1]}.{name}", []).append(time.time() - t)
            depth[0] -= 1
    setattr(mod, name, inner)


t0 = time.time()
torch.cuda.init()
torch.zeros(1, device="cuda")
print(f"cuda context {time.time() - t0:.3f} s")
for name in dir(ops):
    f = getattr(ops, name)
    if callable(f) and getattr(f, "__module__", "") == ops.__name__ and not isinstance(f, type) and not name.startswith("_"):
        wrap(ops, name)
dev = torch.device("cuda", 0)
p = pipeline.PipelineParams()
t0 = time.time()
host = api.HostImage.from_scene(sc)
print(f"pinned host image {time.time() - t0:.3f} s")
t0 = time.time()
tables = api.TileTables(sc.tiles, dev, p.shift)
p1_out = torch.empty((tables.p1_floats,), dtype=torch.float32, device=dev)
torch.cuda.synchronize()
print(f"tile tables + P1 buffer {time.time() - t0:.3f} s")
runner = pipeline.ChainRunner(p)
for k in range(3):
    acc.clear()
    torch.cuda.synchronize()
    t0 = time.time()
    out, _ = api.run_image(host, p, dev, tables, p1_out, runner=runner, want_table=True)
    dt = time.time() - t0
    inside = sum(sum(v) for v in acc.values())
    print(f"run_image call {k + 1}: {dt:.3f} s (wrapped ops {inside:.3f} s)")
    for name, v in sorted(acc.items(), key=lambda kv: -sum(kv[1]))[:14]:
        print(f"    {name:28s} {sum(v):.4f} s  x{len(v)}")
