"""Diagnostic: cProfile of process_files on N synthetic images (bench.e2e_files), main thread only.
usage: python scripts/profile_files.py [n_images] [size_px]"""
import cProfile
import io
import json
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from treedetection_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
size = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
sc = synth.make_scene(seed=1234, size_px=size, px=0.2, ndsm_px=0.2, density_per_km2=2500.0, stem="FDOP20_000000_rgbi")
pr = cProfile.Profile()
pr.enable()
out = bench.e2e_files(sc, n, "profile")
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(70)
print(s.getvalue()[:14000])
print(json.dumps({k: out[k] for k in ("value", "wall_s", "steady_state", "stage_s")}))
