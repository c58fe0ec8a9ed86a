"""Golden of the FULL BASELINE config 2 scene -- the bench workload itself: synthetic 10 000 x 10 000 px RGBI
(0.2 m) + nDSM, seed 1234, 1 600 tiles, 34 374 ROI-head instances -- through the CPU oracle (oracle/port.py:
the reference's path restated, pinned to the reference's own functions by the other goldens; the full-size
forms of the statistics / containment / NMS are proven equal to the literal ones in
tests/test_oracle_windowed.py).  Both nDSM variants of the config: 0.2 m (split statistics path) and 1 m
(combined path).  Stored: counts, the crown ids, areas, heights, containment columns, ring lengths of the
final layer, and SHA-256 digests of the stitched table's and the final layer's vertex arrays.

    python tests/golden/make_golden_config2.py        (about 15 minutes on one core)
"""
import hashlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import port  # noqa: E402
from treedetection_b200 import geo, pipeline, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SIZE, SEED, DENSITY = 10000, 1234, 2500.0


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    t0 = time.time()
    p = pipeline.PipelineParams()
    cfg = {k: getattr(p, k) for k in p.__dataclass_fields__}
    g = {}
    rings = conf = None
    for name, ndsm_px in (("split", 0.2), ("combined", 1.0)):
        sc = synth.make_scene(seed=SEED, size_px=SIZE, px=0.2, ndsm_px=ndsm_px, density_per_km2=DENSITY)
        print(name, "scene", round(time.time() - t0), "s", flush=True)
        if rings is None:
            rings, conf = port.predict_stage(sc.det, sc.tiles)
            print("predict_stage", len(rings), "rings", round(time.time() - t0), "s", flush=True)
            tv = np.array([q for r in rings for q in r], dtype=np.float64).reshape(-1, 2)
            g["n_instances"] = np.array([len(sc.det.scores)]); g["n_candidates"] = np.array([len(rings)])
            g["table_ring_len"] = np.array([len(r) for r in rings], dtype=np.int32)
            g["table_conf"] = np.array(conf, dtype=np.float64)
            g["table_verts_sha256"] = np.array(sha(tv))
        H, W = sc.rgbi.shape[1:]
        oh, ow = int(H * p.ndvi_scaling_factor), int(W * p.ndvi_scaling_factor)
        dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in (0, 3)])
        ndvi = port.ndvi_from_rgbi(np.stack([dec[0], dec[0], dec[0], dec[1]])).astype(np.float32)
        ndvi_tf = geo.compose(sc.transform, geo.scale(W / ow, H / oh))
        h, w = sc.ndsm.shape
        out, dbg = port.post_process(rings, conf, ndvi, ndvi_tf, tuple(geo.raster_bounds(sc.transform, W, H)), sc.ndsm,
                                     sc.ndsm_transform, tuple(geo.raster_bounds(sc.ndsm_transform, w, h)), 0.2, 0.2,
                                     cfg, large=True)
        assert dbg["combined"] == (name == "combined")
        print(name, "post_process", len(out), "crowns", round(time.time() - t0), "s", flush=True)
        ov = np.array([q for o in out for q in o["coords"]], dtype=np.float64).reshape(-1, 2)
        g[f"{name}_ndvi_sha256"] = np.array(sha(ndvi))
        g[f"{name}_ids_after_nms"] = np.array(dbg["ids_after_nms"], dtype=np.int64)
        g[f"{name}_poly_id"] = np.array([int(o["poly_id"]) for o in out], dtype=np.int64)
        g[f"{name}_area"] = np.array([o["Area"] for o in out], dtype=np.float64)
        g[f"{name}_height"] = np.array([o["TreeHeight"] for o in out], dtype=np.float32)
        g[f"{name}_centroid"] = np.array([o["Centroid"] for o in out], dtype=np.float32).reshape(-1, 2)
        g[f"{name}_is_contained"] = np.array([o["is_contained"] for o in out])
        g[f"{name}_num_contained"] = np.array([o["num_contained"] for o in out], dtype=np.int32)
        g[f"{name}_ring_len"] = np.array([len(o["coords"]) for o in out], dtype=np.int32)
        g[f"{name}_verts_sha256"] = np.array(sha(ov))
    g["workload"] = np.array([SIZE, SEED, int(DENSITY)])
    np.savez_compressed(os.path.join(OUT, "config2.npz"), **g)
    print("done", round(time.time() - t0), "s")


if __name__ == "__main__":
    main()
