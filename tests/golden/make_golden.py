"""Generate tests/golden/*.npz by running the REFERENCE'S OWN functions
(/root/reference/TreeDetection/postprocessing.py, utilities.py, helpers.py, loaded
unmodified through oracle.refshim) on seeded synthetic inputs.

Run in the build container only (the reference is not on the GPU box):
    python tests/golden/make_golden.py
The fixtures hold inputs AND reference outputs, so the tests that consume them
(tests/test_oracle_golden.py on CPU, tests/test_gpu_pipeline.py on the B200) need
neither the reference nor this script.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import port, refshim  # noqa: E402
from treedetection_b200 import geo, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

CFG = dict(tile_width=50, tile_height=50, buffer=20, use_overlap=True, overlapping_tiles_width=3,
           overlapping_tiles_height=3, confidence_threshold=0.3, containment_threshold=0.75, height_threshold=3,
           ndvi_mean_threshold=0.1, ndvi_var_threshold=0.1, iou_threshold=0.6, area_threshold=1,
           ndvi_scaling_factor=0.2, height_scaling_factor=1.0)


def ragged(rings):
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.array([p for r in rings for p in r], dtype=np.float64).reshape(-1, 2)
    return verts, off


class _Poly:
    def __init__(self, b):
        self.bounds = tuple(b)


def golden_nms(ns):
    rng = np.random.default_rng(42)
    cases = {}
    for name, n, ext, iou, athr in (("a", 400, 60.0, 0.6, 1), ("b", 900, 80.0, 0.2, 1), ("c", 700, 60.0, 0.05, 0.4)):
        cx = 412000 + rng.uniform(0, ext, n); cy = 5318000 + rng.uniform(0, ext, n)
        rx = rng.uniform(1.5, 6, n); ry = rx * rng.uniform(0.8, 1.2, n)
        bounds = np.stack([cx - rx, cy - ry, cx + rx, cy + ry], 1)
        conf = np.round(rng.uniform(0.3, 1.0, n), 2)
        area = np.pi * rx * ry
        ids = [str(i) for i in range(n)]
        ret = ns.postprocessing.filter_polygons_by_iou_and_area(
            {i: _Poly(b) for i, b in zip(ids, bounds)}, dict(zip(ids, area.tolist())), dict(zip(ids, conf.tolist())),
            iou, athr)
        removed = np.array([i not in ret for i in ids])
        cases[f"nms_{name}_bounds"] = bounds; cases[f"nms_{name}_conf"] = conf; cases[f"nms_{name}_area"] = area
        cases[f"nms_{name}_params"] = np.array([iou, athr], dtype=np.float64)
        cases[f"nms_{name}_removed"] = removed
    return cases


def features_from(rings, conf, ids):
    return [{"type": "Feature", "properties": {"poly_id": str(i), "Confidence_score": c},
             "geometry": {"type": "Polygon", "coordinates": [[list(p) for p in r]]}}
            for r, c, i in zip(rings, conf, ids)]


def golden_scene(ns, name, ndsm_px, seed):
    """Full post-processing of one small scene with the reference's process_features."""
    sc = synth.make_scene(seed=seed, size_px=1000, px=0.2, ndsm_px=ndsm_px, density_per_km2=6000.0)
    cfg = dict(CFG)
    refshim.set_config(ns, **cfg)
    rings, conf = port.predict_stage(sc.det, sc.tiles)
    # rasters as process_geojson reads them (postprocessing.py:780-800); decimation = oracle restatement
    H, W = sc.rgbi.shape[1:]
    oh, ow = int(H * cfg["ndvi_scaling_factor"]), int(W * cfg["ndvi_scaling_factor"])
    dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in range(4)])
    ndvi64 = ns.helpers.ndvi_array_from_rgbi(dec)                       # reference (numba)
    ndvi_tf = geo.compose(sc.transform, geo.scale(W / ow, H / oh))
    ndvi_bounds = geo.raster_bounds(sc.transform, W, H)
    height = sc.ndsm
    h, w = height.shape
    height_tf = geo.compose(sc.ndsm_transform, geo.scale(1.0, 1.0))
    height_bounds = geo.raster_bounds(sc.ndsm_transform, w, h)
    # head of process_geojson (postprocessing.py:736-778), reference geometry via the shim
    feats = [(r, c) for r, c in zip(rings, conf) if c >= cfg["confidence_threshold"]]
    ids = list(range(len(feats)))
    features = features_from([f[0] for f in feats], [f[1] for f in feats], ids)
    id_to_area = {}
    for f in features:
        poly = ns.postprocessing.shape(f["geometry"]).simplify(2)
        id_to_area[f["properties"]["poly_id"]] = ns.postprocessing.calculate_area(poly)
    polygon_dict = {f["properties"]["poly_id"]: ns.postprocessing.shape(f["geometry"]) for f in features}
    features = [f for f in features if id_to_area[f["properties"]["poly_id"]] >= cfg["area_threshold"]]
    features = [f for f in features if id_to_area[f["properties"]["poly_id"]] <= 1000]
    fids = {f["properties"]["poly_id"] for f in features}
    polygon_dict = {k: v for k, v in polygon_dict.items() if k in fids}
    confs = {f["properties"]["poly_id"]: f["properties"]["Confidence_score"] for f in features}
    retained = ns.postprocessing.filter_polygons_by_iou_and_area(polygon_dict, id_to_area, confs,
                                                                 cfg["iou_threshold"], cfg["area_threshold"])
    features = [f for f in features if f["properties"]["poly_id"] in retained]
    A = ns.Affine
    out = ns.postprocessing.process_features(
        features, id_to_area, height, A(*height_tf), ns.BoundingBox(*height_bounds), ndvi64, A(*ndvi_tf),
        ns.BoundingBox(*ndvi_bounds), abs(sc.transform[0]), abs(sc.transform[4]))
    verts, off = ragged(rings)
    overts, ooff = ragged([f["geometry"]["coordinates"][0] for f in out])
    g = {
        "rings_verts": verts, "rings_off": off, "conf": np.array(conf, dtype=np.float64),
        "ndvi": ndvi64.astype(np.float32), "ndvi_transform": np.array(ndvi_tf), "ndvi_bounds": np.array(ndvi_bounds),
        "height": height, "height_transform": np.array(height_tf), "height_bounds": np.array(height_bounds),
        "pixel": np.array([abs(sc.transform[0]), abs(sc.transform[4])]),
        "ids_after_nms": np.array([int(f["properties"]["poly_id"]) for f in features], dtype=np.int64),
        "out_poly_id": np.array([int(f["properties"]["poly_id"]) for f in out], dtype=np.int64),
        "out_area": np.array([f["properties"]["Area"] for f in out], dtype=np.float64),
        "out_height": np.array([f["properties"]["TreeHeight"] for f in out], dtype=np.float32),
        "out_centroid": np.array([[f["properties"]["Centroid"]["x"], f["properties"]["Centroid"]["y"]] for f in out],
                                 dtype=np.float64).reshape(-1, 2),
        "out_is_contained": np.array([bool(f["properties"]["is_contained"]) for f in out]),
        "out_num_contained": np.array([int(f["properties"]["num_contained"]) for f in out], dtype=np.int64),
        "out_verts": overts, "out_off": ooff,
        "seed": np.array([seed]), "ndsm_px": np.array([ndsm_px]),
    }
    # per-stage reference numbers for the kernels: stats + containment of the post-NMS set
    Frings = [f["geometry"]["coordinates"][0] for f in features]
    px32, py32 = port.pad_polygons(Frings)
    cp = ns.cupy
    if g["ndsm_px"][0] == 1.0:
        hv, nv = ns.postprocessing.get_metadata_within_polygon(
            cp.array(px32), cp.array(py32), cp.array(g["ndvi"]), cp.array(height), A(*ndvi_tf), height.shape[0],
            height.shape[1], ns.BoundingBox(*ndvi_bounds))
        g["stat_max_h"] = np.asarray(hv[0]); g["stat_hxy"] = np.asarray(hv[1])
        g["stat_ndvi"] = np.stack([np.asarray(v) for v in nv], axis=1)
    else:
        mh, mc = ns.postprocessing.get_height_within_polygon(cp.array(px32), cp.array(py32), cp.array(height),
                                                             A(*height_tf), height.shape[0], height.shape[1],
                                                             ns.BoundingBox(*height_bounds))
        nv = ns.postprocessing.get_ndvi_within_polygon(cp.array(px32), cp.array(py32), cp.array(g["ndvi"]),
                                                       A(*ndvi_tf), g["ndvi"].shape[0], g["ndvi"].shape[1],
                                                       ns.BoundingBox(*ndvi_bounds))
        g["stat_max_h"] = np.asarray(mh); g["stat_hxy"] = np.asarray(mc)
        g["stat_ndvi"] = np.stack([np.asarray(v) for v in nv], axis=1)
    g["stat_centroid"] = np.asarray(ns.utilities.get_centroids(cp.array(px32), cp.array(py32)))
    np.savez_compressed(os.path.join(OUT, f"scene_{name}.npz"), **g)
    print(name, "rings", len(rings), "after nms", len(features), "out", len(out))


def main():
    assert refshim.available(), "needs /root/reference"
    ns = refshim.load()
    refshim.set_config(ns, **CFG)
    np.savez_compressed(os.path.join(OUT, "nms.npz"), **golden_nms(ns))
    golden_scene(ns, "combined", 1.0, 21)
    golden_scene(ns, "split", 0.2, 22)
    # NDVI: all uint8 pairs through the reference's numba kernel
    r = np.arange(256, dtype=np.uint8)
    R, N = np.meshgrid(r, r, indexing="ij")
    rgbi = np.zeros((4, 256, 256), np.uint8); rgbi[0] = R; rgbi[3] = N
    np.savez_compressed(os.path.join(OUT, "ndvi_u8.npz"), ndvi=ns.helpers.ndvi_array_from_rgbi(rgbi).astype(np.float32))


if __name__ == "__main__":
    main()
