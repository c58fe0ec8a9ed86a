"""Golden vectors for the parts of process_features the seeded tree fields of make_golden.py never reach
(VERDICT r01): crowns that contain 1, 2, 3 and more other crowns, crowns whose statistics set is EMPTY (-1),
and the threshold grid of /root/reference/supplementary/postprocessing_hyperparams.py:6-11.

A hand-made "nested" crown table (big crowns with k = 1..5 small crowns inside, ordinary crowns, crowns
outside the rasters) is post-processed by the REFERENCE'S OWN functions (filter_polygons_by_iou_and_area,
process_features incl. get_metadata_within_polygon / the split pair, process_containment_features; loaded
unmodified through oracle.refshim) under 9 parameter combinations, on a combined-path and a split-path
raster pair.  Inputs are stored once, outputs per combination -> tests/golden/grid_{combined,split}.npz.

    python tests/golden/make_golden_grid.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import port, refshim  # noqa: E402
from treedetection_b200 import geo, synth  # noqa: E402
from tests.golden.make_golden import CFG, features_from, ragged  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))

# (confidence, containment, iou, ndvi_mean, ndvi_var, use_overlap): rows of the hyper-parameter grid
# (postprocessing_hyperparams.py:6-11: conf {0.3,0.4,0.5}, containment {0.6,0.75,0.9}, iou {0.6,0.4,0.8},
# ndvi_mean {0.05,0.15,0.1,0.2}, ndvi_var {0.1,0.05,0.15}); use_overlap False keeps the crowns outside the rasters
COMBOS = [
    (0.3, 0.75, 0.6, 0.1, 0.1, True),
    (0.4, 0.6, 0.4, 0.05, 0.05, True),
    (0.5, 0.9, 0.8, 0.15, 0.15, True),
    (0.3, 0.6, 0.8, 0.2, 0.1, True),
    (0.3, 0.9, 0.4, 0.15, 0.05, True),
    (0.4, 0.9, 0.4, 0.1, 0.05, False),
    (0.3, 0.75, 0.6, 0.2, 0.15, False),
    (0.5, 0.6, 0.6, 0.05, 0.1, False),
    (0.3, 0.6, 0.8, 0.05, 0.15, False),
]


def nested_field(seed, size_m):
    """TreeField with big crowns holding k = 1..5 small ones, ordinary crowns, and crowns off the rasters."""
    rng = np.random.default_rng(seed)
    L, B = synth.ORIGIN_X, synth.ORIGIN_Y
    xs, ys, rs, hs, ss, es = [], [], [], [], [], []

    def add(x, y, r, h, s, e=1.0):
        xs.append(x); ys.append(y); rs.append(r); hs.append(h); ss.append(s); es.append(e)
    # big crowns on a coarse lattice inside the overlap band, k small crowns inside each
    k_of = [1, 2, 3, 4, 5, 1, 2, 3, 1, 3, 2, 5]
    for q, k in enumerate(k_of):
        cx = L + 45 + 27 * (q % 4) + rng.uniform(-2, 2)
        cy = B + 50 + 30 * (q // 4) + rng.uniform(-2, 2)
        R = rng.uniform(8.5, 11.0)
        add(cx, cy, R, rng.uniform(12, 30), rng.uniform(0.35, 0.95))
        for j in range(k):
            ang = 2 * np.pi * j / k + rng.uniform(-0.2, 0.2)
            d = rng.uniform(0.25, 0.5) * R
            add(cx + d * np.cos(ang), cy + d * np.sin(ang), rng.uniform(1.3, 2.2), rng.uniform(2.0, 25.0),
                rng.uniform(0.3, 0.99))
    # ordinary crowns (some overlapping each other: NMS groups), some low (height filter), some near the border
    for _ in range(140):
        add(L + rng.uniform(3, size_m - 3), B + rng.uniform(3, size_m - 3), rng.uniform(1.5, 5.0),
            rng.uniform(0.5, 30.0), rng.uniform(0.25, 0.99), rng.uniform(0.85, 1.2))
    for _ in range(25):   # near-duplicates of existing crowns (what overlapping tiles produce)
        j = int(rng.integers(0, len(xs)))
        add(xs[j] + rng.normal(0, 0.3), ys[j] + rng.normal(0, 0.3), rs[j] * rng.uniform(0.9, 1.1), hs[j],
            rng.uniform(0.3, 0.99), es[j])
    n_in = len(xs)
    # crowns completely outside the rasters: empty statistics sets
    for _ in range(6):
        side = rng.integers(0, 4)
        off = rng.uniform(15, 40)
        x = L - off if side == 0 else (L + size_m + off if side == 1 else L + rng.uniform(0, size_m))
        y = B - off if side == 2 else (B + size_m + off if side == 3 else B + rng.uniform(0, size_m))
        add(x, y, rng.uniform(2, 4), 10.0, rng.uniform(0.4, 0.9))
    f = synth.TreeField(*[np.array(v, dtype=np.float64) for v in (xs, ys, rs, hs, ss, es)], L, B, size_m, size_m)
    return f, n_in


def crown_rings(field, seed):
    """closed polygon per tree: ellipse with 14-26 jittered vertices (what a simplified mask outline looks like)"""
    rng = np.random.default_rng(seed + 1)
    rings, conf = [], []
    for k in range(len(field.x)):
        nv = int(rng.integers(14, 27))
        ang = np.sort(rng.uniform(0, 2 * np.pi, nv))
        rad = 1.0 + rng.normal(0, 0.04, nv)
        x = field.x[k] + field.r[k] * rad * np.cos(ang)
        y = field.y[k] + field.r[k] * field.ecc[k] * rad * np.sin(ang)
        ring = [(float(np.round(a, 1)), float(np.round(b, 1))) for a, b in zip(x, y)]
        ring.append(ring[0])
        rings.append(ring)
        conf.append(float(np.round(field.score[k], 3)))
    order = rng.permutation(len(rings))          # table order is unrelated to the spatial layout
    return [rings[i] for i in order], [conf[i] for i in order]


def run_reference(ns, rings, conf, cfg, height, height_tf, height_bounds, ndvi64, ndvi_tf, ndvi_bounds, px):
    """head of process_geojson (postprocessing.py:736-778) + process_features, all reference code"""
    refshim.set_config(ns, **cfg)
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
        ns.BoundingBox(*ndvi_bounds), px, px)
    overts, ooff = ragged([f["geometry"]["coordinates"][0] for f in out])
    # P8 on ALL post-NMS crowns, as process_features calls it (postprocessing.py:617-622): crowns that contain
    # 2, 3 or more others never reach the output, so their counts are recorded here
    pb = [polygon_dict[f["properties"]["poly_id"]].bounds for f in features]
    cont = ns.postprocessing.process_containment_features(features, [f["properties"]["poly_id"] for f in features], pb,
                                                          cfg["containment_threshold"])
    return {
        "p8_num_contained": np.array([f["properties"]["num_contained"] for f in cont], dtype=np.int64),
        "p8_is_contained": np.array([f["properties"]["is_contained"] for f in cont]),
        "p8_ratio": np.array([f["properties"]["containment_ratio"] for f in cont], dtype=np.float32),
        "ids_after_nms": np.array([int(f["properties"]["poly_id"]) for f in features], dtype=np.int64),
        "out_poly_id": np.array([int(f["properties"]["poly_id"]) for f in out], dtype=np.int64),
        "out_area": np.array([f["properties"]["Area"] for f in out], dtype=np.float64),
        "out_height": np.array([f["properties"]["TreeHeight"] for f in out], dtype=np.float32),
        "out_centroid": np.array([[f["properties"]["Centroid"]["x"], f["properties"]["Centroid"]["y"]] for f in out],
                                 dtype=np.float64).reshape(-1, 2),
        "out_is_contained": np.array([bool(f["properties"]["is_contained"]) for f in out]),
        "out_num_contained": np.array([int(f["properties"]["num_contained"]) for f in out], dtype=np.int64),
        "out_verts": overts, "out_off": ooff,
    }


def grid(ns, name, ndsm_px, seed):
    px, size_px = 0.2, 1000
    size_m = size_px * px
    field, _ = nested_field(seed, size_m)
    rings, conf = crown_rings(field, seed)
    top = synth.ORIGIN_Y + size_m
    tf = synth.image_transform(synth.ORIGIN_X, top, px)
    rgbi = synth.make_rgbi(field, px, seed)
    height = synth.make_ndsm(field, ndsm_px, seed)
    H, W = rgbi.shape[1:]
    oh, ow = int(H * CFG["ndvi_scaling_factor"]), int(W * CFG["ndvi_scaling_factor"])
    dec = np.stack([port.decimate_bilinear(rgbi[b], oh, ow) for b in range(4)])
    ndvi64 = ns.helpers.ndvi_array_from_rgbi(dec)
    ndvi_tf = geo.compose(tf, geo.scale(W / ow, H / oh))
    ndvi_bounds = geo.raster_bounds(tf, W, H)
    h, w = height.shape
    height_tf = synth.image_transform(synth.ORIGIN_X, top, ndsm_px)
    height_bounds = geo.raster_bounds(height_tf, w, h)
    verts, off = ragged(rings)
    g = {"rings_verts": verts, "rings_off": off, "conf": np.array(conf, dtype=np.float64),
         "ndvi": ndvi64.astype(np.float32), "ndvi_transform": np.array(ndvi_tf), "ndvi_bounds": np.array(ndvi_bounds),
         "height": height, "height_transform": np.array(height_tf), "height_bounds": np.array(height_bounds),
         "pixel": np.array([px, px]), "combos": np.array([[float(v) for v in c] for c in COMBOS])}
    for k, (conf_t, cont, iou, nm, nv, ov) in enumerate(COMBOS):
        cfg = dict(CFG, confidence_threshold=conf_t, containment_threshold=cont, iou_threshold=iou,
                   ndvi_mean_threshold=nm, ndvi_var_threshold=nv, use_overlap=ov)
        r = run_reference(ns, rings, conf, cfg, height, height_tf, height_bounds, ndvi64, ndvi_tf, ndvi_bounds, px)
        for key, v in r.items():
            g[f"c{k}_{key}"] = v
        print(name, k, "after nms", len(r["ids_after_nms"]), "out", len(r["out_poly_id"]),
              "num_contained hist (all post-NMS crowns)", np.bincount(r["p8_num_contained"], minlength=6).tolist(),
              "contained", int(r["p8_is_contained"].sum()), "empty (-1)", int((r["out_height"] == -1).sum()))
    np.savez_compressed(os.path.join(OUT, f"grid_{name}.npz"), **g)


def main():
    assert refshim.available(), "needs /root/reference"
    ns = refshim.load()
    grid(ns, "combined", 1.0, 77)
    grid(ns, "split", 0.5, 78)


if __name__ == "__main__":
    main()
