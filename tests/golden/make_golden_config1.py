"""Config 1 of BASELINE.json: the reference's own example (example/config.yml) over the bundled
nDSM tile data/nDSM/324125317.tif (1000 x 1000 float32, 1 m) and a SYNTHETIC 5000 x 5000 RGBI
companion (the bundled RGB is absent: .MISSING_LARGE_BLOBS), ROI-head outputs replayed from
seeded fixtures.  The post-processing is run by the REFERENCE'S OWN process_features /
filter_polygons_by_iou_and_area through oracle.refshim; only the outputs are stored
(tests/golden/config1.npz) -- the inputs are regenerated from the seed and from
tests/golden/ndsm_324125317.npz (a lossless copy of the bundled raster's pixels).

    python tests/golden/make_golden_config1.py       (build container only)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import port, refshim  # noqa: E402
from tests.golden.make_golden import CFG, features_from, ragged  # noqa: E402
from treedetection_b200 import geo, geotiff, synth  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))
SEED = 31


def config1_scene():
    """RGBI 5000^2 @0.2 m + detections over the bundled tile's extent (top-left 412000, 5318000)."""
    return synth.config1_scene(os.path.join(OUT, "ndsm_324125317.npz"), SEED)


def main():
    assert refshim.available()
    ref_tif = "/root/reference/data/nDSM/324125317.tif"
    arr, info = geotiff.read(ref_tif)
    assert np.array_equal(arr[0], np.load(os.path.join(OUT, "ndsm_324125317.npz"))["ndsm"])
    assert info.transform == (1.0, 0.0, 412000.0, 0.0, -1.0, 5318000.0)
    ns = refshim.load()
    cfg = dict(CFG)
    cfg.update(containment_threshold=0.75, iou_threshold=0.6)      # example/config.yml
    refshim.set_config(ns, **cfg)
    sc = config1_scene()
    rings, conf = port.predict_stage(sc.det, sc.tiles)
    H, W = sc.rgbi.shape[1:]
    oh, ow = int(H * 0.2), int(W * 0.2)
    dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in range(4)])
    ndvi64 = ns.helpers.ndvi_array_from_rgbi(dec)
    ndvi_tf = geo.compose(sc.transform, geo.scale(W / ow, H / oh))
    ndvi_bounds = geo.raster_bounds(sc.transform, W, H)
    h, w = sc.ndsm.shape
    height_tf = sc.ndsm_transform
    height_bounds = geo.raster_bounds(height_tf, w, h)
    feats = [(r, c) for r, c in zip(rings, conf) if c >= cfg["confidence_threshold"]]
    features = features_from([f[0] for f in feats], [f[1] for f in feats], list(range(len(feats))))
    id_to_area = {}
    for f in features:
        id_to_area[f["properties"]["poly_id"]] = ns.postprocessing.calculate_area(
            ns.postprocessing.shape(f["geometry"]).simplify(2))
    polygon_dict = {f["properties"]["poly_id"]: ns.postprocessing.shape(f["geometry"]) for f in features}
    features = [f for f in features if cfg["area_threshold"] <= id_to_area[f["properties"]["poly_id"]] <= 1000]
    fids = {f["properties"]["poly_id"] for f in features}
    polygon_dict = {k: v for k, v in polygon_dict.items() if k in fids}
    confs = {f["properties"]["poly_id"]: f["properties"]["Confidence_score"] for f in features}
    retained = ns.postprocessing.filter_polygons_by_iou_and_area(polygon_dict, id_to_area, confs, cfg["iou_threshold"],
                                                                 cfg["area_threshold"])
    features = [f for f in features if f["properties"]["poly_id"] in retained]
    A = ns.Affine
    out = ns.postprocessing.process_features(features, id_to_area, sc.ndsm, A(*height_tf), ns.BoundingBox(*height_bounds),
                                             ndvi64, A(*ndvi_tf), ns.BoundingBox(*ndvi_bounds), 0.2, 0.2)
    sv, so = ragged(rings)
    ov, oo = ragged([f["geometry"]["coordinates"][0] for f in out])
    np.savez_compressed(
        os.path.join(OUT, "config1.npz"), seed=np.array([SEED]),
        stitched_off=so, stitched_verts_sum=np.array([sv.sum(axis=0)]), stitched_conf=np.array(conf),
        stitched_first=sv[:64], n_instances=np.array([len(sc.det.scores)]),
        ids_after_nms=np.array([int(f["properties"]["poly_id"]) for f in features], dtype=np.int64),
        out_poly_id=np.array([int(f["properties"]["poly_id"]) for f in out], dtype=np.int64),
        out_area=np.array([f["properties"]["Area"] for f in out]),
        out_height=np.array([f["properties"]["TreeHeight"] for f in out], dtype=np.float32),
        out_centroid=np.array([[f["properties"]["Centroid"]["x"], f["properties"]["Centroid"]["y"]] for f in out]),
        out_is_contained=np.array([bool(f["properties"]["is_contained"]) for f in out]),
        out_num_contained=np.array([int(f["properties"]["num_contained"]) for f in out], dtype=np.int64),
        out_verts=ov, out_off=oo)
    print("instances", len(sc.det.scores), "stitched", len(rings), "after nms", len(features), "out", len(out))


if __name__ == "__main__":
    main()
