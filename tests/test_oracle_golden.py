"""The oracle restatement (oracle/port.py) against the golden vectors produced by the
reference's own functions (tests/golden/make_golden.py).  Runs on CPU everywhere."""
import os

import numpy as np
import pytest

from oracle import port

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

CFG = dict(tile_width=50, tile_height=50, buffer=20, use_overlap=True, overlapping_tiles_width=3,
           overlapping_tiles_height=3, confidence_threshold=0.3, containment_threshold=0.75, height_threshold=3,
           ndvi_mean_threshold=0.1, ndvi_var_threshold=0.1, iou_threshold=0.6, area_threshold=1,
           ndvi_scaling_factor=0.2, height_scaling_factor=1.0)


def rings_of(verts, off):
    return [[(float(x), float(y)) for x, y in verts[off[i]:off[i + 1]]] for i in range(len(off) - 1)]


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_nms_port_matches_reference(case):
    g = np.load(os.path.join(G, "nms.npz"))
    iou, athr = g[f"nms_{case}_params"]
    for fn in (port.nms_bbox, port.nms_bbox_sparse):
        got = fn(g[f"nms_{case}_bounds"], g[f"nms_{case}_conf"], g[f"nms_{case}_area"], float(iou), float(athr))
        np.testing.assert_array_equal(got, g[f"nms_{case}_removed"])
    assert g[f"nms_{case}_removed"].sum() > 0


def test_ndvi_port_matches_reference_on_all_uint8_pairs():
    g = np.load(os.path.join(G, "ndvi_u8.npz"))
    r = np.arange(256, dtype=np.uint8)
    R, N = np.meshgrid(r, r, indexing="ij")
    rgbi = np.zeros((4, 256, 256), np.uint8); rgbi[0] = R; rgbi[3] = N
    np.testing.assert_array_equal(port.ndvi_from_rgbi(rgbi).astype(np.float32), g["ndvi"])


@pytest.mark.parametrize("name", ["combined", "split"])
def test_post_process_port_matches_reference(name):
    g = np.load(os.path.join(G, f"scene_{name}.npz"))
    rings = rings_of(g["rings_verts"], g["rings_off"])
    out, dbg = port.post_process(rings, g["conf"].tolist(), g["ndvi"], tuple(g["ndvi_transform"]),
                                 tuple(g["ndvi_bounds"]), g["height"], tuple(g["height_transform"]),
                                 tuple(g["height_bounds"]), float(g["pixel"][0]), float(g["pixel"][1]), CFG)
    assert dbg["combined"] == (name == "combined")
    np.testing.assert_array_equal(np.array(dbg["ids_after_nms"]), g["ids_after_nms"])
    st = dbg["stats"]
    np.testing.assert_array_equal(st["max_h"], g["stat_max_h"])
    np.testing.assert_array_equal(np.stack([st["hx"], st["hy"]], 1), g["stat_hxy"])
    np.testing.assert_array_equal(np.stack([st["ndvi_min"], st["ndvi_max"], st["ndvi_mean"], st["ndvi_var"]], 1),
                                  g["stat_ndvi"])
    np.testing.assert_array_equal(dbg["centroid"].astype(np.float32), g["stat_centroid"].astype(np.float32))
    np.testing.assert_array_equal(np.array([int(f["poly_id"]) for f in out]), g["out_poly_id"])
    np.testing.assert_array_equal(np.array([f["Area"] for f in out]), g["out_area"])
    np.testing.assert_array_equal(np.array([f["TreeHeight"] for f in out], dtype=np.float32), g["out_height"])
    np.testing.assert_array_equal(np.array([f["Centroid"] for f in out]).reshape(-1, 2), g["out_centroid"])
    np.testing.assert_array_equal(np.array([f["is_contained"] for f in out]), g["out_is_contained"])
    np.testing.assert_array_equal(np.array([f["num_contained"] for f in out]), g["out_num_contained"])
    got_verts = np.array([p for f in out for p in f["coords"]]).reshape(-1, 2)
    np.testing.assert_array_equal(got_verts, g["out_verts"])
    assert len(out) > 50


def test_config1_port_matches_reference():
    """BASELINE config 1 (bundled nDSM tile + synthetic RGBI): the oracle port against the
    outputs of the reference's own process_features (tests/golden/make_golden_config1.py)."""
    from tests.golden.make_golden_config1 import config1_scene
    from treedetection_b200 import geo
    g = np.load(os.path.join(G, "config1.npz"))
    sc = config1_scene()
    assert len(sc.det.scores) == int(g["n_instances"][0])
    rings, conf = port.predict_stage(sc.det, sc.tiles)
    np.testing.assert_array_equal(np.array(conf), g["stitched_conf"])
    H, W = sc.rgbi.shape[1:]
    oh, ow = int(H * 0.2), int(W * 0.2)
    dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in (0, 0, 0, 3)])
    ndvi = port.ndvi_from_rgbi(dec).astype(np.float32)
    ndvi_tf = geo.compose(sc.transform, geo.scale(W / ow, H / oh))
    h, w = sc.ndsm.shape
    cfg = dict(CFG)
    out, dbg = port.post_process(rings, conf, ndvi, ndvi_tf, tuple(geo.raster_bounds(sc.transform, W, H)), sc.ndsm,
                                 sc.ndsm_transform, tuple(geo.raster_bounds(sc.ndsm_transform, w, h)), 0.2, 0.2, cfg)
    assert dbg["combined"]
    np.testing.assert_array_equal(np.array(dbg["ids_after_nms"]), g["ids_after_nms"])
    np.testing.assert_array_equal(np.array([int(f["poly_id"]) for f in out]), g["out_poly_id"])
    np.testing.assert_array_equal(np.array([f["Area"] for f in out]), g["out_area"])
    np.testing.assert_array_equal(np.array([f["TreeHeight"] for f in out], dtype=np.float32), g["out_height"])
    np.testing.assert_array_equal(np.array([p for f in out for p in f["coords"]]).reshape(-1, 2), g["out_verts"])


GRID_KEYS = ("confidence_threshold", "containment_threshold", "iou_threshold", "ndvi_mean_threshold", "ndvi_var_threshold",
             "use_overlap")


def grid_cfg(combo):
    cfg = dict(CFG)
    cfg.update({k: (bool(v) if k == "use_overlap" else float(v)) for k, v in zip(GRID_KEYS, combo)})
    return cfg


@pytest.mark.parametrize("name", ["combined", "split"])
def test_post_process_port_matches_reference_on_threshold_grid(name):
    """nested crowns (num_contained up to 7), empty statistics sets (-1), 9 rows of the reference's
    hyper-parameter grid (supplementary/postprocessing_hyperparams.py:6-11): tests/golden/make_golden_grid.py"""
    g = np.load(os.path.join(G, f"grid_{name}.npz"))
    rings = rings_of(g["rings_verts"], g["rings_off"])
    seen_nc, seen_empty = set(), 0
    for k, combo in enumerate(g["combos"]):
        cfg = grid_cfg(combo)
        out, dbg = port.post_process(rings, g["conf"].tolist(), g["ndvi"], tuple(g["ndvi_transform"]),
                                     tuple(g["ndvi_bounds"]), g["height"], tuple(g["height_transform"]),
                                     tuple(g["height_bounds"]), float(g["pixel"][0]), float(g["pixel"][1]), cfg)
        assert dbg["combined"] == (name == "combined")
        np.testing.assert_array_equal(np.array(dbg["ids_after_nms"]), g[f"c{k}_ids_after_nms"])
        np.testing.assert_array_equal(np.asarray(dbg["num_contained"]), g[f"c{k}_p8_num_contained"])
        np.testing.assert_array_equal(np.asarray(dbg["is_contained"]).astype(bool), g[f"c{k}_p8_is_contained"])
        np.testing.assert_array_equal(np.array([int(f["poly_id"]) for f in out]), g[f"c{k}_out_poly_id"])
        np.testing.assert_array_equal(np.array([f["Area"] for f in out]), g[f"c{k}_out_area"])
        np.testing.assert_array_equal(np.array([f["TreeHeight"] for f in out], dtype=np.float32), g[f"c{k}_out_height"])
        np.testing.assert_array_equal(np.array([f["Centroid"] for f in out]).reshape(-1, 2), g[f"c{k}_out_centroid"])
        np.testing.assert_array_equal(np.array([f["is_contained"] for f in out]), g[f"c{k}_out_is_contained"])
        np.testing.assert_array_equal(np.array([f["num_contained"] for f in out]), g[f"c{k}_out_num_contained"])
        np.testing.assert_array_equal(np.array([p for f in out for p in f["coords"]]).reshape(-1, 2), g[f"c{k}_out_verts"])
        seen_nc |= set(int(v) for v in g[f"c{k}_p8_num_contained"])
        seen_empty += int((g[f"c{k}_out_height"] == -1).sum())
    assert {0, 1, 2, 3, 4, 5} <= seen_nc and seen_empty > 0
