"""GPU parity of the composed path (C-ABI kernels through treedetection_b200.pipeline):
* predict side (P2+P3+P4) against the oracle restatement (torch-CPU paste, cv2 contours,
  GEOS-semantics simplify) -- ring vertices bit-exact, same order;
* post-processing side (P6..P9) against golden vectors produced by the REFERENCE'S OWN
  process_features / filter_polygons_by_iou_and_area (tests/golden/make_golden.py)."""
import os

import numpy as np
import pytest
import torch

from oracle import port
from treedetection_b200 import geo, ops, pipeline, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
ATOL = 1e-5


def _dev_det(sc, dev):
    d = sc.det
    tile_tf, boxes_int = pipeline.tile_tables(sc.tiles, dev)
    return dict(boxes_net=torch.from_numpy(d.boxes_net).to(dev), scores=torch.from_numpy(d.scores).to(dev),
                probs=torch.from_numpy(d.probs).to(dev), inst_tile=torch.from_numpy(d.inst_tile).to(dev),
                tile_dims=torch.from_numpy(d.tile_dims).to(dev), tile_tf=tile_tf,
                tile_boxes=pipeline.filter_boxes(boxes_int, 1, dev))


def _rings_of(verts, off):
    verts = np.asarray(verts); off = np.asarray(off)
    return [[(float(x), float(y)) for x, y in verts[off[i]:off[i + 1]]] for i in range(len(off) - 1)]


@pytest.mark.parametrize("seed,px", [(5, 0.2), (6, 0.25)])
def test_predict_stage_matches_oracle(dev, seed, px):
    sc = synth.make_scene(seed=seed, size_px=1200, px=px, ndsm_px=1.0, density_per_km2=5000.0, with_rasters=False)
    p = pipeline.PipelineParams()
    table = pipeline.predict_stage(**_dev_det(sc, dev), p=p)
    rings_ref, conf_ref = port.predict_stage(sc.det, sc.tiles, paste="torch")
    got = _rings_of(table.verts.cpu().numpy(), table.ring_off.cpu().numpy())
    assert len(got) == len(rings_ref) and len(got) > 100
    for a, b in zip(got, rings_ref):
        assert a == [tuple(q) for q in b]
    np.testing.assert_array_equal(table.conf.cpu().numpy(), np.array(conf_ref))


def test_trace_rings_noise_masks_match_cv2(dev):
    """P3 alone on hand-made packed rasters with holes, islands and noise."""
    rng = np.random.default_rng(0)
    masks = []
    for k in range(40):
        h, w = int(rng.integers(1, 90)), int(rng.integers(1, 130))
        masks.append((rng.uniform(size=(h, w)) < [0.3, 0.55, 0.8, 0.97][k % 4]).astype(np.uint8))
    m = np.zeros((60, 70), np.uint8); m[5:55, 5:65] = 1; m[10:50, 10:60] = 0; m[20:40, 20:50] = 1; m[25:35, 30:40] = 0
    masks.append(m)
    win, words = [], []
    for m in masks:
        h, w = m.shape
        wpr = (w + 31) // 32
        pad = np.zeros((h, wpr * 32), np.uint8); pad[:, :w] = m
        wd = (pad.reshape(h, wpr, 32).astype(np.uint64) << np.arange(32, dtype=np.uint64)).sum(axis=2).astype(np.uint32)
        words.append(wd.reshape(-1)); win.append([3, 4, w, h])
    nwords = np.array([len(w) for w in words], dtype=np.int64)
    off = np.zeros(len(words) + 1, dtype=np.int64); off[1:] = np.cumsum(nwords)
    bits = torch.from_numpy(np.concatenate(words).view(np.int32)).to(dev)
    tf = (0.2, 0.0, 412000.0, 0.0, -0.2, 5318100.0)
    rings = ops.trace_rings(bits, torch.tensor(win, dtype=torch.int32, device=dev), torch.from_numpy(off).to(dev),
                            torch.zeros(len(masks), dtype=torch.int32, device=dev),
                            torch.tensor([tf], dtype=torch.float64, device=dev))
    got = _rings_of(rings.verts.cpu().numpy(), rings.ring_off.cpu().numpy())
    inst = rings.ring_inst.cpu().numpy()
    want, want_inst = [], []
    for i, m in enumerate(masks):
        full = np.zeros((m.shape[0] + 4, m.shape[1] + 3), np.uint8); full[4:, 3:] = m
        for ring in port.mask_to_polygons(full.astype(bool), tf):
            want.append(ring); want_inst.append(i)
    assert len(got) == len(want)
    np.testing.assert_array_equal(inst, np.array(want_inst))
    for a, b in zip(got, want):
        assert a == [tuple(q) for q in b]
    # the single-pass form (one walk into per-instance slots): identical rings when the slots are big
    # enough, bit 2 of the flag when an instance outgrows its slot (noise masks have hundreds of borders)
    d_win = torch.tensor(win, dtype=torch.int32, device=dev)
    d_off = torch.from_numpy(off).to(dev)
    hw = np.array([[m.shape[1], m.shape[0]] for m in masks], dtype=np.int64)
    px_off = np.zeros(len(masks) + 1, dtype=np.int64); px_off[1:] = np.cumsum(hw[:, 0] * hw[:, 1])
    for slot_pts, slot_contours, expect_flag in [(20000, 4096, 0), (4 * 220 + 64, 16, 4)]:
        slot_off = torch.arange(len(masks) + 1, dtype=torch.int64, device=dev) * slot_pts
        caps = {"words": int(off[-1]), "px": int(px_off[-1]), "ptslots": slot_pts * len(masks),
                "rings": len(want) + 5, "verts": int(rings.verts.shape[0]) + 7}
        flag = torch.zeros(1, dtype=torch.int64, device=dev)
        totals = torch.zeros(2, dtype=torch.int64, device=dev)
        old = ops.TRACE_SLOT_CONTOURS
        ops.TRACE_SLOT_CONTOURS = slot_contours
        try:
            r2 = ops.trace_rings_slots(bits, d_win, d_off, torch.from_numpy(px_off).to(dev), slot_off,
                                       torch.zeros(len(masks), dtype=torch.int32, device=dev),
                                       torch.tensor([tf], dtype=torch.float64, device=dev), caps, flag, totals)
        finally:
            ops.TRACE_SLOT_CONTOURS = old
        assert int(flag) == expect_flag
        if expect_flag == 0:
            nr, nv = [int(v) for v in totals.tolist()]
            assert nr == len(want) and nv == rings.verts.shape[0]
            assert torch.equal(r2.verts[:nv], rings.verts) and torch.equal(r2.ring_off[:nr + 1], rings.ring_off)
            assert torch.equal(r2.ring_inst[:nr], rings.ring_inst)


@pytest.mark.parametrize("name", ["combined", "split"])
def test_postprocess_stage_matches_reference_golden(dev, name):
    g = np.load(os.path.join(G, f"scene_{name}.npz"))
    p = pipeline.PipelineParams()
    table = pipeline.CrownTable(torch.from_numpy(g["rings_verts"]).to(dev), torch.from_numpy(g["rings_off"]).to(dev),
                                torch.from_numpy(g["conf"]).to(dev))
    rasters = {"ndvi": torch.from_numpy(g["ndvi"]).to(dev), "ndvi_transform": tuple(g["ndvi_transform"]),
               "ndvi_bounds": geo.BoundingBox(*g["ndvi_bounds"]), "height": torch.from_numpy(g["height"]).to(dev),
               "height_transform": tuple(g["height_transform"]), "height_bounds": geo.BoundingBox(*g["height_bounds"]),
               "pixel_x": float(g["pixel"][0]), "pixel_y": float(g["pixel"][1])}
    f = pipeline.postprocess_stage(table, rasters, p, keep_debug=True)
    ex = f.extras
    assert ex["combined"] == (name == "combined")
    np.testing.assert_array_equal(ex["pid_after_nms"].cpu().numpy(), g["ids_after_nms"])
    np.testing.assert_array_equal(ex["max_h"].cpu().numpy(), g["stat_max_h"])
    np.testing.assert_array_equal(ex["hxy"].cpu().numpy(), g["stat_hxy"])
    st = ex["ndvi_stats"].cpu().numpy()
    np.testing.assert_array_equal(st[:, :2], g["stat_ndvi"][:, :2])
    np.testing.assert_allclose(st[:, 2:], g["stat_ndvi"][:, 2:], atol=ATOL, rtol=0)
    np.testing.assert_array_equal(ex["centroid"].cpu().numpy(), g["stat_centroid"].astype(np.float32))
    # the crown set, its order and every attribute
    np.testing.assert_array_equal(f.poly_id.cpu().numpy(), g["out_poly_id"])
    np.testing.assert_array_equal(f.area.cpu().numpy(), g["out_area"])
    np.testing.assert_array_equal(f.tree_height.cpu().numpy(), g["out_height"])
    np.testing.assert_array_equal(f.centroid.cpu().numpy().astype(np.float64), g["out_centroid"])
    np.testing.assert_array_equal(f.is_contained.cpu().numpy().astype(bool), g["out_is_contained"])
    np.testing.assert_array_equal(f.num_contained.cpu().numpy(), g["out_num_contained"])
    np.testing.assert_array_equal(f.ring_off.cpu().numpy(), g["out_off"])
    np.testing.assert_array_equal(f.verts.cpu().numpy(), g["out_verts"])


@pytest.mark.parametrize("name", ["combined", "split"])
def test_postprocess_stage_matches_reference_on_threshold_grid(dev, name):
    """nested crowns (num_contained up to 7), empty statistics sets (-1) and 9 rows of the reference's
    hyper-parameter grid (supplementary/postprocessing_hyperparams.py:6-11), against the outputs of the
    reference's own process_features / process_containment_features (tests/golden/make_golden_grid.py)"""
    from tests.test_oracle_golden import grid_cfg
    g = np.load(os.path.join(G, f"grid_{name}.npz"))
    table = pipeline.CrownTable(torch.from_numpy(g["rings_verts"]).to(dev), torch.from_numpy(g["rings_off"]).to(dev),
                                torch.from_numpy(g["conf"]).to(dev))
    rasters = {"ndvi": torch.from_numpy(g["ndvi"]).to(dev), "ndvi_transform": tuple(g["ndvi_transform"]),
               "ndvi_bounds": geo.BoundingBox(*g["ndvi_bounds"]), "height": torch.from_numpy(g["height"]).to(dev),
               "height_transform": tuple(g["height_transform"]), "height_bounds": geo.BoundingBox(*g["height_bounds"]),
               "pixel_x": float(g["pixel"][0]), "pixel_y": float(g["pixel"][1])}
    for k, combo in enumerate(g["combos"]):
        p = pipeline.PipelineParams.from_config(grid_cfg(combo))
        f = pipeline.postprocess_stage(table, rasters, p, keep_debug=True)
        ex = f.extras
        assert ex["combined"] == (name == "combined")
        np.testing.assert_array_equal(ex["pid_after_nms"].cpu().numpy(), g[f"c{k}_ids_after_nms"])
        np.testing.assert_array_equal(ex["num_contained"].cpu().numpy(), g[f"c{k}_p8_num_contained"])
        np.testing.assert_array_equal(ex["is_contained"].cpu().numpy().astype(bool), g[f"c{k}_p8_is_contained"])
        np.testing.assert_array_equal(ex["containment_ratio"].cpu().numpy(), g[f"c{k}_p8_ratio"])
        np.testing.assert_array_equal(f.poly_id.cpu().numpy(), g[f"c{k}_out_poly_id"])
        np.testing.assert_array_equal(f.area.cpu().numpy(), g[f"c{k}_out_area"])
        np.testing.assert_array_equal(f.tree_height.cpu().numpy(), g[f"c{k}_out_height"])
        np.testing.assert_array_equal(f.centroid.cpu().numpy().astype(np.float64), g[f"c{k}_out_centroid"])
        np.testing.assert_array_equal(f.is_contained.cpu().numpy().astype(bool), g[f"c{k}_out_is_contained"])
        np.testing.assert_array_equal(f.num_contained.cpu().numpy(), g[f"c{k}_out_num_contained"])
        np.testing.assert_array_equal(f.ring_off.cpu().numpy(), g[f"c{k}_out_off"])
        np.testing.assert_array_equal(f.verts.cpu().numpy(), g[f"c{k}_out_verts"])


def test_nms_golden(dev):
    g = np.load(os.path.join(G, "nms.npz"))
    for case in "abc":
        iou, athr = g[f"nms_{case}_params"]
        got = ops.bbox_nms_ordered(torch.from_numpy(g[f"nms_{case}_bounds"]).to(dev),
                                   torch.from_numpy(g[f"nms_{case}_conf"]).to(dev),
                                   torch.from_numpy(g[f"nms_{case}_area"]).to(dev), float(iou), float(athr))
        np.testing.assert_array_equal(got.cpu().numpy().astype(bool), g[f"nms_{case}_removed"])


def test_ndvi_all_uint8_pairs_match_reference(dev):
    g = np.load(os.path.join(G, "ndvi_u8.npz"))
    r = np.arange(256, dtype=np.uint8)
    R, N = np.meshgrid(r, r, indexing="ij")
    rgbi = np.zeros((4, 256, 256), np.uint8); rgbi[0] = R; rgbi[3] = N
    got = ops.ndvi_decimate(torch.from_numpy(rgbi).to(dev), 256, 256).cpu().numpy()
    np.testing.assert_array_equal(got, g["ndvi"])


def test_full_chain_matches_oracle(dev):
    """rasters + detections -> final crowns, GPU vs the oracle end to end (split path)."""
    sc = synth.make_scene(seed=9, size_px=1500, px=0.2, ndsm_px=0.2, density_per_km2=5000.0)
    p = pipeline.PipelineParams()
    table = pipeline.predict_stage(**_dev_det(sc, dev), p=p)
    rasters = pipeline.raster_stage(torch.from_numpy(sc.rgbi).to(dev), sc.transform, torch.from_numpy(sc.ndsm).to(dev),
                                    sc.ndsm_transform, p)
    f = pipeline.postprocess_stage(table, rasters, p, keep_debug=True)
    # oracle
    rings, conf = port.predict_stage(sc.det, sc.tiles)
    H, W = sc.rgbi.shape[1:]
    oh, ow = int(H * 0.2), int(W * 0.2)
    dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in (0, 0, 0, 3)])
    ndvi = port.ndvi_from_rgbi(dec).astype(np.float32)
    np.testing.assert_array_equal(rasters["ndvi"].cpu().numpy(), ndvi)
    cfg = {k: getattr(p, k) for k in p.__dataclass_fields__}
    out, dbg = port.post_process(rings, conf, ndvi, rasters["ndvi_transform"], tuple(rasters["ndvi_bounds"]), sc.ndsm,
                                 rasters["height_transform"], tuple(rasters["height_bounds"]), 0.2, 0.2, cfg)
    assert not dbg["combined"] and len(out) > 30
    np.testing.assert_array_equal(f.poly_id.cpu().numpy(), np.array([int(o["poly_id"]) for o in out]))
    np.testing.assert_array_equal(f.area.cpu().numpy(), np.array([o["Area"] for o in out]))
    np.testing.assert_array_equal(f.tree_height.cpu().numpy(), np.array([o["TreeHeight"] for o in out], np.float32))
    np.testing.assert_array_equal(f.verts.cpu().numpy(), np.array([q for o in out for q in o["coords"]]).reshape(-1, 2))


def test_predict_stage_mask_iou_mode(dev):
    """opt-in ``iou_mode: mask``: instances dropped by the mask-IoU cleaner leave no rings, the others go through
    P3 / P4 unchanged -- equal to the oracle's predict stage on the surviving instances"""
    sc = synth.make_scene(seed=15, size_px=1000, px=0.2, ndsm_px=1.0, density_per_km2=6000.0, with_rasters=False)
    p = pipeline.PipelineParams(iou_mode="mask", mask_iou_threshold=0.5, confidence_threshold_stitching=0.3)
    dd = _dev_det(sc, dev)
    tile_org = torch.tensor([m["window"][:2] for m in sc.tiles.values()], dtype=torch.int32, device=dev)
    table = pipeline.predict_stage(**dd, p=p, tile_org=tile_org)
    plain = pipeline.predict_stage(**dd, p=pipeline.PipelineParams())
    boxes_px, win, nwords = ops.paste_plan(dd["boxes_net"], dd["inst_tile"], dd["tile_dims"])
    off = ops.exclusive_offsets(nwords)
    bits = ops.paste_threshold_pack(boxes_px, win, off, dd["probs"])
    keep, match, _ = ops.mask_iou_clean(bits, off, win, tile_org, dd["inst_tile"], dd["scores"], 0.5, 0.3)
    keep, match = keep.cpu().numpy().astype(bool), match.cpu().numpy()
    d = sc.det
    sub = synth.Detections(d.boxes_net[keep], d.scores[match[keep]], d.probs[keep], d.inst_tile[keep], d.tile_dims,
                           d.tile_ids, d.tiles)
    rings_ref, conf_ref = port.predict_stage(sub, sc.tiles)
    got = _rings_of(table.verts.cpu().numpy(), table.ring_off.cpu().numpy())
    assert len(got) == len(rings_ref) and 50 < len(got) < len(plain)
    for a, b in zip(got, rings_ref):
        assert a == [tuple(q) for q in b]
    np.testing.assert_array_equal(table.conf.cpu().numpy(), np.array(conf_ref))
