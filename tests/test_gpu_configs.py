"""BASELINE.json configs beyond the bench line, on the GPU against the oracle, plus size-independent
properties at the full bench size:

* config 4 (dense-forest stress, ~2,000 overlapping crowns per tile, 25 % tile overlap): the whole chain on a
  small dense scene against the oracle; ordered NMS + containment on 20 k dense boxes against the O(N^2) oracle;
* config 5 (row-sharded mosaic): the down-seam strip assembled on the device from the lower neighbour's halo
  rows equals the centre crop of the two-image mosaic (oracle: merging.py restated), and the strip runs through
  the chain like any image (seam rules of P9) with the oracle's result;
* full size (10 000 x 10 000 px): sampled tiles of P1 against PIL, sync-free chain == exact-size chain,
  filter invariants of the final crown table."""
import numpy as np
import pytest
import torch

from oracle import port
from treedetection_b200 import api, geo, ops, pipeline, sharding, synth, tiling

pytestmark = pytest.mark.gpu


def _det(sc, dev):
    d = sc.det
    return {k: torch.from_numpy(getattr(d, k)).to(dev) for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}


def _oracle_chain(sc, p, rgbi, ndsm, transform, ndsm_transform):
    rings, conf = port.predict_stage(sc.det, sc.tiles)
    H, W = rgbi.shape[1:]
    oh, ow = int(H * p.ndvi_scaling_factor), int(W * p.ndvi_scaling_factor)
    dec = np.stack([port.decimate_bilinear(rgbi[b], oh, ow) for b in (0, 0, 0, 3)])
    ndvi = port.ndvi_from_rgbi(dec).astype(np.float32)
    ndvi_tf = geo.compose(transform, geo.scale(W / ow, H / oh))
    h, w = ndsm.shape
    cfg = {k: getattr(p, k) for k in p.__dataclass_fields__}
    out, dbg = port.post_process(rings, conf, ndvi, ndvi_tf, tuple(geo.raster_bounds(transform, W, H)), ndsm,
                                 ndsm_transform, tuple(geo.raster_bounds(ndsm_transform, w, h)), abs(transform[0]),
                                 abs(transform[4]), cfg)
    return rings, out


def _assert_same(f, out):
    np.testing.assert_array_equal(f.poly_id.cpu().numpy(), np.array([int(o["poly_id"]) for o in out], dtype=np.int64))
    np.testing.assert_array_equal(f.area.cpu().numpy(), np.array([o["Area"] for o in out]))
    np.testing.assert_array_equal(f.tree_height.cpu().numpy(), np.array([o["TreeHeight"] for o in out], np.float32))
    np.testing.assert_array_equal(f.verts.cpu().numpy(), np.array([q for o in out for q in o["coords"]]).reshape(-1, 2))


# ---------------------------------------------------------------------------------------------- config 4
def test_config4_dense_forest_chain(dev):
    """tile 50 m, buffer 12.5 m (25 % overlap), small crowns at 60 000 / km^2, no per-tile cap"""
    sc = synth.make_scene(seed=41, size_px=750, px=0.2, ndsm_px=0.2, density_per_km2=60000.0, cap=10**6,
                          buffer=12.5, r_range=(0.6, 1.6))
    assert max(np.bincount(sc.det.inst_tile)) > 300
    p = pipeline.PipelineParams(buffer=12.5)
    tile_tf, boxes_int = pipeline.tile_tables(sc.tiles, dev)
    table = pipeline.predict_stage(**_det(sc, dev), tile_tf=tile_tf, tile_boxes=pipeline.filter_boxes(boxes_int, 1, dev),
                                   p=p)
    rasters = pipeline.raster_stage(torch.from_numpy(sc.rgbi).to(dev), sc.transform, torch.from_numpy(sc.ndsm).to(dev),
                                    sc.ndsm_transform, p)
    f = pipeline.postprocess_stage(table, rasters, p)
    rings, out = _oracle_chain(sc, p, sc.rgbi, sc.ndsm, sc.transform, sc.ndsm_transform)
    assert len(table) == len(rings) and len(rings) > 1500 and len(out) > 100
    _assert_same(f, out)


def test_config4_dense_nms_and_containment(dev):
    rng = np.random.default_rng(4)
    n = 8000                                    # dense oracle: N^2 float32 temporaries
    cx, cy = 412000 + rng.uniform(0, 100.0, n), 5318000 + rng.uniform(0, 100.0, n)      # ~2,000 per 50 m tile
    rx = rng.uniform(0.5, 2.0, n); ry = rx * rng.uniform(0.8, 1.25, n)
    bounds = np.stack([cx - rx, cy - ry, cx + rx, cy + ry], 1)
    conf = np.round(rng.uniform(0.3, 1.0, n), 3)
    area = np.pi * rx * ry
    removed = ops.bbox_nms_ordered(torch.from_numpy(bounds).to(dev), torch.from_numpy(conf).to(dev),
                                   torch.from_numpy(area).to(dev), 0.6, 1.0).cpu().numpy().astype(bool)
    ref = port.nms_bbox(bounds, conf, area, 0.6, 1.0)
    assert ref.sum() > 300
    np.testing.assert_array_equal(removed, ref)
    np.testing.assert_array_equal(port.nms_bbox_sparse(bounds, conf, area, 0.6, 1.0), ref)
    b32 = bounds.astype(np.float32)
    ratio, isc, num = ops.containment(torch.from_numpy(b32).to(dev), 0.75)
    r_ref, isc_ref, num_ref = port.containment(b32, 0.75)
    np.testing.assert_array_equal(isc.cpu().numpy().astype(bool), isc_ref)
    np.testing.assert_array_equal(num.cpu().numpy(), num_ref)
    np.testing.assert_array_equal(ratio.cpu().numpy(), r_ref)


# ---------------------------------------------------------------------------------------------- config 5
def test_config5_down_seam_strip(dev):
    """two vertically adjacent images: strip = centre crop of their mosaic; the strip is an image of its own"""
    px, size = 0.2, 1500
    p = pipeline.PipelineParams()
    top = synth.make_scene(seed=51, size_px=size, px=px, ndsm_px=px, density_per_km2=4000.0)
    low = synth.make_scene(seed=52, size_px=size, px=px, ndsm_px=px, density_per_km2=4000.0,
                           bottom=synth.ORIGIN_Y - size * px)
    rows = sharding.halo_rows(p.tile_height, p.buffer, p.overlapping_tiles_height)
    assert rows == 135
    # device assembly from the halo (what rank r does with the rows rank r + 1 sent)
    own = torch.from_numpy(top.rgbi).to(dev)
    halo = torch.from_numpy(np.ascontiguousarray(low.rgbi[:, :rows])).to(dev)
    strip = sharding.assemble_down_strip(own, halo)
    ref, (left, top_row) = port.seam_crop(top.rgbi, low.rgbi, 1, size, 2 * rows)
    ref = np.ascontiguousarray(ref)
    assert (left, top_row) == (0, size - rows)
    np.testing.assert_array_equal(strip.cpu().numpy(), ref)
    s_ndsm = sharding.assemble_down_strip(torch.from_numpy(top.ndsm).to(dev)[None],
                                          torch.from_numpy(np.ascontiguousarray(low.ndsm[:rows])).to(dev)[None])[0]
    ndsm_ref = np.ascontiguousarray(port.seam_crop(top.ndsm[None], low.ndsm[None], 1, size, 2 * rows)[0][0])
    np.testing.assert_array_equal(s_ndsm.cpu().numpy(), ndsm_ref)
    # the strip through the chain: trees of both images, seam rules of P9 (is_seam_image)
    both = synth.TreeField(*[np.concatenate([getattr(top.field, k), getattr(low.field, k)]) for k in
                             ("x", "y", "r", "h", "score", "ecc")], top.field.left, low.field.bottom, top.field.width_m,
                           2 * top.field.height_m)
    s_top = top.field.bottom + rows * px
    s_tf = synth.image_transform(top.field.left, s_top, px)
    s_tiles = tiling.tile_grid("FDOP20_seam_rgbi", s_tf, size, 2 * rows, synth.EPSG, p.tile_width, p.tile_height, p.buffer)
    s_det = synth.make_detections(both, s_tiles, px, 51)
    sc = synth.Scene("FDOP20_seam_rgbi", s_tf, px, ref, ndsm_ref, s_tf, both, s_tiles, s_det)
    tile_tf, boxes_int = pipeline.tile_tables(s_tiles, dev)
    table = pipeline.predict_stage(**_det(sc, dev), tile_tf=tile_tf, tile_boxes=pipeline.filter_boxes(boxes_int, 1, dev),
                                   p=p)
    rasters = pipeline.raster_stage(strip, s_tf, s_ndsm, s_tf, p)
    sp = pipeline.select_params(p, tuple(rasters["ndvi"].shape), rasters["ndvi_bounds"], rasters["pixel_x"],
                                rasters["pixel_y"])
    assert sp[1] == 1.0, "a 270-row strip must be recognised as a seam image (postprocessing.py:580-589)"
    f = pipeline.postprocess_stage(table, rasters, p)
    rings, out = _oracle_chain(sc, p, ref, ndsm_ref, s_tf, s_tf)
    assert len(table) == len(rings) and len(out) > 10
    _assert_same(f, out)


# ---------------------------------------------------------------------------------------------- full size
@pytest.fixture(scope="module")
def big():
    return synth.make_scene(seed=1234, size_px=10000, px=0.2, ndsm_px=0.2, density_per_km2=2500.0)


def test_full_size_p1_sampled_tiles_match_pil(dev, big):
    tables = api.TileTables(big.tiles, dev, 1)
    rgbi = torch.from_numpy(big.rgbi).to(dev)
    out, off, _ = tables.plan(rgbi).run(rgbi)
    torch.cuda.synchronize()
    wins = tables.win.tolist()
    rng = np.random.default_rng(0)
    edge = [t for t, w in enumerate(wins) if w[2] != 450 or w[3] != 450]
    pick = sorted(set(rng.choice(len(wins), 10, replace=False).tolist() + edge[:3] + edge[-3:] + [0, len(wins) - 1]))
    for t in pick:
        ref = port.tile_cut_normalize(big.rgbi, tuple(wins[t]))
        got = out[int(off[t]):int(off[t + 1])].cpu().numpy().reshape(ref.shape)
        np.testing.assert_array_equal(got, ref, err_msg=f"tile {t} window {wins[t]}")
    # every output is an integer 0..255 (uint8 path) and no tile was left unwritten
    chk = out[:: 4099]
    assert float(chk.min()) >= 0 and float(chk.max()) <= 255 and bool((chk == chk.round()).all())


def test_full_size_chain_dyn_equals_exact_and_invariants(dev, big):
    p = pipeline.PipelineParams()
    tables = api.TileTables(big.tiles, dev, p.shift)
    det = _det(big, dev)
    rgbi, ndsm = torch.from_numpy(big.rgbi).to(dev), torch.from_numpy(big.ndsm).to(dev)
    rasters = lambda: pipeline.raster_stage(rgbi, big.transform, ndsm, big.ndsm_transform, p)
    run = pipeline.ChainRunner(p)
    n0, f0 = run.collect(run.submit(det, tables.tile_tf, tables.tile_boxes, rasters))      # exact sizes
    n1, f1 = run.collect(run.submit(det, tables.tile_tf, tables.tile_boxes, rasters))      # capacity buffers
    assert run.fallbacks == 0 and n0 == n1 and len(f0) == len(f1) > 5000
    for name in ("verts", "ring_off", "poly_id", "conf", "area", "tree_height", "centroid", "is_contained",
                 "num_contained"):
        assert torch.equal(getattr(f0, name), getattr(f1, name)), name
    # invariants of process_features (postprocessing.py:571-667) on the final table
    pid = f1.poly_id.cpu().numpy()
    assert (pid >= 0).all() and pid.max() < len(big.det.scores) * 3
    assert float(f1.conf.min()) >= p.confidence_threshold
    area = f1.area.cpu().numpy()
    assert (area >= p.area_threshold).all() and (area <= 1000).all()
    h = f1.tree_height.cpu().numpy()
    assert ((h >= p.height_threshold) | (h <= -1.0)).all()
    assert int(f1.num_contained.max()) <= 1
    off = f1.ring_off.cpu().numpy()
    v = f1.verts.cpu().numpy()
    assert (np.diff(off) >= 4).all() and np.array_equal(v[off[:-1]], v[off[1:] - 1])       # closed rings
    assert np.array_equal(v, np.round(v, 3))                                              # round_coordinates
    b = geo.raster_bounds(big.transform, 10000, 10000)
    assert v[:, 0].min() >= b.left + 1 and v[:, 0].max() <= b.right - 1                    # border rule (use_overlap)


@pytest.mark.parametrize("variant,ndsm_px", [("split", 0.2), ("combined", 1.0)])
def test_full_size_config2_equals_oracle_golden(dev, big, variant, ndsm_px):
    """The bench workload itself (10 000 x 10 000 px, 34 374 instances): stitched table and final crown layer of
    the CUDA path -- exact-size kernels AND the CUDA-graph chain -- against the golden the CPU oracle produced for
    the same scene (tests/golden/make_golden_config2.py), for both nDSM resolutions of BASELINE config 2."""
    from treedetection_b200 import golden_check
    assert golden_check.golden_matches_workload(10000, 1234, 2500)
    p = pipeline.PipelineParams()
    tables = api.TileTables(big.tiles, dev, p.shift)
    det = _det(big, dev)
    rgbi = torch.from_numpy(big.rgbi).to(dev)
    if ndsm_px == 0.2:
        ndsm_np, ndsm_tf = big.ndsm, big.ndsm_transform
    else:
        ndsm_np = synth.make_ndsm(big.field, 1.0, 1234)
        ndsm_tf = synth.image_transform(big.field.left, big.field.bottom + big.field.height_m, 1.0)
    ndsm = torch.from_numpy(ndsm_np).to(dev)
    rasters = lambda: pipeline.raster_stage(rgbi, big.transform, ndsm, ndsm_tf, p)
    table = pipeline.predict_stage(**det, tile_tf=tables.tile_tf, tile_boxes=tables.tile_boxes, p=p)
    golden_check.check_table(table.verts.cpu().numpy(), table.ring_off.cpu().numpy(), table.conf.cpu().numpy())
    run = pipeline.ChainRunner(p)
    for which in ("exact", "graph", "graph replay"):
        n, f = run.collect(run.submit(det, tables.tile_tf, tables.tile_boxes, rasters))
        golden_check.check_layer(api.features_to_host(f), variant, n_candidates=n)
    assert run.fallbacks == 0


def test_config4_dense_merge_at_bench_scale(dev):
    """the ``crowns_merged`` workload of bench.py (dense-forest stress, ~2,000 crowns per 50 m tile) at 200 000 crowns
    against the oracle's sparse NMS (itself equal to the reference-pinned dense form up to 8 000, above) and the
    chunked containment"""
    rng = np.random.default_rng(4)
    n = 200_000
    cx, cy = 412000 + rng.uniform(0, 500.0, n), 5318000 + rng.uniform(0, 500.0, n)
    rx = rng.uniform(0.5, 2.0, n); ry = rx * rng.uniform(0.8, 1.25, n)
    bounds = np.stack([cx - rx, cy - ry, cx + rx, cy + ry], 1)
    conf = np.round(rng.uniform(0.3, 1.0, n), 3)
    area = np.pi * rx * ry
    removed = ops.bbox_nms_ordered(torch.from_numpy(bounds).to(dev), torch.from_numpy(conf).to(dev),
                                   torch.from_numpy(area).to(dev), 0.6, 1.0).cpu().numpy().astype(bool)
    ref = port.nms_bbox_sparse(bounds, conf, area, 0.6, 1.0)
    assert ref.sum() > 5000
    np.testing.assert_array_equal(removed, ref)
    sub = bounds[:40_000].astype(np.float32)
    ratio, isc, num = ops.containment(torch.from_numpy(sub).to(dev), 0.75)
    r_ref, isc_ref, num_ref = port.containment_chunked(sub, 0.75, chunk=256)
    np.testing.assert_array_equal(isc.cpu().numpy().astype(bool), isc_ref)
    np.testing.assert_array_equal(num.cpu().numpy(), num_ref)
    np.testing.assert_array_equal(ratio.cpu().numpy(), r_ref)
