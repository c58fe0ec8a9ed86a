"""GPU parity: every kernel family called through the C-ABI (ctypes) against the CPU
oracle on the same seeded inputs.  Bit-exact for masks / keep flags / indices,
1e-5 absolute for float statistics (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import port
from treedetection_b200 import ops, synth

pytestmark = pytest.mark.gpu

ATOL = 1e-5  # float statistics tolerance stated by north_star


def _unpack(bits, off, win, i):
    x0, y0, w, h = [int(v) for v in win[i]]
    wpr = (w + 31) // 32
    words = bits[off[i]:off[i] + wpr * h].astype(np.uint32).reshape(h, wpr)
    b = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)[None, None, :]) & 1).reshape(h, wpr * 32)
    return b[:, :w].astype(bool)


def _small_scene():
    return synth.make_scene(seed=7, size_px=1500, px=0.2, ndsm_px=1.0, density_per_km2=4000.0)


def test_paste_bits_and_values(dev):
    sc = _small_scene()
    d = sc.det
    n = d.boxes_net.shape[0]
    assert n > 200
    boxes_px, win, nwords = ops.paste_plan(torch.from_numpy(d.boxes_net).to(dev), torch.from_numpy(d.inst_tile).to(dev),
                                           torch.from_numpy(d.tile_dims).to(dev))
    off = ops.exclusive_offsets(nwords)
    bits = ops.paste_threshold_pack(boxes_px, win, off, torch.from_numpy(d.probs).to(dev))
    vals, voff = ops.paste_values(boxes_px, win, torch.from_numpy(d.probs).to(dev))
    boxes_px = boxes_px.cpu().numpy(); win = win.cpu().numpy(); off = off.cpu().numpy()
    bits = bits.cpu().numpy(); vals = vals.cpu().numpy(); voff = voff.cpu().numpy()
    worst = 0.0
    for i in range(0, n, 3):
        t = d.inst_tile[i]
        th, tw, nh, nw = d.tile_dims[t]
        b_ref, keep = port.scale_clip_boxes(d.boxes_net[i:i + 1], (nh, nw), (th, tw))
        assert keep[0]
        np.testing.assert_array_equal(b_ref[0], boxes_px[i])
        v_ref, (x0, y0, x1, y1) = port.paste_probs(b_ref[0], d.probs[i], th, tw)
        assert (x0, y0, x1 - x0, y1 - y0) == tuple(int(v) for v in win[i])
        got = vals[voff[i]:voff[i + 1]].reshape(y1 - y0, x1 - x0)
        worst = max(worst, float(np.abs(got - v_ref).max()))
        v_cf, _ = port.paste_probs_closed_form(b_ref[0], d.probs[i], th, tw)
        np.testing.assert_array_equal(got, v_cf)             # bit-identical to the closed form
        np.testing.assert_array_equal(_unpack(bits, off, win, i), v_cf >= np.float32(0.5))
        np.testing.assert_array_equal(_unpack(bits, off, win, i), v_ref >= np.float32(0.5))
    assert worst <= ATOL


def test_paste_drops_empty_boxes(dev):
    boxes = np.array([[10, 10, 10, 50], [5, 5, 60, 70], [900, 900, 950, 950]], dtype=np.float32)
    tile_dims = np.array([[450, 450, 800, 800]], dtype=np.int32)
    bpx, win, nwords = ops.paste_plan(torch.from_numpy(boxes).to(dev), torch.zeros(3, dtype=torch.int32, device=dev),
                                      torch.from_numpy(tile_dims).to(dev))
    b_ref, keep = port.scale_clip_boxes(boxes, (800, 800), (450, 450))
    np.testing.assert_array_equal(bpx.cpu().numpy(), b_ref)
    np.testing.assert_array_equal((nwords.cpu().numpy() > 0), keep)


def _random_crowns(rng, n, extent=300.0):
    cx = 412000 + rng.uniform(0, extent, n); cy = 5318000 + rng.uniform(0, extent, n)
    rx = rng.uniform(1.5, 6, n); ry = rx * rng.uniform(0.8, 1.2, n)
    bounds = np.stack([cx - rx, cy - ry, cx + rx, cy + ry], 1)
    conf = rng.uniform(0.3, 1.0, n)
    area = np.pi * rx * ry
    return bounds, conf, area


@pytest.mark.parametrize("n,iou,athr", [(1, 0.6, 1), (2, 0.1, 1), (700, 0.6, 1), (1500, 0.2, 1), (1500, 0.05, 0.4),
                                        (900, 0.0, 1), (300, -1.0, 1)])
def test_nms_matches_oracle(dev, n, iou, athr):
    rng = np.random.default_rng(n + int(iou * 100))
    bounds, conf, area = _random_crowns(rng, n, extent=120.0 if n > 1000 else 80.0)
    # duplicated confidences exercise the first-max tie rule in float16
    conf = np.round(conf, 2)
    ref = port.nms_bbox(bounds, conf, area, iou, athr)
    got = ops.bbox_nms_ordered(torch.from_numpy(bounds).to(dev), torch.from_numpy(conf).to(dev),
                               torch.from_numpy(area).to(dev), iou, athr).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(got, ref)
    if n > 100:
        assert ref.sum() > 0


def test_nms_large_sparse(dev):
    rng = np.random.default_rng(5)
    bounds, conf, area = _random_crowns(rng, 60000, extent=1500.0)
    ref = port.nms_bbox_sparse(bounds, conf, area, 0.3, 1)
    got = ops.bbox_nms_ordered(torch.from_numpy(bounds).to(dev), torch.from_numpy(conf).to(dev),
                               torch.from_numpy(area).to(dev), 0.3, 1).cpu().numpy().astype(bool)
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("n,thr", [(1, 0.75), (600, 0.75), (1500, 0.3), (400, 0.0)])
def test_containment_matches_oracle(dev, n, thr):
    rng = np.random.default_rng(n)
    bounds, _, _ = _random_crowns(rng, n, extent=100.0)
    b32 = bounds.astype(np.float32)
    r_ref, c_ref, n_ref = port.containment(b32, thr)
    ratio, isc, num = ops.containment(torch.from_numpy(b32).to(dev), thr)
    np.testing.assert_array_equal(isc.cpu().numpy().astype(bool), c_ref)
    np.testing.assert_array_equal(num.cpu().numpy(), n_ref)
    np.testing.assert_array_equal(ratio.cpu().numpy(), r_ref)


def _rings(rng, n, left, bottom, extent):
    rings = []
    for _ in range(n):
        cx = left + rng.uniform(0, extent); cy = bottom + rng.uniform(0, extent)
        r = rng.uniform(1.5, 6.0); k = int(rng.integers(5, 40))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = r * rng.uniform(0.6, 1.0, k)
        pts = [(cx + rad[j] * np.cos(ang[j]), cy + rad[j] * np.sin(ang[j])) for j in range(k)]
        pts.append(pts[0])
        rings.append(pts)
    return rings


def _ragged(rings, dev):
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.array([p for r in rings for p in r], dtype=np.float64).reshape(-1, 2)
    return torch.from_numpy(verts).to(dev), torch.from_numpy(off).to(dev)


@pytest.mark.parametrize("px", [1.0, 0.2])
def test_crown_stats_all_modes(dev, px):
    rng = np.random.default_rng(11)
    size = 300 if px == 1.0 else 600
    left, top = 412000.0, 5318000.0 + size * px
    tf = (px, 0.0, left, 0.0, -px, top)
    ndvi = rng.uniform(-1, 1, (size, size)).astype(np.float32)
    height = np.round(rng.uniform(0, 30, (size, size)), 1).astype(np.float32)  # ties for the arg-max rule
    rings = _rings(rng, 60, left - 5, top - size * px - 5, size * px + 10)     # some crowns leave the raster
    rings += _rings(rng, 3, left - 500, top + 300, 20)                           # entirely outside: empty sets
    px32, py32 = port.pad_polygons(rings)
    verts, off = _ragged(rings, dev)
    nd = torch.from_numpy(ndvi).to(dev); hd = torch.from_numpy(height).to(dev)
    ref = port.crown_stats_combined(px32, py32, ndvi, height, tf)
    got = ops.crown_stats(verts, off, nd, hd, tf, ops.STATS_COMBINED)
    np.testing.assert_array_equal(got["max_h"].cpu().numpy(), ref["max_h"])
    np.testing.assert_array_equal(got["hxy"].cpu().numpy()[:, 0], ref["hx"])
    np.testing.assert_array_equal(got["hxy"].cpu().numpy()[:, 1], ref["hy"])
    st = got["ndvi"].cpu().numpy()
    np.testing.assert_array_equal(st[:, 0], ref["ndvi_min"])
    np.testing.assert_array_equal(st[:, 1], ref["ndvi_max"])
    np.testing.assert_allclose(st[:, 2], ref["ndvi_mean"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(st[:, 3], ref["ndvi_var"], atol=ATOL, rtol=0)
    ref_h = port.crown_stats_height(px32, py32, height, tf)
    got_h = ops.crown_stats(verts, off, None, hd, tf, ops.STATS_HEIGHT_ONLY)
    np.testing.assert_array_equal(got_h["max_h"].cpu().numpy(), ref_h["max_h"])
    np.testing.assert_array_equal(got_h["hxy"].cpu().numpy()[:, 0], ref_h["hx"])
    np.testing.assert_array_equal(got_h["hxy"].cpu().numpy()[:, 1], ref_h["hy"])
    ref_n = port.crown_stats_ndvi(px32, py32, ndvi, tf)
    got_n = ops.crown_stats(verts, off, nd, None, tf, ops.STATS_NDVI_ONLY)["ndvi"].cpu().numpy()
    np.testing.assert_array_equal(got_n[:, 0], ref_n["ndvi_min"])
    np.testing.assert_array_equal(got_n[:, 1], ref_n["ndvi_max"])
    np.testing.assert_allclose(got_n[:, 2], ref_n["ndvi_mean"], atol=ATOL, rtol=0)
    np.testing.assert_allclose(got_n[:, 3], ref_n["ndvi_var"], atol=ATOL, rtol=0)
    assert (ref["max_h"] == -1).any() and (ref["max_h"] > 0).any()


def test_centroids_bit_exact(dev):
    rng = np.random.default_rng(3)
    rings = _rings(rng, 300, 412000.0, 5318000.0, 500.0)
    # one long ring forces V > 128 and numpy's recursive pairwise blocking
    k = 333
    ang = np.linspace(0, 2 * np.pi, k)
    rings.append([(412100 + 9 * np.cos(a), 5318100 + 9 * np.sin(a)) for a in ang])
    px32, py32 = port.pad_polygons(rings)
    ref = port.centroids(px32, py32)
    verts, off = _ragged(rings, dev)
    got = ops.centroids(verts, off).cpu().numpy()
    np.testing.assert_array_equal(got, ref.astype(np.float32))


@pytest.mark.parametrize("shape,factor", [((1000, 1000), 0.2), ((730, 1210), 0.2), ((500, 640), 0.5), ((300, 300), 1.0),
                                          ((611, 977), 1 / 3), ((700, 910), 1 / 7), ((1500, 1100), 0.1),
                                          ((333, 20), 0.25)])
def test_ndvi_decimate(dev, shape, factor):
    """decimation factors 2:1 ... 10:1, ragged sizes, a raster narrower than one CTA tile"""
    rng = np.random.default_rng(shape[0])
    rgbi = rng.integers(0, 256, size=(4,) + shape, dtype=np.uint8)
    oh, ow = int(shape[0] * factor), int(shape[1] * factor)
    red = port.decimate_bilinear(rgbi[0], oh, ow); nir = port.decimate_bilinear(rgbi[3], oh, ow)
    dec = np.stack([red, rgbi[1][:oh, :ow], rgbi[2][:oh, :ow], nir])
    ref = port.ndvi_from_rgbi(dec).astype(np.float32)
    got = ops.ndvi_decimate(torch.from_numpy(rgbi).to(dev), oh, ow).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


def test_ndvi_decimate_unaligned_view(dev):
    """the fast kernel fetches aligned words around each run of taps: a raster that starts at an odd byte offset
    inside its allocation (a band slice of a larger tensor, odd band size) must give the same NDVI"""
    rng = np.random.default_rng(77)
    big = rng.integers(0, 256, size=(6, 335, 501), dtype=np.uint8)
    rgbi = big[1:5]
    oh, ow = 67, 100
    red = port.decimate_bilinear(rgbi[0], oh, ow); nir = port.decimate_bilinear(rgbi[3], oh, ow)
    dec = np.stack([red, rgbi[1][:oh, :ow], rgbi[2][:oh, :ow], nir])
    ref = port.ndvi_from_rgbi(dec).astype(np.float32)
    d = torch.from_numpy(big).to(dev)[1:5]
    assert d.data_ptr() % 4 != 0 and d.is_contiguous()
    got = ops.ndvi_decimate(d, oh, ow).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


def test_decimate_f32(dev):
    rng = np.random.default_rng(2)
    band = rng.uniform(0, 40, (900, 1100)).astype(np.float32)
    ref = port.decimate_bilinear(band, 180, 220)
    got = ops.decimate_f32(torch.from_numpy(band).to(dev), 180, 220).cpu().numpy()
    np.testing.assert_array_equal(got, ref)


@pytest.mark.parametrize("out_hw", [(180, 220), (300, 367), (129, 157)])
def test_decimate_f32_with_nodata_and_nan(dev, out_hw):
    """nDSM rasters carry nodata (-3.4e38) and may carry NaN / Inf: they spread exactly as far as the oracle's tap
    loop spreads them (a weight-0 tap at the edge of the support included)"""
    rng = np.random.default_rng(5)
    band = rng.uniform(0, 40, (900, 1100)).astype(np.float32)
    band[rng.integers(0, 900, 300), rng.integers(0, 1100, 300)] = np.nan
    band[rng.integers(0, 900, 100), rng.integers(0, 1100, 100)] = np.inf
    band[rng.integers(0, 900, 300), rng.integers(0, 1100, 300)] = np.float32(-3.4028234663852886e38)
    band[:, -1] = np.nan           # the last column: the padded taps of the last output must stay inside the row
    with np.errstate(invalid="ignore", over="ignore"):
        ref = port.decimate_bilinear(band, *out_hw)
    got = ops.decimate_f32(torch.from_numpy(band).to(dev), *out_hw).cpu().numpy()
    np.testing.assert_array_equal(np.isnan(got), np.isnan(ref))
    np.testing.assert_array_equal(got[~np.isnan(ref)], ref[~np.isnan(ref)])


def test_mask_iou_clean_matches_restated_clean_crowns(dev):
    """opt-in ``iou_mode: mask``: popcount IoU over the packed P2 rasters + the rule of the reference's (uncalled)
    clean_crowns (helpers.py:602-701) against its NumPy restatement on the same masks -- instances of overlapping
    tiles (every tree is seen by ~3 tiles), exact duplicates, empty masks"""
    sc = synth.make_scene(seed=13, size_px=900, px=0.2, ndsm_px=1.0, density_per_km2=6000.0, with_rasters=False)
    d = sc.det
    # duplicate a few instances (same mask, another confidence) and blank one out
    dup = np.array([3, 40, 41, 200])
    boxes = np.concatenate([d.boxes_net, d.boxes_net[dup]]); probs = np.concatenate([d.probs, d.probs[dup]])
    inst_tile = np.concatenate([d.inst_tile, d.inst_tile[dup]])
    scores = np.concatenate([d.scores, np.array([0.99, d.scores[40], 0.31, 0.5], np.float32)])
    order = np.argsort(inst_tile, kind="stable")          # keep the tile-major layout
    boxes, probs, inst_tile, scores = boxes[order], probs[order], inst_tile[order], scores[order]
    probs = probs.copy(); probs[7] = 0.0                   # an instance whose mask is empty
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    boxes_px, win, nwords = ops.paste_plan(t(boxes), t(inst_tile), t(d.tile_dims))
    off = ops.exclusive_offsets(nwords)
    bits = ops.paste_threshold_pack(boxes_px, win, off, t(probs))
    tile_org = np.array([m["window"][:2] for m in sc.tiles.values()], dtype=np.int32)
    for thr, conf in ((0.7, 0.2), (0.3, 0.45), (0.95, 0.0)):
        keep, match, best = ops.mask_iou_clean(bits, off, win, t(tile_org), t(inst_tile), t(scores), thr, conf)
        # restatement on the unpacked masks
        b, o, w = bits.cpu().numpy().view(np.uint32), off.cpu().numpy(), win.cpu().numpy()
        masks, orgs = [], []
        for i in range(len(scores)):
            x0, y0, ww, hh = w[i]
            wpr = (ww + 31) // 32
            words = b[o[i]:o[i] + wpr * hh].reshape(hh, wpr) if ww * hh else np.zeros((0, 0), np.uint32)
            m = ((words[:, :, None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(hh, -1)[:, :ww].astype(bool) \
                if ww * hh else np.zeros((0, 0), bool)
            masks.append(m); orgs.append((tile_org[inst_tile[i]][0] + x0, tile_org[inst_tile[i]][1] + y0))
        wkeep, wmatch = port.mask_iou_clean(masks, orgs, scores, thr, conf)
        np.testing.assert_array_equal(match.cpu().numpy(), wmatch)
        np.testing.assert_array_equal(keep.cpu().numpy().astype(bool), wkeep)
        assert 0 < wkeep.sum() < len(wkeep)
    assert not wkeep[7] and wmatch[7] == -1


@pytest.mark.parametrize("q", [50.0, 95.0, 0.0, 100.0, 33.3])
def test_crown_height_summary_matches_numpy(dev, q):
    """optional nDSM min / mean / percentile per crown (radix select + numpy's linear interpolation) against NumPy
    on the pixel set of get_height_within_polygon; crowns outside the raster (-1), ties and NaN pixels included"""
    rng = np.random.default_rng(int(q))
    left, bottom, size_m, px = synth.ORIGIN_X + 100.0, synth.ORIGIN_Y + 50.0, 80.0, 0.5
    n_px = int(size_m / px)
    tf = synth.image_transform(left, bottom + size_m, px)
    height = np.round(rng.uniform(0, 30, (n_px, n_px)), 1).astype(np.float32)      # rounded: many equal values
    height[20:24, 30:40] = np.nan
    rings = []
    for _ in range(120):
        cx, cy = left + rng.uniform(-10, size_m + 10), bottom + rng.uniform(-10, size_m + 10)
        r = rng.uniform(0.3, 9.0)
        ang = np.sort(rng.uniform(0, 2 * np.pi, 12))
        ring = [(float(cx + r * np.cos(a)), float(cy + r * np.sin(a))) for a in ang]
        rings.append(ring + [ring[0]])
    px32, py32 = port.pad_polygons(rings)
    want = port.crown_height_summary(px32, py32, height, tf, q)
    off = np.zeros(len(rings) + 1, dtype=np.int64); off[1:] = np.cumsum([len(r) for r in rings])
    verts = torch.from_numpy(np.array([p for r in rings for p in r], dtype=np.float64)).to(dev)
    got = ops.crown_height_summary(verts, torch.from_numpy(off).to(dev), torch.from_numpy(height).to(dev), tf, q).cpu().numpy()
    np.testing.assert_array_equal(got[:, 3], want[:, 3])                 # the pixel sets
    np.testing.assert_array_equal(got[:, 0], want[:, 0])                 # min (NaN where the set holds a NaN)
    np.testing.assert_array_equal(got[:, 2], want[:, 2])                 # percentile: exact order statistics
    np.testing.assert_allclose(got[:, 1], want[:, 1], atol=ATOL, rtol=0, equal_nan=True)
    assert (want[:, 3] == 0).any() and np.isnan(want[:, 0]).any() and (want[:, 3] > 50).any()
