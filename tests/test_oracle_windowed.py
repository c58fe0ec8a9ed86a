"""The full-size forms of the oracle (windowed crown statistics, chunked containment, sparse NMS: the
``large=True`` switch of oracle.port.post_process) against the literal N x P / N x N forms they replace
-- the licence for freezing a golden of the full 10 000 x 10 000 px BASELINE config 2 scene
(tests/golden/make_golden_config2.py)."""
import numpy as np
import pytest

from oracle import port
from treedetection_b200 import geo, pipeline, synth


def _crowns(rng, n, left, bottom, size_m):
    rings = []
    for _ in range(n):
        cx = left + rng.uniform(-8, size_m + 8); cy = bottom + rng.uniform(-8, size_m + 8)   # some stick out / lie outside
        r = rng.uniform(0.4, 9.0)
        k = int(rng.integers(5, 30))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        ring = [(float(cx + r * np.cos(a)), float(cy + r * 1.1 * np.sin(a))) for a in ang]
        rings.append(ring + [ring[0]])
    return rings


@pytest.mark.parametrize("px_h", [1.0, 0.2, 0.5])
def test_windowed_stats_equal_literal_forms(px_h):
    rng = np.random.default_rng(int(px_h * 10))
    size_m = 60.0
    left, bottom = synth.ORIGIN_X + 333.0, synth.ORIGIN_Y + 777.0
    rings = _crowns(rng, 70, left, bottom, size_m)
    px32, py32 = port.pad_polygons(rings)
    top = bottom + size_m
    n_px = int(round(size_m / px_h))
    tf = synth.image_transform(left, top, px_h)
    height = rng.uniform(0, 30, (n_px, n_px)).astype(np.float32)
    height[rng.uniform(size=height.shape) < 0.01] = np.nan                 # NaN beats numbers in argmax
    height[5:9, 5:9] = 30.0                                                # ties: first index wins
    ndvi = rng.uniform(-1, 1, (n_px, n_px)).astype(np.float32)
    a = port.crown_stats_combined(px32, py32, ndvi, height, tf)
    b = port.crown_stats_combined_windowed(px32, py32, ndvi, height, tf)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    a = port.crown_stats_height(px32, py32, height, tf)
    b = port.crown_stats_height_windowed(px32, py32, height, tf)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    a = port.crown_stats_ndvi(px32, py32, ndvi, tf)
    b = port.crown_stats_ndvi_windowed(px32, py32, ndvi, tf)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k], err_msg=k)
    assert (a["ndvi_mean"] == -1).any() and (a["ndvi_mean"] != -1).any()    # empty and non-empty sets


def test_chunked_containment_equals_literal_form():
    rng = np.random.default_rng(3)
    n = 1300
    c = rng.uniform(0, 300, (n, 2)); r = rng.uniform(0.5, 12, (n, 2))
    b = np.concatenate([c - r, c + r], axis=1)
    b[7] = b[8]                           # identical boxes
    b[11, 2] = b[11, 0]                   # degenerate (zero-area) inner box -> NaN / inf ratios
    for thr in (0.6, 0.75, 0.9):
        want = port.containment(b, thr)
        got = port.containment_chunked(b, thr, chunk=97)
        for x, y in zip(want, got):
            np.testing.assert_array_equal(x, y)


def test_large_post_process_equals_literal_post_process():
    """the whole post-processing with and without the full-size forms, split and combined rasters"""
    for ndsm_px, seed in ((0.2, 41), (1.0, 42)):
        sc = synth.make_scene(seed=seed, size_px=700, px=0.2, ndsm_px=ndsm_px, density_per_km2=7000.0)
        p = pipeline.PipelineParams()
        rings, conf = port.predict_stage(sc.det, sc.tiles)
        H, W = sc.rgbi.shape[1:]
        oh, ow = int(H * 0.2), int(W * 0.2)
        dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in (0, 0, 0, 3)])
        ndvi = port.ndvi_from_rgbi(dec).astype(np.float32)
        ndvi_tf = geo.compose(sc.transform, geo.scale(W / ow, H / oh))
        h, w = sc.ndsm.shape
        cfg = {k: getattr(p, k) for k in p.__dataclass_fields__}
        args = (rings, conf, ndvi, ndvi_tf, tuple(geo.raster_bounds(sc.transform, W, H)), sc.ndsm, sc.ndsm_transform,
                tuple(geo.raster_bounds(sc.ndsm_transform, w, h)), 0.2, 0.2, cfg)
        a, da = port.post_process(*args)
        b, db = port.post_process(*args, large=True)
        assert a == b and len(a) > 20
        assert da["ids_after_nms"] == db["ids_after_nms"] and da["combined"] == db["combined"] == (ndsm_px == 1.0)
