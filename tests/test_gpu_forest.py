"""GPU parity of P10: forest-outline predicates, tile flags and the file-level fusion against
the oracle restatement (GEOS semantics, parity with a real GEOS build unpinned)."""
import os

import numpy as np
import pytest
import torch

from oracle import port
from treedetection_b200 import fusion, gpkg, ops, synth, tiling

pytestmark = pytest.mark.gpu


def convex(rng, cx, cy, r, k):
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    pts = [(float(cx + r * np.cos(a)), float(cy + r * np.sin(a))) for a in ang]
    return pts + [pts[0]]


def rect(x0, y0, x1, y1):
    return [(x1, y0), (x1, y1), (x0, y1), (x0, y0), (x1, y0)]


def make_forest(rng, n, extent, left=412000.0, bottom=5318000.0):
    out = []
    for _ in range(n):
        cx, cy = left + rng.uniform(0, extent), bottom + rng.uniform(0, extent)
        if rng.uniform() < 0.4:
            w, h = rng.uniform(20, 120, 2)
            out.append(rect(float(cx), float(cy), float(cx + w), float(cy + h)))
        else:
            out.append(convex(rng, cx, cy, rng.uniform(20, 90), int(rng.integers(5, 14))))
    return out


def ragged(rings, dev):
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rings])
    xy = np.array([p for r in rings for p in r], dtype=np.float64).reshape(-1, 2)
    return torch.from_numpy(np.ascontiguousarray(xy)).to(dev), torch.from_numpy(off).to(dev)


@pytest.mark.parametrize("seed", [0, 1])
def test_forest_predicates_match_oracle(dev, seed):
    rng = np.random.default_rng(seed)
    forest = make_forest(rng, 30, 700.0)
    forest += [rect(412100.0, 5318100.0, 412200.0, 5318200.0), rect(412200.0, 5318100.0, 412300.0, 5318200.0)]
    crowns = [convex(rng, 412000 + rng.uniform(-20, 720), 5318000 + rng.uniform(-20, 720), rng.uniform(1.5, 8), 9)
              for _ in range(1500)]
    crowns += [rect(412190.0, 5318120.0, 412210.0, 5318130.0), rect(412100.0, 5318100.0, 412200.0, 5318200.0)]
    wi, ww = port.forest_predicates(crowns, forest)
    av, ao = ragged(crowns, dev); fv, fo = ragged(forest, dev)
    gi, gw = ops.forest_predicates(av, ao, fv, fo)
    np.testing.assert_array_equal(gi.cpu().numpy().astype(bool), wi)
    np.testing.assert_array_equal(gw.cpu().numpy().astype(bool), ww)
    assert ww[-2] and ww[-1] and (wi & ~ww).any() and (~wi).any()


def test_tile_flags_match_oracle(dev):
    rng = np.random.default_rng(3)
    forest = make_forest(rng, 12, 500.0) + [rect(412000.0 - 50, 5318000.0 - 50, 412260.0, 5318260.0)]
    idx = fusion.ForestIndex([np.array(f) for f in forest], dev)
    tf = synth.image_transform(412000.0, 5318500.0, 0.25)
    tiles = tiling.tile_grid("img", tf, 2000, 2000, 25832, 50, 50, 20, forest=idx)
    n_forest = n_urban = 0
    for tid, m in tiles.items():
        parts = [int(p) for p in tid.split("_")[-5:]]
        box = (float(parts[0]), float(parts[1]), float(parts[0] + 50), float(parts[1] + 50))
        b = m["bounds"]
        want = port.tile_flags(box, rect(b[0], b[1], b[2], b[3]), forest)
        assert (m["only_forest"], m["only_urban"]) == want, tid
        n_forest += m["only_forest"]; n_urban += m["only_urban"]
    assert n_forest > 0 and n_urban > 0 and n_forest + n_urban < len(tiles)


def test_fuse_predictions_files(dev, tmp_path):
    rng = np.random.default_rng(9)
    forest = make_forest(rng, 20, 400.0)
    urban = [convex(rng, 412000 + rng.uniform(0, 400), 5318000 + rng.uniform(0, 400), rng.uniform(2, 6), 8)
             for _ in range(300)]
    fcrowns = [convex(rng, 412000 + rng.uniform(0, 400), 5318000 + rng.uniform(0, 400), rng.uniform(2, 6), 8)
               for _ in range(250)]
    uconf = rng.uniform(0.3, 1, len(urban)); fconf = rng.uniform(0.3, 1, len(fcrowns))
    ud, fd, od = tmp_path / "urban_geojson", tmp_path / "forrest_geojson", tmp_path / "geojson_predictions"
    ud.mkdir(); fd.mkdir()

    def dump(path, layer, rings, cols=None):
        off = np.zeros(len(rings) + 1, dtype=np.int64); off[1:] = np.cumsum([len(r) for r in rings])
        xy = np.array([p for r in rings for p in r]).reshape(-1, 2)
        gpkg.write_layer(str(path), layer, xy, off, cols or {}, gpkg.STITCHED_SCHEMA if cols else {})

    dump(ud / "img.gpkg", "img", urban, {"Confidence_score": uconf})
    dump(fd / "img.gpkg", "img", fcrowns, {"Confidence_score": fconf})
    dump(tmp_path / "forest_outline.gpkg", "forest", forest)
    fusion.fuse_predictions(str(ud), str(fd), str(tmp_path / "forest_outline.gpkg"), str(od), device=dev)
    v, o, cols, _ = gpkg.read_layer(str(od / "img.gpkg"))
    keep_f, keep_u = port.fuse(urban, fcrowns, forest)
    want = [fcrowns[i] for i in keep_f] + [urban[i] for i in keep_u]
    assert len(o) - 1 == len(want) and 0 < len(keep_f) < len(fcrowns) and 0 < len(keep_u) < len(urban)
    np.testing.assert_array_equal(v, np.array([p for r in want for p in r]).reshape(-1, 2))
    np.testing.assert_array_equal(np.array(cols["Confidence_score"]), np.concatenate([fconf[keep_f], uconf[keep_u]]))
    assert os.path.exists(od / "fusion_recovery.yaml")
