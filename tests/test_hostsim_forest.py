"""forest_core.cuh (host build) against the oracle restatement of the P10 predicates."""
import ctypes as C

import numpy as np
import pytest

from oracle import port
from tests import hostsim


def ragged(rings):
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(r) for r in rings])
    xy = np.array([p for r in rings for p in r], dtype=np.float64).reshape(-1, 2)
    return np.ascontiguousarray(xy), off


def flatten(forest):
    """forest polygons (bare rings or [shell, hole, ...]) -> (rings, poly_off)"""
    polys = [port._as_polygon(F) for F in forest]
    rings = [r for p in polys for r in p]
    poly_off = np.zeros(len(polys) + 1, dtype=np.int64)
    poly_off[1:] = np.cumsum([len(p) for p in polys])
    return rings, poly_off


def ours(rings, forest):
    lib = hostsim.load()
    frings, poly_off = flatten(forest)
    a, ao = ragged(rings); f, fo = ragged(frings)
    inter = np.zeros(len(rings), dtype=np.uint8); within = np.zeros(len(rings), dtype=np.uint8)
    vp = lambda x: x.ctypes.data_as(C.c_void_p)
    lib.hs_forest_predicates(vp(a), vp(ao), len(rings), vp(f), vp(fo), vp(poly_off), len(forest), vp(inter), vp(within))
    return inter.astype(bool), within.astype(bool)


def convex(rng, cx, cy, r, k):
    ang = np.sort(rng.uniform(0, 2 * np.pi, k))
    pts = [(float(cx + r * np.cos(a)), float(cy + r * np.sin(a))) for a in ang]
    return pts + [pts[0]]


def rect(x0, y0, x1, y1):
    return [(x1, y0), (x1, y1), (x0, y1), (x0, y0), (x1, y0)]


def make_forest(rng, n, extent):
    out = []
    for _ in range(n):
        cx, cy = 412000 + rng.uniform(0, extent), 5318000 + rng.uniform(0, extent)
        if rng.uniform() < 0.4:
            w, h = rng.uniform(20, 120, 2)
            out.append(rect(float(cx), float(cy), float(cx + w), float(cy + h)))
        else:
            out.append(convex(rng, cx, cy, rng.uniform(20, 90), int(rng.integers(5, 14))))
    return out


@pytest.mark.parametrize("seed", range(4))
def test_random_crowns_against_forest(seed):
    rng = np.random.default_rng(seed)
    forest = make_forest(rng, 25, 600.0)
    crowns = [convex(rng, 412000 + rng.uniform(-20, 620), 5318000 + rng.uniform(-20, 620), rng.uniform(1.5, 8), 9)
              for _ in range(400)]
    wi, ww = port.forest_predicates(crowns, forest)
    gi, gw = ours(crowns, forest)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gw, ww)
    assert wi.any() and (~wi).any() and ww.any() and (wi & ~ww).any()


def test_touching_and_shared_edges():
    forest = [rect(0.0, 0.0, 10.0, 10.0), rect(10.0, 0.0, 20.0, 10.0)]       # share an edge
    crowns = [rect(8.0, 2.0, 12.0, 4.0),        # spans both: within the union, not within either
              rect(18.0, 2.0, 22.0, 4.0),       # sticks out
              rect(20.0, 2.0, 24.0, 4.0),       # touches from outside: intersects, not within
              rect(30.0, 2.0, 34.0, 4.0),       # disjoint
              rect(0.0, 0.0, 10.0, 10.0)]       # identical to a forest polygon: within
    wi, ww = port.forest_predicates(crowns, forest)
    assert wi.tolist() == [True, True, True, False, True]
    assert ww.tolist() == [True, False, False, False, True]
    gi, gw = ours(crowns, forest)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gw, ww)


def test_polygons_with_holes():
    """holes of the outline (helpers.py:802-807 tests against unary_union, which keeps them): a crown inside a
    hole does not intersect the forest, a crown around a hole is not within it, a crown on the rim is both"""
    shell, hole = rect(0.0, 0.0, 100.0, 100.0), rect(40.0, 40.0, 60.0, 60.0)
    forest = [[shell, hole], rect(200.0, 0.0, 260.0, 60.0)]
    crowns = [rect(45.0, 45.0, 55.0, 55.0),      # strictly inside the hole: disjoint from the forest
              rect(30.0, 30.0, 70.0, 70.0),      # swallows the hole: intersects, not within
              rect(35.0, 45.0, 45.0, 55.0),      # straddles the hole's rim: intersects, not within
              rect(10.0, 10.0, 30.0, 30.0),      # solid part: within
              rect(39.0, 39.0, 61.0, 61.0),      # a little larger than the hole: intersects, not within
              rect(60.0, 45.0, 70.0, 55.0),      # touches the hole from the solid side: within
              rect(210.0, 10.0, 220.0, 20.0)]    # the second polygon (no holes): within
    wi, ww = port.forest_predicates(crowns, forest)
    assert wi.tolist() == [False, True, True, True, True, True, True]
    assert ww.tolist() == [False, False, False, True, False, True, True]
    gi, gw = ours(crowns, forest)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gw, ww)


@pytest.mark.parametrize("seed", range(3))
def test_random_crowns_against_forest_with_holes(seed):
    rng = np.random.default_rng(100 + seed)
    forest = []
    for F in make_forest(rng, 18, 600.0):
        xs = [p[0] for p in F]; ys = [p[1] for p in F]
        cx, cy = (min(xs) + max(xs)) / 2, (min(ys) + max(ys)) / 2
        r = min(max(xs) - min(xs), max(ys) - min(ys)) / 6
        # a hole around the centre (convex shells: the centre region is inside), orientation opposite to the shell
        forest.append([F, convex(rng, cx, cy, r, 7)[::-1]] if rng.uniform() < 0.7 else F)
    crowns = [convex(rng, 412000 + rng.uniform(-20, 620), 5318000 + rng.uniform(-20, 620), rng.uniform(1.5, 12), 9)
              for _ in range(500)]
    wi, ww = port.forest_predicates(crowns, forest)
    gi, gw = ours(crowns, forest)
    np.testing.assert_array_equal(gi, wi)
    np.testing.assert_array_equal(gw, ww)
    plain_i, plain_w = port.forest_predicates(crowns, [port._as_polygon(F)[0] for F in forest])
    assert (plain_w & ~ww).any()        # the holes matter (test_polygons_with_holes covers `intersects`)
