"""simplify_core.cuh (host build) against the GEOS-semantics oracle (oracle/geom.py):
identical kept-vertex lists, bit-identical areas, identical predicates."""
import ctypes as C
from fractions import Fraction

import cv2
import numpy as np
import pytest

from oracle import geom, port
from tests import hostsim


def _lib():
    lib = hostsim.load()
    lib.hs_orientation.argtypes = [C.c_double] * 6
    lib.hs_orientation_exact.argtypes = [C.c_double] * 6
    lib.hs_orientation_nonzero.argtypes = [C.c_double] * 6
    lib.hs_simplify_ring.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
    return lib


def ours(ring, tol):
    lib = _lib()
    xy = np.ascontiguousarray(np.asarray(ring, dtype=np.float64).reshape(-1, 2))
    keep = np.zeros(len(xy) + 1, dtype=np.int32)
    area = C.c_double(0.0)
    m = lib.hs_simplify_ring(xy.ctypes.data_as(C.c_void_p), len(xy), float(tol), keep.ctypes.data_as(C.c_void_p),
                             C.byref(area))
    return [tuple(xy[k]) for k in keep[:m]], area.value


def test_orientation_exact_on_near_degenerate_inputs():
    lib = _lib()
    rng = np.random.default_rng(0)
    n_fallback = 0
    for _ in range(4000):
        ax, ay = 412000 + rng.uniform(0, 100), 5318000 + rng.uniform(0, 100)
        dx, dy = rng.uniform(-5, 5), rng.uniform(-5, 5)
        t1, t2 = rng.uniform(0, 3, 2)
        bx, by = ax + t1 * dx, ay + t1 * dy
        cx, cy = ax + t2 * dx, ay + t2 * dy          # collinear up to rounding
        if rng.uniform() < 0.3:
            cx = np.nextafter(cx, np.inf)
        ax, ay, bx, by, cx, cy = [float(v) for v in (ax, ay, bx, by, cx, cy)]
        fa = [Fraction(v) for v in (ax, ay, bx, by, cx, cy)]
        d = (fa[0] - fa[4]) * (fa[3] - fa[5]) - (fa[1] - fa[5]) * (fa[2] - fa[4])
        want = (d > 0) - (d < 0)
        assert lib.hs_orientation(ax, ay, bx, by, cx, cy) == want
        assert lib.hs_orientation_exact(ax, ay, bx, by, cx, cy) == want
        assert lib.hs_orientation_nonzero(ax, ay, bx, by, cx, cy) == (want != 0)   # the guard's neighbour test
        assert geom.orientation(ax, ay, bx, by, cx, cy) == want
        n_fallback += 1
    assert n_fallback


def _contour_rings(seed, tf):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(40, 160)), int(rng.integers(40, 160))
    low = rng.normal(size=(h // 8 + 2, w // 8 + 2)).astype(np.float32)
    f = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
    return port.mask_to_polygons(f > 0.1, tf)


@pytest.mark.parametrize("seed", range(10))
@pytest.mark.parametrize("tol", [0.2, 2.0])
def test_simplify_matches_oracle_on_contour_rings(seed, tol):
    tf = (0.2, 0.0, 412030.0, 0.0, -0.2, 5318090.0)
    rings = _contour_rings(seed, tf)
    assert rings
    for ring in rings:
        want = geom.simplify_ring([tuple(p) for p in ring], tol)
        got, area = ours(ring, tol)
        assert got == want
        assert area == abs(geom.ring_signed_area(want))


def test_simplify_random_star_polygons():
    rng = np.random.default_rng(5)
    for _ in range(40):
        k = int(rng.integers(4, 120))
        ang = np.sort(rng.uniform(0, 2 * np.pi, k))
        rad = rng.uniform(0.5, 8.0, k)
        ring = [(float(412000 + 50 + rad[j] * np.cos(ang[j])), float(5318000 + 50 + rad[j] * np.sin(ang[j])))
                for j in range(k)]
        ring.append(ring[0])
        for tol in (0.2, 1.0, 2.0, 5.0):
            want = geom.simplify_ring(ring, tol)
            got, area = ours(ring, tol)
            assert got == want
            assert area == abs(geom.ring_signed_area(want))


def test_simplify_keeps_minimum_ring():
    sq = [(0.0, 0.0), (1.0, 0.0), (1.0, 1.0), (0.0, 1.0), (0.0, 0.0)]
    got, area = ours(sq, 10.0)
    assert got == geom.simplify_ring(sq, 10.0)
    assert len(got) >= 4 and area > 0


def _degenerate_rings(rng):
    """Rings that drive the intersection guard through its collinear / touching branches: runs of
    collinear vertices, spikes (A -> B -> A), repeated points, axis-parallel combs whose teeth are
    thinner than the tolerance.  The neighbour shortcut of simplify_ring (one orientation instead of
    the full predicate when a neighbour segment is not collinear with the chord) must not change any
    of them with respect to the full GEOS-semantics restatement in oracle/geom.py."""
    out = []
    for k in range(60):
        pts = []
        x, y = 412000.0 + k, 5318000.0
        n = int(rng.integers(6, 40))
        kind = k % 4
        for i in range(n):
            if kind == 0:      # staircase with long collinear runs
                x += float(rng.choice([0.0, 0.2, 0.4])); y += float(rng.choice([0.0, 0.2])) if i % 3 == 0 else 0.0
            elif kind == 1:    # comb: teeth of 0.1 m on a base line
                x += 0.2; y = 5318000.0 + (0.1 if i % 2 else 0.0)
            elif kind == 2:    # spikes and repeated points
                step = rng.choice([-0.2, 0.0, 0.2], size=2)
                x += float(step[0]); y += float(step[1])
            else:              # diagonal runs (collinear at 45 degrees) with kinks
                d = 0.2 * float(rng.integers(1, 4))
                x += d; y += d if i % 5 else -d
            pts.append((x, y))
        # come back below the start so that the ring has area, then close
        pts.append((pts[-1][0], 5317999.0 - (k % 3)))
        pts.append((pts[0][0], 5317999.0 - (k % 3)))
        pts.append(pts[0])
        out.append(pts)
    return out


@pytest.mark.parametrize("tol", [0.05, 0.2, 2.0])
def test_simplify_collinear_and_touching_neighbours(tol):
    rng = np.random.default_rng(99)
    checked = 0
    for ring in _degenerate_rings(rng):
        want = geom.simplify_ring([tuple(p) for p in ring], tol)
        got, _ = ours(ring, tol)
        assert got == [tuple(map(float, q)) for q in want], (tol, ring)
        checked += 1
    assert checked == 60


@pytest.mark.parametrize("tol", [0.2, 2.0])
def test_tile_box_prefilter_is_conservative(tol):
    """simplify_kernel rejects a ring WITHOUT simplifying it when its bounds stick out of the tile box by
    more than the tolerance (geometry.cu).  That is only allowed if such a ring can never be `within` the
    box after simplification: checked here on contour rings against boxes cutting through them."""
    tf = (0.2, 0.0, 412000.0, 0.0, -0.2, 5318100.0)
    rng = np.random.default_rng(7)
    n_pre = n_rings = 0
    for seed in range(12):
        for ring in _contour_rings(seed, tf):
            xy = np.asarray(ring, dtype=np.float64)
            simp = np.asarray(geom.simplify_ring([tuple(p) for p in ring], tol), dtype=np.float64)
            x0, y0, x1, y1 = xy[:, 0].min(), xy[:, 1].min(), xy[:, 0].max(), xy[:, 1].max()
            for _ in range(6):
                # a box edge somewhere near the ring's own extent
                j = tol + 1.0
                bx0 = x0 + rng.uniform(-j, j); by0 = y0 + rng.uniform(-j, j)
                bx1 = x1 + rng.uniform(-j, j); by1 = y1 + rng.uniform(-j, j)
                slack = tol * (1.0 + 1e-9) + 1e-9
                pre_reject = x0 < bx0 - slack or y0 < by0 - slack or x1 > bx1 + slack or y1 > by1 + slack
                within = bool(len(simp)) and simp[:, 0].min() >= bx0 and simp[:, 1].min() >= by0 and \
                    simp[:, 0].max() <= bx1 and simp[:, 1].max() <= by1
                assert not (pre_reject and within)
                n_pre += pre_reject
                n_rings += 1
    assert n_pre > 50 and n_rings - n_pre > 50


def test_orientation_nonzero_equals_orientation_on_grid_and_random_points():
    """the branch-free neighbour test of the simplifier's guard: same answer as orientation() != 0 on exactly
    collinear grid points (axis-parallel, diagonal), zero products, mixed signs and random triples"""
    lib = _lib()
    rng = np.random.default_rng(5)
    pts = []
    for _ in range(3000):
        ox, oy = 412000.0 + float(rng.integers(0, 5000)) * 0.2, 5318000.0 - float(rng.integers(0, 5000)) * 0.2
        a = (ox, oy)
        kind = int(rng.integers(0, 5))
        if kind == 0:      # horizontal / vertical run (one product exactly zero)
            b, c = (ox + 0.2 * float(rng.integers(1, 9)), oy), (ox - 0.2 * float(rng.integers(0, 9)), oy)
        elif kind == 1:
            b, c = (ox, oy + 0.2 * float(rng.integers(1, 9))), (ox, oy - 0.2 * float(rng.integers(0, 9)))
        elif kind == 2:    # diagonal on the pixel grid: collinear up to the rounding of the coordinates
            k, m = float(rng.integers(1, 9)), float(rng.integers(-9, 9))
            b, c = (ox + 0.2 * k, oy + 0.2 * k), (ox + 0.2 * m, oy + 0.2 * m)
        elif kind == 3:    # c == a or c == b
            b = (ox + float(rng.uniform(-3, 3)), oy + float(rng.uniform(-3, 3)))
            c = a if rng.uniform() < 0.5 else b
        else:
            b = (ox + float(rng.uniform(-3, 3)), oy + float(rng.uniform(-3, 3)))
            c = (ox + float(rng.uniform(-3, 3)), oy + float(rng.uniform(-3, 3)))
        pts.append((a, b, c))
    n_zero = 0
    for a, b, c in pts:
        want = lib.hs_orientation(a[0], a[1], b[0], b[1], c[0], c[1])
        assert lib.hs_orientation_nonzero(a[0], a[1], b[0], b[1], c[0], c[1]) == (want != 0)
        n_zero += want == 0
    assert n_zero > 300
