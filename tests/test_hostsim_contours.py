"""The border-following core (treedetection_b200/csrc/contour_core.cuh), compiled for
the host, against cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE) -- the very call the
reference makes (TreeDetection/prediction.py:232-234): same contours, same points,
same order."""
import ctypes as C

import cv2
import numpy as np
import pytest

from tests import hostsim


def ours(mask, lockstep=False):
    lib = hostsim.load()
    h, w = mask.shape
    maxc, maxp = 70000, 4 * h * w + 16
    npts = np.zeros(maxc, dtype=np.int32)
    pts = np.zeros(2 * maxp, dtype=np.int16)
    counts = np.zeros(4, dtype=np.int32)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    if lockstep:
        n = lib.hs_find_contours_lockstep(m.ctypes.data_as(C.c_void_p), h, w, npts.ctypes.data_as(C.c_void_p), maxc,
                                          pts.ctypes.data_as(C.c_void_p), maxp, counts.ctypes.data_as(C.c_void_p), None)
    else:
        n = lib.hs_find_contours(m.ctypes.data_as(C.c_void_p), h, w, npts.ctypes.data_as(C.c_void_p), maxc,
                                 pts.ctypes.data_as(C.c_void_p), maxp, counts.ctypes.data_as(C.c_void_p))
    assert n >= 0, n
    out, k = [], 0
    for i in range(n):
        out.append(pts[2 * k:2 * (k + npts[i])].reshape(-1, 2).astype(np.int32))
        k += npts[i]
    return out, counts


def theirs(mask):
    cs, _ = cv2.findContours(np.ascontiguousarray(mask, dtype=np.uint8), cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
    return [c.reshape(-1, 2) for c in cs]


def check(mask):
    b = theirs(mask)
    rings = [c for c in b if c.size >= 8]
    for lockstep in (False, True):      # sequential core and the lock-step state machine
        a, counts = ours(mask, lockstep)
        assert len(a) == len(b), (lockstep, len(a), len(b))
        for p, q in zip(a, b):
            np.testing.assert_array_equal(p, q)
        assert counts[2] == len(rings)
        assert counts[3] == sum(len(c) + (0 if (c[0] == c[-1]).all() else 1) for c in rings)


def test_nested_holes_and_islands():
    m = np.zeros((20, 30), np.uint8)
    m[1:4, 1:4] = 1; m[1:4, 10:13] = 1; m[6:18, 2:20] = 1; m[8:16, 4:18] = 0
    m[10:14, 6:10] = 1; m[11:13, 7:9] = 0; m[10:12, 12:14] = 1; m[18, 25] = 1
    check(m)


def test_edges_and_full():
    check(np.ones((7, 9), np.uint8))
    check(np.zeros((5, 5), np.uint8))
    m = np.zeros((40, 70), np.uint8); m[0, :] = 1; m[:, 0] = 1; m[-1, :] = 1; m[:, -1] = 1
    check(m)
    m = np.zeros((3, 100), np.uint8); m[1, 5:95] = 1
    check(m)
    m = np.ones((33, 65), np.uint8); m[10:20, 30:40] = 0; m[14:16, 33:36] = 1
    check(m)


@pytest.mark.parametrize("seed", range(12))
def test_random_noise(seed):
    rng = np.random.default_rng(seed)
    h, w = int(rng.integers(1, 80)), int(rng.integers(1, 140))
    p = [0.2, 0.5, 0.8, 0.95][seed % 4]
    check((rng.uniform(size=(h, w)) < p).astype(np.uint8))


@pytest.mark.parametrize("seed", range(8))
def test_random_blobs(seed):
    rng = np.random.default_rng(100 + seed)
    h, w = int(rng.integers(30, 200)), int(rng.integers(30, 200))
    low = rng.normal(size=(h // 6 + 2, w // 6 + 2)).astype(np.float32)
    f = cv2.resize(low, (w, h), interpolation=cv2.INTER_CUBIC)
    check((f > 0.2).astype(np.uint8))
    check((np.abs(f) < 0.5).astype(np.uint8))


def test_diagonal_and_checkerboard():
    m = np.eye(40, dtype=np.uint8)
    check(m)
    check(m[:, ::-1])
    yy, xx = np.mgrid[0:30, 0:45]
    check(((yy + xx) % 2 == 0).astype(np.uint8))
    check(((yy % 2 == 0) & (xx % 2 == 0)).astype(np.uint8))


def _walk_slot(mask, cap_contours, cap_points):
    lib = hostsim.load()
    h, w = mask.shape
    npts = np.zeros(max(cap_contours, 1), dtype=np.int32)
    pts = np.zeros(2 * max(cap_points, 1), dtype=np.int16)
    counts = np.zeros(4, dtype=np.int32)
    m = np.ascontiguousarray(mask, dtype=np.uint8)
    rc = lib.hs_walk_slot(m.ctypes.data_as(C.c_void_p), h, w, cap_contours, cap_points,
                          counts.ctypes.data_as(C.c_void_p), npts.ctypes.data_as(C.c_void_p),
                          pts.ctypes.data_as(C.c_void_p))
    out, k = [], 0
    if rc == 0:
        for i in range(counts[0]):
            out.append(pts[2 * k:2 * (k + npts[i])].reshape(-1, 2).astype(np.int32))
            k += npts[i]
    return rc, counts, out


@pytest.mark.parametrize("seed", range(10))
def test_single_pass_walk_into_slots(seed):
    """trace_walk_kernel's walk (contours.cu): one pass with labels into capacity slots.  A slot that is
    big enough gives cv2's contours; one that is too small still counts everything, writes nothing outside
    the slot (canaries) and reports the overflow."""
    rng = np.random.default_rng(100 + seed)
    h, w = int(rng.integers(5, 70)), int(rng.integers(5, 100))
    mask = (rng.uniform(size=(h, w)) < [0.35, 0.6, 0.85][seed % 3]).astype(np.uint8)
    want = theirs(mask)
    n_pts = sum(len(c) for c in want)
    rc, counts, got = _walk_slot(mask, len(want), n_pts)            # exact fit
    assert rc == 0 and counts[0] == len(want) and counts[1] == n_pts
    for p, q in zip(got, want):
        np.testing.assert_array_equal(p, q)
    for cc, cp in [(max(len(want) // 2, 0), n_pts), (len(want), max(n_pts // 3, 0)), (0, 0), (1, 1)]:
        if cc >= len(want) and cp >= n_pts:
            continue
        rc2, counts2, _ = _walk_slot(mask, cc, cp)
        assert rc2 == 1, "overflow must be reported and no canary touched"
        np.testing.assert_array_equal(counts2, counts)


@pytest.mark.parametrize("w", [1, 2, 31, 32, 33, 63, 64, 65, 97])
def test_neighbour_masks_across_word_boundaries(w):
    """neighbours() reads one word per row plus a predicated neighbour word at bit 0 / bit 31: every pixel of random
    masks whose widths straddle the 32-bit word boundaries against a plain pixel lookup (E, NE, N, NW, W, SW, S, SE)"""
    lib = hostsim.load()
    rng = np.random.default_rng(w)
    h = 9
    mask = (rng.uniform(size=(h, w)) < 0.55).astype(np.uint8)
    got = np.zeros((h, w), dtype=np.uint8)
    lib.hs_neighbour_masks(mask.ctypes.data_as(C.c_void_p), h, w, got.ctypes.data_as(C.c_void_p))
    pad = np.zeros((h + 2, w + 2), dtype=np.uint8)
    pad[1:-1, 1:-1] = mask
    want = np.zeros((h, w), dtype=np.uint8)
    for k, (dx, dy) in enumerate([(1, 0), (1, -1), (0, -1), (-1, -1), (-1, 0), (-1, 1), (0, 1), (1, 1)]):
        want |= (pad[1 + dy:1 + dy + h, 1 + dx:1 + dx + w] << k).astype(np.uint8)
    np.testing.assert_array_equal(got, want)
