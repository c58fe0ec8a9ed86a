"""GEOS-semantics planar geometry restated in pure Python (TEST INFRASTRUCTURE).

shapely / GEOS are not installed in this image and the reference pins neither
(``/root/reference/setup.py:22``), so this module *defines* the results of the
handful of shapely calls the hot path makes:

* ``Polygon(coords)``, ``.bounds``, ``.exterior.xy``     helpers.py:457, postprocessing.py:497-503
* ``.simplify(tol, preserve_topology=True)``              helpers.py:464, postprocessing.py:749
* ``.area``                                               postprocessing.py:750 (utilities.py:100-110)
* ``within(box)`` for a rectangle                         helpers.py:466-468
* ``shape(geojson)``, ``box(...)``                        postprocessing.py:490, helpers.py:296

The algorithms follow GEOS' published sources (3.11 line, the one bundled with
shapely 2.0 wheels): ``TopologyPreservingSimplifier`` /
``TaggedLineStringSimplifier`` (Douglas-Peucker with a minimum ring size of 4
and an interior-intersection guard against the input and output segment
indexes), ``Distance::pointToSegment``, ``Area::ofRingSigned``,
``LineIntersector::computeIntersect`` + ``isInteriorIntersection`` with an
exact orientation predicate, and ``RectangleContains`` (envelope containment
for areal geometries).  Parity with a real GEOS build is unpinned.

All arithmetic is IEEE double without fused multiply-add, in the written
order, so the CUDA implementation can reproduce it bit for bit.
"""
from __future__ import annotations

import math
from fractions import Fraction

# ----------------------------------------------------------------------------
# predicates
# ----------------------------------------------------------------------------
_ORIENT_ERRBOUND = 3.3306690738754716e-16  # Shewchuk ccwerrboundA = (3 + 16 eps) eps


def orientation(ax, ay, bx, by, cx, cy) -> int:
    """Sign of the exact determinant |b-a, c-a| on the double inputs.

    GEOS ``Orientation::index`` (CGAlgorithmsDD with a floating filter).  A
    Shewchuk-style static filter decides almost every call; the fallback is
    exact rational arithmetic.
    """
    detleft = (ax - cx) * (by - cy)
    detright = (ay - cy) * (bx - cx)
    det = detleft - detright
    if detleft > 0.0:
        if detright <= 0.0:
            return (det > 0) - (det < 0)
        detsum = detleft + detright
    elif detleft < 0.0:
        if detright >= 0.0:
            return (det > 0) - (det < 0)
        detsum = -detleft - detright
    else:
        return (det > 0) - (det < 0)
    errbound = _ORIENT_ERRBOUND * detsum
    if det >= errbound or -det >= errbound:
        return (det > 0) - (det < 0)
    # exact
    fax, fay, fbx, fby, fcx, fcy = map(Fraction, (ax, ay, bx, by, cx, cy))
    d = (fax - fcx) * (fby - fcy) - (fay - fcy) * (fbx - fcx)
    return (d > 0) - (d < 0)


def _env_intersects_pt(p1, p2, q) -> bool:
    """GEOS Envelope::intersects(p1, p2, q): q inside the envelope of p1-p2."""
    return (q[0] >= min(p1[0], p2[0]) and q[0] <= max(p1[0], p2[0]) and
            q[1] >= min(p1[1], p2[1]) and q[1] <= max(p1[1], p2[1]))


def _env_intersects_seg(p1, p2, q1, q2) -> bool:
    """GEOS Envelope::intersects(p1, p2, q1, q2)."""
    minq = min(q1[0], q2[0]); maxq = max(q1[0], q2[0])
    minp = min(p1[0], p2[0]); maxp = max(p1[0], p2[0])
    if minp > maxq or maxp < minq:
        return False
    minq = min(q1[1], q2[1]); maxq = max(q1[1], q2[1])
    minp = min(p1[1], p2[1]); maxp = max(p1[1], p2[1])
    if minp > maxq or maxp < minq:
        return False
    return True


def interior_intersection(p1, p2, q1, q2) -> bool:
    """GEOS LineIntersector::computeIntersect + isInteriorIntersection().

    True when the two closed segments share a point that is not an endpoint of
    both of them.
    """
    if not _env_intersects_seg(p1, p2, q1, q2):
        return False
    Pq1 = orientation(p1[0], p1[1], p2[0], p2[1], q1[0], q1[1])
    Pq2 = orientation(p1[0], p1[1], p2[0], p2[1], q2[0], q2[1])
    if (Pq1 > 0 and Pq2 > 0) or (Pq1 < 0 and Pq2 < 0):
        return False
    Qp1 = orientation(q1[0], q1[1], q2[0], q2[1], p1[0], p1[1])
    Qp2 = orientation(q1[0], q1[1], q2[0], q2[1], p2[0], p2[1])
    if (Qp1 > 0 and Qp2 > 0) or (Qp1 < 0 and Qp2 < 0):
        return False
    if Pq1 == 0 and Pq2 == 0 and Qp1 == 0 and Qp2 == 0:
        pts = _collinear_points(p1, p2, q1, q2)
    elif Pq1 == 0 or Pq2 == 0 or Qp1 == 0 or Qp2 == 0:
        if p1 == q1 or p1 == q2:
            pts = [p1]
        elif p2 == q1 or p2 == q2:
            pts = [p2]
        elif Pq1 == 0:
            pts = [q1]
        elif Pq2 == 0:
            pts = [q2]
        elif Qp1 == 0:
            pts = [p1]
        else:
            pts = [p2]
    else:
        return True  # proper crossing: interior to both
    for ip in pts:
        if not (ip == p1 or ip == p2):
            return True
        if not (ip == q1 or ip == q2):
            return True
    return False


def _collinear_points(p1, p2, q1, q2):
    a = _env_intersects_pt(p1, p2, q1)   # p1q1p2
    b = _env_intersects_pt(p1, p2, q2)   # p1q2p2
    c = _env_intersects_pt(q1, q2, p1)   # q1p1q2
    d = _env_intersects_pt(q1, q2, p2)   # q1p2q2
    if a and b:
        return [q1, q2]
    if c and d:
        return [p1, p2]
    if a and c:
        return [q1] if (q1 == p1 and not b and not d) else [q1, p1]
    if a and d:
        return [q1] if (q1 == p2 and not b and not c) else [q1, p2]
    if b and c:
        return [q2] if (q2 == p1 and not a and not d) else [q2, p1]
    if b and d:
        return [q2] if (q2 == p2 and not a and not c) else [q2, p2]
    return []


def point_segment_distance(p, A, B) -> float:
    """GEOS Distance::pointToSegment."""
    if A[0] == B[0] and A[1] == B[1]:
        dx = p[0] - A[0]; dy = p[1] - A[1]
        return math.sqrt(dx * dx + dy * dy)
    len2 = (B[0] - A[0]) * (B[0] - A[0]) + (B[1] - A[1]) * (B[1] - A[1])
    r = ((p[0] - A[0]) * (B[0] - A[0]) + (p[1] - A[1]) * (B[1] - A[1])) / len2
    if r <= 0.0:
        dx = p[0] - A[0]; dy = p[1] - A[1]
        return math.sqrt(dx * dx + dy * dy)
    if r >= 1.0:
        dx = p[0] - B[0]; dy = p[1] - B[1]
        return math.sqrt(dx * dx + dy * dy)
    s = ((A[1] - p[1]) * (B[0] - A[0]) - (A[0] - p[0]) * (B[1] - A[1])) / len2
    return abs(s) * math.sqrt(len2)


# ----------------------------------------------------------------------------
# TopologyPreservingSimplifier on one closed ring
# ----------------------------------------------------------------------------
def simplify_ring(pts, tol, min_size=4):
    """GEOS TaggedLineStringSimplifier::simplify on a single ring.

    ``pts``: list of (x, y) with pts[0] == pts[-1].  Returns the simplified
    list (closed).  Input index = all original segments not yet flattened,
    output index = flattened segments (single-ring polygon: one line).
    """
    n = len(pts)
    if n == 0:
        return []
    in_alive = [True] * (n - 1)          # input segment k = (pts[k], pts[k+1])
    out_segs = []                         # flattened output segments (i, j)
    result = []                           # result segments (i, j) in order

    def has_bad_intersection(i, j):
        a = pts[i]; b = pts[j]
        for (oi, oj) in out_segs:
            if interior_intersection(pts[oi], pts[oj], a, b):
                return True
        for k in range(n - 1):
            if not in_alive[k]:
                continue
            if i <= k < j:
                continue                  # isInLineSection
            if interior_intersection(pts[k], pts[k + 1], a, b):
                return True
        return False

    # explicit stack instead of recursion (depth is carried along)
    stack = [(0, n - 1, 0)]
    while stack:
        i, j, depth = stack.pop()
        depth += 1
        if i + 1 == j:
            result.append((i, j))
            continue
        valid = True
        rsize = 0 if not result else len(result) + 1
        if rsize < min_size:
            if depth + 1 < min_size:
                valid = False
        maxd = -1.0
        far = i
        A = pts[i]; B = pts[j]
        for k in range(i + 1, j):
            d = point_segment_distance(pts[k], A, B)
            if d > maxd:
                maxd = d
                far = k
        if maxd > tol:
            valid = False
        # GEOS evaluates hasBadIntersection unconditionally; it has no side
        # effects, so short-circuiting keeps the result identical.
        if valid and has_bad_intersection(i, j):
            valid = False
        if valid:
            for k in range(i, j):
                in_alive[k] = False
            out_segs.append((i, j))
            result.append((i, j))
            continue
        stack.append((far, j, depth))     # processed second
        stack.append((i, far, depth))     # processed first
    out = [pts[s[0]] for s in result]
    out.append(pts[result[-1][1]])
    return out


def ring_signed_area(pts) -> float:
    """GEOS Area::ofRingSigned (shoelace relative to the first x)."""
    n = len(pts)
    if n < 3:
        return 0.0
    x0 = pts[0][0]
    p1x = pts[0][0]; p1y = pts[0][1]
    p2x = pts[1][0] - x0; p2y = pts[1][1]
    s = 0.0
    for i in range(1, n - 1):
        p0y = p1y
        p1x = p2x; p1y = p2y
        p2x = pts[i + 1][0] - x0; p2y = pts[i + 1][1]
        s += p1x * (p0y - p2y)
    return s / 2.0


# ----------------------------------------------------------------------------
# minimal shapely object model
# ----------------------------------------------------------------------------
class _Ring:
    def __init__(self, coords):
        self.coords = [(float(c[0]), float(c[1])) for c in coords]

    @property
    def xy(self):
        return [c[0] for c in self.coords], [c[1] for c in self.coords]


class Polygon:
    geom_type = "Polygon"

    def __init__(self, shell=None, holes=None):
        shell = [] if shell is None else [(float(c[0]), float(c[1])) for c in shell]
        if shell and shell[0] != shell[-1]:
            shell = shell + [shell[0]]     # shapely closes rings implicitly
        self.exterior = _Ring(shell)
        self.interiors = [_Ring(h) for h in (holes or [])]

    @property
    def is_empty(self):
        return len(self.exterior.coords) == 0

    @property
    def bounds(self):
        xs, ys = self.exterior.xy
        return (min(xs), min(ys), max(xs), max(ys))

    @property
    def area(self):
        a = abs(ring_signed_area(self.exterior.coords))
        for h in self.interiors:
            a -= abs(ring_signed_area(h.coords))
        return a

    def simplify(self, tolerance, preserve_topology=True):
        if not preserve_topology:
            raise NotImplementedError("only preserve_topology=True is on the hot path")
        # single-ring fast path is exact; with holes GEOS shares the indexes
        # across rings -- handled by simplify_rings below.
        rings = [self.exterior.coords] + [h.coords for h in self.interiors]
        if len(rings) == 1:
            return Polygon(simplify_ring(rings[0], tolerance))
        out = simplify_rings(rings, tolerance)
        return Polygon(out[0], out[1:])

    def within(self, other):
        # other must be a rectangle built by box(): GEOS RectangleContains
        ob = other.bounds
        b = self.bounds
        return b[0] >= ob[0] and b[1] >= ob[1] and b[2] <= ob[2] and b[3] <= ob[3]

    @property
    def __geo_interface__(self):
        return {"type": "Polygon",
                "coordinates": [list(self.exterior.coords)] + [list(h.coords) for h in self.interiors]}


class MultiPolygon:
    geom_type = "MultiPolygon"

    def __init__(self, polys=()):
        self.geoms = list(polys)

    @property
    def bounds(self):
        bs = [p.bounds for p in self.geoms]
        return (min(b[0] for b in bs), min(b[1] for b in bs), max(b[2] for b in bs), max(b[3] for b in bs))


def simplify_rings(rings, tol):
    """TaggedLinesSimplifier over several rings of one polygon: rings are
    simplified in order, sharing the input (all rings) and output indexes."""
    n_r = len(rings)
    alive = [[True] * (len(r) - 1) for r in rings]
    out_segs = []      # (ring, i, j)
    results = []
    for ri, pts in enumerate(rings):
        n = len(pts)
        result = []
        stack = [(0, n - 1, 0)]
        while stack:
            i, j, depth = stack.pop()
            depth += 1
            if i + 1 == j:
                result.append((i, j)); continue
            valid = True
            rsize = 0 if not result else len(result) + 1
            if rsize < 4 and depth + 1 < 4:
                valid = False
            maxd = -1.0; far = i
            for k in range(i + 1, j):
                d = point_segment_distance(pts[k], pts[i], pts[j])
                if d > maxd:
                    maxd = d; far = k
            if maxd > tol:
                valid = False
            if valid:
                a = pts[i]; b = pts[j]
                bad = False
                for (orr, oi, oj) in out_segs:
                    if interior_intersection(rings[orr][oi], rings[orr][oj], a, b):
                        bad = True; break
                if not bad:
                    for rr in range(n_r):
                        for k in range(len(rings[rr]) - 1):
                            if not alive[rr][k]:
                                continue
                            if rr == ri and i <= k < j:
                                continue
                            if interior_intersection(rings[rr][k], rings[rr][k + 1], a, b):
                                bad = True; break
                        if bad:
                            break
                if bad:
                    valid = False
            if valid:
                for k in range(i, j):
                    alive[ri][k] = False
                out_segs.append((ri, i, j))
                result.append((i, j)); continue
            stack.append((far, j, depth))
            stack.append((i, far, depth))
        out = [pts[s[0]] for s in result]
        out.append(pts[result[-1][1]])
        results.append(out)
    return results


def box(minx, miny, maxx, maxy):
    return Polygon([(maxx, miny), (maxx, maxy), (minx, maxy), (minx, miny), (maxx, miny)])


def shape(geom):
    t = geom["type"]
    if t == "Polygon":
        c = geom["coordinates"]
        return Polygon(c[0], c[1:])
    if t == "MultiPolygon":
        return MultiPolygon([Polygon(c[0], c[1:]) for c in geom["coordinates"]])
    raise ValueError(f"unsupported geometry type {t}")
