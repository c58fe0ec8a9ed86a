// P10 core -- crown polygon vs forest-outline predicates.
//
// Restates the shapely/GEOS calls of the two-model fusion and of the tile flags:
//     forest_shapes.geometry.intersects(forest_union)          TreeDetection/helpers.py:806
//     ~urban_shapes.geometry.within(forest_union)              TreeDetection/helpers.py:807
//     candidates.intersects(bbox) / unary_union.contains(bbox) TreeDetection/preprocessing.py:86-91
// for single-ring polygons (crowns, tile boxes) against the UNION of single-ring forest
// polygons, without ever building the union:
//   intersects(A, U)  <=>  A intersects some F_i: an edge pair intersects, or a vertex of one
//                          lies inside the other (closed sets, touching counts);
//   within(A, U)      <=>  every piece of A's boundary lies in some closed F_i: each edge of A
//                          is split at its crossings with the forest edges and the midpoint
//                          of every piece is located by ray crossing.
// Point location and segment crossing tests use the exact orientation predicate of
// simplify_core.cuh; split parameters and midpoints are ordinary doubles (identical
// arithmetic in oracle/port.py, so GPU and oracle agree bit for bit).  GEOS is an un-pinned
// third-party dependency of the reference: parity with a real GEOS build is unpinned, and
// two documented simplifications apply -- forest polygons are taken without holes, and a
// hole of the union that lies entirely inside a crown is not detected.
//
// Plain C++ (tests/hostsim compiles it with g++).
#pragma once
#include "simplify_core.cuh"

namespace td {

// 1 inside, 0 on the boundary, -1 outside (GEOS RayCrossingCounter on a closed ring)
TD_HD inline int locate_in_ring(const P2& p, const P2* ring, int n) {
  int crossings = 0;
  for (int k = 0; k + 1 < n; ++k) {
    const P2 p1 = ring[k], p2 = ring[k + 1];
    if (p1.x < p.x && p2.x < p.x) continue;
    if (p.x == p2.x && p.y == p2.y) return 0;
    if (p1.y == p.y && p2.y == p.y) {
      const double minx = fmin(p1.x, p2.x), maxx = fmax(p1.x, p2.x);
      if (p.x >= minx && p.x <= maxx) return 0;
      continue;
    }
    if ((p1.y > p.y && p2.y <= p.y) || (p2.y > p.y && p1.y <= p.y)) {
      int sign = orientation(p1.x, p1.y, p2.x, p2.y, p.x, p.y);
      if (sign == 0) return 0;
      if (p2.y < p1.y) sign = -sign;
      if (sign > 0) ++crossings;
    }
  }
  return (crossings & 1) ? 1 : -1;
}

// closed segments share at least one point
TD_HD inline bool segments_touch(const P2& p1, const P2& p2, const P2& q1, const P2& q2) {
  if (!env_overlap(p1, p2, q1, q2)) return false;
  const int o1 = orientation(p1.x, p1.y, p2.x, p2.y, q1.x, q1.y);
  const int o2 = orientation(p1.x, p1.y, p2.x, p2.y, q2.x, q2.y);
  if ((o1 > 0 && o2 > 0) || (o1 < 0 && o2 < 0)) return false;
  const int o3 = orientation(q1.x, q1.y, q2.x, q2.y, p1.x, p1.y);
  const int o4 = orientation(q1.x, q1.y, q2.x, q2.y, p2.x, p2.y);
  if ((o3 > 0 && o4 > 0) || (o3 < 0 && o4 < 0)) return false;
  return true;
}

struct Box2 {
  double minx, miny, maxx, maxy;
};
TD_HD inline bool boxes_overlap(const Box2& a, const Box2& b) {
  return !(a.minx > b.maxx || a.maxx < b.minx || a.miny > b.maxy || a.maxy < b.miny);
}

TD_HD inline bool ring_intersects_ring(const P2* A, int na, const P2* F, int nf) {
  for (int i = 0; i + 1 < na; ++i)
    for (int j = 0; j + 1 < nf; ++j)
      if (segments_touch(A[i], A[i + 1], F[j], F[j + 1])) return true;
  if (na > 0 && locate_in_ring(A[0], F, nf) >= 0) return true;
  if (nf > 0 && locate_in_ring(F[0], A, na) >= 0) return true;
  return false;
}

// ---- forest polygons with holes -----------------------------------------------------------------
// Polygon k owns the rings poly_off[k] .. poly_off[k + 1): the first is its shell, the others are holes.
// ring r = fverts[foff[r] .. foff[r + 1]).
struct ForestSet {
  const P2* fverts;
  const long long* foff;       // ring offsets
  const long long* poly_off;   // polygon -> first ring; null: every ring is a polygon of its own
  TD_HD long long ring0(int k) const { return poly_off ? poly_off[k] : k; }
  TD_HD long long ring1(int k) const { return poly_off ? poly_off[k + 1] : k + 1; }
  TD_HD const P2* ring(long long r) const { return fverts + foff[r]; }
  TD_HD int ring_len(long long r) const { return (int)(foff[r + 1] - foff[r]); }
};

// 1 interior, 0 boundary (shell or hole), -1 exterior (outside the shell or strictly inside a hole)
TD_HD inline int locate_in_polygon(const P2& p, const ForestSet& S, int k) {
  const long long r0 = S.ring0(k), r1 = S.ring1(k);
  const int s = locate_in_ring(p, S.ring(r0), S.ring_len(r0));
  if (s <= 0) return s;
  for (long long r = r0 + 1; r < r1; ++r) {
    const int h = locate_in_ring(p, S.ring(r), S.ring_len(r));
    if (h == 1) return -1;
    if (h == 0) return 0;
  }
  return 1;
}

// closed ring A (as an areal polygon) shares a point with polygon k
TD_HD inline bool ring_intersects_polygon(const P2* A, int na, const ForestSet& S, int k) {
  const long long r0 = S.ring0(k), r1 = S.ring1(k);
  for (long long r = r0; r < r1; ++r) {
    const P2* F = S.ring(r);
    const int nf = S.ring_len(r);
    for (int i = 0; i + 1 < na; ++i)
      for (int j = 0; j + 1 < nf; ++j)
        if (segments_touch(A[i], A[i + 1], F[j], F[j + 1])) return true;
  }
  // no boundary contact: A lies in one face of the polygon's ring arrangement, or swallows the shell
  if (na > 0 && locate_in_polygon(A[0], S, k) >= 0) return true;
  if (S.ring_len(r0) > 0 && locate_in_ring(S.ring(r0)[0], A, na) >= 0) return true;
  return false;
}

// candidate polygons of a query: a list, or (list == null) every polygon whose bounds overlap `q`
struct CandSet {
  const int* list;
  int n;
  const double* bounds;   // (n_poly, 4)
  Box2 q;
  int n_poly;
  bool strict;            // strict inequalities (the tile-flag rule)
  TD_HD bool overlaps(int k) const {
    const double* b = bounds + 4 * (size_t)k;
    if (strict) return b[2] > q.minx && b[0] < q.maxx && b[3] > q.miny && b[1] < q.maxy;
    return !(q.minx > b[2] || q.maxx < b[0] || q.miny > b[3] || q.maxy < b[1]);
  }
  TD_HD int scan(int k) const {
    for (; k < n_poly; ++k)
      if (overlaps(k)) return k;
    return -1;
  }
  TD_HD int first() const { return list ? (n > 0 ? 0 : -1) : scan(0); }
  TD_HD int next(int cur) const { return list ? (cur + 1 < n ? cur + 1 : -1) : scan(cur + 1); }
  TD_HD int poly(int cur) const { return list ? list[cur] : cur; }
};

// GEOS validity of a closed ring taken as a polygon shell: no two non-adjacent segments share a point, and
// adjacent segments share only their common vertex (they may not fold back onto each other).  This is the
// `is_valid` test that decides which geometries the reference hands to buffer(0) (helpers.py:816-821).
TD_HD inline bool ring_is_simple(const P2* A, int n) {
  if (n < 4 || !same(A[0], A[n - 1])) return false;
  const int ns = n - 1;
  for (int i = 0; i < ns; ++i) {
    for (int j = i + 1; j < ns; ++j) {
      const bool adjacent = (j == i + 1) || (i == 0 && j == ns - 1);
      if (!adjacent) {
        if (segments_touch(A[i], A[i + 1], A[j], A[j + 1])) return false;
        continue;
      }
      // shared vertex v, the two far ends a, b: a fold back means a, v, b collinear with a and b on the same side
      const P2 v = (j == i + 1) ? A[j] : A[0];
      const P2 a = (j == i + 1) ? A[i] : A[1];
      const P2 b = (j == i + 1) ? A[j + 1] : A[ns - 1];
      if (same(a, v) || same(b, v)) continue;          // repeated point: a zero-length segment, harmless
      if (orientation(a.x, a.y, v.x, v.y, b.x, b.y) != 0) continue;
      if ((a.x - v.x) * (b.x - v.x) + (a.y - v.y) * (b.y - v.y) > 0.0) return false;
    }
  }
  return true;
}

constexpr int kMaxSplits = 62;

// parameter of the intersection of segment a0->a1 with q1->q2 along a0->a1 (doubles).
// Returns the number of parameters written (0, 1, or 2 for a collinear overlap).
TD_HD inline int split_params(const P2& a0, const P2& a1, const P2& q1, const P2& q2, double* t) {
  if (!segments_touch(a0, a1, q1, q2)) return 0;
  const double dx = a1.x - a0.x, dy = a1.y - a0.y;
  const double ex = q2.x - q1.x, ey = q2.y - q1.y;
  const double den = dx * ey - dy * ex;
  if (den != 0.0) {
    t[0] = ((q1.x - a0.x) * ey - (q1.y - a0.y) * ex) / den;
    return 1;
  }
  // parallel and touching: collinear overlap -- project the other segment's ends
  const double len2 = dx * dx + dy * dy;
  if (len2 == 0.0) return 0;
  t[0] = ((q1.x - a0.x) * dx + (q1.y - a0.y) * dy) / len2;
  t[1] = ((q2.x - a0.x) * dx + (q2.y - a0.y) * dy) / len2;
  return 2;
}

// A within the union of the candidate forest polygons.  Every edge of A is split at its crossings with
// the forest edges (shells AND holes) and the midpoint of every piece must lie in some closed polygon;
// a hole with a vertex strictly inside A means A's interior leaves the forest there (not recognised: a hole
// that another forest polygon happens to cover, and the measure-zero case of a hole all of whose vertices
// lie exactly ON A's boundary).
// Returns 1 within, 0 not within, 2 when an edge has more than kMaxSplits crossings.
TD_HD inline int ring_within_union(const P2* A, int na, const ForestSet& S, const CandSet& C) {
  if (na < 2 || C.first() < 0) return 0;
  for (int i = 0; i + 1 < na; ++i) {
    const P2 a0 = A[i], a1 = A[i + 1];
    double ts[kMaxSplits + 2];
    int nt = 0;
    ts[nt++] = 0.0;
    ts[nt++] = 1.0;
    for (int c = C.first(); c >= 0; c = C.next(c)) {
      const int k = C.poly(c);
      for (long long r = S.ring0(k); r < S.ring1(k); ++r) {
        const P2* F = S.ring(r);
        const int nf = S.ring_len(r);
        for (int j = 0; j + 1 < nf; ++j) {
          double t2[2];
          const int m = split_params(a0, a1, F[j], F[j + 1], t2);
          for (int q = 0; q < m; ++q) {
            if (!(t2[q] > 0.0 && t2[q] < 1.0)) continue;
            if (nt >= kMaxSplits + 2) return 2;
            // insertion keeps ts sorted
            int pos = nt++;
            while (pos > 0 && ts[pos - 1] > t2[q]) { ts[pos] = ts[pos - 1]; --pos; }
            ts[pos] = t2[q];
          }
        }
      }
    }
    for (int k = 0; k + 1 < nt; ++k) {
      if (!(ts[k + 1] > ts[k])) continue;
      const double tm = (ts[k] + ts[k + 1]) / 2.0;
      P2 m;
      m.x = a0.x + (a1.x - a0.x) * tm;
      m.y = a0.y + (a1.y - a0.y) * tm;
      bool covered = false;
      for (int c = C.first(); c >= 0 && !covered; c = C.next(c)) covered = locate_in_polygon(m, S, C.poly(c)) >= 0;
      if (!covered) return 0;
    }
  }
  for (int c = C.first(); c >= 0; c = C.next(c)) {
    const int k = C.poly(c);
    for (long long r = S.ring0(k) + 1; r < S.ring1(k); ++r) {
      const P2* Hh = S.ring(r);
      const int nh = S.ring_len(r);
      for (int j = 0; j + 1 < nh; ++j)
        if (locate_in_ring(Hh[j], A, na) == 1) return 0;
    }
  }
  return 1;
}

}  // namespace td
