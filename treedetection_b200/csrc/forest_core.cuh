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

// A within the union of the forest rings F_k (k in cand[0..nc)).  ring k = fverts[foff[k] .. foff[k+1]).
// Returns 1 within, 0 not within, 2 when an edge has more than kMaxSplits crossings.
TD_HD inline int ring_within_union(const P2* A, int na, const P2* fverts, const long long* foff, const int* cand,
                                   int nc) {
  if (na < 2 || nc == 0) return 0;
  for (int i = 0; i + 1 < na; ++i) {
    const P2 a0 = A[i], a1 = A[i + 1];
    double ts[kMaxSplits + 2];
    int nt = 0;
    ts[nt++] = 0.0;
    ts[nt++] = 1.0;
    for (int c = 0; c < nc; ++c) {
      const P2* F = fverts + foff[cand[c]];
      const int nf = (int)(foff[cand[c] + 1] - foff[cand[c]]);
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
    for (int k = 0; k + 1 < nt; ++k) {
      if (!(ts[k + 1] > ts[k])) continue;
      const double tm = (ts[k] + ts[k + 1]) / 2.0;
      P2 m;
      m.x = a0.x + (a1.x - a0.x) * tm;
      m.y = a0.y + (a1.y - a0.y) * tm;
      bool covered = false;
      for (int c = 0; c < nc && !covered; ++c) {
        const P2* F = fverts + foff[cand[c]];
        const int nf = (int)(foff[cand[c] + 1] - foff[cand[c]]);
        covered = locate_in_ring(m, F, nf) >= 0;
      }
      if (!covered) return 0;
    }
  }
  return 1;
}

}  // namespace td
