// P4 / P9 core -- GEOS-semantics ring simplification, area and bounds.
//
// Restates the shapely calls the hot path makes:
//     Polygon(coords).simplify(tol, preserve_topology=True)   TreeDetection/helpers.py:464
//     shape(geom).simplify(2) ... .area                       TreeDetection/postprocessing.py:749-750
//     polygon.bounds                                          postprocessing.py:497-503
// i.e. GEOS TopologyPreservingSimplifier on a single closed ring (Douglas-Peucker with
// a minimum ring size of 4 and the interior-intersection guard against the input and
// output segment sets), Distance::pointToSegment, Area::ofRingSigned and the robust
// LineIntersector with an exact orientation predicate.  shapely / GEOS are un-pinned
// third-party dependencies of the reference and absent here: oracle/geom.py defines
// the results and this file reproduces them bit for bit (IEEE double, no contraction;
// the orientation predicate falls back to exact expansion arithmetic).
//
// Plain C++ so that tests/hostsim compiles the same code with g++.
#pragma once
#include <cmath>

#include "common.cuh"

namespace td {

struct P2 {
  double x, y;
};

// ---- exact orientation ------------------------------------------------------
TD_HD inline void two_sum(double a, double b, double& s, double& e) {
  s = a + b;
  const double bb = s - a;
  e = (a - (s - bb)) + (b - bb);
}
TD_HD inline void two_prod(double a, double b, double& p, double& e) {
  p = a * b;
  e = fma(a, b, -p);
}

// adds the double v to the non-overlapping expansion e[0..n) (increasing magnitude)
TD_HD inline int grow_expansion(double* e, int n, double v) {
  double q = v;
  int m = 0;
  for (int i = 0; i < n; ++i) {
    double s, err;
    two_sum(q, e[i], s, err);
    if (err != 0.0) e[m++] = err;
    q = s;
  }
  if (q != 0.0 || m == 0) e[m++] = q;
  return m;
}

// sign of (ax-cx)(by-cy) - (ay-cy)(bx-cx) evaluated exactly on the double inputs
// (rare fallback: kept out of line on the device so that the callers stay small)
#if defined(__CUDACC__)
__noinline__
#endif
TD_HD inline int orientation_exact(double ax, double ay, double bx, double by, double cx, double cy) {
  // = ax*by - ax*cy - cx*by - ay*bx + ay*cx + cy*bx   (cx*cy cancels)
  const double pa[6] = {ax, -ax, -cx, -ay, ay, cy};
  const double pb[6] = {by, cy, by, bx, cx, bx};
  double e[16];
  int n = 0;
  for (int k = 0; k < 6; ++k) {
    double p, err;
    two_prod(pa[k], pb[k], p, err);
    n = grow_expansion(e, n, err);
    n = grow_expansion(e, n, p);
  }
  const double top = e[n - 1];
  return (top > 0.0) - (top < 0.0);
}

TD_HD inline int orientation(double ax, double ay, double bx, double by, double cx, double cy) {
  const double detleft = (ax - cx) * (by - cy);
  const double detright = (ay - cy) * (bx - cx);
  const double det = detleft - detright;
  double detsum;
  if (detleft > 0.0) {
    if (detright <= 0.0) return (det > 0.0) - (det < 0.0);
    detsum = detleft + detright;
  } else if (detleft < 0.0) {
    if (detright >= 0.0) return (det > 0.0) - (det < 0.0);
    detsum = -detleft - detright;
  } else {
    return (det > 0.0) - (det < 0.0);
  }
  const double errbound = 3.3306690738754716e-16 * detsum;
  if (det >= errbound || -det >= errbound) return (det > 0.0) - (det < 0.0);
  return orientation_exact(ax, ay, bx, by, cx, cy);
}

// orientation(...) != 0 without the sign cases.  With detsum = |detleft| + |detright| the filter of
// orientation() reads |det| >= 3.33e-16 * detsum in every branch: where the two products differ in sign (or
// one is zero) det = +-detsum exactly, so the test holds whenever orientation() returns sign(det) unfiltered,
// and where they agree in sign it is orientation()'s own error bound.
TD_HD inline bool orientation_nonzero(double ax, double ay, double bx, double by, double cx, double cy) {
  const double detleft = (ax - cx) * (by - cy);
  const double detright = (ay - cy) * (bx - cx);
  const double det = detleft - detright;
  const double detsum = fabs(detleft) + fabs(detright);
  if (fabs(det) >= 3.3306690738754716e-16 * detsum) return det != 0.0;
  return orientation_exact(ax, ay, bx, by, cx, cy) != 0;
}

// ---- robust segment intersection: "is there an interior intersection" --------
TD_HD inline bool env_has_pt(const P2& p1, const P2& p2, const P2& q) {
  return q.x >= fmin(p1.x, p2.x) && q.x <= fmax(p1.x, p2.x) && q.y >= fmin(p1.y, p2.y) && q.y <= fmax(p1.y, p2.y);
}
TD_HD inline bool env_overlap(const P2& p1, const P2& p2, const P2& q1, const P2& q2) {
  double minq = fmin(q1.x, q2.x), maxq = fmax(q1.x, q2.x);
  double minp = fmin(p1.x, p2.x), maxp = fmax(p1.x, p2.x);
  if (minp > maxq || maxp < minq) return false;
  minq = fmin(q1.y, q2.y); maxq = fmax(q1.y, q2.y);
  minp = fmin(p1.y, p2.y); maxp = fmax(p1.y, p2.y);
  if (minp > maxq || maxp < minq) return false;
  return true;
}
TD_HD inline bool same(const P2& a, const P2& b) { return a.x == b.x && a.y == b.y; }

// an intersection point ip is "interior" if it is not an endpoint of both segments
TD_HD inline bool ip_interior(const P2& ip, const P2& p1, const P2& p2, const P2& q1, const P2& q2) {
  if (!(same(ip, p1) || same(ip, p2))) return true;
  if (!(same(ip, q1) || same(ip, q2))) return true;
  return false;
}

TD_HD inline bool interior_intersection(const P2& p1, const P2& p2, const P2& q1, const P2& q2) {
  if (!env_overlap(p1, p2, q1, q2)) return false;
  const int Pq1 = orientation(p1.x, p1.y, p2.x, p2.y, q1.x, q1.y);
  const int Pq2 = orientation(p1.x, p1.y, p2.x, p2.y, q2.x, q2.y);
  if ((Pq1 > 0 && Pq2 > 0) || (Pq1 < 0 && Pq2 < 0)) return false;
  const int Qp1 = orientation(q1.x, q1.y, q2.x, q2.y, p1.x, p1.y);
  const int Qp2 = orientation(q1.x, q1.y, q2.x, q2.y, p2.x, p2.y);
  if ((Qp1 > 0 && Qp2 > 0) || (Qp1 < 0 && Qp2 < 0)) return false;
  if (Pq1 == 0 && Pq2 == 0 && Qp1 == 0 && Qp2 == 0) {
    // collinear: up to two intersection points
    const bool a = env_has_pt(p1, p2, q1), b = env_has_pt(p1, p2, q2);
    const bool c = env_has_pt(q1, q2, p1), d = env_has_pt(q1, q2, p2);
    P2 i0, i1;
    int n = 0;
    if (a && b) { i0 = q1; i1 = q2; n = 2; }
    else if (c && d) { i0 = p1; i1 = p2; n = 2; }
    else if (a && c) { i0 = q1; i1 = p1; n = (same(q1, p1) && !b && !d) ? 1 : 2; }
    else if (a && d) { i0 = q1; i1 = p2; n = (same(q1, p2) && !b && !c) ? 1 : 2; }
    else if (b && c) { i0 = q2; i1 = p1; n = (same(q2, p1) && !a && !d) ? 1 : 2; }
    else if (b && d) { i0 = q2; i1 = p2; n = (same(q2, p2) && !a && !c) ? 1 : 2; }
    if (n >= 1 && ip_interior(i0, p1, p2, q1, q2)) return true;
    if (n >= 2 && ip_interior(i1, p1, p2, q1, q2)) return true;
    return false;
  }
  if (Pq1 == 0 || Pq2 == 0 || Qp1 == 0 || Qp2 == 0) {
    P2 ip;
    if (same(p1, q1) || same(p1, q2)) ip = p1;
    else if (same(p2, q1) || same(p2, q2)) ip = p2;
    else if (Pq1 == 0) ip = q1;
    else if (Pq2 == 0) ip = q2;
    else if (Qp1 == 0) ip = p1;
    else ip = p2;
    return ip_interior(ip, p1, p2, q1, q2);
  }
  return true;  // proper crossing
}

// The simplifier's guard calls the predicate for a handful of segments per ring only (those whose
// envelope meets the chord's): out of line on the device, so that the guard's hot loop is the
// envelope test and nothing else (the inlined predicate carried four orientation bodies per call site).
#if defined(__CUDACC__)
__noinline__
#endif
TD_HD inline bool interior_intersection_cold(double p1x, double p1y, double p2x, double p2y, double q1x, double q1y,
                                             double q2x, double q2y) {
  return interior_intersection(P2{p1x, p1y}, P2{p2x, p2y}, P2{q1x, q1y}, P2{q2x, q2y});
}

// env_overlap(p1, p2, A, B) with the chord's envelope computed once: min(p1, p2) > max  <=>  both are,
// so eight comparisons decide it (same booleans as the fmin / fmax form for non-NaN input)
struct ChordEnv {
  double minx, maxx, miny, maxy;
};
TD_HD inline ChordEnv chord_envelope(const P2& A, const P2& B) {
  // one comparison per axis (coordinates are never NaN; a tie selects equal values either way)
  const bool xl = A.x < B.x, yl = A.y < B.y;
  return ChordEnv{xl ? A.x : B.x, xl ? B.x : A.x, yl ? A.y : B.y, yl ? B.y : A.y};
}
TD_HD inline bool env_overlap(const ChordEnv& c, const P2& p1, const P2& p2) {
#if defined(__CUDA_ARCH__)
  // spelled out as eight predicate-chained comparisons: the compiler otherwise folds each pair back into
  // a NaN-aware 64-bit min / max (five times the instructions)
  int r;
  asm("{\n\t.reg .pred a, b, c, d;\n\t"
      "setp.le.f64 a, %1, %6;\n\tsetp.le.or.f64 a, %3, %6, a;\n\t"
      "setp.ge.f64 b, %1, %5;\n\tsetp.ge.or.f64 b, %3, %5, b;\n\t"
      "setp.le.f64 c, %2, %8;\n\tsetp.le.or.f64 c, %4, %8, c;\n\t"
      "setp.ge.f64 d, %2, %7;\n\tsetp.ge.or.f64 d, %4, %7, d;\n\t"
      "and.pred a, a, b;\n\tand.pred c, c, d;\n\tand.pred a, a, c;\n\t"
      "selp.s32 %0, 1, 0, a;\n\t}"
      : "=r"(r)
      : "d"(p1.x), "d"(p1.y), "d"(p2.x), "d"(p2.y), "d"(c.minx), "d"(c.maxx), "d"(c.miny), "d"(c.maxy));
  return r != 0;
#else
  const bool x_lo = (p1.x <= c.maxx) | (p2.x <= c.maxx), x_hi = (p1.x >= c.minx) | (p2.x >= c.minx);
  const bool y_lo = (p1.y <= c.maxy) | (p2.y <= c.maxy), y_hi = (p1.y >= c.miny) | (p2.y >= c.miny);
  return x_lo & x_hi & y_lo & y_hi;
#endif
}

TD_HD inline double point_segment_distance(const P2& p, const P2& A, const P2& B) {
  if (A.x == B.x && A.y == B.y) {
    const double dx = p.x - A.x, dy = p.y - A.y;
    return sqrt(dx * dx + dy * dy);
  }
  const double len2 = (B.x - A.x) * (B.x - A.x) + (B.y - A.y) * (B.y - A.y);
  const double r = ((p.x - A.x) * (B.x - A.x) + (p.y - A.y) * (B.y - A.y)) / len2;
  if (r <= 0.0) {
    const double dx = p.x - A.x, dy = p.y - A.y;
    return sqrt(dx * dx + dy * dy);
  }
  if (r >= 1.0) {
    const double dx = p.x - B.x, dy = p.y - B.y;
    return sqrt(dx * dx + dy * dy);
  }
  const double s = ((A.y - p.y) * (B.x - A.x) - (A.x - p.x) * (B.y - A.y)) / len2;
  return fabs(s) * sqrt(len2);
}

// The same distance with everything that depends on the segment alone computed once (the farthest-
// point scan of a section evaluates many points against one chord): identical operations on
// identical values, so identical results.
struct SegPrep {
  P2 A, B;
  double ux, uy;       // B - A
  double len2, root;   // |B - A|^2 and its square root
  bool degenerate;
};
TD_HD inline SegPrep prepare_segment(const P2& A, const P2& B) {
  SegPrep s;
  s.A = A; s.B = B;
  s.ux = B.x - A.x; s.uy = B.y - A.y;
  s.degenerate = (A.x == B.x && A.y == B.y);
  s.len2 = s.ux * s.ux + s.uy * s.uy;
  s.root = sqrt(s.len2);
  return s;
}
TD_HD inline double point_segment_distance(const P2& p, const SegPrep& g) {
  // r = dot / len2 is only ever compared with 0 and 1:
  //   r <= 0  <=>  dot <= 0    (unless the quotient of a tiny positive dot underflows -> divide)
  //   r >= 1  <=>  dot >= len2 (a quotient of two doubles that is below 1 is at most 1 - 2^-53, which is
  //                             representable, so it never rounds up to 1)
  // so the scan of a section pays one division per point, not two.
  const double dot = (p.x - g.A.x) * g.ux + (p.y - g.A.y) * g.uy;
  bool before = dot <= 0.0, after = dot >= g.len2;
  if (!(g.len2 < 1e300) || (dot > 0.0 && dot < 1e-280)) {
    const double r = dot / g.len2;
    before = r <= 0.0;
    after = r >= 1.0;
  }
  if (g.degenerate) { before = true; after = false; }
  if (before || after) {   // distance to an end point (one square-root site for the three cases)
    const P2 e = before ? g.A : g.B;
    const double dx = p.x - e.x, dy = p.y - e.y;
    return sqrt(dx * dx + dy * dy);
  }
  const double s = ((g.A.y - p.y) * g.ux - (g.A.x - p.x) * g.uy) / g.len2;
  return fabs(s) * g.root;
}

// ---- cooperation policy ---------------------------------------------------------------------
// The simplifier is a sequential stack machine, but its two inner loops (farthest point of a
// section, interior-intersection scan over the segment sets) are data parallel.  `Coop` says
// how many lanes run the function together; every lane executes the control flow redundantly
// on identical values (identical writes to the shared scratch are benign), only the loops are
// strided over the lanes and combined with the two collectives below.
struct SerialCoop {
  TD_HD int lane() const { return 0; }
  TD_HD int size() const { return 1; }
  // all lanes receive the maximum d and, among equal maxima, the lowest k
  TD_HD void argmax_first(double&, int&) const {}
  TD_HD bool any(bool b) const { return b; }
  TD_HD void sync() const {}
};

#if defined(__CUDACC__)
struct WarpCoop {
  __device__ int lane() const { return threadIdx.x & 31; }
  __device__ int size() const { return 32; }
  // max d (d >= 0, or -1 for a lane without work), lowest k among equal maxima: non-negative doubles
  // order like their bit patterns, so three warp-wide integer reductions (REDUX) do it
  __device__ void argmax_first(double& d, int& k) const {
    const unsigned full = 0xffffffffu;
    const unsigned long long bits = d < 0.0 ? 0ull : (unsigned long long)__double_as_longlong(d);
    const unsigned hi = (unsigned)(bits >> 32);
    const unsigned mhi = __reduce_max_sync(full, hi);
    const unsigned lo = hi == mhi ? (unsigned)bits : 0u;
    const unsigned mlo = __reduce_max_sync(full, lo);
    const bool top = d >= 0.0 && hi == mhi && (unsigned)bits == mlo;
    const unsigned kk = __reduce_min_sync(full, top ? (unsigned)k : 0x7fffffffu);
    if (kk != 0x7fffffffu) {       // some lane had work
      d = __longlong_as_double((long long)(((unsigned long long)mhi << 32) | mlo));
      k = (int)kk;
    }
  }
  __device__ bool any(bool b) const { return __any_sync(0xffffffffu, b); }
  __device__ void sync() const { __syncwarp(); }   // orders the lanes' scratch writes
};
#endif

// ---- TopologyPreservingSimplifier on one closed ring ---------------------------
// pts[0..n) with pts[0] == pts[n-1].  scratch: 5 * n ints.  Writes the indices of the
// kept vertices (including the closing one) to res[0..m) and returns m.
//   scratch layout: res[n] | cend[n] | stack[3n]   (cend[k] = end vertex + 1 of the flattened section
//   that starts in vertex k, i.e. of a member of the output segment index, else 0)
//   alive: n bits in (n + 31) / 32 words (input segment k = pts[k], pts[k+1] still indexed)
//   IdxT: int in general; unsigned char when n <= 254 (every stored value is an index <= n)
//
// The guard ("the chord of a flattened section must not meet any other segment in an interior point")
// consults GEOS's two segment sets -- output chords, live input segments -- in ONE sweep over the start
// vertices: vertex k starts either a live input segment (k, k+1) or, when a section starting there was
// flattened, the chord (k, cend[k] - 1); flattening a section clears the live bits of its segments, so the
// two never coexist.  The sets are the same as in the two-loop form, and the answer is an OR over them.
template <typename Coop, typename IdxT>
TD_HD inline int simplify_ring(const P2* pts, int n, double tol, IdxT* scratch, uint32_t* alive, const Coop& co) {
  if (n <= 0) return 0;
  IdxT* res = scratch;
  IdxT* cend = scratch + n;
  IdxT* stack = scratch + 2 * n;
  const int nseg = n - 1;
  const int lane = co.lane(), nl = co.size();
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = lane; k < (n + 31) / 32; k += nl) alive[k] = 0xffffffffu;
#if defined(__CUDACC__)
#pragma unroll 1
#endif
  for (int k = lane; k < n; k += nl) cend[k] = 0;
  co.sync();
  int m = 0;        // number of result segments so far (res[k] = start vertex of result segment k)
  int sp = 0;
  stack[0] = 0; stack[1] = (IdxT)(n - 1); stack[2] = 0;
  sp = 1;
  const int min_size = 4;
  while (sp > 0) {
    --sp;
    const int i = stack[3 * sp], j = stack[3 * sp + 1];
    const int depth = stack[3 * sp + 2] + 1;
    if (i + 1 == j) {
      res[m] = (IdxT)i; ++m;
      continue;
    }
    bool valid = true;
    const int rsize = m == 0 ? 0 : m + 1;
    if (rsize < min_size && depth + 1 < min_size) valid = false;
    double maxd = -1.0;
    int far = i;
    const P2 A = pts[i], B = pts[j];
    const SegPrep chord = prepare_segment(A, B);
    for (int k = i + 1 + lane; k < j; k += nl) {
      const double d = point_segment_distance(pts[k], chord);
      if (d > maxd) { maxd = d; far = k; }
    }
    if (maxd < 0.0) far = 0x7fffffff;      // lane without work: loses every tie
    co.argmax_first(maxd, far);
    if (maxd > tol) valid = false;
    if (valid) {
      const ChordEnv env = chord_envelope(A, B);
      bool bad = false;
      for (int k = lane; k < nseg; k += nl) {
        const bool live = (alive[k >> 5] >> (k & 31)) & 1u;
        int v = k + 1;
        if (live) {
          if (k >= i && k < j) continue;       // the section's own segments
        } else {
          v = (int)cend[k] - 1;
          if (v < 0) continue;                 // vertex inside a flattened section
        }
        const P2 p = pts[k], q = pts[v];
        // A segment next to the section shares an end point with its chord (A or B).  Two segments that
        // share an end point and are not collinear meet in that point only, which is interior to
        // neither: interior_intersection is false (its "touching" branch picks the shared point).
        // One exact orientation decides that; only collinear neighbours take the full predicate.
        const bool ends_in_a = v == i;                 // input segment or output chord ending in A
        const bool starts_in_b = live && k == j;       // input segment starting in B
        if (ends_in_a || starts_in_b) {
          const P2 o = ends_in_a ? p : q;
          if (orientation_nonzero(A.x, A.y, B.x, B.y, o.x, o.y)) continue;
        }
        if (env_overlap(env, p, q) && interior_intersection_cold(p.x, p.y, q.x, q.y, A.x, A.y, B.x, B.y)) bad = true;
      }
      if (co.any(bad)) valid = false;
    }
    if (valid) {
      // clear bits [i, j): whole words strided over the lanes (identical result for 1 lane)
      for (int w = (i >> 5) + lane; w <= ((j - 1) >> 5); w += nl) {
        const int lo = w == (i >> 5) ? (i & 31) : 0;
        const int hi = w == ((j - 1) >> 5) ? ((j - 1) & 31) : 31;
        const uint32_t mask = (hi == 31 ? 0xffffffffu : ((2u << hi) - 1u)) & ~((1u << lo) - 1u);
        alive[w] &= ~mask;
      }
      co.sync();
      res[m] = (IdxT)i; cend[i] = (IdxT)(j + 1); ++m;
      continue;
    }
    // right section is processed second
    stack[3 * sp] = (IdxT)far; stack[3 * sp + 1] = (IdxT)j; stack[3 * sp + 2] = (IdxT)depth; ++sp;
    stack[3 * sp] = (IdxT)i; stack[3 * sp + 1] = (IdxT)far; stack[3 * sp + 2] = (IdxT)depth; ++sp;
  }
  res[m] = (IdxT)(n - 1);
  return m + 1;
}

TD_HD inline int simplify_ring(const P2* pts, int n, double tol, int* scratch, uint32_t* alive) {
  return simplify_ring(pts, n, tol, scratch, alive, SerialCoop());
}

// GEOS Area::ofRingSigned via an index accessor (so that it can run on the
// simplified index list without materialising the ring)
template <typename Get>
TD_HD inline double ring_signed_area(int n, Get get) {
  if (n < 3) return 0.0;
  const P2 first = get(0);
  const double x0 = first.x;
  double p1y = first.y;
  P2 q = get(1);
  double p2x = q.x - x0, p2y = q.y;
  double s = 0.0;
  for (int i = 1; i < n - 1; ++i) {
    const double p0y = p1y;
    const double p1x = p2x;
    p1y = p2y;
    q = get(i + 1);
    p2x = q.x - x0;
    p2y = q.y;
    s += p1x * (p0y - p2y);
  }
  return s / 2.0;
}

}  // namespace td
