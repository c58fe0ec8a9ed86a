// TEST INFRASTRUCTURE ONLY -- host simulation of the per-item sequential device
// algorithms.  The headers under treedetection_b200/csrc/*_core.cuh are plain C++ when
// compiled without nvcc; this file wraps them so that pytest can compare them with
// cv2 / the oracle in a container without a GPU.  The product (treedetection_b200)
// never loads this library; the GPU tests check the same functions compiled by nvcc.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "contour_core.cuh"
#include "contour_lockstep.cuh"
#if __has_include("simplify_core.cuh")
#include "simplify_core.cuh"
#define HS_HAVE_SIMPLIFY 1
#endif
#if __has_include("forest_core.cuh")
#include "forest_core.cuh"
#define HS_HAVE_FOREST 1
#endif

extern "C" {

// mask: (h, w) uint8 (non-zero = foreground).  Outputs (caller allocated, sized
// generously): contour point counts in OpenCV order, flattened (x, y) points,
// hierarchy parent (index in discovery order) for debugging.  Returns the number
// of contours, or -1 on overflow of the provided buffers.
int hs_find_contours(const uint8_t* mask, int h, int w, int* npts_out, int max_contours, short* pts_out,
                     int max_points, int* counts4) {
  const int wpr = (w + 31) / 32;
  std::vector<uint32_t> fg((size_t)wpr * h, 0u), vis((size_t)wpr * h, 0u), rgt((size_t)wpr * h, 0u);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (mask[(size_t)y * w + x]) fg[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
  td::Raster R;
  R.fg = fg.data(); R.visited = vis.data(); R.right = rgt.data(); R.label = nullptr;
  R.w = w; R.h = h; R.wpr = wpr;
  td::ContourCounts cc = td::scan_instance(R, nullptr);
  if (counts4) { counts4[0] = cc.n_contours; counts4[1] = cc.n_points; counts4[2] = cc.n_rings; counts4[3] = cc.n_ring_verts; }
  if (cc.n_contours < 0 || cc.n_contours > max_contours || cc.n_points > max_points) return -1;
  std::fill(vis.begin(), vis.end(), 0u);
  std::fill(rgt.begin(), rgt.end(), 0u);
  std::vector<unsigned short> lab((size_t)w * h, 0);
  R.label = lab.data();
  const int nc = cc.n_contours;
  std::vector<int> parent(nc + 1), npts(nc + 1), ptoff(nc + 1), lc(nc + 1), ps(nc + 1), order(nc + 1);
  std::vector<unsigned char> hole(nc + 1);
  std::vector<short> pts(2 * (size_t)cc.n_points + 2);
  td::ContourOut out;
  out.parent = parent.data(); out.npts = npts.data(); out.pt_off = ptoff.data(); out.is_hole = hole.data();
  out.pts = pts.data();
  td::ContourCounts c2 = td::scan_instance(R, &out);
  if (c2.n_contours != cc.n_contours || c2.n_points != cc.n_points) return -2;
  td::contour_order(nc, parent.data(), lc.data(), ps.data(), order.data());
  size_t k = 0;
  for (int q = 0; q < nc; ++q) {
    const int c = order[q];
    npts_out[q] = npts[c];
    std::memcpy(pts_out + 2 * k, pts.data() + 2 * (size_t)ptoff[c], sizeof(short) * 2 * npts[c]);
    k += npts[c];
  }
  return nc;
}

// same contract as hs_find_contours, through the lock-step state machine (one lane)
int hs_find_contours_lockstep(const uint8_t* mask, int h, int w, int* npts_out, int max_contours, short* pts_out,
                              int max_points, int* counts4, long long* steps_out) {
  const int wpr = (w + 31) / 32;
  std::vector<uint32_t> fg((size_t)wpr * h, 0u), vis((size_t)wpr * h, 0u), rgt((size_t)wpr * h, 0u);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (mask[(size_t)y * w + x]) fg[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
  td::LaneState<unsigned short> S;
  S.R.fg = fg.data(); S.R.visited = vis.data(); S.R.right = rgt.data(); S.R.label = nullptr;
  S.R.w = w; S.R.h = h; S.R.wpr = wpr;
  td::lane_init(S, nullptr);
  long long steps = 0;
  while (S.mode != td::kDone) { td::lane_step(S); ++steps; }
  td::ContourCounts cc = S.cc;
  if (counts4) { counts4[0] = cc.n_contours; counts4[1] = cc.n_points; counts4[2] = cc.n_rings; counts4[3] = cc.n_ring_verts; }
  if (steps_out) *steps_out = steps;
  if (cc.n_contours < 0 || cc.n_contours > max_contours || cc.n_points > max_points) return -1;
  std::fill(vis.begin(), vis.end(), 0u);
  std::fill(rgt.begin(), rgt.end(), 0u);
  std::vector<unsigned short> lab((size_t)w * h + 1, 0);
  S.R.label = lab.data();
  const int nc = cc.n_contours;
  std::vector<int> parent(nc + 1), npts(nc + 1), ptoff(nc + 1), lc(nc + 1), ps(nc + 1), order(nc + 1);
  std::vector<unsigned char> hole(nc + 1);
  std::vector<short> pts(2 * (size_t)cc.n_points + 2);
  td::ContourOut out;
  out.parent = parent.data(); out.npts = npts.data(); out.pt_off = ptoff.data(); out.is_hole = hole.data();
  out.pts = pts.data();
  td::lane_init(S, &out);
  while (S.mode != td::kDone) td::lane_step(S);
  if (S.cc.n_contours != cc.n_contours || S.cc.n_points != cc.n_points) return -2;
  td::contour_order(nc, parent.data(), lc.data(), ps.data(), order.data());
  size_t k = 0;
  for (int q = 0; q < nc; ++q) {
    const int c = order[q];
    npts_out[q] = npts[c];
    std::memcpy(pts_out + 2 * k, pts.data() + 2 * (size_t)ptoff[c], sizeof(short) * 2 * npts[c]);
    k += npts[c];
  }
  return nc;
}

// 8-neighbour masks of every pixel of a mask (contour_lockstep.cuh neighbours()): out[y * w + x], bit k = direction k
void hs_neighbour_masks(const uint8_t* mask, int h, int w, uint8_t* out) {
  const int wpr = (w + 31) / 32;
  std::vector<uint32_t> fg((size_t)wpr * h, 0u);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (mask[(size_t)y * w + x]) fg[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
  td::Raster R;
  R.fg = fg.data(); R.visited = nullptr; R.right = nullptr; R.label = nullptr;
  R.w = w; R.h = h; R.wpr = wpr;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) out[(size_t)y * w + x] = (uint8_t)td::neighbours(R, x, y);
}

// Single-pass walk into a SLOT (contours.cu trace_walk_kernel): tables of cap_contours rows and
// cap_points points, guarded by canaries.  Returns 0 when everything fitted, 1 when the instance
// outgrew its slot (the counts are still complete), -1 when a canary was overwritten.
int hs_walk_slot(const uint8_t* mask, int h, int w, int cap_contours, int cap_points, int* counts4, int* npts_out,
                 short* pts_out) {
  const int wpr = (w + 31) / 32;
  std::vector<uint32_t> fg((size_t)wpr * h, 0u), vis((size_t)wpr * h, 0u), rgt((size_t)wpr * h, 0u);
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x)
      if (mask[(size_t)y * w + x]) fg[(size_t)y * wpr + (x >> 5)] |= 1u << (x & 31);
  const int kCanary = 0x5a5a5a5a;
  std::vector<int> parent(cap_contours + 4, kCanary), npts(cap_contours + 4, kCanary), ptoff(cap_contours + 4, kCanary);
  std::vector<unsigned char> hole(cap_contours + 4, 0x5a);
  std::vector<short> pts(2 * (size_t)cap_points + 8, (short)0x5a5a);
  std::vector<unsigned short> lab((size_t)w * h + 1, 0);
  td::LaneState<unsigned short> S;
  S.R.fg = fg.data(); S.R.visited = vis.data(); S.R.right = rgt.data(); S.R.label = lab.data();
  S.R.w = w; S.R.h = h; S.R.wpr = wpr;
  td::ContourOut out;
  out.parent = parent.data(); out.npts = npts.data(); out.pt_off = ptoff.data(); out.is_hole = hole.data();
  out.pts = pts.data();
  out.cap_contours = cap_contours; out.cap_points = cap_points;
  td::lane_init(S, &out);
  while (S.mode != td::kDone) td::lane_step(S);
  counts4[0] = S.cc.n_contours; counts4[1] = S.cc.n_points; counts4[2] = S.cc.n_rings; counts4[3] = S.cc.n_ring_verts;
  for (int k = cap_contours; k < cap_contours + 4; ++k)
    if (parent[k] != kCanary || npts[k] != kCanary || ptoff[k] != kCanary || hole[k] != 0x5a) return -1;
  for (size_t k = 2 * (size_t)cap_points; k < pts.size(); ++k)
    if (pts[k] != (short)0x5a5a) return -1;
  const bool over = S.cc.n_contours < 0 || S.cc.n_contours > cap_contours || S.cc.n_points > cap_points;
  if (over) return 1;
  const int nc = S.cc.n_contours;
  std::vector<int> lc(nc + 1), ps(nc + 1), order(nc + 1);
  td::contour_order(nc, parent.data(), lc.data(), ps.data(), order.data());
  size_t k = 0;
  for (int q = 0; q < nc; ++q) {
    const int c = order[q];
    npts_out[q] = npts[c];
    std::memcpy(pts_out + 2 * k, pts.data() + 2 * (size_t)ptoff[c], sizeof(short) * 2 * npts[c]);
    k += npts[c];
  }
  return 0;
}

#ifdef HS_HAVE_SIMPLIFY
// ring: n points (x, y interleaved), closed.  Returns the number of kept vertices and
// writes their indices; *area = |signed area| of the simplified ring.
int hs_simplify_ring(const double* xy, int n, double tol, int* keep_idx, double* area) {
  std::vector<int> scratch(5 * (size_t)n + 8);
  std::vector<uint32_t> alive((n + 31) / 32 + 1);
  const td::P2* pts = reinterpret_cast<const td::P2*>(xy);
  const int m = td::simplify_ring(pts, n, tol, scratch.data(), alive.data());
  for (int k = 0; k < m; ++k) keep_idx[k] = scratch[k];
  if (area) *area = std::fabs(td::ring_signed_area(m, [&](int k) { return pts[scratch[k]]; }));
  return m;
}

int hs_orientation(double ax, double ay, double bx, double by, double cx, double cy) {
  return td::orientation(ax, ay, bx, by, cx, cy);
}
int hs_orientation_nonzero(double ax, double ay, double bx, double by, double cx, double cy) {
  return td::orientation_nonzero(ax, ay, bx, by, cx, cy) ? 1 : 0;
}
int hs_orientation_exact(double ax, double ay, double bx, double by, double cx, double cy) {
  return td::orientation_exact(ax, ay, bx, by, cx, cy);
}
int hs_interior_intersection(const double* p) {
  td::P2 a{p[0], p[1]}, b{p[2], p[3]}, c{p[4], p[5]}, d{p[6], p[7]};
  return td::interior_intersection(a, b, c, d) ? 1 : 0;
}
#endif

#ifdef HS_HAVE_FOREST
// query rings vs forest polygons (rings grouped by poly_off: first ring = shell, others = holes; poly_off may be
// null: every ring is a polygon); candidates = every polygon whose shell bounds overlap, as the kernel does.
void hs_forest_predicates(const double* a_xy, const long long* a_off, int n_a, const double* f_xy,
                          const long long* f_off, const long long* poly_off, int n_poly, unsigned char* inter,
                          unsigned char* within) {
  const td::P2* AV = reinterpret_cast<const td::P2*>(a_xy);
  td::ForestSet S;
  S.fverts = reinterpret_cast<const td::P2*>(f_xy); S.foff = f_off; S.poly_off = poly_off;
  std::vector<double> fb(4 * (size_t)n_poly);
  for (int k = 0; k < n_poly; ++k) {
    td::Box2 b = {1e300, 1e300, -1e300, -1e300};
    const long long r = S.ring0(k);
    for (int v = 0; v < S.ring_len(r); ++v) {
      const td::P2 q = S.ring(r)[v];
      b.minx = std::fmin(b.minx, q.x); b.maxx = std::fmax(b.maxx, q.x);
      b.miny = std::fmin(b.miny, q.y); b.maxy = std::fmax(b.maxy, q.y);
    }
    fb[4 * k] = b.minx; fb[4 * k + 1] = b.miny; fb[4 * k + 2] = b.maxx; fb[4 * k + 3] = b.maxy;
  }
  for (int r = 0; r < n_a; ++r) {
    const td::P2* A = AV + a_off[r];
    const int na = (int)(a_off[r + 1] - a_off[r]);
    td::Box2 ab = {1e300, 1e300, -1e300, -1e300};
    for (int k = 0; k < na; ++k) {
      ab.minx = std::fmin(ab.minx, A[k].x); ab.maxx = std::fmax(ab.maxx, A[k].x);
      ab.miny = std::fmin(ab.miny, A[k].y); ab.maxy = std::fmax(ab.maxy, A[k].y);
    }
    td::CandSet C;
    C.list = nullptr; C.n = 0; C.bounds = fb.data(); C.q = ab; C.n_poly = n_poly; C.strict = false;
    bool hit = false;
    for (int c = C.first(); c >= 0 && !hit; c = C.next(c)) hit = td::ring_intersects_polygon(A, na, S, C.poly(c));
    inter[r] = hit ? 1 : 0;
    within[r] = hit ? (unsigned char)td::ring_within_union(A, na, S, C) : 0;
  }
}
#endif

}  // extern "C"
