// P10 -- forest-outline predicates for the two-model fusion (config 3) and the tile flags.
//
// Replaces the GEOS predicates of fuse_predictions (TreeDetection/helpers.py:795-811:
// forest crowns that intersect the forest union are kept, urban crowns within it are
// dropped) and of tile_single_file (TreeDetection/preprocessing.py:67-96: only_forest /
// only_urban).  One thread per query ring (crown or tile box); forest polygons are
// pre-filtered by bounding box, the union is never built (forest_core.cuh).
#include "common.cuh"
#include "forest_core.cuh"

namespace {

constexpr int kMaxCand = 128;   // candidate polygons kept in a list; more than that are re-found by scanning the bounds

__global__ void __launch_bounds__(64)
forest_predicates_kernel(const double* __restrict__ a_verts, const long long* __restrict__ a_off, int n_a,
                         const double* __restrict__ f_verts, const long long* __restrict__ f_off,
                         const long long* __restrict__ f_poly_off, const double* __restrict__ f_bounds, int n_poly,
                         const double* __restrict__ a_filter, unsigned char* __restrict__ out_intersects,
                         unsigned char* __restrict__ out_within) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_a) return;
  const td::P2* A = reinterpret_cast<const td::P2*>(a_verts) + a_off[r];
  const int na = (int)(a_off[r + 1] - a_off[r]);
  td::Box2 ab = {INFINITY, INFINITY, -INFINITY, -INFINITY};
  for (int k = 0; k < na; ++k) {
    ab.minx = fmin(ab.minx, A[k].x); ab.maxx = fmax(ab.maxx, A[k].x);
    ab.miny = fmin(ab.miny, A[k].y); ab.maxy = fmax(ab.maxy, A[k].y);
  }
  td::CandSet C;
  C.bounds = f_bounds; C.n_poly = n_poly; C.list = nullptr; C.n = 0;
  if (a_filter) {
    // tile flags: candidates by STRICT overlap with the un-buffered tile box (preprocessing.py:71-79)
    const double* q = a_filter + 4 * (size_t)r;
    C.q = td::Box2{q[0], q[1], q[2], q[3]};
    C.strict = true;
  } else {
    C.q = ab;
    C.strict = false;
  }
  int cand[kMaxCand];
  int nc = 0;
  bool listed = true;
  for (int k = C.scan(0); k >= 0; k = C.scan(k + 1)) {
    if (nc >= kMaxCand) { listed = false; break; }     // too many for the list: the set is re-scanned on every use
    cand[nc++] = k;
  }
  if (listed) { C.list = cand; C.n = nc; }
  td::ForestSet S;
  S.fverts = reinterpret_cast<const td::P2*>(f_verts); S.foff = f_off; S.poly_off = f_poly_off;
  bool hit = false;
  for (int c = C.first(); c >= 0 && !hit; c = C.next(c)) hit = td::ring_intersects_polygon(A, na, S, C.poly(c));
  out_intersects[r] = hit ? 1 : 0;
  out_within[r] = hit ? (unsigned char)td::ring_within_union(A, na, S, C) : 0;
}

__global__ void ring_simple_kernel(const double* __restrict__ verts, const long long* __restrict__ ring_off, int n,
                                   unsigned char* __restrict__ out) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  const td::P2* A = reinterpret_cast<const td::P2*>(verts) + ring_off[r];
  out[r] = td::ring_is_simple(A, (int)(ring_off[r + 1] - ring_off[r])) ? 1 : 0;
}

}  // namespace

// out (n) u8: 1 when ring r is a valid polygon shell (simple closed ring), 0 otherwise -- the geometries the
// reference would repair with buffer(0) / make_valid (TreeDetection/helpers.py:816-821).
extern "C" int td_ring_is_simple(const double* verts, const long long* ring_off, int n_rings, unsigned char* out,
                                 void* stream) {
  TD_ARG(n_rings >= 0);
  if (n_rings == 0) return TD_OK;
  TD_ARG(verts && ring_off && out);
  ring_simple_kernel<<<td_div_up(n_rings, 128), 128, 0, (cudaStream_t)stream>>>(verts, ring_off, n_rings, out);
  TD_CHECK_LAUNCH("td_ring_is_simple");
  return TD_OK;
}

// a_*: query rings (crowns / tile boxes); f_*: forest rings, f_poly_off (n_poly + 1, nullable) groups them into
// polygons (first ring = shell, the others = holes; null: every ring is a polygon without holes);
// f_bounds (n_poly,4) f64 = bounds of every polygon's shell.
// a_filter (n_a,4) f64 or null: candidate forest polygons by strict bbox overlap with this box
// instead of the query ring's own bounds (the tile-flag rule of the reference).
// out_intersects / out_within (n_a) u8: 0 / 1; out_within 2 when one edge of the query crosses more than 62
// forest edges (caller must treat as error).
extern "C" int td_forest_predicates(const double* a_verts, const long long* a_off, int n_a, const double* f_verts,
                                    const long long* f_off, const long long* f_poly_off, const double* f_bounds,
                                    int n_poly, const double* a_filter, unsigned char* out_intersects,
                                    unsigned char* out_within, void* stream) {
  TD_ARG(n_a >= 0 && n_poly >= 0);
  if (n_a == 0) return TD_OK;
  TD_ARG(a_verts && a_off && out_intersects && out_within);
  TD_ARG(n_poly == 0 || (f_verts && f_off && f_bounds));
  forest_predicates_kernel<<<td_div_up(n_a, 64), 64, 0, (cudaStream_t)stream>>>(
      a_verts, a_off, n_a, f_verts, f_off, f_poly_off, f_bounds, n_poly, a_filter, out_intersects, out_within);
  TD_CHECK_LAUNCH("td_forest_predicates");
  return TD_OK;
}
