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

constexpr int kMaxCand = 128;

__global__ void __launch_bounds__(64)
forest_predicates_kernel(const double* __restrict__ a_verts, const long long* __restrict__ a_off, int n_a,
                         const double* __restrict__ f_verts, const long long* __restrict__ f_off,
                         const double* __restrict__ f_bounds, int n_f, const double* __restrict__ a_filter,
                         unsigned char* __restrict__ out_intersects, unsigned char* __restrict__ out_within) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n_a) return;
  const td::P2* A = reinterpret_cast<const td::P2*>(a_verts) + a_off[r];
  const int na = (int)(a_off[r + 1] - a_off[r]);
  td::Box2 ab = {INFINITY, INFINITY, -INFINITY, -INFINITY};
  for (int k = 0; k < na; ++k) {
    ab.minx = fmin(ab.minx, A[k].x); ab.maxx = fmax(ab.maxx, A[k].x);
    ab.miny = fmin(ab.miny, A[k].y); ab.maxy = fmax(ab.maxy, A[k].y);
  }
  int cand[kMaxCand];
  int nc = 0;
  bool overflow = false;
  for (int k = 0; k < n_f; ++k) {
    const td::Box2 fb = {f_bounds[4 * k], f_bounds[4 * k + 1], f_bounds[4 * k + 2], f_bounds[4 * k + 3]};
    if (a_filter) {
      // tile flags: candidates by STRICT overlap with the un-buffered tile box (preprocessing.py:71-79)
      const double* q = a_filter + 4 * (size_t)r;
      if (!(fb.maxx > q[0] && fb.minx < q[2] && fb.maxy > q[1] && fb.miny < q[3])) continue;
    } else if (!td::boxes_overlap(ab, fb)) continue;
    if (nc >= kMaxCand) { overflow = true; break; }
    cand[nc++] = k;
  }
  const td::P2* F = reinterpret_cast<const td::P2*>(f_verts);
  bool hit = false;
  for (int c = 0; c < nc && !hit; ++c)
    hit = td::ring_intersects_ring(A, na, F + f_off[cand[c]], (int)(f_off[cand[c] + 1] - f_off[cand[c]]));
  out_intersects[r] = overflow ? 2 : (hit ? 1 : 0);
  int w = 0;
  if (!overflow && hit) w = td::ring_within_union(A, na, F, f_off, cand, nc);
  out_within[r] = overflow ? 2 : (unsigned char)w;
}

}  // namespace

// a_*: query rings (crowns / tile boxes); f_*: forest polygons (one closed ring each, no holes);
// f_bounds (n_f,4) f64 = bounds of every forest ring (td_simplify_rings with tolerance 0).
// a_filter (n_a,4) f64 or null: candidate forest polygons by strict bbox overlap with this box
// instead of the query ring's own bounds (the tile-flag rule of the reference).
// out_intersects / out_within (n_a) u8: 0 / 1, or 2 when a query overlaps more than 128 forest
// polygons or one of its edges crosses more than 62 forest edges (caller must treat as error).
extern "C" int td_forest_predicates(const double* a_verts, const long long* a_off, int n_a, const double* f_verts,
                                    const long long* f_off, const double* f_bounds, int n_f, const double* a_filter,
                                    unsigned char* out_intersects, unsigned char* out_within, void* stream) {
  TD_ARG(n_a >= 0 && n_f >= 0);
  if (n_a == 0) return TD_OK;
  TD_ARG(a_verts && a_off && out_intersects && out_within);
  TD_ARG(n_f == 0 || (f_verts && f_off && f_bounds));
  forest_predicates_kernel<<<td_div_up(n_a, 64), 64, 0, (cudaStream_t)stream>>>(
      a_verts, a_off, n_a, f_verts, f_off, f_bounds, n_f, a_filter, out_intersects, out_within);
  TD_CHECK_LAUNCH("td_forest_predicates");
  return TD_OK;
}
