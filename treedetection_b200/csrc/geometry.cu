// P4 -- stitch geometry: simplify(tol, preserve_topology=True) + the tile box filter;
// and the simplify(2) area of P9's head.
//
// Replaces process_prediction_file_sync (TreeDetection/helpers.py:419-476: shapely
// Polygon -> simplify(simplify_tolerance) -> geopandas sjoin "within" against the tile
// box shrunk by shift) and the area computation of process_geojson
// (TreeDetection/postprocessing.py:747-754: shape(geom).simplify(2).area).
// `within` a rectangle is envelope containment for an areal geometry (GEOS
// RectangleContains), so the filter is four comparisons on the simplified ring's bounds.
//
// One warp simplifies one ring: the algorithm is a sequential stack machine (run redundantly by
// all lanes) whose inner scans are split over the lanes; there are ~10^5 rings per image.  Results are index lists into the input ring, so a second tiny
// kernel (td_take_rings) gathers the surviving vertices once the caller has scanned the counts.
#include <cstdlib>

#include "chain_internal.cuh"
#include "common.cuh"
#include "simplify_core.cuh"

namespace {

constexpr int kSmemRing = 128;                              // vertices of a ring whose scratch fits shared memory
constexpr int kSmemInts = 5 * kSmemRing + kSmemRing / 32 + 4;   // per warp: res | cend | stack | alive bits
constexpr int kSmemWarpBytes = 16 * kSmemRing + 4 * kSmemInts;   // the ring's vertices in front of them

// Warp-wide min / max of doubles through the integer reductions (REDUX): a double's bits, with the sign
// bit flipped for positive values and all bits flipped for negative ones, order like unsigned integers.
// Two 32-bit reductions per value instead of a five-step shuffle butterfly of 64-bit fmin / fmax.
__device__ __forceinline__ unsigned long long ordered_key(double d) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(d);
  return b ^ ((b >> 63) ? 0xffffffffffffffffull : 0x8000000000000000ull);
}
__device__ __forceinline__ double from_ordered_key(unsigned long long k) {
  return __longlong_as_double((long long)(k ^ ((k >> 63) ? 0x8000000000000000ull : 0xffffffffffffffffull)));
}
__device__ __forceinline__ double warp_max(double d) {
  const unsigned long long k = ordered_key(d);
  const unsigned hi = __reduce_max_sync(0xffffffffu, (unsigned)(k >> 32));
  const unsigned lo = __reduce_max_sync(0xffffffffu, (unsigned)(k >> 32) == hi ? (unsigned)k : 0u);
  return from_ordered_key(((unsigned long long)hi << 32) | lo);
}
__device__ __forceinline__ double warp_min(double d) {
  const unsigned long long k = ordered_key(d);
  const unsigned hi = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32));
  const unsigned lo = __reduce_min_sync(0xffffffffu, (unsigned)(k >> 32) == hi ? (unsigned)k : 0xffffffffu);
  return from_ordered_key(((unsigned long long)hi << 32) | lo);
}

// bounds of pts[idx[k]] (idx == nullptr: pts[k]), k < cnt, on every lane (coordinates are never NaN)
__device__ __forceinline__ void warp_bounds(const td::P2* __restrict__ pts, const int* idx, int cnt, int lane,
                                            double& minx, double& miny, double& maxx, double& maxy) {
  minx = INFINITY; miny = INFINITY; maxx = -INFINITY; maxy = -INFINITY;
#pragma unroll 1
  for (int k = lane; k < cnt; k += 32) {
    const td::P2 p = pts[idx ? idx[k] : k];
    minx = p.x < minx ? p.x : minx; maxx = p.x > maxx ? p.x : maxx;
    miny = p.y < miny ? p.y : miny; maxy = p.y > maxy ? p.y : maxy;
  }
  minx = warp_min(minx); maxx = warp_max(maxx);
  miny = warp_min(miny); maxy = warp_max(maxy);
}

// Cold path, out of line (so that the kernel's hot path holds ONE copy of the simplifier -- the instruction
// cache was its top stall): rings of more than kSmemRing vertices (a handful per image) with their scratch in
// global memory, and the non-positive tolerance of helpers.py:463 (no simplification).  Returns the kept
// count; bounds of the kept vertices and the area on request.
struct ColdOut {
  double minx, miny, maxx, maxy, area;
};
__device__ __noinline__ int cold_ring(const td::P2* pts, int len, double tol, int* sc, uint32_t* al, int lane,
                                      bool want_bounds, bool want_area, ColdOut* o) {
  int m;
  if (tol > 0.0 && len > 0) {
    m = td::simplify_ring(pts, len, tol, sc, al, td::WarpCoop());
  } else {
#pragma unroll 1
    for (int k = lane; k < len; k += 32) sc[k] = k;
    m = len;
  }
  __syncwarp();
  if (want_bounds) warp_bounds(pts, sc, m, lane, o->minx, o->miny, o->maxx, o->maxy);
  if (want_area) o->area = fabs(td::ring_signed_area(m, [&](int k) { return pts[sc[k]]; }));
  return m;
}

// one warp per ring: the stack machine runs redundantly on all lanes, the farthest-point
// and intersection scans are strided over the lanes (td::WarpCoop).  Rings of up to kSmemRing vertices
// (nearly all) are staged in shared memory together with the simplifier's scratch (result list, chord
// ends, stack, live-segment bits): the stack machine touches them at every node, and in global memory
// each access is a 64-bit address computation plus an L1 / L2 round trip that gets long when a
// bandwidth-bound kernel runs next to this one.  The kept-vertex list is copied out at the end for
// td_take_rings.
template <int kMinBlocks>
__global__ void __launch_bounds__(64, kMinBlocks)
simplify_kernel(const double* __restrict__ verts, const long long* __restrict__ ring_off, int n, double tol,
                int* __restrict__ scratch, uint32_t* __restrict__ alive, const double* __restrict__ boxes,
                const int* __restrict__ ring_box, int* __restrict__ out_count, double* __restrict__ out_bounds,
                double* __restrict__ out_area, unsigned char* __restrict__ out_keep, int bounds_of_input,
                const long long* __restrict__ n_dev) {
  extern __shared__ __align__(16) unsigned char s_raw[];
  const int lane = threadIdx.x & 31;
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= n || (n_dev && r >= *n_dev)) return;
  const long long v0 = ring_off[r];
  const int len = (int)(ring_off[r + 1] - v0);
  const td::P2* pts = reinterpret_cast<const td::P2*>(verts) + v0;
  int* sc = scratch + 5 * v0;
  unsigned char* wbase = s_raw + (threadIdx.x >> 5) * kSmemWarpBytes;
  td::P2* spts = reinterpret_cast<td::P2*>(wbase);
  int* my = reinterpret_cast<int*>(wbase + sizeof(td::P2) * kSmemRing);
  const bool hot = len <= kSmemRing && len > 0 && tol > 0.0;
  // bounds of the input ring (needed for the pre-filter below and for bounds_of_input); the hot path
  // stages the ring in the same sweep
  double minx = INFINITY, miny = INFINITY, maxx = -INFINITY, maxy = -INFINITY;
  if (hot) {
#pragma unroll 1
    for (int k = lane; k < len; k += 32) {
      const td::P2 p = pts[k];
      spts[k] = p;
      minx = p.x < minx ? p.x : minx; maxx = p.x > maxx ? p.x : maxx;
      miny = p.y < miny ? p.y : miny; maxy = p.y > maxy ? p.y : maxy;
    }
    if (boxes || bounds_of_input) {
      minx = warp_min(minx); maxx = warp_max(maxx);
      miny = warp_min(miny); maxy = warp_max(maxy);
    }
    __syncwarp();
  } else if (boxes || bounds_of_input) {
    warp_bounds(pts, nullptr, len, lane, minx, miny, maxx, maxy);
  }
  // Pre-filter of the tile box test: every vertex the simplifier drops lies within `tol` of the chord
  // that replaces it, and a chord point outside the (convex) box means a kept end point outside it.
  // So a ring that sticks out of the box by more than tol cannot be `within` it after simplification
  // either: it is rejected without being simplified (a quarter of the rings of a tiled image).
  if (boxes && !out_bounds && !out_area && out_keep && len > 0) {
    const double* b = boxes + 4 * (size_t)ring_box[r];
    const double slack = (tol > 0.0 ? tol : 0.0) * (1.0 + 1e-9) + 1e-9;
    if (minx < b[0] - slack || miny < b[1] - slack || maxx > b[2] + slack || maxy > b[3] + slack) {
      if (lane == 0) { out_count[r] = 0; out_keep[r] = 0; }
      return;
    }
  }
  int m;
  double area = 0.0;
  if (hot) {
    m = td::simplify_ring(spts, len, tol, my, reinterpret_cast<uint32_t*>(my + 5 * kSmemRing), td::WarpCoop());
    __syncwarp();
#pragma unroll 1
    for (int k = lane; k < m; k += 32) sc[k] = my[k];
    if (!bounds_of_input) warp_bounds(spts, my, m, lane, minx, miny, maxx, maxy);   // kept vertices
    // the shoelace sum is order dependent: one lane, in ring order
    if (out_area && lane == 0) area = fabs(td::ring_signed_area(m, [&](int k) { return spts[my[k]]; }));
  } else {
    // alive words: ring r owns (len + 31) / 32 words starting at v0 / 32 + r  (disjoint)
    ColdOut o;
    m = cold_ring(pts, len, tol, sc, alive + (v0 >> 5) + r, lane, !bounds_of_input, out_area != nullptr && lane == 0, &o);
    if (!bounds_of_input) { minx = o.minx; miny = o.miny; maxx = o.maxx; maxy = o.maxy; }
    area = o.area;
  }
  if (lane != 0) return;
  out_count[r] = m;
  if (out_bounds) {
    out_bounds[4 * r + 0] = minx; out_bounds[4 * r + 1] = miny;
    out_bounds[4 * r + 2] = maxx; out_bounds[4 * r + 3] = maxy;
  }
  if (out_area) out_area[r] = area;
  if (out_keep) {
    bool keep = true;
    if (boxes) {
      const double* b = boxes + 4 * (size_t)ring_box[r];
      keep = (m > 0) && minx >= b[0] && miny >= b[1] && maxx <= b[2] && maxy <= b[3];
    }
    out_keep[r] = keep ? 1 : 0;
  }
}

// one warp per OUTPUT ring q: source ring sel[q]; with idx lists (scratch != null) only
// the kept vertices are copied, otherwise the whole ring
__global__ void take_rings_kernel(const double* __restrict__ verts, const long long* __restrict__ ring_off,
                                  const long long* __restrict__ sel, int n_out, const int* __restrict__ scratch,
                                  const long long* __restrict__ dst_off, double* __restrict__ out_verts,
                                  const long long* __restrict__ n_dev, int round3) {
  const int lane = threadIdx.x & 31;
  const int q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_out || (n_dev && q >= *n_dev)) return;
  const long long r = sel[q];
  const long long v0 = ring_off[r];
  const long long o = dst_off[q];
  const int cnt = (int)(dst_off[q + 1] - o);
  const int* sc = scratch ? scratch + 5 * v0 : nullptr;
  for (int k = lane; k < cnt; k += 32) {
    const long long s = v0 + (sc ? sc[k] : k);
    double x = verts[2 * s], y = verts[2 * s + 1];
    if (round3) {   // round_coordinates (utilities.py:146-161): round(v * 1000) / 1000, half to even
      x = __ddiv_rn(rint(__dmul_rn(x, 1000.0)), 1000.0);
      y = __ddiv_rn(rint(__dmul_rn(y, 1000.0)), 1000.0);
    }
    out_verts[2 * (o + k)] = x;
    out_verts[2 * (o + k) + 1] = y;
  }
}

}  // namespace

// scratch: 5 ints per input vertex; alive: (V / 32 + n + 1) uint32.
// n_dev (nullable, device): live ring count; rings r >= *n_dev are not touched (n_rings is then the capacity).
// boxes / ring_box / out_bounds / out_area / out_keep may be null.
extern "C" int td_simplify_rings(const double* verts, const long long* ring_off, int n_rings, double tolerance,
                                 int* scratch, uint32_t* alive, const double* boxes, const int* ring_box,
                                 int* out_count, double* out_bounds, double* out_area, unsigned char* out_keep,
                                 int bounds_of_input, const long long* n_dev, void* stream) {
  TD_ARG(n_rings >= 0);
  if (n_rings == 0) return TD_OK;
  TD_ARG(verts && ring_off && scratch && alive && out_count);
  TD_ARG((boxes == nullptr) == (ring_box == nullptr));
  static int bs = 0;
  if (bs == 0) { const char* e = getenv("TREEDET_SIMPLIFY_BLOCK"); bs = e && atoi(e) > 0 ? atoi(e) : 64; }
  // residency: 10 (96 registers), 12 (80, default: ~2 % faster in the chain) or 16 (64, spills: slower) CTAs of 2 warps per SM
  static int mb = 0;
  if (mb == 0) { const char* e = getenv("TREEDET_SIMPLIFY_MINBLOCKS"); mb = e && atoi(e) > 0 ? atoi(e) : 12; }
  if (bs > 64) bs = 64;
  const dim3 grid(td_div_up((long long)n_rings * 32, bs));
  const size_t smem = (size_t)(bs / 32) * kSmemWarpBytes;
  cudaStream_t st = (cudaStream_t)stream;
#define TD_SIMPLIFY_LAUNCH(MB)                                                                                          \
  simplify_kernel<MB><<<grid, bs, smem, st>>>(verts, ring_off, n_rings, tolerance, scratch, alive, boxes, ring_box,   \
                                              out_count, out_bounds, out_area, out_keep, bounds_of_input, n_dev)
  if (mb >= 16) TD_SIMPLIFY_LAUNCH(16);
  else if (mb >= 12) TD_SIMPLIFY_LAUNCH(12);
  else TD_SIMPLIFY_LAUNCH(10);
#undef TD_SIMPLIFY_LAUNCH
  TD_CHECK_LAUNCH("td_simplify_rings");
  return TD_OK;
}

// out ring q = ring sel[q] of the input; dst_off (n_out + 1) = offsets of the output rings
// (lengths = kept counts when `scratch` holds the index lists of td_simplify_rings, else
// the source ring lengths).
int td_take_rings_ex(const double* verts, const long long* ring_off, const long long* sel, int n_out,
                     const int* scratch, const long long* dst_off, double* out_verts, const long long* n_dev,
                     int round3, cudaStream_t st) {
  TD_ARG(n_out >= 0);
  if (n_out == 0) return TD_OK;
  TD_ARG(verts && ring_off && sel && dst_off && out_verts);
  take_rings_kernel<<<td_div_up((long long)n_out * 32, 256), 256, 0, st>>>(verts, ring_off, sel, n_out, scratch,
                                                                           dst_off, out_verts, n_dev, round3);
  TD_CHECK_LAUNCH("td_take_rings");
  return TD_OK;
}

extern "C" int td_take_rings(const double* verts, const long long* ring_off, const long long* sel, int n_out,
                             const int* scratch, const long long* dst_off, double* out_verts,
                             const long long* n_dev, void* stream) {
  return td_take_rings_ex(verts, ring_off, sel, n_out, scratch, dst_off, out_verts, n_dev, 0, (cudaStream_t)stream);
}
