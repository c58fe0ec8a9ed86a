// P6 -- ordered bbox NMS, and P8 -- bbox containment.
//
// P6 replaces filter_polygons_by_iou_and_area (TreeDetection/postprocessing.py:349-406)
// + calculate_iou (TreeDetection/utilities.py:112-144): the reference builds N x N
// float32 IoU and float16 area matrices and then walks the crowns in INDEX order:
//     for i in 0..N-1: if removed[i]: continue
//         group = where(mask[i]) ++ [i]; best = group[argmax(conf16[group])]
//         removed[group \ {best}] = True
// P8 replaces process_containment_features (postprocessing.py:408-476).
//
// Here no N^2 object exists.  Boxes are binned by their min corner into square
// cells twice as wide as the widest box (so every overlapping pair lies in
// adjacent cells), sorted by cell key, and every crown enumerates the 3x3
// neighbourhood.  The order-dependent loop is reproduced exactly by noting that
// best(j) is static and that crown i "fires" iff no lower-indexed neighbour j with
// best(j) != i fires; that recurrence is resolved by a cooperative kernel that
// iterates to the fixed point (chain length rounds, a handful in practice), after
// which removed[k] = OR_{j in group(k)} fires[j] && best(j) != k.
//
// dtypes are part of the contract (postprocessing.py:367-369):
// boxes float32, confidence and area float16, thresholds compared in the array dtype.
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include "chain_internal.cuh"
#include "common.cuh"

namespace cg = cooperative_groups;

namespace {

typedef unsigned int KeyT;   // sort key of a grid cell

struct GridParams {
  int minx_enc, miny_enc;  // order-preserving int encodings (atomicMin/Max)
  int maxw_enc, maxh_enc;
};

TD_D int enc_f(float f) {
  int i = __float_as_int(f);
  return i >= 0 ? i : i ^ 0x7fffffff;
}
TD_D float dec_f(int i) { return __int_as_float(i >= 0 ? i : i ^ 0x7fffffff); }

__global__ void grid_init_kernel(GridParams* gp) {
  gp->minx_enc = 0x7fffffff;
  gp->miny_enc = 0x7fffffff;
  gp->maxw_enc = enc_f(0.f);
  gp->maxh_enc = enc_f(0.f);
}

// bounds f64 (N,4) -> f32 boxes (the reference casts every coordinate with
// np.float32, postprocessing.py:366) + extent reduction
__global__ void prep_boxes_kernel(const double* __restrict__ bounds, int n, float4* __restrict__ box32,
                                  GridParams* gp, const long long* __restrict__ n_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = (int)*n_dev < n ? (int)*n_dev : n;
  float x0 = 0, y0 = 0, w = 0, h = 0;
  bool ok = false;
  if (i < n) {
    float4 b;
    b.x = __double2float_rn(bounds[4 * i + 0]);
    b.y = __double2float_rn(bounds[4 * i + 1]);
    b.z = __double2float_rn(bounds[4 * i + 2]);
    b.w = __double2float_rn(bounds[4 * i + 3]);
    box32[i] = b;
    x0 = b.x; y0 = b.y; w = b.z - b.x; h = b.w - b.y;
    ok = isfinite(x0) && isfinite(y0) && isfinite(w) && isfinite(h);
  }
  int ex = ok ? enc_f(x0) : 0x7fffffff, ey = ok ? enc_f(y0) : 0x7fffffff;
  int ew = ok ? enc_f(fmaxf(w, 0.f)) : enc_f(0.f), eh = ok ? enc_f(fmaxf(h, 0.f)) : enc_f(0.f);
  for (int o = 16; o > 0; o >>= 1) {
    ex = min(ex, __shfl_xor_sync(0xffffffffu, ex, o));
    ey = min(ey, __shfl_xor_sync(0xffffffffu, ey, o));
    ew = max(ew, __shfl_xor_sync(0xffffffffu, ew, o));
    eh = max(eh, __shfl_xor_sync(0xffffffffu, eh, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&gp->minx_enc, ex);
    atomicMin(&gp->miny_enc, ey);
    atomicMax(&gp->maxw_enc, ew);
    atomicMax(&gp->maxh_enc, eh);
  }
}

struct Cell {
  int cx, cy;
};

TD_D float cell_size(const GridParams* gp) {
  float c = 2.f * fmaxf(dec_f(gp->maxw_enc), dec_f(gp->maxh_enc));
  return (c > 0.f && isfinite(c)) ? c : 1.f;
}

TD_D Cell cell_of(float x0, float y0, const GridParams* gp, float c) {
  Cell r;
  float fx = floorf((x0 - dec_f(gp->minx_enc)) / c);
  float fy = floorf((y0 - dec_f(gp->miny_enc)) / c);
  // non-finite boxes never overlap anything: park them in a far cell
  if (!isfinite(fx) || !isfinite(fy)) { fx = 2.0e9f; fy = 2.0e9f; }
  r.cx = (int)fminf(fmaxf(fx, 0.f), 2.0e9f);
  r.cy = (int)fminf(fmaxf(fy, 0.f), 2.0e9f);
  return r;
}

// 16 bits per axis: cells beyond 65535 alias into the last one, which only adds candidates (every
// pair is tested exactly); a 32-bit key halves the radix-sort passes
TD_D unsigned int cell_key(int cx, int cy) {
  return ((unsigned)min(cy, 65535) << 16) | (unsigned)min(cx, 65535);
}

__global__ void cell_keys_kernel(const float4* __restrict__ box32, int n, const GridParams* __restrict__ gp,
                                 KeyT* __restrict__ keys, int* __restrict__ idx,
                                 const long long* __restrict__ n_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n_dev && i >= *n_dev) { keys[i] = ~(KeyT)0; idx[i] = i; return; }   // capacity tail sorts to the end
  const float c = cell_size(gp);
  const Cell ce = cell_of(box32[i].x, box32[i].y, gp, c);
  keys[i] = cell_key(ce.cx, ce.cy);
  idx[i] = i;
}

TD_D int lower_bound_key(const KeyT* __restrict__ keys, int n, KeyT k) {
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) >> 1;
    if (keys[mid] < k) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct PairGrid {
  const float4* box32;
  const KeyT* keys;  // sorted
  const int* idx;                  // sorted order -> crown index
  const GridParams* gp;
  int n;
  int all_pairs;  // threshold <= 0: zero-overlap pairs matter, enumerate everything
  const long long* n_dev;   // nullable: live count on the device (n is then the capacity)
};

TD_D void live_count(PairGrid& g) {
  if (g.n_dev) g.n = (int)*g.n_dev < g.n ? (int)*g.n_dev : g.n;
}

// calls f(j) for every candidate partner j of crown i (j == i included)
template <typename F>
TD_D void for_each_candidate(const PairGrid& g, int i, F f) {
  if (g.all_pairs) {
    for (int j = 0; j < g.n; ++j) f(j);
    return;
  }
  const float c = cell_size(g.gp);
  const Cell ce = cell_of(g.box32[i].x, g.box32[i].y, g.gp, c);
  for (int dy = -1; dy <= 1; ++dy) {
    const long long cy = (long long)ce.cy + dy;
    if (cy < 0) continue;
    const int cx_lo = max(ce.cx - 1, 0);
    const long long cx_hi = (long long)ce.cx + 1;
    const int s = lower_bound_key(g.keys, g.n, cell_key(cx_lo, (int)cy));
    const KeyT kend = cell_key((int)cx_hi, (int)cy);
    for (int p = s; p < g.n && g.keys[p] <= kend; ++p) f(g.idx[p]);
  }
}

// ---- P6 pair predicate ------------------------------------------------------
TD_D float box_area(const float4& b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

TD_D float box_inter(const float4& a, const float4& b) {
  const float xA = fmaxf(a.x, b.x), yA = fmaxf(a.y, b.y);
  const float xB = fminf(a.z, b.z), yB = fminf(a.w, b.w);
  return __fmul_rn(fmaxf(0.f, __fsub_rn(xB, xA)), fmaxf(0.f, __fsub_rn(yB, yA)));
}

TD_D bool nms_connected(const float4& bi, const float4& bj, __half ai, __half aj, float iou_thr, __half area_thr) {
  const float inter = box_inter(bi, bj);
  const float uni = __fsub_rn(__fadd_rn(box_area(bi), box_area(bj)), inter);
  const float iou = __fdiv_rn(inter, uni);
  if (!(iou > iou_thr)) return false;
  // float16 arithmetic, each operation rounded to half (numpy/cupy semantics)
  const float fa = __half2float(ai), fb = __half2float(aj);
  const __half diff = __float2half_rn(fabsf(__half2float(__float2half_rn(fa - fb))));
  const __half mx = __float2half_rn(fmaxf(fa, fb));
  // np.maximum propagates NaN
  const float fmx = (isnan(fa) || isnan(fb)) ? nanf("") : __half2float(mx);
  const __half rel = __float2half_rn(__fdiv_rn(__half2float(diff), fmx));
  return __half2float(rel) < __half2float(area_thr);
}

__global__ void to_half_kernel(const double* __restrict__ conf, const double* __restrict__ area, int n,
                               __half* __restrict__ c16, __half* __restrict__ a16,
                               const long long* __restrict__ n_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (n_dev && i >= *n_dev)) return;
  c16[i] = __double2half(conf[i]);
  a16[i] = __double2half(area[i]);
}

// pass 0: count neighbours (excluding self); pass 1: fill CSR, compute best()
template <bool kFill>
__global__ void nms_adjacency_kernel(PairGrid g, const __half* __restrict__ c16, const __half* __restrict__ a16,
                                     float iou_thr, __half area_thr, long long* __restrict__ deg_or_off,
                                     int* __restrict__ nbr, int* __restrict__ best, int slots,
                                     long long* __restrict__ end, long long* __restrict__ flag) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  live_count(g);
  if (i >= g.n) return;
  // slots > 0 (capacity form): no count pass, crown i owns nbr[i * slots .. (i + 1) * slots); a list
  // that does not fit raises bit 1 of *flag (the resolve kernel then does nothing and the caller
  // discards the result)
  const float4 bi = g.box32[i];
  const __half ai = a16[i];
  long long cnt = 0;
  const long long off = !kFill ? 0 : (slots > 0 ? (long long)i * slots : deg_or_off[i]);
  // argmax(conf16[group]) with group = sorted(where(mask[i])) ++ [i]: first maximum wins,
  // i.e. the smallest index among the maxima; the appended i only wins ties when it is
  // also a regular member (mask[i][i]), otherwise it is last.  NaN counts as the maximum.
  float bestc = 0.f;
  int besti = -1;
  bool bestnan = false;
  bool self_in = false;
  for_each_candidate(g, i, [&](int j) {
    if (!nms_connected(bi, g.box32[j], ai, a16[j], iou_thr, area_thr)) return;
    if (j == i) { self_in = true; }
    else {
      if (kFill && (slots <= 0 || cnt < slots)) nbr[off + cnt] = j;
      ++cnt;
    }
    if (kFill) {
      const float cj = __half2float(c16[j]);
      const bool jn = isnan(cj);
      bool better;
      if (besti < 0) better = true;
      else if (bestnan) better = jn && j < besti;
      else if (jn) better = true;
      else better = (cj > bestc) || (cj == bestc && j < besti);
      if (better) { bestc = cj; besti = j; bestnan = jn; }
    }
  });
  if (!kFill) { deg_or_off[i] = cnt; return; }
  if (slots > 0) {
    if (cnt > slots) atomicOr((unsigned long long*)flag, 2ull);
    deg_or_off[i] = off;
    end[i] = off + (cnt < slots ? cnt : slots);
  }
  if (!self_in) {  // i appended last: wins only with a strictly larger confidence
    const float ci = __half2float(c16[i]);
    const bool in_ = isnan(ci);
    bool better;
    if (besti < 0) better = true;
    else if (bestnan) better = false;
    else if (in_) better = true;
    else better = ci > bestc;
    if (better) besti = i;
  }
  best[i] = besti;
}

// state: 0 undecided, 1 fires, 2 skipped (was already removed when visited)
__global__ void nms_resolve_kernel(int n, const long long* __restrict__ off, const long long* __restrict__ end,
                                   const int* __restrict__ nbr,
                                   const int* __restrict__ best, volatile int* state, int* pending,
                                   unsigned char* __restrict__ removed, const long long* __restrict__ n_dev,
                                   const long long* __restrict__ flag) {
  cg::grid_group grid = cg::this_grid();
  if (flag && *((const volatile long long*)flag) != 0) return;   // uniform: an earlier stage overflowed
  if (n_dev) n = (int)*n_dev < n ? (int)*n_dev : n;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nth = gridDim.x * blockDim.x;
  for (int i = tid; i < n; i += nth) state[i] = 0;
  if (tid == 0) { pending[0] = 0; pending[1] = 0; }
  grid.sync();
  for (int round = 0;; ++round) {
    int* cur = pending + (round & 1);
    int* nxt = pending + ((round + 1) & 1);
    int undecided = 0;
    for (int i = tid; i < n; i += nth) {
      if (state[i] != 0) continue;
      bool killed = false, wait = false;
      for (long long p = off[i]; p < end[i]; ++p) {
        const int j = nbr[p];
        if (j >= i || best[j] == i) continue;
        const int s = state[j];
        if (s == 1) { killed = true; break; }
        if (s == 0) wait = true;
      }
      if (killed) state[i] = 2;
      else if (!wait) state[i] = 1;
      else ++undecided;
    }
    if (undecided) atomicAdd(cur, undecided);
    if (tid == 0) *nxt = 0;
    grid.sync();
    const int left = *((volatile int*)cur);
    grid.sync();
    if (left == 0) break;
  }
  for (int k = tid; k < n; k += nth) {
    bool rem = (state[k] == 1 && best[k] != k);
    for (long long p = off[k]; p < end[k] && !rem; ++p) {
      const int j = nbr[p];
      rem = (state[j] == 1 && best[j] != k);
    }
    removed[k] = rem ? 1 : 0;
  }
}

// ---- P8 ---------------------------------------------------------------------
__global__ void box32_from_f32_kernel(const float* __restrict__ b, int n, float4* __restrict__ box32, GridParams* gp,
                                      const long long* __restrict__ n_dev) {
  // containment takes bounds that the reference already holds as float32
  // (cp.array(polygon_bounds, dtype=float32), postprocessing.py:621)
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = (int)*n_dev < n ? (int)*n_dev : n;
  float x0 = 0, y0 = 0, w = 0, h = 0;
  bool ok = false;
  if (i < n) {
    float4 v = make_float4(b[4 * i], b[4 * i + 1], b[4 * i + 2], b[4 * i + 3]);
    box32[i] = v;
    x0 = v.x; y0 = v.y; w = v.z - v.x; h = v.w - v.y;
    ok = isfinite(x0) && isfinite(y0) && isfinite(w) && isfinite(h);
  }
  int ex = ok ? enc_f(x0) : 0x7fffffff, ey = ok ? enc_f(y0) : 0x7fffffff;
  int ew = ok ? enc_f(fmaxf(w, 0.f)) : enc_f(0.f), eh = ok ? enc_f(fmaxf(h, 0.f)) : enc_f(0.f);
  for (int o = 16; o > 0; o >>= 1) {
    ex = min(ex, __shfl_xor_sync(0xffffffffu, ex, o));
    ey = min(ey, __shfl_xor_sync(0xffffffffu, ey, o));
    ew = max(ew, __shfl_xor_sync(0xffffffffu, ew, o));
    eh = max(eh, __shfl_xor_sync(0xffffffffu, eh, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&gp->minx_enc, ex);
    atomicMin(&gp->miny_enc, ey);
    atomicMax(&gp->maxw_enc, ew);
    atomicMax(&gp->maxh_enc, eh);
  }
}

__global__ void containment_kernel(PairGrid g, float thr, float* __restrict__ ratio_max,
                                   unsigned char* __restrict__ is_contained, int* __restrict__ num_contained) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  live_count(g);
  if (t >= g.n) return;
  const float4 bt = g.box32[t];
  const float at = box_area(bt);
  float rmax = -INFINITY;
  bool rnan = false, contained = false;
  int num = 0;
  for_each_candidate(g, t, [&](int j) {
    const float4 bj = g.box32[j];
    const float inter = box_inter(bt, bj);
    // t as the inner box: ratio[j, t]
    const float r_in = __fdiv_rn(inter, at);
    if (isnan(r_in)) rnan = true; else rmax = fmaxf(rmax, r_in);
    if (j != t) {
      if (r_in >= thr) contained = true;
      // t as the outer box: ratio[t, j]
      const float r_out = __fdiv_rn(inter, box_area(bj));
      if (r_out >= thr) ++num;
    }
  });
  if (!g.all_pairs && !rnan) {
    // pairs outside the neighbourhood have intersection 0: ratio 0 (or NaN for a
    // degenerate inner box, which the self pair has already reported)
    if (g.n > 1) rmax = fmaxf(rmax, 0.f);
  }
  ratio_max[t] = rnan ? nanf("") : rmax;
  is_contained[t] = contained ? 1 : 0;
  num_contained[t] = num;
}

// ---- opt-in mask-IoU de-duplication over the packed crown rasters of P2 (iou_mode: mask) -----------------
// The reference ships a polygon-IoU cleaner it never calls (clean_crowns, TreeDetection/helpers.py:602-701,
// detectree2's): for every crown, among the crowns whose IoU with it exceeds the threshold (itself included)
// the one with the highest confidence is taken, and the crown survives only if that one coincides with it
// (IoU == 1).  Every crown is decided on its own -- no order dependence.  Here the IoU is the PIXEL IoU of
// the 1-bit rasters P2 already left on the device: popcount(a & b) / popcount(a | b) on the image's pixel
// grid (tile windows are pixel aligned, so rasters of different tiles compare exactly), candidates from the
// same uniform grid as the bbox NMS.  One warp per instance: lanes take rows of the overlap rectangle, a
// row's bits are fetched 32 at a time at an arbitrary bit offset (funnel shift of two words), AND + POPC,
// warp-reduced with REDUX.
struct MaskSet {
  const uint32_t* bits;
  const long long* word_off;
  const int4* gbox;   // per instance: global x0, y0, width, height (pixels of the image)
};

// 32 bits of instance k's row `row` (window coordinates) starting at window column x (any x, zero outside)
TD_D uint32_t mask_chunk(const MaskSet& M, int k, int row, int x) {
  const int4 b = M.gbox[k];
  if (row < 0 || row >= b.w || x >= b.z || x <= -32) return 0u;
  const int wpr = (b.z + 31) >> 5;
  const uint32_t* r = M.bits + M.word_off[k] + (size_t)row * wpr;
  const int wi = x >> 5;               // floor division also for negative x
  const int sh = x & 31;
  const uint32_t lo = (wi >= 0 && wi < wpr) ? r[wi] : 0u;
  const uint32_t hi = (wi + 1 >= 0 && wi + 1 < wpr) ? r[wi + 1] : 0u;
  return __funnelshift_r(lo, hi, sh);
}

__global__ void mask_gbox_kernel(const int* __restrict__ win, const int* __restrict__ tile_org,
                                 const int* __restrict__ inst_tile, int n, int4* __restrict__ gbox,
                                 float4* __restrict__ box32, GridParams* gp) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  float x0 = 0, y0 = 0, w = 0, h = 0;
  bool ok = false;
  if (i < n) {
    const int t = inst_tile[i];
    const int4 b = make_int4(tile_org[2 * t] + win[4 * i], tile_org[2 * t + 1] + win[4 * i + 1], win[4 * i + 2],
                             win[4 * i + 3]);
    gbox[i] = b;
    // pixel boxes as float boxes [x0, x0 + w) for the grid (exact below 2^24)
    box32[i] = make_float4((float)b.x, (float)b.y, (float)(b.x + b.z), (float)(b.y + b.w));
    x0 = (float)b.x; y0 = (float)b.y; w = (float)b.z; h = (float)b.w;
    ok = b.z > 0 && b.w > 0;
    if (!ok) box32[i] = make_float4(nanf(""), nanf(""), nanf(""), nanf(""));   // parked in the far cell
  }
  int ex = ok ? enc_f(x0) : 0x7fffffff, ey = ok ? enc_f(y0) : 0x7fffffff;
  int ew = ok ? enc_f(w) : enc_f(0.f), eh = ok ? enc_f(h) : enc_f(0.f);
  for (int o = 16; o > 0; o >>= 1) {
    ex = min(ex, __shfl_xor_sync(0xffffffffu, ex, o));
    ey = min(ey, __shfl_xor_sync(0xffffffffu, ey, o));
    ew = max(ew, __shfl_xor_sync(0xffffffffu, ew, o));
    eh = max(eh, __shfl_xor_sync(0xffffffffu, eh, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(&gp->minx_enc, ex);
    atomicMin(&gp->miny_enc, ey);
    atomicMax(&gp->maxw_enc, ew);
    atomicMax(&gp->maxh_enc, eh);
  }
}

__global__ void mask_area_kernel(MaskSet M, int n, int* __restrict__ area) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int4 b = M.gbox[i];
  const long long nw = (long long)((b.z + 31) >> 5) * b.w;
  const uint32_t* p = M.bits + M.word_off[i];
  int a = 0;
  for (long long k = lane; k < nw; k += 32) a += __popc(p[k]);
  a = __reduce_add_sync(0xffffffffu, a);
  if (lane == 0) area[i] = a;
}

__global__ void __launch_bounds__(128)
mask_iou_kernel(PairGrid g, MaskSet M, const int* __restrict__ area, const float* __restrict__ scores, float iou_thr,
                float confidence, unsigned char* __restrict__ keep, int* __restrict__ match,
                float* __restrict__ best_iou) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= g.n) return;
  const int4 bi = M.gbox[i];
  const int ai = area[i];
  int best = -1, best_inter = 0, best_uni = 1;
  float best_conf = 0.f;
  if (bi.z > 0 && bi.w > 0 && ai > 0) {
    for_each_candidate(g, i, [&](int j) {          // every lane walks the same candidates
      const int4 bj = M.gbox[j];
      const int ox0 = max(bi.x, bj.x), oy0 = max(bi.y, bj.y);
      const int ox1 = min(bi.x + bi.z, bj.x + bj.z), oy1 = min(bi.y + bi.w, bj.y + bj.w);
      if (ox0 >= ox1 || oy0 >= oy1 || area[j] <= 0) return;
      int inter = 0;
      for (int y = oy0 + lane; y < oy1; y += 32) {
        for (int x = ox0; x < ox1; x += 32) {
          uint32_t w = mask_chunk(M, i, y - bi.y, x - bi.x) & mask_chunk(M, j, y - bj.y, x - bj.x);
          if (ox1 - x < 32) w &= (1u << (ox1 - x)) - 1u;
          inter += __popc(w);
        }
      }
      inter = __reduce_add_sync(0xffffffffu, inter);
      const int uni = ai + area[j] - inter;
      const float iou = __fdiv_rn((float)inter, (float)uni);
      if (!(iou > iou_thr)) return;
      const float cj = scores[j];
      // nlargest(1, field): the highest confidence, the lowest index among equals
      if (best < 0 || cj > best_conf || (cj == best_conf && j < best)) {
        best = j; best_conf = cj; best_inter = inter; best_uni = uni;
      }
    });
  }
  if (lane != 0) return;
  const bool kept = best >= 0 && best_inter == best_uni && best_conf > confidence;
  keep[i] = kept ? 1 : 0;
  match[i] = best;
  if (best_iou) best_iou[i] = best >= 0 ? __fdiv_rn((float)best_inter, (float)best_uni) : 0.f;
}

struct Scratch {
  cudaStream_t s;
  void* ptrs[16];
  int n = 0;
  explicit Scratch(cudaStream_t st) : s(st) {}
  void* get(size_t bytes) {
    void* p = nullptr;
    if (td_tmp_alloc(&p, bytes, s) != cudaSuccess) return nullptr;
    ptrs[n++] = p;
    return p;
  }
  ~Scratch() {
    for (int i = 0; i < n; ++i) td_tmp_free(ptrs[i], s);
  }
};

int build_grid(Scratch& sc, float4* box32, GridParams* gp, int n, KeyT** keys_out, int** idx_out,
               const long long* n_dev) {
  cudaStream_t st = sc.s;
  auto* keys_a = (KeyT*)sc.get(sizeof(KeyT) * n);
  auto* keys_b = (KeyT*)sc.get(sizeof(KeyT) * n);
  int* idx_a = (int*)sc.get(sizeof(int) * n);
  int* idx_b = (int*)sc.get(sizeof(int) * n);
  if (!keys_a || !keys_b || !idx_a || !idx_b) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
  cell_keys_kernel<<<td_div_up(n, 256), 256, 0, st>>>(box32, n, gp, keys_a, idx_a, n_dev);
  TD_CHECK_LAUNCH("cell_keys");
  size_t tmp_bytes = 0;
  TD_CUDA(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, keys_a, keys_b, idx_a, idx_b, n, 0, 32, st));
  void* tmp = sc.get(tmp_bytes);
  if (!tmp) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
  TD_CUDA(cub::DeviceRadixSort::SortPairs(tmp, tmp_bytes, keys_a, keys_b, idx_a, idx_b, n, 0, 32, st));
  *keys_out = keys_b;
  *idx_out = idx_b;
  return TD_OK;
}

}  // namespace

// nbr_cap < 0: exact mode (one read back sizes the neighbour lists); >= 0: capacity mode, no read back
static int nms_impl(const double* bounds, const double* conf, const double* area, int n, const long long* n_dev,
                    double iou_threshold, double area_threshold, long long nbr_cap, long long* flag,
                    unsigned char* removed, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(bounds && conf && area && removed);
  cudaStream_t st = (cudaStream_t)stream;
  td_ensure_pool();
  Scratch sc(st);
  float4* box32 = (float4*)sc.get(sizeof(float4) * n);
  GridParams* gp = (GridParams*)sc.get(sizeof(GridParams));
  __half* c16 = (__half*)sc.get(sizeof(__half) * n);
  __half* a16 = (__half*)sc.get(sizeof(__half) * n);
  long long* off = (long long*)sc.get(sizeof(long long) * (n + 1));
  long long* deg = (long long*)sc.get(sizeof(long long) * (n + 1));
  int* best = (int*)sc.get(sizeof(int) * n);
  int* state = (int*)sc.get(sizeof(int) * n);
  int* pending = (int*)sc.get(sizeof(int) * 2);
  if (!box32 || !gp || !c16 || !a16 || !off || !deg || !best || !state || !pending) {
    td_set_error("scratch allocation failed");
    return TD_ERR_CUDA;
  }
  const int blocks = td_div_up(n, 256);
  grid_init_kernel<<<1, 1, 0, st>>>(gp);
  prep_boxes_kernel<<<blocks, 256, 0, st>>>(bounds, n, box32, gp, n_dev);
  to_half_kernel<<<blocks, 256, 0, st>>>(conf, area, n, c16, a16, n_dev);
  TD_CHECK_LAUNCH("nms prep");
  PairGrid g;
  g.box32 = box32; g.gp = gp; g.n = n; g.n_dev = n_dev;
  // thresholds are python scalars: compared in the array dtype (float32 / float16)
  const float iou_thr = (float)iou_threshold;
  const __half area_thr = __double2half(area_threshold);
  g.all_pairs = !(iou_thr >= 0.f);
  KeyT* keys = nullptr;
  int* idx = nullptr;
  int rc = build_grid(sc, box32, gp, n, &keys, &idx, n_dev);
  if (rc != TD_OK) return rc;
  g.keys = keys; g.idx = idx;
  int* nbr = nullptr;
  const long long* end_c = off + 1;
  if (nbr_cap >= 0) {
    // capacity form: fixed slots per crown, one pass (deg doubles as the end-offset array)
    const int slots = (int)(nbr_cap / n > 0 ? nbr_cap / n : 1);
    nbr = (int*)sc.get(sizeof(int) * (size_t)n * slots);
    if (!nbr) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
    nms_adjacency_kernel<true><<<blocks, 256, 0, st>>>(g, c16, a16, iou_thr, area_thr, off, nbr, best, slots, deg, flag);
    TD_CHECK_LAUNCH("nms fill (slots)");
    end_c = deg;
  } else {
    TD_CUDA(cudaMemsetAsync(deg, 0, sizeof(long long) * (n + 1), st));
    nms_adjacency_kernel<false><<<blocks, 256, 0, st>>>(g, c16, a16, iou_thr, area_thr, deg, nullptr, nullptr, 0, nullptr,
                                                        nullptr);
    TD_CHECK_LAUNCH("nms count");
    size_t tmp_bytes = 0;
    TD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, deg, off, n + 1, st));
    void* tmp = sc.get(tmp_bytes);
    if (!tmp) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
    TD_CUDA(cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, deg, off, n + 1, st));
    // the neighbour count is data dependent: one 8-byte read back sizes the CSR array
    long long total = 0;
    TD_CUDA(cudaMemcpyAsync(&total, off + n, sizeof(long long), cudaMemcpyDeviceToHost, st));
    TD_CUDA(cudaStreamSynchronize(st));
    nbr = (int*)sc.get(sizeof(int) * (size_t)(total > 0 ? total : 1));
    if (!nbr) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
    nms_adjacency_kernel<true><<<blocks, 256, 0, st>>>(g, c16, a16, iou_thr, area_thr, off, nbr, best, 0, nullptr, nullptr);
    TD_CHECK_LAUNCH("nms fill");
  }
  // cooperative fixed-point resolution: one resident wave
  int per_sm = 0;
  TD_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, nms_resolve_kernel, 256, 0));
  if (per_sm < 1) per_sm = 1;
  int grid = td_num_sms() * per_sm;
  if (grid > blocks) grid = blocks;
  if (grid < 1) grid = 1;
  int n_ = n;
  const long long* off_c = off;
  const int* nbr_c = nbr;
  const int* best_c = best;
  const long long* flag_c = flag;
  void* args[] = {&n_, &off_c, &end_c, &nbr_c, &best_c, &state, &pending, &removed, &n_dev, &flag_c};
  TD_CUDA(cudaLaunchCooperativeKernel((void*)nms_resolve_kernel, dim3(grid), dim3(256), args, 0, st));
  return TD_OK;
}

extern "C" int td_bbox_nms_ordered(const double* bounds, const double* conf, const double* area, int n,
                                   double iou_threshold, double area_threshold, unsigned char* removed,
                                   void* stream) {
  return nms_impl(bounds, conf, area, n, nullptr, iou_threshold, area_threshold, -1, nullptr, removed, stream);
}

// Capacity form of the same operation: n = capacity, *n_dev = live count, neighbour lists limited to
// nbr_cap entries in total.  No host synchronisation; bit 1 of *flag is raised when nbr_cap was too
// small (removed[] is then undefined) and nothing is resolved when *flag is already non-zero.
extern "C" int td_bbox_nms_ordered_dyn(const double* bounds, const double* conf, const double* area, int n,
                                       const long long* n_dev, double iou_threshold, double area_threshold,
                                       long long nbr_cap, long long* flag, unsigned char* removed, void* stream) {
  TD_ARG(n_dev && flag && nbr_cap >= 0);
  return nms_impl(bounds, conf, area, n, n_dev, iou_threshold, area_threshold, nbr_cap, flag, removed, stream);
}

// bounds64 (N,4) f64 (cast to float32 here, as cp.array(..., dtype=float32) does) or bounds32 (N,4) f32
int td_containment_ex(const double* bounds64, const float* bounds32, int n, double threshold, float* ratio_max,
                      unsigned char* is_contained, int* num_contained, const long long* n_dev, cudaStream_t st) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG((bounds32 || bounds64) && ratio_max && is_contained && num_contained);
  td_ensure_pool();
  Scratch sc(st);
  float4* box32 = (float4*)sc.get(sizeof(float4) * n);
  GridParams* gp = (GridParams*)sc.get(sizeof(GridParams));
  if (!box32 || !gp) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
  const int blocks = td_div_up(n, 256);
  grid_init_kernel<<<1, 1, 0, st>>>(gp);
  if (bounds64) prep_boxes_kernel<<<blocks, 256, 0, st>>>(bounds64, n, box32, gp, n_dev);
  else box32_from_f32_kernel<<<blocks, 256, 0, st>>>(bounds32, n, box32, gp, n_dev);
  TD_CHECK_LAUNCH("containment prep");
  PairGrid g;
  g.box32 = box32; g.gp = gp; g.n = n; g.n_dev = n_dev;
  // ratio is float32 and the python threshold is compared in float32
  const float thr = (float)threshold;
  g.all_pairs = !(thr > 0.f);
  KeyT* keys = nullptr;
  int* idx = nullptr;
  int rc = build_grid(sc, box32, gp, n, &keys, &idx, n_dev);
  if (rc != TD_OK) return rc;
  g.keys = keys; g.idx = idx;
  containment_kernel<<<blocks, 256, 0, st>>>(g, thr, ratio_max, is_contained, num_contained);
  TD_CHECK_LAUNCH("containment");
  return TD_OK;
}

extern "C" int td_containment(const float* bounds32, int n, double threshold, float* ratio_max,
                              unsigned char* is_contained, int* num_contained, const long long* n_dev, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(bounds32);
  return td_containment_ex(nullptr, bounds32, n, threshold, ratio_max, is_contained, num_contained, n_dev,
                           (cudaStream_t)stream);
}

// Opt-in mask-IoU cleaner on the packed rasters of P2 (the rule of clean_crowns, TreeDetection/helpers.py:602-701,
// with pixel IoU instead of polygon IoU): keep[i] = the best-confidence crown among those with IoU(i, .) >
// iou_thr coincides with crown i (IoU == 1) and its confidence exceeds `confidence`; match[i] = that crown
// (-1: none, e.g. an empty mask).  win (N,4) [x0,y0,w,h] in tile pixels, tile_org (T,2) [col_off,row_off] of
// each tile window in the image, best_iou nullable.
extern "C" int td_mask_iou_clean(const uint32_t* bits, const long long* word_off, const int* win, const int* tile_org,
                                 const int* inst_tile, const float* scores, int n, float iou_thr, float confidence,
                                 unsigned char* keep, int* match, float* best_iou, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(bits && word_off && win && tile_org && inst_tile && scores && keep && match);
  cudaStream_t st = (cudaStream_t)stream;
  td_ensure_pool();
  Scratch sc(st);
  int4* gbox = (int4*)sc.get(sizeof(int4) * n);
  float4* box32 = (float4*)sc.get(sizeof(float4) * n);
  GridParams* gp = (GridParams*)sc.get(sizeof(GridParams));
  int* area = (int*)sc.get(sizeof(int) * n);
  if (!gbox || !box32 || !gp || !area) { td_set_error("scratch allocation failed"); return TD_ERR_CUDA; }
  const int blocks = td_div_up(n, 256);
  grid_init_kernel<<<1, 1, 0, st>>>(gp);
  mask_gbox_kernel<<<blocks, 256, 0, st>>>(win, tile_org, inst_tile, n, gbox, box32, gp);
  MaskSet M{bits, word_off, gbox};
  mask_area_kernel<<<td_div_up((long long)n * 32, 256), 256, 0, st>>>(M, n, area);
  TD_CHECK_LAUNCH("mask iou prep");
  PairGrid g;
  g.box32 = box32; g.gp = gp; g.n = n; g.n_dev = nullptr; g.all_pairs = 0;
  KeyT* keys = nullptr;
  int* idx = nullptr;
  int rc = build_grid(sc, box32, gp, n, &keys, &idx, nullptr);
  if (rc != TD_OK) return rc;
  g.keys = keys; g.idx = idx;
  mask_iou_kernel<<<td_div_up((long long)n * 32, 128), 128, 0, st>>>(g, M, area, scores, iou_thr, confidence, keep, match,
                                                                   best_iou);
  TD_CHECK_LAUNCH("td_mask_iou_clean");
  return TD_OK;
}
