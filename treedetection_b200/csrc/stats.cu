// P7 -- per-crown raster statistics, and the crown centroids (K11).
//
// Replaces get_metadata_within_polygon (TreeDetection/postprocessing.py:221-347), the
// split pair get_height_within_polygon (:25-115) / get_ndvi_within_polygon (:117-219),
// is_point_in_polygon_batch (utilities.py:78-98) and get_centroids (utilities.py:163-180).
//
// The reference tests EVERY raster pixel against every crown's circle (N x P work,
// ~40 launches per crown).  Here one warp owns one crown and visits only a
// conservative pixel window around the circle; inside that window the membership
// test is evaluated in exactly the reference's arithmetic, so the selected pixel
// set is identical:
//     c = (min + max) / 2 of the float32 vertices, r = max sqrt(dx^2 + dy^2)   (float32)
//     pixel coordinate = a*col + b*row + c0 (float64, corner convention) cast to
//         float32 in the combined and NDVI paths (:266, :160), kept float64 in the
//         height-only path (:67)
//     inside <=> (px - cx)^2 + (py - cy)^2 <= radius^2
//     radius = r for heights, 0.5 r for NDVI in the combined path (:304), r in the
//         NDVI-only path (:195)
// Outputs: max height + coordinates of the FIRST arg-max in row-major order,
// NDVI min / max / mean / population variance; empty set -> -1.
// mean / var are accumulated in float64 and rounded once (the reference sums in
// float32; agreement is to ~1e-7, the contract is 1e-5 absolute).
#include <cstdlib>

#include "chain_internal.cuh"
#include "common.cuh"

namespace {

typedef TdAffine6 Affine6;

enum StatsMode { kCombined = 0, kHeightOnly = 1, kNdviOnly = 2 };

// NaN-aware "numpy argmax" ordering on (value, flat index): NaN beats numbers,
// ties go to the lower index
TD_D bool better_max(float v, long long k, float bv, long long bk) {
  if (bk < 0) return true;
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn || bn) {
    if (vn && bn) return k < bk;
    return vn;
  }
  return v > bv || (v == bv && k < bk);
}
TD_D bool better_min(float v, long long k, float bv, long long bk) {
  if (bk < 0) return true;
  const bool vn = isnan(v), bn = isnan(bv);
  if (vn || bn) {
    if (vn && bn) return k < bk;
    return vn;
  }
  return v < bv || (v == bv && k < bk);
}

template <int MODE, int kRows = 4>
__global__ void __launch_bounds__(256)
crown_stats_kernel(const double* __restrict__ verts, const long long* __restrict__ ring_off,
                   const long long* __restrict__ ring_idx, int n,
                   const float* __restrict__ ndvi, const float* __restrict__ height, int rows, int cols, Affine6 T,
                   const Affine6* __restrict__ T_dev,
                   float* __restrict__ max_h, float* __restrict__ hxy, float* __restrict__ ndvi_stats,
                   const long long* __restrict__ n_dev) {
  const int lane = threadIdx.x & 31;
  const int crown = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (crown >= n || (n_dev && crown >= *n_dev)) return;
  const unsigned full = 0xffffffffu;
  if (T_dev) T = *T_dev;     // per-image transform of the chain (device memory)

  // ---- circle from the float32 vertices ------------------------------------
  const long long ring = ring_idx ? ring_idx[crown] : crown;
  const long long v0 = ring_off[ring], v1 = ring_off[ring + 1];
  float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
  for (long long k = v0 + lane; k < v1; k += 32) {
    const float x = __double2float_rn(verts[2 * k]), y = __double2float_rn(verts[2 * k + 1]);
    if (isnan(x) || isnan(y)) continue;
    mnx = fminf(mnx, x); mxx = fmaxf(mxx, x);
    mny = fminf(mny, y); mxy = fmaxf(mxy, y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(full, mnx, o));
    mxx = fmaxf(mxx, __shfl_xor_sync(full, mxx, o));
    mny = fminf(mny, __shfl_xor_sync(full, mny, o));
    mxy = fmaxf(mxy, __shfl_xor_sync(full, mxy, o));
  }
  const float cx = __fdiv_rn(__fadd_rn(mnx, mxx), 2.f);
  const float cy = __fdiv_rn(__fadd_rn(mny, mxy), 2.f);
  float r = -INFINITY;
  for (long long k = v0 + lane; k < v1; k += 32) {
    const float x = __double2float_rn(verts[2 * k]), y = __double2float_rn(verts[2 * k + 1]);
    if (isnan(x) || isnan(y)) continue;
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy);
    r = fmaxf(r, __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
  }
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(full, r, o));
  const float r_h = r;
  const float r_n = (MODE == kCombined) ? __fmul_rn(r, 0.5f) : r;
  const float r2_h = __fmul_rn(r_h, r_h);
  const float r2_n = __fmul_rn(r_n, r_n);

  // ---- conservative pixel window -------------------------------------------
  int c_lo = 0, c_hi = cols - 1, r_lo = 0, r_hi = rows - 1;
  const bool axis_aligned = (T.b == 0.0 && T.d == 0.0 && T.a != 0.0 && T.e != 0.0);
  if (axis_aligned && isfinite(r) && isfinite(cx) && isfinite(cy)) {
    // margin: float32 rounding of the pixel coordinate (1 ulp of the magnitude)
    // and of the squared-distance arithmetic, generously over-estimated
    const double mag = fmax(fmax(fabs((double)cx), fabs((double)cy)), 1.0);
    const double m = (double)r * 1e-3 + mag * 4.8e-7 + 1e-6;
    const double xa = ((double)cx - (double)r - m - T.c) / T.a, xb = ((double)cx + (double)r + m - T.c) / T.a;
    const double ya = ((double)cy - (double)r - m - T.f) / T.e, yb = ((double)cy + (double)r + m - T.f) / T.e;
    const double cl = floor(fmin(xa, xb)) - 1.0, ch = ceil(fmax(xa, xb)) + 1.0;
    const double rl = floor(fmin(ya, yb)) - 1.0, rh = ceil(fmax(ya, yb)) + 1.0;
    c_lo = (int)fmax(cl, 0.0); c_hi = (int)fmin(ch, (double)(cols - 1));
    r_lo = (int)fmax(rl, 0.0); r_hi = (int)fmin(rh, (double)(rows - 1));
  }
  const int ww = c_hi - c_lo + 1, wh = r_hi - r_lo + 1;

  float bh = 0.f; long long bhk = -1;          // height arg-max
  float nmin = 0.f; long long nmink = -1;      // NDVI arg-min / arg-max
  float nmax = 0.f; long long nmaxk = -1;
  double s1 = 0.0, s2 = 0.0; long long cnt = 0;
  if (ww > 0 && wh > 0 && !isnan(r)) {
    // rows outer, lanes over the columns of a row (no per-pixel division); the arithmetic of the
    // pixel coordinates is the reference's, with the row terms hoisted (same operations, same rounding)
    if (MODE == kHeightOnly) {
      // kRows (four) rows per step: the four raster loads of a lane are issued together, before the distance
      // arithmetic (one dependent load per row left every warp waiting on DRAM for most of its life).
      // The loads are unconditional -- the window lies inside the raster -- and the arg-max carries its
      // flat index, so neither the extra reads nor the order of evaluation change the result.
      for (int rb = r_lo; rb <= r_hi; rb += kRows) {
        for (int cc = c_lo + lane; cc <= c_hi; cc += 32) {
          float v[kRows];
#pragma unroll
          for (int q = 0; q < kRows; ++q)
            v[q] = rb + q <= r_hi ? __ldg(height + ((long long)(rb + q) * cols + cc)) : 0.f;
          const double ax = T.a * (double)cc, dxc = T.d * (double)cc;
#pragma unroll
          for (int q = 0; q < kRows; ++q) {
            const int rr = rb + q;
            if (rr > r_hi) break;
            const double xd = ax + T.b * (double)rr + T.c;
            const double yd = dxc + T.e * (double)rr + T.f;
            const double dx = xd - (double)cx, dy = yd - (double)cy;
            const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
            const long long flat = (long long)rr * cols + cc;
            if (d2 <= (double)r2_h && better_max(v[q], flat, bh, bhk)) { bh = v[q]; bhk = flat; }
          }
        }
      }
    } else
    for (int rr = r_lo; rr <= r_hi; ++rr) {
      const double brr = T.b * (double)rr, err = T.e * (double)rr;
      const long long row_flat = (long long)rr * cols;
      for (int cc = c_lo + lane; cc <= c_hi; cc += 32) {
        const double xd = T.a * (double)cc + brr + T.c;
        const double yd = T.d * (double)cc + err + T.f;
        const long long flat = row_flat + cc;
        if (MODE == kHeightOnly) {
          const double dx = xd - (double)cx, dy = yd - (double)cy;
          const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          if (d2 <= (double)r2_h) {
            const float v = height[flat];
            if (better_max(v, flat, bh, bhk)) { bh = v; bhk = flat; }
          }
        } else {
          const float x = __double2float_rn(xd), y = __double2float_rn(yd);
          const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy);
          const float d2 = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
          if (MODE == kCombined && d2 <= r2_h) {
            const float v = height[flat];
            if (better_max(v, flat, bh, bhk)) { bh = v; bhk = flat; }
          }
          if (d2 <= r2_n) {
            const float v = ndvi[flat];
            if (better_min(v, flat, nmin, nmink)) { nmin = v; nmink = flat; }
            if (better_max(v, flat, nmax, nmaxk)) { nmax = v; nmaxk = flat; }
            s1 += (double)v; s2 += (double)v * (double)v; ++cnt;
          }
        }
      }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    float ov = __shfl_xor_sync(full, bh, o); long long ok = __shfl_xor_sync(full, bhk, o);
    if (ok >= 0 && better_max(ov, ok, bh, bhk)) { bh = ov; bhk = ok; }
    ov = __shfl_xor_sync(full, nmin, o); ok = __shfl_xor_sync(full, nmink, o);
    if (ok >= 0 && better_min(ov, ok, nmin, nmink)) { nmin = ov; nmink = ok; }
    ov = __shfl_xor_sync(full, nmax, o); ok = __shfl_xor_sync(full, nmaxk, o);
    if (ok >= 0 && better_max(ov, ok, nmax, nmaxk)) { nmax = ov; nmaxk = ok; }
    s1 += __shfl_xor_sync(full, s1, o);
    s2 += __shfl_xor_sync(full, s2, o);
    cnt += __shfl_xor_sync(full, cnt, o);
  }
  if (lane != 0) return;
  if (MODE != kNdviOnly) {
    if (bhk < 0) {
      max_h[crown] = -1.f; hxy[2 * crown] = -1.f; hxy[2 * crown + 1] = -1.f;
    } else {
      const int rr = (int)(bhk / cols), cc = (int)(bhk % cols);
      max_h[crown] = bh;
      hxy[2 * crown] = __double2float_rn(T.a * (double)cc + T.b * (double)rr + T.c);
      hxy[2 * crown + 1] = __double2float_rn(T.d * (double)cc + T.e * (double)rr + T.f);
    }
  }
  if (MODE != kHeightOnly) {
    float* o = ndvi_stats + 4 * (size_t)crown;
    if (cnt == 0) {
      o[0] = o[1] = o[2] = o[3] = -1.f;
    } else {
      const double mean = s1 / (double)cnt;
      double var = s2 / (double)cnt - mean * mean;
      if (var < 0.0) var = 0.0;
      o[0] = nmin; o[1] = nmax; o[2] = (float)mean; o[3] = (float)var;
    }
  }
}

// ---- optional nDSM summary per crown: min / mean / percentile (north_star; the reference computes the maximum
// only, postprocessing.py:25-115) -------------------------------------------------------------------------------
// Pixel set = that of get_height_within_polygon (float64 pixel coordinates, full radius).  The percentile is
// numpy's default ("linear"): exact order statistics by a warp-wide radix select on the order-preserving
// integer image of the float32 values (4 passes of 8 bits, 256-bin histograms in shared memory), then
// numpy's _lerp in float64.  out (N,4) f32 = [min, mean, percentile, number of pixels]; empty set -> -1 (count 0),
// a NaN in the set -> NaN (as numpy).
TD_D unsigned f32_key(float v) {
  const unsigned u = __float_as_uint(v);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
TD_D float key_f32(unsigned k) { return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k); }

__global__ void __launch_bounds__(128)
crown_height_summary_kernel(const double* __restrict__ verts, const long long* __restrict__ ring_off,
                            const long long* __restrict__ ring_idx, int n, const float* __restrict__ height, int rows,
                            int cols, Affine6 T, double q01, float* __restrict__ out,
                            const long long* __restrict__ n_dev) {
  __shared__ int s_hist[4][256];
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const int crown = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (crown >= n || (n_dev && crown >= *n_dev)) return;
  int* hist = s_hist[wib];
  const long long ring = ring_idx ? ring_idx[crown] : crown;
  const long long v0 = ring_off[ring], v1 = ring_off[ring + 1];
  float mnx = INFINITY, mxx = -INFINITY, mny = INFINITY, mxy = -INFINITY;
  for (long long k = v0 + lane; k < v1; k += 32) {
    const float x = __double2float_rn(verts[2 * k]), y = __double2float_rn(verts[2 * k + 1]);
    if (isnan(x) || isnan(y)) continue;
    mnx = fminf(mnx, x); mxx = fmaxf(mxx, x);
    mny = fminf(mny, y); mxy = fmaxf(mxy, y);
  }
  for (int o = 16; o > 0; o >>= 1) {
    mnx = fminf(mnx, __shfl_xor_sync(full, mnx, o)); mxx = fmaxf(mxx, __shfl_xor_sync(full, mxx, o));
    mny = fminf(mny, __shfl_xor_sync(full, mny, o)); mxy = fmaxf(mxy, __shfl_xor_sync(full, mxy, o));
  }
  const float cx = __fdiv_rn(__fadd_rn(mnx, mxx), 2.f), cy = __fdiv_rn(__fadd_rn(mny, mxy), 2.f);
  float r = -INFINITY;
  for (long long k = v0 + lane; k < v1; k += 32) {
    const float x = __double2float_rn(verts[2 * k]), y = __double2float_rn(verts[2 * k + 1]);
    if (isnan(x) || isnan(y)) continue;
    const float dx = __fsub_rn(x, cx), dy = __fsub_rn(y, cy);
    r = fmaxf(r, __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy))));
  }
  for (int o = 16; o > 0; o >>= 1) r = fmaxf(r, __shfl_xor_sync(full, r, o));
  const double r2 = (double)__fmul_rn(r, r);
  int c_lo = 0, c_hi = cols - 1, r_lo = 0, r_hi = rows - 1;
  if (T.b == 0.0 && T.d == 0.0 && T.a != 0.0 && T.e != 0.0 && isfinite(r) && isfinite(cx) && isfinite(cy)) {
    const double mag = fmax(fmax(fabs((double)cx), fabs((double)cy)), 1.0);
    const double m = (double)r * 1e-3 + mag * 4.8e-7 + 1e-6;
    const double xa = ((double)cx - (double)r - m - T.c) / T.a, xb = ((double)cx + (double)r + m - T.c) / T.a;
    const double ya = ((double)cy - (double)r - m - T.f) / T.e, yb = ((double)cy + (double)r + m - T.f) / T.e;
    c_lo = (int)fmax(floor(fmin(xa, xb)) - 1.0, 0.0); c_hi = (int)fmin(ceil(fmax(xa, xb)) + 1.0, (double)(cols - 1));
    r_lo = (int)fmax(floor(fmin(ya, yb)) - 1.0, 0.0); r_hi = (int)fmin(ceil(fmax(ya, yb)) + 1.0, (double)(rows - 1));
  }
  // visits every pixel of the set: f(value)
  auto for_each_pixel = [&](auto f) {
    if (isnan(r)) return;
    for (int rr = r_lo; rr <= r_hi; ++rr) {
      const double brr = T.b * (double)rr, err = T.e * (double)rr;
      for (int cc = c_lo + lane; cc <= c_hi; cc += 32) {
        const double dx = (T.a * (double)cc + brr + T.c) - (double)cx, dy = (T.d * (double)cc + err + T.f) - (double)cy;
        if (__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)) <= r2) f(height[(long long)rr * cols + cc]);
      }
    }
  };
  // pass 1: count, sum, min, NaN
  long long cnt = 0;
  double sum = 0.0;
  float vmin = INFINITY;
  int has_nan = 0;
  for_each_pixel([&](float v) {
    ++cnt; sum += (double)v;
    if (isnan(v)) has_nan = 1; else vmin = fminf(vmin, v);
  });
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(full, cnt, o);
    sum += __shfl_xor_sync(full, sum, o);
    vmin = fminf(vmin, __shfl_xor_sync(full, vmin, o));
    has_nan |= __shfl_xor_sync(full, has_nan, o);
  }
  float* o4 = out + 4 * (size_t)crown;
  if (cnt == 0) { if (lane == 0) { o4[0] = o4[1] = o4[2] = -1.f; o4[3] = 0.f; } return; }
  if (has_nan) { if (lane == 0) { o4[0] = o4[1] = o4[2] = nanf(""); o4[3] = (float)cnt; } return; }
  // numpy "linear": virtual index (n - 1) * q, neighbours k and k + 1, weight gamma
  const double virt = (double)(cnt - 1) * q01;
  const long long k = (long long)floor(virt);
  const double gamma = virt - (double)k;
  // radix select of the k-th smallest key
  unsigned prefix = 0u, pmask = 0u;
  long long rank = k;
  for (int shift = 24; shift >= 0; shift -= 8) {
    for (int b = lane; b < 256; b += 32) hist[b] = 0;
    __syncwarp();
    for_each_pixel([&](float v) {
      const unsigned key = f32_key(v);
      if ((key & pmask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1);
    });
    __syncwarp();
    // bin holding the element of rank `rank`: lane l owns bins 8l .. 8l+7
    int mine = 0;
    for (int b = 0; b < 8; ++b) mine += hist[8 * lane + b];
    int incl = mine;
    for (int o = 1; o < 32; o <<= 1) {
      const int t = __shfl_up_sync(full, incl, o);
      if (lane >= o) incl += t;
    }
    const int excl = incl - mine;
    const bool here = rank >= excl && rank < incl;
    int bin = -1;
    long long new_rank = 0;
    if (here) {
      long long acc = excl;
      for (int b = 0; b < 8; ++b) {
        const int h = hist[8 * lane + b];
        if (rank < acc + h) { bin = 8 * lane + b; new_rank = rank - acc; break; }
        acc += h;
      }
    }
    const unsigned owner = __ballot_sync(full, here);
    const int src = __ffs(owner) - 1;
    bin = __shfl_sync(full, bin, src);
    new_rank = __shfl_sync(full, new_rank, src);
    prefix |= (unsigned)bin << shift;
    pmask |= 255u << shift;
    rank = new_rank;
    __syncwarp();
  }
  const float vk = key_f32(prefix);
  // the next order statistic: vk again when it is repeated beyond rank k, else the smallest larger value
  long long le = 0;
  float vnext = INFINITY;
  for_each_pixel([&](float v) {
    if (v <= vk) ++le; else vnext = fminf(vnext, v);
  });
  for (int o = 16; o > 0; o >>= 1) {
    le += __shfl_xor_sync(full, le, o);
    vnext = fminf(vnext, __shfl_xor_sync(full, vnext, o));
  }
  if (lane != 0) return;
  const double a = (double)vk, b = (k + 1 < cnt) ? (double)(le > k + 1 ? vk : vnext) : (double)vk;
  const double diff = b - a;
  double res = a + diff * gamma;                     // numpy _lerp
  if (gamma >= 0.5) res = b - diff * (1.0 - gamma);
  o4[0] = vmin; o4[1] = (float)(sum / (double)cnt); o4[2] = (float)res; o4[3] = (float)cnt;
}

// ---- centroids: np.nanmean over the NaN-padded (N, V) float32 arrays ---------
// numpy reduces every row with its pairwise summation over the PADDED length V
// (NaN -> 0), so the blocking depends on V = the longest ring of the batch.
struct RowGetter {
  const double* verts;  // interleaved x,y
  long long v0;
  int len;    // ring length
  int comp;   // 0 = x, 1 = y
  TD_D float operator()(int i) const {
    if (i >= len) return 0.f;
    const float v = __double2float_rn(verts[2 * (v0 + i) + comp]);
    return isnan(v) ? 0.f : v;
  }
};

__device__ float np_pairwise_sum(const RowGetter& a, int lo, int n) {
  if (n < 8) {
    float res = 0.f;
    for (int i = 0; i < n; ++i) res = __fadd_rn(res, a(lo + i));
    return res;
  }
  if (n <= 128) {
    float r[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = a(lo + j);
    int i;
    for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a(lo + i + j));
    }
    float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                          __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
    for (; i < n; ++i) res = __fadd_rn(res, a(lo + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return __fadd_rn(np_pairwise_sum(a, lo, n2), np_pairwise_sum(a, lo + n2, n - n2));
}

__global__ void max_ring_len_kernel(const long long* __restrict__ ring_off, const long long* __restrict__ ring_idx,
                                    int n, int* __restrict__ vmax, const long long* __restrict__ n_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int len = 0;
  if (i < n && !(n_dev && i >= *n_dev)) {
    const long long r = ring_idx ? ring_idx[i] : i;
    len = (int)(ring_off[r + 1] - ring_off[r]);
  }
  for (int o = 16; o > 0; o >>= 1) len = max(len, __shfl_xor_sync(0xffffffffu, len, o));
  if ((threadIdx.x & 31) == 0) atomicMax(vmax, len);
}

__global__ void centroid_kernel(const double* __restrict__ verts, const long long* __restrict__ ring_off,
                                const long long* __restrict__ ring_idx, int n,
                                const int* __restrict__ vmax, float* __restrict__ centroid,
                                const long long* __restrict__ n_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (n_dev && i >= *n_dev)) return;
  const int V = *vmax;
  RowGetter g;
  const long long ring = ring_idx ? ring_idx[i] : i;
  g.verts = verts; g.v0 = ring_off[ring]; g.len = (int)(ring_off[ring + 1] - ring_off[ring]);
  long long cntx = 0, cnty = 0;
  for (int k = 0; k < g.len; ++k) {
    if (!isnan(__double2float_rn(verts[2 * (g.v0 + k)]))) ++cntx;
    if (!isnan(__double2float_rn(verts[2 * (g.v0 + k) + 1]))) ++cnty;
  }
  g.comp = 0;
  const float sx = np_pairwise_sum(g, 0, V);
  g.comp = 1;
  const float sy = np_pairwise_sum(g, 0, V);
  // _divide_by_count: true_divide(float32 total, int64 count) evaluated in float64
  centroid[2 * i] = (float)((double)sx / (double)cntx);
  centroid[2 * i + 1] = (float)((double)sy / (double)cnty);
}

}  // namespace

int td_crown_stats_ex(const double* verts, const long long* ring_off, const long long* ring_idx, int n,
                      const float* ndvi, const float* height, int rows, int cols, const TdAffine6* tf_host,
                      const TdAffine6* tf_dev, int mode, float* max_h, float* hxy, float* ndvi_stats,
                      const long long* n_dev, cudaStream_t st) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(verts && ring_off && (tf_host || tf_dev) && rows > 0 && cols > 0);
  TD_ARG(mode >= 0 && mode <= 2);
  if (mode != kNdviOnly) TD_ARG(height && max_h && hxy);
  if (mode != kHeightOnly) TD_ARG(ndvi && ndvi_stats);
  Affine6 T = tf_host ? *tf_host : Affine6{1, 0, 0, 0, 1, 0};
  const int threads = 256;                       // 8 crowns per CTA
  const int blocks = td_div_up((long long)n * 32, threads);
  switch (mode) {
    case kCombined:
      crown_stats_kernel<kCombined><<<blocks, threads, 0, st>>>(verts, ring_off, ring_idx, n, ndvi, height, rows, cols,
                                                                 T, tf_dev, max_h, hxy, ndvi_stats, n_dev);
      break;
    case kHeightOnly: {
      static int rows_per_step = 0;    // raster rows in flight per lane (TREEDET_STATS_ROWS: 4 or 8)
      if (rows_per_step == 0) { const char* e = getenv("TREEDET_STATS_ROWS"); rows_per_step = e && atoi(e) >= 8 ? 8 : 4; }
      if (rows_per_step == 8)
        crown_stats_kernel<kHeightOnly, 8><<<blocks, threads, 0, st>>>(verts, ring_off, ring_idx, n, ndvi, height, rows,
                                                                        cols, T, tf_dev, max_h, hxy, ndvi_stats, n_dev);
      else
        crown_stats_kernel<kHeightOnly, 4><<<blocks, threads, 0, st>>>(verts, ring_off, ring_idx, n, ndvi, height, rows,
                                                                        cols, T, tf_dev, max_h, hxy, ndvi_stats, n_dev);
      break;
    }
    default:
      crown_stats_kernel<kNdviOnly><<<blocks, threads, 0, st>>>(verts, ring_off, ring_idx, n, ndvi, height, rows, cols,
                                                                 T, tf_dev, max_h, hxy, ndvi_stats, n_dev);
  }
  TD_CHECK_LAUNCH("td_crown_stats");
  return TD_OK;
}

extern "C" int td_crown_stats(const double* verts, const long long* ring_off, int n, const float* ndvi,
                              const float* height, int rows, int cols, const double* transform6, int mode,
                              float* max_h, float* hxy, float* ndvi_stats, const long long* n_dev, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(transform6);
  const Affine6 T{transform6[0], transform6[1], transform6[2], transform6[3], transform6[4], transform6[5]};
  return td_crown_stats_ex(verts, ring_off, nullptr, n, ndvi, height, rows, cols, &T, nullptr, mode, max_h, hxy,
                           ndvi_stats, n_dev, (cudaStream_t)stream);
}

// vmax_scratch: one int of device scratch (null: taken from the stream-ordered pool)
int td_centroids_ex(const double* verts, const long long* ring_off, const long long* ring_idx, int n, float* centroid,
                    int* vmax_scratch, const long long* n_dev, cudaStream_t st) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(verts && ring_off && centroid);
  int* vmax = vmax_scratch;
  if (!vmax) {
    td_ensure_pool();
    TD_CUDA(td_tmp_alloc((void**)&vmax, sizeof(int), st));
  }
  TD_CUDA(cudaMemsetAsync(vmax, 0, sizeof(int), st));
  max_ring_len_kernel<<<td_div_up(n, 256), 256, 0, st>>>(ring_off, ring_idx, n, vmax, n_dev);
  centroid_kernel<<<td_div_up(n, 128), 128, 0, st>>>(verts, ring_off, ring_idx, n, vmax, centroid, n_dev);
  cudaError_t e = cudaGetLastError();
  if (!vmax_scratch) td_tmp_free(vmax, st);
  if (e != cudaSuccess) { td_set_error("td_centroids: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

extern "C" int td_centroids(const double* verts, const long long* ring_off, int n, float* centroid,
                            const long long* n_dev, void* stream) {
  return td_centroids_ex(verts, ring_off, nullptr, n, centroid, nullptr, n_dev, (cudaStream_t)stream);
}

// Optional nDSM summary per crown over the pixel set of get_height_within_polygon (postprocessing.py:25-115):
// out (N,4) f32 = [min, mean, percentile `q` (0..100, numpy "linear"), number of pixels]; -1 when the set is empty.
extern "C" int td_crown_height_summary(const double* verts, const long long* ring_off, int n, const float* height,
                                       int rows, int cols, const double* transform6, double q, float* out,
                                       const long long* n_dev, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(verts && ring_off && height && transform6 && out && rows > 0 && cols > 0 && q >= 0.0 && q <= 100.0);
  const Affine6 T{transform6[0], transform6[1], transform6[2], transform6[3], transform6[4], transform6[5]};
  crown_height_summary_kernel<<<td_div_up((long long)n * 32, 128), 128, 0, (cudaStream_t)stream>>>(
      verts, ring_off, nullptr, n, height, rows, cols, T, q / 100.0, out, n_dev);
  TD_CHECK_LAUNCH("td_crown_height_summary");
  return TD_OK;
}
