// P1 -- tile cut + normalise, and P0a -- seam strips.
//
// P1 replaces Predictor._process_tile (TreeDetection/prediction.py:159-176):
//   rasterio.mask(crop=True) of the tile box  -> the pixel window (tiling.tile_grid)
//   np.dstack((b2, b1, b0))                    -> channel c of the output = band 2 - c
//   255 * rgb / 65535 if max(band 1) > 255     -> 16-bit branch, float64 (the uint16
//                                                 product wraps modulo 2^16, as numpy does)
//   ResizeShortestEdge(800, 1333)              -> uint8: PIL Image.resize(BILINEAR), i.e. a
//        separable triangle filter in 22-bit fixed point, horizontal pass rounded to uint8,
//        then vertical pass; float: F.interpolate(bilinear, align_corners=False) in float64
//   astype(float32).transpose(2, 0, 1)         -> float32 CHW
// The reference does this per tile on the CPU and ships 7.68 MB per tile to the GPU;
// here the image is resident in HBM, every CTA produces a 128 x 32 block of output pixels
// for the three channels: the source window and the horizontal pass of the few source rows
// it needs live in shared memory, the vertical pass streams 128-bit coalesced float32 rows out.  The kernel is
// write bound: 0.81 MB in / 7.68 MB out per interior tile.
//
// P0a replaces crop_single_image / merge_images / crop_image (TreeDetection/merging.py:34-110,
// TreeDetection/helpers.py:1023-1085): the strip is gathered straight from the two source
// rasters; the 2-image mosaic (800 MB at 10k x 10k) is never built.
#include <cuda.h>

#include <map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kBX = 128;           // output columns per CTA
constexpr int kBY = 32;            // output rows per CTA
constexpr int kThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;   // PIL: 22-bit fixed-point coefficients
constexpr int kMaxK = 8;           // taps per axis supported (down-scaling up to ~3.5x)

struct TileDesc {
  int c_off, r_off, w, h;   // source window
  int nh, nw;               // output size
  long long out_off;        // float offset of the (3, nh, nw) block
  int xtab, ytab;           // offsets into the coefficient tables (entries)
  int kx, ky;               // taps per output index
  int blk0, bx;             // first CTA of the tile in the flat grid, CTAs per row of CTAs
};

// host: PIL precompute_coeffs + normalize_coeffs_8bpc for the bilinear (triangle) filter
void pil_coeffs(int in_size, int out_size, std::vector<int>& xmin, std::vector<int>& cnt, std::vector<int>& kk,
                std::vector<double>& kd, int& ksize) {
  const double scale = (double)in_size / (double)out_size;
  double filterscale = scale;
  if (filterscale < 1.0) filterscale = 1.0;
  const double support = 1.0 * filterscale;
  ksize = (int)ceil(support) * 2 + 1;
  xmin.resize(out_size); cnt.resize(out_size);
  kk.assign((size_t)out_size * ksize, 0);
  kd.assign((size_t)out_size * ksize, 0.0);
  const double ss = 1.0 / filterscale;
  for (int xx = 0; xx < out_size; ++xx) {
    const double center = (xx + 0.5) * scale;
    double ww = 0.0;
    int x0 = (int)(center - support + 0.5);
    if (x0 < 0) x0 = 0;
    int x1 = (int)(center + support + 0.5);
    if (x1 > in_size) x1 = in_size;
    const int n = x1 - x0;
    double* k = &kd[(size_t)xx * ksize];
    for (int x = 0; x < n; ++x) {
      double t = (x + x0 - center + 0.5) * ss;
      if (t < 0.0) t = -t;
      const double w = t < 1.0 ? 1.0 - t : 0.0;
      k[x] = w;
      ww += w;
    }
    for (int x = 0; x < n; ++x)
      if (ww != 0.0) k[x] /= ww;
    for (int x = 0; x < n; ++x)
      kk[(size_t)xx * ksize + x] = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << kPrecisionBits))
                                            : (int)(0.5 + k[x] * (1 << kPrecisionBits));
    xmin[xx] = x0;
    cnt[xx] = n;
  }
}

TD_D int clip8(int v) {
  v >>= kPrecisionBits;
  return min(max(v, 0), 255);
}

// ---- uint8 tiles: PIL fixed-point separable resize ---------------------------------------
// One CTA = 128 x 32 output pixels x 3 channels of one tile.
//   phase 0: the source window (<= max_rows rows x src_stride bytes x 3 bands) is staged in
//            shared memory with coalesced 32-bit loads (one warp per source row);
//   phase 1: horizontal pass, one thread per output column, rounded to uint8 (as PIL does);
//   phase 2: vertical pass, one warp per output row, 4 adjacent pixels per lane from one
//            packed 32-bit shared load per tap, 128-bit coalesced float stores.
// K = compile-time tap bound (3: up-scaling, the usual case; 8: moderate down-scaling).
template <int K>
__global__ void __launch_bounds__(kThreads)
tile_resize_u8_kernel(const unsigned char* __restrict__ image, long long image_bytes, int H, int W,
                      const TileDesc* __restrict__ tiles, int n_tiles, const int* __restrict__ tab_min,
                      const int* __restrict__ tab_cnt, const int* __restrict__ tab_k, float* __restrict__ out,
                      int max_rows, int src_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* s_src = smem;                                             // [3][max_rows][src_stride]
  unsigned char* s_tmp = smem + (size_t)3 * max_rows * src_stride;         // [3][max_rows][kBX]
  int* s_ytab = reinterpret_cast<int*>(s_tmp + (size_t)3 * max_rows * kBX);  // [kBY][2 + K]
  // ---- which tile / block --------------------------------------------------------------
  int lo = 0, hi = n_tiles - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (tiles[mid].blk0 <= (int)blockIdx.x) lo = mid; else hi = mid - 1;
  }
  const TileDesc T = tiles[lo];
  const int b = blockIdx.x - T.blk0;
  const int ox0 = (b % T.bx) * kBX, oy0 = (b / T.bx) * kBY;
  const int ox_last = min(ox0 + kBX, T.nw) - 1, oy_last = min(oy0 + kBY, T.nh) - 1;
  const int row_lo = tab_min[T.ytab + oy0];
  const int nrows = tab_min[T.ytab + oy_last] + tab_cnt[T.ytab + oy_last] - row_lo;
  const int col_lo = tab_min[T.xtab + ox0];
  const int ncols = tab_min[T.xtab + ox_last] + tab_cnt[T.xtab + ox_last] - col_lo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const size_t plane = (size_t)H * W;
  // ---- phase 0: stage the window -----------------------------------------------------------
  for (int rr = warp; rr < 3 * nrows; rr += kThreads / 32) {
    const int c = rr / nrows, r = rr - c * nrows;
    // output channel c reads band 2 - c (BGR order)
    const size_t a = (size_t)(2 - c) * plane + (size_t)(T.r_off + row_lo + r) * W + T.c_off + col_lo;
    const size_t a0 = a & ~(size_t)3;
    const int shift = (int)(a - a0);
    const int nwords = (shift + ncols + 3) >> 2;
    unsigned char* dst = s_src + ((size_t)c * max_rows + r) * src_stride;
    for (int q = lane; q < nwords; q += 32) {
      const size_t wa = a0 + 4 * (size_t)q;
      uint32_t v;
      if (wa + 4 <= (size_t)image_bytes) {
        v = *reinterpret_cast<const uint32_t*>(image + wa);
      } else {
        v = 0;
        for (int z = 0; z < 4; ++z)
          if (wa + z < (size_t)image_bytes) v |= (uint32_t)image[wa + z] << (8 * z);
      }
      *reinterpret_cast<uint32_t*>(dst + 4 * q) = v;
    }
  }
  // y taps of the CTA's rows
  if (threadIdx.x < kBY) {
    const int oy = min(oy0 + (int)threadIdx.x, T.nh - 1);
    int* e = s_ytab + threadIdx.x * (2 + K);
    e[0] = tab_min[T.ytab + oy] - row_lo;
    e[1] = tab_cnt[T.ytab + oy];
    const int* k = tab_k + (size_t)(T.ytab + oy) * kMaxK;
#pragma unroll
    for (int q = 0; q < K; ++q) e[2 + q] = k[q];
  }
  __syncthreads();
  // ---- phase 1: horizontal pass ------------------------------------------------------------
  {
    const int x = threadIdx.x & (kBX - 1), g = threadIdx.x >> 7;
    const int ox = ox0 + x;
    int xs = 0, kx[K];
#pragma unroll
    for (int q = 0; q < K; ++q) kx[q] = 0;
    if (ox < T.nw) {
      xs = tab_min[T.xtab + ox] - col_lo;
      const int* k = tab_k + (size_t)(T.xtab + ox) * kMaxK;
#pragma unroll
      for (int q = 0; q < K; ++q) kx[q] = k[q];      // taps beyond the count are 0
    }
    const size_t a_row0 = (size_t)T.c_off + col_lo;   // only the low 2 bits matter below
    for (int c = 0; c < 3; ++c) {
      for (int r = g; r < nrows; r += kThreads / kBX) {
        const size_t a = (size_t)(2 - c) * plane + (size_t)(T.r_off + row_lo + r) * W + a_row0;
        const int shift = (int)(a & 3);
        const unsigned char* p = s_src + ((size_t)c * max_rows + r) * src_stride + shift + xs;
        int ss = 1 << (kPrecisionBits - 1);
#pragma unroll
        for (int q = 0; q < K; ++q) ss += (int)p[q] * kx[q];
        s_tmp[((size_t)c * max_rows + r) * kBX + x] = (unsigned char)(ox < T.nw ? clip8(ss) : 0);
      }
    }
  }
  __syncthreads();
  // ---- phase 2: vertical pass ----------------------------------------------------------------
  float* o = out + T.out_off;
  const size_t oplane = (size_t)T.nh * T.nw;
  const bool vec = ((T.nw & 3) == 0) && ((T.out_off & 3) == 0);
  for (int y = warp; y < kBY; y += kThreads / 32) {
    const int oy = oy0 + y;
    if (oy >= T.nh) break;
    const int* e = s_ytab + y * (2 + K);
    const int ys = e[0];
    int ky[K];
#pragma unroll
    for (int q = 0; q < K; ++q) ky[q] = e[2 + q];
    if (vec) {
      const int x4 = 4 * lane;
      if (ox0 + x4 >= T.nw) continue;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        int s0 = 1 << (kPrecisionBits - 1), s1 = s0, s2 = s0, s3 = s0;
        const unsigned char* col = s_tmp + ((size_t)c * max_rows + ys) * kBX + x4;
#pragma unroll
        for (int q = 0; q < K; ++q) {
          // rows beyond the tap count have coefficient 0; stay inside the staged rows
          const uint32_t u = (K <= 3 || q < e[1]) ? *reinterpret_cast<const uint32_t*>(col + (size_t)q * kBX) : 0u;
          s0 += (int)(u & 255u) * ky[q];
          s1 += (int)((u >> 8) & 255u) * ky[q];
          s2 += (int)((u >> 16) & 255u) * ky[q];
          s3 += (int)(u >> 24) * ky[q];
        }
        float4 v = make_float4((float)clip8(s0), (float)clip8(s1), (float)clip8(s2), (float)clip8(s3));
        *reinterpret_cast<float4*>(o + c * oplane + (size_t)oy * T.nw + ox0 + x4) = v;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int x = lane + 32 * j;
        if (ox0 + x >= T.nw) break;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          int ss = 1 << (kPrecisionBits - 1);
          const unsigned char* col = s_tmp + ((size_t)c * max_rows + ys) * kBX + x;
#pragma unroll
          for (int q = 0; q < K; ++q)
            if (K <= 3 || q < e[1]) ss += (int)col[(size_t)q * kBX] * ky[q];
          o[c * oplane + (size_t)oy * T.nw + ox0 + x] = (float)clip8(ss);
        }
      }
    }
  }
}

// ---- shared phases of the up-scaling fast path ----------------------------------------------
struct XTaps {
  int xs[4], k0[4], k1[4];
};

// taps of the 4 adjacent output columns of this lane; `bias` is added to the source offsets
TD_D XTaps load_xtaps(const TileDesc& T, const int* __restrict__ tab_min, const int* __restrict__ tab_k, int ox,
                      int col_lo, int bias) {
  XTaps t;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    t.xs[j] = 0; t.k0[j] = 0; t.k1[j] = 0;
    if (ox + j < T.nw) {
      t.xs[j] = tab_min[T.xtab + ox + j] - col_lo + bias;
      const int2 kk = *reinterpret_cast<const int2*>(tab_k + (size_t)(T.xtab + ox + j) * kMaxK);
      t.k0[j] = kk.x; t.k1[j] = kk.y;
    }
  }
  return t;
}

// phase 1: horizontal pass, 4 columns per thread, one packed word per (row, channel)
TD_D void up_phase1(const unsigned char* s_src, int chan_stride, int row_stride, unsigned char* s_tmp, int max_rows,
                    const XTaps& t, int nrows, int warp, int x4) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kHalf = 1 << (kPrecisionBits - 1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const unsigned char* src_c = s_src + (size_t)c * chan_stride;
    unsigned char* tmp_c = s_tmp + (size_t)c * max_rows * kBX + x4;
    for (int r = warp; r < nrows; r += kWarps) {
      const unsigned char* p = src_c + r * row_stride;
      const uint32_t v0 = (uint32_t)((int)p[t.xs[0]] * t.k0[0] + (int)p[t.xs[0] + 1] * t.k1[0] + kHalf) >> kPrecisionBits;
      const uint32_t v1 = (uint32_t)((int)p[t.xs[1]] * t.k0[1] + (int)p[t.xs[1] + 1] * t.k1[1] + kHalf) >> kPrecisionBits;
      const uint32_t v2 = (uint32_t)((int)p[t.xs[2]] * t.k0[2] + (int)p[t.xs[2] + 1] * t.k1[2] + kHalf) >> kPrecisionBits;
      const uint32_t v3 = (uint32_t)((int)p[t.xs[3]] * t.k0[3] + (int)p[t.xs[3] + 1] * t.k1[3] + kHalf) >> kPrecisionBits;
      *reinterpret_cast<uint32_t*>(tmp_c + r * kBX) =
          __byte_perm(__byte_perm(v0, v1, 0x0040), __byte_perm(v2, v3, 0x0040), 0x5410);
    }
  }
}

// phase 2: vertical pass, 4 consecutive output rows per warp, 4 pixels per lane
TD_D void up_phase2(const TileDesc& T, const unsigned char* s_tmp, const int4* s_ytab, int max_rows, int ox0, int oy0,
                    int warp, int x4, float* __restrict__ out) {
  constexpr int kWarps = kThreads / 32;
  constexpr int kHalf = 1 << (kPrecisionBits - 1);
  constexpr int kRows = kBY / kWarps;
  const size_t oplane = (size_t)T.nh * T.nw;
  const bool vec = ((T.nw & 3) == 0) && ((T.out_off & 3) == 0);
  const int y_begin = warp * kRows;
  float* orow = out + T.out_off + (size_t)(oy0 + y_begin) * T.nw + ox0 + x4;
  const unsigned char* colbase = s_tmp + x4;
  const int cstride = max_rows * kBX;
  int cur = -2;
  int a[3][4], bb[3][4];
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) { a[c][j] = 0; bb[c][j] = 0; }
#pragma unroll
  for (int yy = 0; yy < kRows; ++yy) {
    if (oy0 + y_begin + yy >= T.nh) break;
    const int4 e = s_ytab[y_begin + yy];
    if (e.x != cur) {
      const bool adv = (e.x == cur + 1);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        if (adv) {
#pragma unroll
          for (int j = 0; j < 4; ++j) a[c][j] = bb[c][j];
        } else {
          const uint32_t u = *reinterpret_cast<const uint32_t*>(colbase + c * cstride + e.x * kBX);
          a[c][0] = __byte_perm(u, 0, 0x4440); a[c][1] = __byte_perm(u, 0, 0x4441);
          a[c][2] = __byte_perm(u, 0, 0x4442); a[c][3] = __byte_perm(u, 0, 0x4443);
        }
        // row e.x + 1 may be one past the staged rows when the tap count is 1: its weight is 0
        const uint32_t u = *reinterpret_cast<const uint32_t*>(colbase + c * cstride + (e.x + 1) * kBX);
        bb[c][0] = __byte_perm(u, 0, 0x4440); bb[c][1] = __byte_perm(u, 0, 0x4441);
        bb[c][2] = __byte_perm(u, 0, 0x4442); bb[c][3] = __byte_perm(u, 0, 0x4443);
      }
      cur = e.x;
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 f;
      // (acc >> 22) | 0x4B000000 in ONE funnel shift (hi word 0x4B000000 >> 10), then minus 2^23:
      // exact int -> float for 0..255; no clipping needed (2 non-negative taps summing to 2^22 +- 1)
      f.x = __uint_as_float(__funnelshift_r((uint32_t)(a[c][0] * e.y + bb[c][0] * e.z + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      f.y = __uint_as_float(__funnelshift_r((uint32_t)(a[c][1] * e.y + bb[c][1] * e.z + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      f.z = __uint_as_float(__funnelshift_r((uint32_t)(a[c][2] * e.y + bb[c][2] * e.z + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      f.w = __uint_as_float(__funnelshift_r((uint32_t)(a[c][3] * e.y + bb[c][3] * e.z + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      float* dst = orow + c * oplane;
      if (vec) {
        __stcs(reinterpret_cast<float4*>(dst), f);
      } else {
        const int rem = T.nw - (ox0 + x4);
        dst[0] = f.x;
        if (rem > 1) dst[1] = f.y;
        if (rem > 2) dst[2] = f.z;
        if (rem > 3) dst[3] = f.w;
      }
    }
    orow += T.nw;
  }
}

// ---- fast path: up-scaling (at most 2 taps per axis), the configuration of every tile
// of the reference's example config (450 -> 800 px).  Same phases as above, tuned for
// instruction count (the first version of this kernel was issue bound, not HBM bound):
//   phase 0 looks its tile up in a per-CTA table, stages the window byte-aligned (funnel
//           shift) so that later phases need no per-row shift;
//   phase 1 makes 4 adjacent columns per thread and stores them as one packed word;
//   phase 2 gives each warp 4 CONSECUTIVE output rows, so the two intermediate rows a
//           row needs are usually already unpacked in registers (0.56 new rows per output
//           row at 450 -> 800); no clipping is needed (2 non-negative taps summing to
//           2^22 +- 1 cannot leave [0, 255]) and int -> float is one LOP + one FADD.
__global__ void __launch_bounds__(kThreads)
tile_resize_u8_up_kernel(const unsigned char* __restrict__ image, long long image_bytes, int H, int W,
                         const TileDesc* __restrict__ tiles, const int2* __restrict__ blk,
                         const int* __restrict__ tab_min, const int* __restrict__ tab_cnt,
                         const int* __restrict__ tab_k, float* __restrict__ out, int max_rows, int src_stride) {
  extern __shared__ __align__(16) unsigned char smem[];
  unsigned char* s_src = smem;                                             // [3][max_rows][src_stride]
  unsigned char* s_tmp = smem + (size_t)3 * max_rows * src_stride;         // [3][max_rows][kBX]
  int4* s_ytab = reinterpret_cast<int4*>(s_tmp + (size_t)3 * max_rows * kBX + kBX);  // [kBY]: ys, k0, k1, -
  const int2 me = blk[blockIdx.x];
  const TileDesc T = tiles[me.x];
  const int ox0 = (me.y & 0xffff) * kBX, oy0 = (me.y >> 16) * kBY;
  const int ox_last = min(ox0 + kBX, T.nw) - 1, oy_last = min(oy0 + kBY, T.nh) - 1;
  const int row_lo = tab_min[T.ytab + oy0];
  const int nrows = tab_min[T.ytab + oy_last] + tab_cnt[T.ytab + oy_last] - row_lo;
  const int col_lo = tab_min[T.xtab + ox0];
  const int ncols = tab_min[T.xtab + ox_last] + tab_cnt[T.xtab + ox_last] - col_lo;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kWarps = kThreads / 32;
  const size_t plane = (size_t)H * W;
  const unsigned char* img_end = image + image_bytes;
  // ---- phase 0: stage the window, one warp per source row, 32-bit loads ---------------------
  {
    const int nwords = (ncols + 3) >> 2;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      // output channel c reads band 2 - c (BGR order)
      const unsigned char* base = image + (size_t)(2 - c) * plane + (size_t)(T.r_off + row_lo) * W + T.c_off + col_lo;
      unsigned char* dst_c = s_src + (size_t)c * max_rows * src_stride;
      for (int r = warp; r < nrows; r += kWarps) {
        const unsigned char* a = base + (size_t)r * W;
        const unsigned char* a0 = reinterpret_cast<const unsigned char*>(reinterpret_cast<uintptr_t>(a) & ~(uintptr_t)3);
        const int sh = 8 * (int)(a - a0);
        for (int q = lane; q < nwords; q += 32) {
          const unsigned char* wa = a0 + 4 * q;
          uint32_t w0 = 0, w1 = 0;
          if (wa + 8 <= img_end) {
            w0 = *reinterpret_cast<const uint32_t*>(wa);
            w1 = *reinterpret_cast<const uint32_t*>(wa + 4);
          } else {
            for (int z = 0; z < 4; ++z) {
              if (wa + z < img_end) w0 |= (uint32_t)wa[z] << (8 * z);
              if (wa + 4 + z < img_end) w1 |= (uint32_t)wa[4 + z] << (8 * z);
            }
          }
          *reinterpret_cast<uint32_t*>(dst_c + (size_t)r * src_stride + 4 * q) = __funnelshift_r(w0, w1, sh);
        }
      }
    }
  }
  if (threadIdx.x < kBY) {
    const int oy = min(oy0 + (int)threadIdx.x, T.nh - 1);
    const int* k = tab_k + (size_t)(T.ytab + oy) * kMaxK;
    s_ytab[threadIdx.x] = make_int4(tab_min[T.ytab + oy] - row_lo, k[0], k[1], 0);
  }
  __syncthreads();
  // ---- phase 1 + 2 ------------------------------------------------------------------------------
  const int x4 = 4 * lane;
  const XTaps taps = load_xtaps(T, tab_min, tab_k, ox0 + x4, col_lo, 0);
  up_phase1(s_src, max_rows * src_stride, src_stride, s_tmp, max_rows, taps, nrows, warp, x4);
  __syncthreads();
  if (ox0 + x4 >= T.nw) return;
  up_phase2(T, s_tmp, s_ytab, max_rows, ox0, oy0, warp, x4, out);
}

TD_D uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp-autonomous variant of the fast path ("v6") -------------------------------------------
// The CTA kernels above spend more than half of their issue slots outside the arithmetic: the
// intermediate (horizontally filtered) rows make a round trip through shared memory (pack, store,
// barrier, load, unpack) and every short-lived CTA pays a serial prologue (table look-ups, TMA
// round trip) that its few sibling CTAs must hide.  Here ONE WARP owns a 128-column strip of
// `nrows` consecutive output rows of one tile and needs nobody else:
//   * source rows stream through a private two-stage ring in shared memory, filled by TMA
//     (one cp.async.bulk.tensor.3d per chunk: box = box_w bytes x kChunk rows x 3 bands; chunk
//     k + 1 is requested when the first row of chunk k is touched, so the copy hides behind the
//     ~14 output rows the chunk feeds);
//   * the horizontal pass of a source row goes straight into registers (4 pixels x 3 channels
//     per lane) exactly once per strip -- the two source rows an output row blends live in two
//     register sets indexed by the parity of the source row, so advancing a row overwrites the
//     stale set and only the two vertical coefficients swap;
//   * the vertical pass streams 128-bit coalesced float rows out (st.global.cs).
// No __syncthreads, no shared-memory intermediates: ~8 instructions per output float instead of ~20.
constexpr int kChunk = 8;          // source rows per TMA chunk
constexpr int kWarpsV6 = 4;        // warps per CTA (independent of each other)
#ifndef TD_P1_MINBLOCKS
#define TD_P1_MINBLOCKS 6
#endif
constexpr int kMinBlocksV6 = TD_P1_MINBLOCKS;   // 6 CTAs x 4 warps per SM: 80 registers, no spills

// One warp's work, fully resolved at plan time (64 bytes, four independent 128-bit loads): the
// kernel's prologue is one round trip for this record and one for the taps it points to.
struct StripItem {
  int x_box, y_box;        // TMA start coordinates of chunk 0 (x_box 16-byte aligned)
  int n_chunks, nrows;     // source chunks / output rows of this item
  int xtab, ytab;          // first entries of the strip's x taps / the item's y taps in the tables
  int xbias, row_lo;       // added to tab_min[x] -> byte offset in a staged row; first source row (table units)
  long long out_off;       // float offset of (channel 0, first row, first column of the strip)
  long long oplane;        // floats per output channel plane
  int nw, ncols;           // output row pitch; live columns of the strip (<= 128)
  int pad0, pad1;
};
static_assert(sizeof(StripItem) == 64, "StripItem is loaded as four int4");

TD_D void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}

// kVec: every tile row is 16-byte aligned in the output (nw % 4 == 0, out_off % 4 == 0);
// kBoxW: bytes per staged source row (compile-time so that every shared-memory offset is an immediate);
// tab_y: per output row {first source row, k0, k1, tap count} (the y tables of tab_min / tab_k, interleaved)
template <bool kVec, int kBoxW>
__global__ void __launch_bounds__(32 * kWarpsV6, kMinBlocksV6)
tile_resize_u8_up_warp_kernel(const __grid_constant__ CUtensorMap tmap, const StripItem* __restrict__ items,
                              int n_items, const int* __restrict__ tab_min, const int* __restrict__ tab_k,
                              const int4* __restrict__ tab_y, float* __restrict__ out) {
  constexpr int box_w = kBoxW;
  extern __shared__ __align__(128) unsigned char smem_v6[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * kWarpsV6 + warp;
  if (item >= n_items) return;
  constexpr int band_bytes = kChunk * box_w;      // one band of one chunk
  constexpr int stage_bytes = 3 * band_bytes;     // multiple of 128 (box_w is a multiple of 16)
  unsigned char* ring = smem_v6 + (size_t)warp * 2 * stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_v6 + (size_t)kWarpsV6 * 2 * stage_bytes) + 2 * warp;
  float* stage = reinterpret_cast<float*>(smem_v6 + (size_t)kWarpsV6 * (2 * stage_bytes + 16)) + kBX * warp;   // !kVec only
  const uint32_t ring_s = smem_u32(ring), bar_s = smem_u32(bars);
  const int4* ip = reinterpret_cast<const int4*>(items + item);
  const int4 i0 = ip[0], i1 = ip[1], i2 = ip[2], i3 = ip[3];
  const int x_box = i0.x, y_box = i0.y, n_chunks = i0.z, nrows = i0.w;
  const int xbias = i1.z, row_lo = i1.w;
  const long long out_off = ((long long)(uint32_t)i2.x) | ((long long)i2.y << 32);
  const size_t oplane = (size_t)(((long long)(uint32_t)i2.z) | ((long long)i2.w << 32));
  const int nw = i3.x, ncols = i3.y;
  const int x4 = 4 * lane;
  const int4* __restrict__ yrec = tab_y + i1.y;
  auto request = [&](int chunk) {   // lane 0 only
    const uint32_t bar = bar_s + 8 * (chunk & 1);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"((uint32_t)stage_bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(ring_s + (chunk & 1) * stage_bytes), "l"(&tmap), "r"(x_box), "r"(y_box + chunk * kChunk), "r"(0),
        "r"(bar)
        : "memory");
  };
  if (lane == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_s + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    request(0);
  }
  __syncwarp();
  // x taps of this lane's 4 columns (byte offsets inside a staged row)
  XTaps taps;
  const bool live = x4 < ncols;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    taps.xs[j] = 0; taps.k0[j] = 0; taps.k1[j] = 0;
    if (x4 + j < ncols) {
      taps.xs[j] = tab_min[i1.x + x4 + j] + xbias;
      const int2 kk = *reinterpret_cast<const int2*>(tab_k + (size_t)(i1.x + x4 + j) * kMaxK);
      taps.k0[j] = kk.x; taps.k1[j] = kk.y;
    }
  }
  float* orow = out + out_off + x4;
  constexpr int kHalf = 1 << (kPrecisionBits - 1);
  int E[3][4], O[3][4];      // horizontally filtered source rows of even / odd parity
#pragma unroll
  for (int c = 0; c < 3; ++c)
#pragma unroll
    for (int j = 0; j < 4; ++j) { E[c][j] = 0; O[c][j] = 0; }
  int have = -1;             // last source row (relative to row_lo) already filtered
  int cur_chunk = -1;
  // horizontal pass of source row r into one register set
  auto hpass = [&](int r, int (&dst)[3][4]) {
    const int chunk = r / kChunk;
    if (chunk != cur_chunk) {
      cur_chunk = chunk;
      __syncwarp();                                   // every lane is done with chunk - 1
      if (lane == 0 && chunk + 1 < n_chunks) request(chunk + 1);
      mbar_wait(bar_s + 8 * (chunk & 1), (uint32_t)((chunk >> 1) & 1));
    }
    const unsigned char* p = ring + (chunk & 1) * stage_bytes + (r & (kChunk - 1)) * box_w;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const unsigned char* q = p + (2 - c) * band_bytes;     // output channel c reads band 2 - c (BGR order)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        dst[c][j] = (int)((uint32_t)((int)q[taps.xs[j]] * taps.k0[j] + (int)q[taps.xs[j] + 1] * taps.k1[j] + kHalf) >>
                          kPrecisionBits);
    }
  };
  int4 rec_n = yrec[0];
  for (int y = 0; y < nrows; ++y) {
    const int4 rec = rec_n;
    if (y + 1 < nrows) rec_n = yrec[y + 1];   // next row's taps are in flight while this row is computed
    const int ys = rec.x - row_lo;
    // rec.w = tap count (1 on the last source row of a window: row ys + 1 is then not needed).  Using
    // all four fields also keeps the register of the in-flight prefetch from being recycled early.
    const int need = ys + rec.w - 1;
    while (have < need) {
      ++have;
      if (have & 1) hpass(have, O); else hpass(have, E);
    }
    const int ke = (ys & 1) ? rec.z : rec.y, ko = (ys & 1) ? rec.y : rec.z;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float4 f;
      // (acc >> 22) | 0x4B000000 in one funnel shift, minus 2^23: exact int -> float for 0..255; no
      // clipping needed (2 non-negative taps summing to 2^22 +- 1)
      f.x = __uint_as_float(__funnelshift_r((uint32_t)(E[c][0] * ke + O[c][0] * ko + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      f.y = __uint_as_float(__funnelshift_r((uint32_t)(E[c][1] * ke + O[c][1] * ko + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      f.z = __uint_as_float(__funnelshift_r((uint32_t)(E[c][2] * ke + O[c][2] * ko + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      f.w = __uint_as_float(__funnelshift_r((uint32_t)(E[c][3] * ke + O[c][3] * ko + kHalf), 0x0012C000u, kPrecisionBits)) - 8388608.0f;
      float* dst = orow + c * oplane;
      if (kVec) {
        if (live) __stcs(reinterpret_cast<float4*>(dst), f);
      } else {
        // rows of this tile are not 16-byte aligned: transpose the warp's 128 floats through shared
        // memory so that each of the four scalar stores writes 32 consecutive floats
        __syncwarp();
        *reinterpret_cast<float4*>(stage + x4) = f;
        __syncwarp();
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int col = lane + 32 * m;
          if (col < ncols) __stcs(dst - x4 + col, stage[col]);
        }
      }
    }
    orow += nw;
  }
}

// host: tensor map of the (W, H, bands) uint8 raster with a (box_w, box_h, box_bands) box
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

bool make_image_tensor_map(CUtensorMap* map, const void* image, int bands, int H, int W, int box_w, int box_h,
                           int box_bands) {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
  }
  if (!fn) return false;
  const cuuint64_t dims[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)bands};
  const cuuint64_t strides[2] = {(cuuint64_t)W, (cuuint64_t)W * (cuuint64_t)H};   // bytes, dims 1 and 2
  const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_bands};
  const cuuint32_t estr[3] = {1u, 1u, 1u};
  return fn(map, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, const_cast<void*>(image), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// ---- uint16 images: per-tile max of band 1 decides the branch (prediction.py:167) ----------
__global__ void tile_band1_max_kernel(const unsigned short* __restrict__ image, int H, int W,
                                      const TileDesc* __restrict__ tiles, int* __restrict__ tile_max) {
  const TileDesc T = tiles[blockIdx.x];
  const unsigned short* b1 = image + (size_t)H * W;
  int m = 0;
  for (int i = threadIdx.x; i < T.w * T.h; i += blockDim.x) {
    const int r = i / T.w, c = i - r * T.w;
    m = max(m, (int)b1[(size_t)(T.r_off + r) * W + T.c_off + c]);
  }
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(&tile_max[blockIdx.x], m);
}

// float branch: v = (255 * x mod 2^16) / 65535 in float64, then torch's bilinear
// (align_corners=False): src = max((dst + 0.5) * in/out - 0.5, 0)
__global__ void tile_resize_u16_kernel(const unsigned short* __restrict__ image, int H, int W,
                                       const TileDesc* __restrict__ tiles, const int* __restrict__ tile_max,
                                       float* __restrict__ out, unsigned char* __restrict__ rescale16) {
  const TileDesc T = tiles[blockIdx.z];
  const int ox = blockIdx.x * blockDim.x + threadIdx.x;
  const int oy = blockIdx.y;
  const bool scaled = tile_max[blockIdx.z] > 255;
  if (ox == 0 && oy == 0 && rescale16) rescale16[blockIdx.z] = scaled ? 1 : 2;  // 2: the reference skips the tile
  if (!scaled || ox >= T.nw || oy >= T.nh) return;
  const double sy = (double)T.h / (double)T.nh, sx = (double)T.w / (double)T.nw;
  double fy = ((double)oy + 0.5) * sy - 0.5; if (fy < 0.0) fy = 0.0;
  double fx = ((double)ox + 0.5) * sx - 0.5; if (fx < 0.0) fx = 0.0;
  const int y0 = (int)fy, x0 = (int)fx;
  const int y1 = y0 + (y0 < T.h - 1 ? 1 : 0), x1 = x0 + (x0 < T.w - 1 ? 1 : 0);
  const double ly = fy - (double)y0, lx = fx - (double)x0;
  const double hy = 1.0 - ly, hx = 1.0 - lx;
  const size_t plane = (size_t)H * W, oplane = (size_t)T.nh * T.nw;
  float* o = out + T.out_off;
  for (int c = 0; c < 3; ++c) {
    const unsigned short* b = image + (size_t)(2 - c) * plane;
    auto px = [&](int yy, int xx) {
      const unsigned v = b[(size_t)(T.r_off + yy) * W + T.c_off + xx];
      return (double)((255u * v) & 0xffffu) / 65535.0;
    };
    const double top = hx * px(y0, x0) + lx * px(y0, x1);
    const double bot = hx * px(y1, x0) + lx * px(y1, x1);
    o[c * oplane + (size_t)oy * T.nw + ox] = (float)(hy * top + ly * bot);
  }
}

// ---- P0a -------------------------------------------------------------------------------
template <typename T>
__global__ void seam_crop_kernel(const T* __restrict__ a, const T* __restrict__ b, int bands, int ha, int wa, int hb,
                                 int wb, int axis, int left, int top, int strip_w, int strip_h, T* __restrict__ out) {
  const long long total = (long long)bands * strip_h * strip_w;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const int x = (int)(i % strip_w);
    const int y = (int)((i / strip_w) % strip_h);
    const int c = (int)(i / ((long long)strip_w * strip_h));
    const int mx = x + left, my = y + top;   // mosaic coordinates
    T v = 0;
    if (axis == 0) {
      if (mx < wa) { if (my < ha) v = a[((size_t)c * ha + my) * wa + mx]; }
      else if (mx - wa < wb && my < hb) v = b[((size_t)c * hb + my) * wb + (mx - wa)];
    } else {
      if (my < ha) { if (mx < wa) v = a[((size_t)c * ha + my) * wa + mx]; }
      else if (my - ha < hb && mx < wb) v = b[((size_t)c * hb + (my - ha)) * wb + mx];
    }
    out[i] = v;
  }
}

}  // namespace

// ---- plan / execute: the tile tables are uploaded once per tiling --------------------------
struct TilePlan {
  int n_tiles = 0, elem_size = 1, H = 0, W = 0;
  int max_nw = 0, max_nh = 0, max_rows = 1, max_cols = 1, max_k = 1, max_cnt = 0;
  long long n_blocks = 0;
  TileDesc* d_td = nullptr;
  int *d_min = nullptr, *d_cnt = nullptr, *d_k = nullptr, *d_max = nullptr;
  int2* d_blk = nullptr;   // per CTA: tile index, (block row << 16) | block column
  int4* d_y = nullptr;            // per table entry {min, k0, k1, count} (v6 kernel's y taps)
  StripItem* d_items = nullptr;   // per warp of the v6 kernel: tile, column strip, first row, rows
  int n_items = 0, n_items_vec = 0;   // the first n_items_vec items belong to tiles with 16-byte aligned rows
  // the (few) unaligned tiles run on a side stream next to the aligned ones
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
};

// tile_win (T,4) [col_off,row_off,w,h], tile_net (T,2) [net_h,net_w], out_off (T+1): HOST pointers
extern "C" int td_tile_plan_create(const int* tile_win, const int* tile_net, const long long* out_off, int n_tiles,
                                   int elem_size, int H, int W, void** plan_out) {
  TD_ARG(n_tiles > 0 && tile_win && tile_net && out_off && plan_out);
  TD_ARG(H > 0 && W > 0 && (elem_size == 1 || elem_size == 2));
  TilePlan* P = new TilePlan();
  P->n_tiles = n_tiles; P->elem_size = elem_size; P->H = H; P->W = W;
  std::vector<TileDesc> td(n_tiles);
  std::map<std::pair<int, int>, std::pair<int, int>> tabs;  // (in,out) -> (offset, ksize)
  std::vector<int> tmin, tcnt, tk;
  // group > 0: also track the widest source span of `group` consecutive outputs
  auto table = [&](int in_size, int out_size, int group, int& max_span) {
    auto key = std::make_pair(in_size, out_size);
    auto it = tabs.find(key);
    std::pair<int, int> res;
    if (it == tabs.end()) {
      std::vector<int> xm, cn, kk;
      std::vector<double> kd;
      int ks = 0;
      pil_coeffs(in_size, out_size, xm, cn, kk, kd, ks);
      res = std::make_pair((int)tmin.size(), ks);
      for (int i = 0; i < out_size; ++i) {
        tmin.push_back(xm[i]); tcnt.push_back(cn[i]);
        for (int q = 0; q < kMaxK; ++q) tk.push_back((q < ks && q < kMaxK) ? kk[(size_t)i * ks + q] : 0);
      }
      tabs[key] = res;
    } else {
      res = it->second;
    }
    for (int i = 0; i < out_size; i += group) {
      const int last = (i + group < out_size ? i + group : out_size) - 1;
      const int span = tmin[res.first + last] + tcnt[res.first + last] - tmin[res.first + i];
      if (span > max_span) max_span = span;
    }
    return res;
  };
  int rc = TD_OK;
  for (int t = 0; t < n_tiles && rc == TD_OK; ++t) {
    TileDesc& d = td[t];
    d.c_off = tile_win[4 * t]; d.r_off = tile_win[4 * t + 1]; d.w = tile_win[4 * t + 2]; d.h = tile_win[4 * t + 3];
    d.nh = tile_net[2 * t]; d.nw = tile_net[2 * t + 1];
    d.out_off = out_off[t];
    if (!(d.w > 0 && d.h > 0 && d.nh > 0 && d.nw > 0 && d.c_off >= 0 && d.r_off >= 0 && d.c_off + d.w <= W &&
          d.r_off + d.h <= H)) {
      td_set_error("td_tile_plan_create: tile %d window outside the raster", t);
      rc = TD_ERR_ARG;
      break;
    }
    d.xtab = d.ytab = d.kx = d.ky = 0;
    if (elem_size == 1) {
      auto tx = table(d.w, d.nw, kBX, P->max_cols);
      auto ty = table(d.h, d.nh, kBY, P->max_rows);
      if (tx.second > kMaxK || ty.second > kMaxK) {
        td_set_error("td_tile_plan_create: down-scaling factor needs %d taps (max %d)",
                     tx.second > ty.second ? tx.second : ty.second, kMaxK);
        rc = TD_ERR_UNSUPPORTED;
        break;
      }
      d.xtab = tx.first; d.kx = tx.second; d.ytab = ty.first; d.ky = ty.second;
      if (d.kx > P->max_k) P->max_k = d.kx;
      if (d.ky > P->max_k) P->max_k = d.ky;
    }
    d.bx = td_div_up(d.nw, kBX);
    d.blk0 = (int)P->n_blocks;
    P->n_blocks += (long long)d.bx * td_div_up(d.nh, kBY);
    if (P->n_blocks >= 0x7fffffffLL) { td_set_error("td_tile_plan_create: too many CTAs"); rc = TD_ERR_OVERFLOW; }
    if (d.nw > P->max_nw) P->max_nw = d.nw;
    if (d.nh > P->max_nh) P->max_nh = d.nh;
  }
  for (int v : tcnt) if (v > P->max_cnt) P->max_cnt = v;
  std::vector<int2> blk;
  if (rc == TD_OK && elem_size == 1) {
    blk.reserve((size_t)P->n_blocks);
    for (int t = 0; t < n_tiles; ++t) {
      const int by = td_div_up(td[t].nh, kBY);
      for (int j = 0; j < by; ++j)
        for (int i = 0; i < td[t].bx; ++i) blk.push_back(make_int2(t, (j << 16) | i));
    }
  }
  std::vector<StripItem> items;
  if (rc == TD_OK && elem_size == 1) {
    int rows_per_item = 20;
    if (const char* e = getenv("TREEDET_P1_ROWS")) rows_per_item = atoi(e) > 0 ? atoi(e) : rows_per_item;
    // tiles whose output rows are 16-byte aligned first (128-bit stores), the others after them
    for (int pass = 0; pass < 2; ++pass) {
      for (int t = 0; t < n_tiles; ++t) {
        const bool vec = ((td[t].nw & 3) == 0) && ((td[t].out_off & 3) == 0);
        if (vec != (pass == 0)) continue;
        const TileDesc& d = td[t];
        for (int r0 = 0; r0 < d.nh; r0 += rows_per_item)
          for (int i = 0; i < d.bx; ++i) {
            StripItem it{};
            const int ox0 = i * kBX;
            it.nrows = d.nh - r0 < rows_per_item ? d.nh - r0 : rows_per_item;
            const int col_lo = tmin[d.xtab + ox0];
            it.x_box = (d.c_off + col_lo) & ~15;
            it.xbias = (d.c_off + col_lo) - it.x_box - col_lo;
            it.row_lo = tmin[d.ytab + r0];
            it.y_box = d.r_off + it.row_lo;
            const int n_src = tmin[d.ytab + r0 + it.nrows - 1] + 2 - it.row_lo;   // rows ys and ys + 1 of every row
            it.n_chunks = (n_src + kChunk - 1) / kChunk;
            it.xtab = d.xtab + ox0;
            it.ytab = d.ytab + r0;
            it.out_off = d.out_off + (long long)r0 * d.nw + ox0;
            it.oplane = (long long)d.nh * d.nw;
            it.nw = d.nw;
            it.ncols = d.nw - ox0 < kBX ? d.nw - ox0 : kBX;
            items.push_back(it);
          }
      }
      if (pass == 0) P->n_items_vec = (int)items.size();
    }
    P->n_items = (int)items.size();
  }
  auto up = [&](void** dst, const void* src, size_t bytes) {
    if (cudaMalloc(dst, bytes ? bytes : 4) != cudaSuccess) return false;
    return bytes == 0 || cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice) == cudaSuccess;
  };
  if (rc == TD_OK) {
    bool ok = up((void**)&P->d_td, td.data(), sizeof(TileDesc) * n_tiles);
    if (elem_size == 1) {
      ok = ok && up((void**)&P->d_min, tmin.data(), sizeof(int) * tmin.size());
      ok = ok && up((void**)&P->d_cnt, tcnt.data(), sizeof(int) * tcnt.size());
      ok = ok && up((void**)&P->d_k, tk.data(), sizeof(int) * tk.size());
      ok = ok && up((void**)&P->d_blk, blk.data(), sizeof(int2) * blk.size());
      ok = ok && up((void**)&P->d_items, items.data(), sizeof(StripItem) * items.size());
      std::vector<int4> ty(tmin.size());
      for (size_t i = 0; i < tmin.size(); ++i) ty[i] = make_int4(tmin[i], tk[i * kMaxK], tk[i * kMaxK + 1], tcnt[i]);
      ok = ok && up((void**)&P->d_y, ty.data(), sizeof(int4) * ty.size());
    } else {
      ok = ok && (cudaMalloc((void**)&P->d_max, sizeof(int) * n_tiles) == cudaSuccess);
    }
    if (!ok) { td_set_error("td_tile_plan_create: %s", cudaGetErrorString(cudaGetLastError())); rc = TD_ERR_CUDA; }
  }
  if (rc != TD_OK) {
    cudaFree(P->d_td); cudaFree(P->d_min); cudaFree(P->d_cnt); cudaFree(P->d_k); cudaFree(P->d_max); cudaFree(P->d_blk); cudaFree(P->d_items); cudaFree(P->d_y);
    delete P;
    return rc;
  }
  if (P->n_items > P->n_items_vec && P->n_items_vec > 0) {
    if (cudaStreamCreateWithFlags(&P->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&P->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&P->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      P->side = nullptr;   // fall back to two launches on the caller's stream
    }
  }
  *plan_out = P;
  return TD_OK;
}

extern "C" int td_tile_plan_destroy(void* plan) {
  if (!plan) return TD_OK;
  TilePlan* P = (TilePlan*)plan;
  if (P->ev_fork) cudaEventDestroy(P->ev_fork);
  if (P->ev_join) cudaEventDestroy(P->ev_join);
  if (P->side) cudaStreamDestroy(P->side);
  cudaFree(P->d_td); cudaFree(P->d_min); cudaFree(P->d_cnt); cudaFree(P->d_k); cudaFree(P->d_max); cudaFree(P->d_blk); cudaFree(P->d_items); cudaFree(P->d_y);
  delete P;
  return TD_OK;
}

extern "C" int td_tile_cut_normalize(const void* plan, const void* image, int bands, float* out,
                                     unsigned char* rescale16, void* stream) {
  TD_ARG(plan && image && out && bands >= 3);
  TD_ARG(((uintptr_t)out & 15) == 0 && ((uintptr_t)image & 3) == 0);
  const TilePlan* P = (const TilePlan*)plan;
  cudaStream_t st = (cudaStream_t)stream;
  const int H = P->H, W = P->W, n_tiles = P->n_tiles;
  if (P->elem_size == 1) {
    if (rescale16) TD_CUDA(cudaMemsetAsync(rescale16, 0, n_tiles, st));
    const int K = P->max_k <= 3 ? 3 : kMaxK;
    const int src_stride = (P->max_cols + K + 8 + 15) & ~15;   // 16-byte multiple keeps every region aligned
    const int max_rows = P->max_rows;
    const size_t smem = (size_t)3 * max_rows * src_stride + (size_t)3 * max_rows * kBX + sizeof(int) * kBY * (2 + K);
    const long long image_bytes = (long long)bands * H * W;
    const unsigned grid = (unsigned)P->n_blocks;
    const unsigned char* img = (const unsigned char*)image;
    bool launched = false;
    if (P->max_cnt <= 2 && P->max_k <= 3 && (W % 16) == 0 && ((size_t)H * W) % 16 == 0 &&
        ((uintptr_t)image & 15) == 0 && max_rows <= 256 && !getenv("TREEDET_NO_TMA")) {
      // TMA-staged fast path (warp-autonomous strips); anything else takes the CTA kernels below
      const int box_w = (P->max_cols + 2 + 15 + 15) & ~15;   // + up to 15 bytes of start alignment
      CUtensorMap tmap;
      // v6 (warp-autonomous strips): staged row pitch 96 bytes (every 450 -> 800 tile) or 160 (any up-scaling)
      const int box_v6 = box_w <= 96 ? 96 : 160;
      if (box_w <= 160 && make_image_tensor_map(&tmap, image, bands, H, W, box_v6, kChunk, 3)) {
        const size_t smem_need = (size_t)kWarpsV6 * (2 * 3 * kChunk * box_v6 + 16 + sizeof(float) * kBX);
        // Residency: 4 CTAs (16 warps) per SM already saturate HBM (2.11 ms alone against 2.14 with 6), and
        // what they leave free -- a third of the registers, ~40 KB of shared memory -- lets the latency-
        // bound P2-P9 kernels of the other stream run NEXT to this kernel instead of waiting for it
        // (step 4.76 -> 4.5 ms).  There is no launch attribute for "at most n CTAs per SM"; asking for
        // 46 KB of dynamic shared memory per CTA does it (4 x 46 <= 227 < 5 x 46).  A persistent grid
        // of 4 CTAs per SM was tried instead and ran 30 % slower.
        size_t smem_v6 = smem_need;
        static int pad_kb = -1;
        if (pad_kb < 0) { const char* e = getenv("TREEDET_P1_SMEM_KB"); pad_kb = e ? atoi(e) : 46; }
        if ((size_t)pad_kb * 1024 > smem_v6) smem_v6 = (size_t)pad_kb * 1024;
        const bool fork = P->side && P->n_items_vec > 0 && P->n_items > P->n_items_vec;
        if (fork) {
          TD_CUDA(cudaEventRecord(P->ev_fork, st));
          TD_CUDA(cudaStreamWaitEvent(P->side, P->ev_fork, 0));
        }
        for (int pass = 1; pass >= 0; --pass) {
          const int first = pass == 0 ? 0 : P->n_items_vec;
          const int count = pass == 0 ? P->n_items_vec : P->n_items - P->n_items_vec;
          if (count == 0) continue;
          cudaStream_t st_pass = (fork && pass == 1) ? P->side : st;
          auto kern = box_v6 == 96 ? (pass == 0 ? tile_resize_u8_up_warp_kernel<true, 96> : tile_resize_u8_up_warp_kernel<false, 96>)
                                   : (pass == 0 ? tile_resize_u8_up_warp_kernel<true, 160> : tile_resize_u8_up_warp_kernel<false, 160>);
          const size_t smem_pass = pass == 0 ? smem_v6 : smem_need;      // the few unaligned tiles: no limit
          if (smem_pass > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_pass);
          kern<<<td_div_up(count, kWarpsV6), 32 * kWarpsV6, smem_pass, st_pass>>>(tmap, P->d_items + first, count, P->d_min,
                                                                                 P->d_k, P->d_y, out);
          if (fork && pass == 1) TD_CUDA(cudaEventRecord(P->ev_join, P->side));
        }
        if (fork) TD_CUDA(cudaStreamWaitEvent(st, P->ev_join, 0));
        launched = true;
      }
    }
    if (launched) {
    } else if (P->max_cnt <= 2 && P->max_k <= 3) {
      // up-scaling fast path; one extra intermediate row is readable (weight 0 taps)
      const size_t smem_up = (size_t)3 * max_rows * src_stride + (size_t)3 * max_rows * kBX + kBX + sizeof(int4) * kBY;
      auto kern = tile_resize_u8_up_kernel;
      if (smem_up > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_up);
      kern<<<grid, kThreads, smem_up, st>>>(img, image_bytes, H, W, P->d_td, P->d_blk, P->d_min, P->d_cnt, P->d_k, out,
                                            max_rows, src_stride);
    } else if (K == 3) {
      auto kern = tile_resize_u8_kernel<3>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, kThreads, smem, st>>>(img, image_bytes, H, W, P->d_td, n_tiles, P->d_min, P->d_cnt, P->d_k, out,
                                         max_rows, src_stride);
    } else {
      auto kern = tile_resize_u8_kernel<kMaxK>;
      if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      kern<<<grid, kThreads, smem, st>>>(img, image_bytes, H, W, P->d_td, n_tiles, P->d_min, P->d_cnt, P->d_k, out,
                                         max_rows, src_stride);
    }
  } else {
    TD_CUDA(cudaMemsetAsync(P->d_max, 0, sizeof(int) * n_tiles, st));
    tile_band1_max_kernel<<<n_tiles, 256, 0, st>>>((const unsigned short*)image, H, W, P->d_td, P->d_max);
    dim3 grid(td_div_up(P->max_nw, 128), P->max_nh, n_tiles);
    tile_resize_u16_kernel<<<grid, 128, 0, st>>>((const unsigned short*)image, H, W, P->d_td, P->d_max, out, rescale16);
  }
  TD_CHECK_LAUNCH("td_tile_cut_normalize");
  return TD_OK;
}

extern "C" int td_seam_crop(const void* a, const void* b, int elem_size, int bands, int ha, int wa, int hb, int wb,
                            int axis, int strip_w, int strip_h, void* out, void* stream) {
  TD_ARG(a && b && out && bands > 0 && ha > 0 && wa > 0 && hb > 0 && wb > 0 && strip_w > 0 && strip_h > 0);
  TD_ARG(axis == 0 || axis == 1);
  TD_ARG(elem_size == 1 || elem_size == 2 || elem_size == 4);
  // crop_image (helpers.py:1053-1085): centre window of the mosaic
  const int mw = axis == 0 ? wa + wb : (wa > wb ? wa : wb);
  const int mh = axis == 0 ? (ha > hb ? ha : hb) : ha + hb;
  int left = mw / 2 - strip_w / 2; if (left < 0) left = 0;
  int top = mh / 2 - strip_h / 2; if (top < 0) top = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const int blocks = td_num_sms() * 8;
  if (elem_size == 1)
    seam_crop_kernel<unsigned char><<<blocks, 256, 0, st>>>((const unsigned char*)a, (const unsigned char*)b, bands, ha,
                                                            wa, hb, wb, axis, left, top, strip_w, strip_h,
                                                            (unsigned char*)out);
  else if (elem_size == 2)
    seam_crop_kernel<unsigned short><<<blocks, 256, 0, st>>>((const unsigned short*)a, (const unsigned short*)b, bands,
                                                             ha, wa, hb, wb, axis, left, top, strip_w, strip_h,
                                                             (unsigned short*)out);
  else
    seam_crop_kernel<unsigned int><<<blocks, 256, 0, st>>>((const unsigned int*)a, (const unsigned int*)b, bands, ha, wa,
                                                           hb, wb, axis, left, top, strip_w, strip_h,
                                                           (unsigned int*)out);
  TD_CHECK_LAUNCH("td_seam_crop");
  return TD_OK;
}
