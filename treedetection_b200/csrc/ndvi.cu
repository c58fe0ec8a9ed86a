// P5 -- decimated raster reads + NDVI.
//
// Replaces the raster part of process_geojson (TreeDetection/postprocessing.py:780-800:
// rasterio read with out_shape + Resampling.bilinear for the nDSM band and the RGBI
// bands) and ndvi_array_from_rgbi / ndvi_index (TreeDetection/helpers.py:862-896).
//
// Decimation is the separable triangle-filter convolution GDAL applies for a
// down-sampling bilinear RasterIO (support = scale, centre = (i + 0.5) * scale,
// weights normalised per output pixel).  GDAL is an un-pinned third-party dependency
// of the reference and absent here, so oracle/port.py decimate_bilinear *defines* the
// arithmetic (parity unpinned at that boundary): float32 weights, float32
// accumulation in tap order, horizontal pass then vertical pass, uint8 results
// rounded half up.  The two passes are fused per CTA tile through shared memory, so
// the full-resolution bands are read once and no intermediate raster touches HBM.
//
// NDVI: (nir/255 - red/255) / (nir/255 + red/255 + 1e-10) in float64 on bands 0 and 3,
// stored as float32 -- the reference's float64 array is only ever consumed through
// cp.array(ndvi_data, dtype=float32) (postprocessing.py:543); the float32 value is
// identical for all 65 536 uint8 input pairs (tests/test_oracle_vs_reference.py).
#include <cstdint>
#include <cstdlib>
#include <map>
#include <mutex>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kTileW = 32;   // output columns per CTA
constexpr int kTileH = 8;    // output rows per CTA
constexpr int kMaxTaps = 64; // per-axis taps supported by the fused kernel (scale <= ~31)

struct AxisTable {
  const int* start;    // first source index per output index
  const int* count;    // taps per output index
  const float* w;      // (n_out, ktaps) weights
  int ktaps;
};

// host: PIL/GDAL-style coefficient table for one axis (float64 math, float32 weights)
void build_axis(int in_size, int out_size, std::vector<int>& start, std::vector<int>& count, std::vector<float>& w,
                int& ktaps) {
  const double scale = (double)in_size / (double)out_size;
  const double fscale = scale < 1.0 ? 1.0 : scale;
  const double support = 1.0 * fscale;
  ktaps = (int)ceil(support) * 2 + 1;
  start.assign(out_size, 0);
  count.assign(out_size, 0);
  w.assign((size_t)out_size * ktaps, 0.f);
  std::vector<double> tmp(ktaps);
  for (int i = 0; i < out_size; ++i) {
    const double center = (i + 0.5) * scale;
    int xmin = (int)(center - support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + support + 0.5);
    if (xmax > in_size) xmax = in_size;
    const int n = xmax - xmin;
    double tot = 0.0;
    for (int x = 0; x < n; ++x) {
      double t = ((double)(x + xmin) - center + 0.5) / fscale;
      if (t < 0) t = -t;
      const double v = t < 1.0 ? 1.0 - t : 0.0;
      tmp[x] = v;
      tot += v;
    }
    for (int x = 0; x < n; ++x) w[(size_t)i * ktaps + x] = (float)(tot != 0.0 ? tmp[x] / tot : tmp[x]);
    start[i] = xmin;
    count[i] = n;
  }
}

template <typename T>
TD_D float load_as_float(const T* p) { return (float)(*p); }

// Fused separable decimation of NB bands of a (bands, H, W) raster.  Each CTA makes a
// kTileW x kTileH block of output pixels: the horizontal pass of every needed source
// row goes to shared memory, the vertical pass reads it back.
// EPI = 0: float32 out (one band);  EPI = 1: uint8-rounded red & nir -> NDVI float32
template <typename T, int EPI>
__global__ void __launch_bounds__(kTileW* kTileH)
decimate_kernel(const T* __restrict__ src0, const T* __restrict__ src1, int in_h, int in_w, int out_h, int out_w,
                AxisTable ax, AxisTable ay, float* __restrict__ out, int max_src_rows) {
  extern __shared__ float sh[];  // [NB][max_src_rows][kTileW]
  constexpr int NB = EPI == 1 ? 2 : 1;
  const int tx = threadIdx.x % kTileW, ty = threadIdx.x / kTileW;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kTileH;
  const int oy_last = min(oy0 + kTileH, out_h) - 1;
  const int row_lo = ay.start[oy0];
  const int row_hi = ay.start[oy_last] + ay.count[oy_last];  // exclusive
  const int nrows = row_hi - row_lo;
  const int ox = ox0 + tx;
  // ---- horizontal pass ------------------------------------------------------
  if (ox < out_w) {
    const int xs = ax.start[ox], xn = ax.count[ox];
    const float* wx = ax.w + (size_t)ox * ax.ktaps;
    for (int r = ty; r < nrows; r += kTileH) {
      const size_t base = (size_t)(row_lo + r) * in_w + xs;
      float acc0 = 0.f, acc1 = 0.f;
      for (int k = 0; k < xn; ++k) {
        const float wk = wx[k];
        acc0 = __fadd_rn(acc0, __fmul_rn(load_as_float(src0 + base + k), wk));
        if (NB == 2) acc1 = __fadd_rn(acc1, __fmul_rn(load_as_float(src1 + base + k), wk));
      }
      sh[(size_t)r * kTileW + tx] = acc0;
      if (NB == 2) sh[(size_t)(max_src_rows + r) * kTileW + tx] = acc1;
    }
  }
  __syncthreads();
  // ---- vertical pass --------------------------------------------------------
  const int oy = oy0 + ty;
  if (ox >= out_w || oy >= out_h) return;
  const int ys = ay.start[oy] - row_lo, yn = ay.count[oy];
  const float* wy = ay.w + (size_t)oy * ay.ktaps;
  float acc0 = 0.f, acc1 = 0.f;
  for (int k = 0; k < yn; ++k) {
    const float wk = wy[k];
    acc0 = __fadd_rn(acc0, __fmul_rn(sh[(size_t)(ys + k) * kTileW + tx], wk));
    if (NB == 2) acc1 = __fadd_rn(acc1, __fmul_rn(sh[(size_t)(max_src_rows + ys + k) * kTileW + tx], wk));
  }
  if (EPI == 0) {
    out[(size_t)oy * out_w + ox] = acc0;
  } else {
    // uint8 band: round half up, clamp; then the NDVI of the two rounded bands
    const double red = (double)fminf(fmaxf(floorf(__fadd_rn(acc0, 0.5f)), 0.f), 255.f) / 255.0;
    const double nir = (double)fminf(fmaxf(floorf(__fadd_rn(acc1, 0.5f)), 0.f), 255.f) / 255.0;
    out[(size_t)oy * out_w + ox] = (float)((nir - red) / (nir + red + 1e-10));
  }
}

// The uint8 / NDVI case of decimate_kernel for <= KT taps per output column (scale <= ~5.5), which is what
// P5 runs on every image (bands 0 and 3 at full resolution -> NDVI at ndvi_scaling_factor).  Same
// arithmetic in the same order; what changes is the instruction count per tap (the generic kernel is
// issue bound at ~9 instructions per tap and band):
//   * a thread keeps the KT weights of its output column in registers and the tap loop is unrolled
//     (taps past the column's count carry weight 0: acc + x * 0 = acc exactly for the uint8 samples);
//   * 16 output rows per CTA instead of 8: the horizontal pass of the vertical halo rows (2 x scale rows
//     per tile) is redone by every tile, 6 % of the rows instead of 25 % at scale 5;
//   * the epilogue's two divisions by 255 come from a 256-entry table built by the CTA (same quotients).
constexpr int kFastTileH = 16;
template <int KT>
__global__ void __launch_bounds__(kTileW* kFastTileH)
decimate_ndvi_fast_kernel(const unsigned char* __restrict__ src0, const unsigned char* __restrict__ src1, int in_h,
                          int in_w, int out_h, int out_w, AxisTable ax, AxisTable ay, float* __restrict__ out,
                          int max_src_rows) {
  extern __shared__ float sh[];  // [2][max_src_rows][kTileW]
  __shared__ double s_unit[256];
  const int tx = threadIdx.x % kTileW, ty = threadIdx.x / kTileW;
  if (threadIdx.x < 256) s_unit[threadIdx.x] = (double)(float)threadIdx.x / 255.0;
  const int ox0 = blockIdx.x * kTileW, oy0 = blockIdx.y * kFastTileH;
  const int oy_last = min(oy0 + kFastTileH, out_h) - 1;
  const int row_lo = ay.start[oy0];
  const int nrows = ay.start[oy_last] + ay.count[oy_last] - row_lo;
  const int ox = ox0 + tx;
  // every lane may read the four aligned words around the KT bytes from its first tap (up to KT + 4 bytes past
  // it) without leaving its row: true unless the tile touches the right edge
  const int ox_last = min(ox0 + kTileW, out_w) - 1;
  const bool wide = ax.start[ox_last] + KT + 4 <= in_w;
  if (ox < out_w) {
    const int xs = ax.start[ox], xn = ax.count[ox];
    const float* wx = ax.w + (size_t)ox * ax.ktaps;
    if (wide) {
      float w[KT];
#pragma unroll
      for (int k = 0; k < KT; ++k) w[k] = k < xn ? wx[k < ax.ktaps ? k : 0] : 0.f;
      // The KT bytes of a row are fetched as the four aligned 32-bit words that hold them (a byte load per
      // tap made the kernel load-issue bound) and shifted into place; `wide` keeps all four inside the row (the
      // first word may start up to 3 bytes before the first tap: inside the row or the previous row / band).
      static_assert(KT == 12, "three realigned words");
#pragma unroll 1
      for (int r = ty; r < nrows; r += kFastTileH) {
        const size_t o = (size_t)(row_lo + r) * in_w + xs;
        float acc[2];
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const unsigned char* p = (b == 0 ? src0 : src1) + o;
          const unsigned sh8 = ((unsigned)(uintptr_t)p & 3u) * 8u;
          const unsigned* q = reinterpret_cast<const unsigned*>((uintptr_t)p & ~(uintptr_t)3);
          const unsigned q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2), q3 = __ldg(q + 3);
          const unsigned u[3] = {__funnelshift_r(q0, q1, sh8), __funnelshift_r(q1, q2, sh8),
                                 __funnelshift_r(q2, q3, sh8)};
          float a = 0.f;
#pragma unroll
          for (int k = 0; k < KT; ++k) a = __fadd_rn(a, __fmul_rn((float)((u[k >> 2] >> (8 * (k & 3))) & 0xffu), w[k]));
          acc[b] = a;
        }
        sh[(size_t)r * kTileW + tx] = acc[0];
        sh[(size_t)(max_src_rows + r) * kTileW + tx] = acc[1];
      }
    } else {
      for (int r = ty; r < nrows; r += kFastTileH) {
        const size_t base = (size_t)(row_lo + r) * in_w + xs;
        float acc0 = 0.f, acc1 = 0.f;
        for (int k = 0; k < xn; ++k) {
          const float wk = wx[k];
          acc0 = __fadd_rn(acc0, __fmul_rn((float)src0[base + k], wk));
          acc1 = __fadd_rn(acc1, __fmul_rn((float)src1[base + k], wk));
        }
        sh[(size_t)r * kTileW + tx] = acc0;
        sh[(size_t)(max_src_rows + r) * kTileW + tx] = acc1;
      }
    }
  }
  __syncthreads();
  const int oy = oy0 + ty;
  if (ox >= out_w || oy >= out_h) return;
  const int ys = ay.start[oy] - row_lo, yn = ay.count[oy];
  const float* wy = ay.w + (size_t)oy * ay.ktaps;
  const float* c0 = sh + (size_t)ys * kTileW + tx;
  const float* c1 = c0 + (size_t)max_src_rows * kTileW;
  float acc0 = 0.f, acc1 = 0.f;
  for (int k = 0; k < yn; ++k) {
    const float wk = wy[k];
    acc0 = __fadd_rn(acc0, __fmul_rn(c0[k * kTileW], wk));
    acc1 = __fadd_rn(acc1, __fmul_rn(c1[k * kTileW], wk));
  }
  // uint8 band: round half up, clamp (the value is an integer in 0..255); then the NDVI of the two bands
  const double red = s_unit[(int)fminf(fmaxf(floorf(__fadd_rn(acc0, 0.5f)), 0.f), 255.f)];
  const double nir = s_unit[(int)fminf(fmaxf(floorf(__fadd_rn(acc1, 0.5f)), 0.f), 255.f)];
  out[(size_t)oy * out_w + ox] = (float)((nir - red) / (nir + red + 1e-10));
}

// no decimation (scale factor 1): straight NDVI of the full-resolution bands
__global__ void ndvi_full_kernel(const unsigned char* __restrict__ red, const unsigned char* __restrict__ nir,
                                 long long n, float* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double r = (double)red[i] / 255.0, q = (double)nir[i] / 255.0;
    out[i] = (float)((q - r) / (q + r + 1e-10));
  }
}

struct DevAxis {
  int* start = nullptr;
  int* count = nullptr;
  float* w = nullptr;
  int ktaps = 0;
  int max_rows = 0;  // max source span of `group` consecutive outputs
  int max_count = 0;
};

// Axis tables depend on (in, out, group) only: built and uploaded once, kept for the life of the
// process (a few hundred KB), so a call launches its kernel and nothing else.
struct AxisKey {
  int dev, in_size, out_size, group;
  bool operator<(const AxisKey& o) const {
    if (dev != o.dev) return dev < o.dev;
    if (in_size != o.in_size) return in_size < o.in_size;
    if (out_size != o.out_size) return out_size < o.out_size;
    return group < o.group;
  }
};

int get_axis(int in_size, int out_size, int group, DevAxis& d) {
  static std::mutex mu;
  static std::map<AxisKey, DevAxis> cache;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  const AxisKey key{dev, in_size, out_size, group};
  auto it = cache.find(key);
  if (it != cache.end()) { d = it->second; return TD_OK; }
  std::vector<int> s, c;
  std::vector<float> w;
  build_axis(in_size, out_size, s, c, w, d.ktaps);
  d.max_rows = 0;
  d.max_count = 0;
  for (int i = 0; i < out_size; ++i)
    if (c[i] > d.max_count) d.max_count = c[i];
  for (int i = 0; i < out_size; i += group) {
    const int last = (i + group < out_size ? i + group : out_size) - 1;
    const int span = s[last] + c[last] - s[i];
    if (span > d.max_rows) d.max_rows = span;
  }
  TD_CUDA(cudaMalloc((void**)&d.start, sizeof(int) * out_size));
  TD_CUDA(cudaMalloc((void**)&d.count, sizeof(int) * out_size));
  TD_CUDA(cudaMalloc((void**)&d.w, sizeof(float) * w.size()));
  TD_CUDA(cudaMemcpy(d.start, s.data(), sizeof(int) * out_size, cudaMemcpyHostToDevice));
  TD_CUDA(cudaMemcpy(d.count, c.data(), sizeof(int) * out_size, cudaMemcpyHostToDevice));
  TD_CUDA(cudaMemcpy(d.w, w.data(), sizeof(float) * w.size(), cudaMemcpyHostToDevice));
  cache[key] = d;
  return TD_OK;
}

template <typename T, int EPI>
int run_decimate(const T* s0, const T* s1, int in_h, int in_w, int out_h, int out_w, float* out, cudaStream_t st) {
  DevAxis dx, dy;
  int rc = get_axis(in_w, out_w, kTileW, dx);
  if (rc != TD_OK) return rc;
  rc = get_axis(in_h, out_h, kTileH, dy);
  if (rc != TD_OK) return rc;
  if (dx.ktaps > kMaxTaps || dy.ktaps > kMaxTaps) {
    td_set_error("decimation factor too large for the fused kernel (taps %d x %d)", dx.ktaps, dy.ktaps);
    return TD_ERR_UNSUPPORTED;
  }
  if constexpr (EPI == 1) {
    constexpr int KT = 12;
    static int generic = -1;
    if (generic < 0) { const char* e = getenv("TREEDET_DECIMATE_GENERIC"); generic = e && atoi(e) > 0 ? 1 : 0; }
    DevAxis dyf;
    if (!generic && dx.max_count <= KT && get_axis(in_h, out_h, kFastTileH, dyf) == TD_OK) {
      const size_t smem_f = sizeof(float) * 2 * (size_t)dyf.max_rows * kTileW;
      if (smem_f <= 160 * 1024) {
        auto kf = decimate_ndvi_fast_kernel<KT>;
        if (smem_f > 40 * 1024) cudaFuncSetAttribute(kf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_f);
        AxisTable fx{dx.start, dx.count, dx.w, dx.ktaps}, fy{dyf.start, dyf.count, dyf.w, dyf.ktaps};
        dim3 gridf(td_div_up(out_w, kTileW), td_div_up(out_h, kFastTileH));
        kf<<<gridf, kTileW * kFastTileH, smem_f, st>>>((const unsigned char*)s0, (const unsigned char*)s1, in_h, in_w,
                                                        out_h, out_w, fx, fy, out, dyf.max_rows);
        TD_CHECK_LAUNCH("decimate");
        return TD_OK;
      }
    }
  }
  constexpr int NB = EPI == 1 ? 2 : 1;
  const size_t smem = sizeof(float) * NB * (size_t)dy.max_rows * kTileW;
  auto kern = decimate_kernel<T, EPI>;
  if (smem > 48 * 1024) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  AxisTable ax{dx.start, dx.count, dx.w, dx.ktaps}, ay{dy.start, dy.count, dy.w, dy.ktaps};
  dim3 grid(td_div_up(out_w, kTileW), td_div_up(out_h, kTileH));
  kern<<<grid, kTileW * kTileH, smem, st>>>(s0, s1, in_h, in_w, out_h, out_w, ax, ay, out, dy.max_rows);
  TD_CHECK_LAUNCH("decimate");
  return TD_OK;
}

}  // namespace

// rgbi: (bands >= 4, H, W) uint8 planar on the device.  out: (out_h, out_w) float32 NDVI.
extern "C" int td_ndvi_decimate(const unsigned char* rgbi, int bands, int in_h, int in_w, int out_h, int out_w,
                                float* ndvi_out, void* stream) {
  TD_ARG(rgbi && ndvi_out && bands >= 4 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0);
  cudaStream_t st = (cudaStream_t)stream;
  const unsigned char* red = rgbi;
  const unsigned char* nir = rgbi + (size_t)3 * in_h * in_w;
  if (out_h == in_h && out_w == in_w) {
    const long long n = (long long)in_h * in_w;
    ndvi_full_kernel<<<td_num_sms() * 8, 256, 0, st>>>(red, nir, n, ndvi_out);
    TD_CHECK_LAUNCH("td_ndvi_decimate");
    return TD_OK;
  }
  return run_decimate<unsigned char, 1>(red, nir, in_h, in_w, out_h, out_w, ndvi_out, st);
}

// one float32 band (the nDSM): (in_h, in_w) -> (out_h, out_w)
extern "C" int td_decimate_f32(const float* src, int in_h, int in_w, int out_h, int out_w, float* out,
                               void* stream) {
  TD_ARG(src && out && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0);
  cudaStream_t st = (cudaStream_t)stream;
  if (out_h == in_h && out_w == in_w) {
    TD_CUDA(cudaMemcpyAsync(out, src, sizeof(float) * (size_t)in_h * in_w, cudaMemcpyDeviceToDevice, st));
    return TD_OK;
  }
  return run_decimate<float, 0>(src, src, in_h, in_w, out_h, out_w, out, st);
}
