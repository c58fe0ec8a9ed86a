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
// here the image is resident in HBM, every CTA produces a 128 x 8 block of output pixels
// for the three channels: the horizontal pass of the few source rows it needs goes to
// shared memory, the vertical pass streams coalesced float32 rows out.  The kernel is
// write bound: 0.81 MB in / 7.68 MB out per interior tile.
//
// P0a replaces crop_single_image / merge_images / crop_image (TreeDetection/merging.py:34-110,
// TreeDetection/helpers.py:1023-1085): the strip is gathered straight from the two source
// rasters; the 2-image mosaic (800 MB at 10k x 10k) is never built.
#include <map>
#include <utility>
#include <vector>

#include "common.cuh"

namespace {

constexpr int kBX = 128;            // output columns per CTA
constexpr int kBY = 8;              // output rows per CTA
constexpr int kThreads = 256;
constexpr int kPrecisionBits = 32 - 8 - 2;   // PIL: 22-bit fixed-point coefficients
constexpr int kMaxK = 8;            // taps per axis supported (down-scaling up to ~3.5x)

struct TileDesc {
  int c_off, r_off, w, h;   // source window
  int nh, nw;               // output size
  long long out_off;        // float offset of the (3, nh, nw) block
  int xtab, ytab;           // offsets into the coefficient tables (entries)
  int kx, ky;               // taps per output index
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
__global__ void __launch_bounds__(kThreads)
tile_resize_u8_kernel(const unsigned char* __restrict__ image, int H, int W, const TileDesc* __restrict__ tiles,
                      const int* __restrict__ tab_min, const int* __restrict__ tab_cnt,
                      const int* __restrict__ tab_k, float* __restrict__ out, int max_rows) {
  extern __shared__ unsigned char tmp[];  // [3][max_rows][kBX] horizontal pass, rounded to uint8
  const TileDesc T = tiles[blockIdx.z];
  const int ox0 = blockIdx.x * kBX, oy0 = blockIdx.y * kBY;
  if (ox0 >= T.nw || oy0 >= T.nh) return;
  const int oy1 = min(oy0 + kBY, T.nh) - 1;
  const int row_lo = tab_min[T.ytab + oy0];
  const int row_hi = tab_min[T.ytab + oy1] + tab_cnt[T.ytab + oy1];
  const int nrows = row_hi - row_lo;
  const size_t plane = (size_t)H * W;
  // horizontal pass: items = (channel, source row, output column)
  const int items = 3 * nrows * kBX;
  for (int it = threadIdx.x; it < items; it += kThreads) {
    const int x = it % kBX;
    const int r = (it / kBX) % nrows;
    const int c = it / (kBX * nrows);
    const int ox = ox0 + x;
    int v = 0;
    if (ox < T.nw) {
      const int xs = tab_min[T.xtab + ox], xn = tab_cnt[T.xtab + ox];
      const int* k = tab_k + (size_t)(T.xtab + ox) * kMaxK;
      // output channel c reads band 2 - c (BGR order)
      const unsigned char* src = image + (size_t)(2 - c) * plane + (size_t)(T.r_off + row_lo + r) * W + T.c_off + xs;
      int ss = 1 << (kPrecisionBits - 1);
      for (int q = 0; q < xn; ++q) ss += (int)src[q] * k[q];
      v = clip8(ss);
    }
    tmp[((size_t)c * max_rows + r) * kBX + x] = (unsigned char)v;
  }
  __syncthreads();
  // vertical pass: thread -> column x, rows y = ty, ty + 2, ...
  const int x = threadIdx.x % kBX, ty = threadIdx.x / kBX;
  const int ox = ox0 + x;
  if (ox >= T.nw) return;
  float* o = out + T.out_off;
  const size_t oplane = (size_t)T.nh * T.nw;
  for (int y = ty; y < kBY; y += kThreads / kBX) {
    const int oy = oy0 + y;
    if (oy >= T.nh) break;
    const int ys = tab_min[T.ytab + oy] - row_lo, yn = tab_cnt[T.ytab + oy];
    const int* k = tab_k + (size_t)(T.ytab + oy) * kMaxK;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      int ss = 1 << (kPrecisionBits - 1);
      for (int q = 0; q < yn; ++q) ss += (int)tmp[((size_t)c * max_rows + ys + q) * kBX + x] * k[q];
      o[c * oplane + (size_t)oy * T.nw + ox] = (float)clip8(ss);
    }
  }
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

// tile_win (T,4) [col_off,row_off,w,h], tile_net (T,2) [net_h,net_w], out_off (T+1): HOST pointers
extern "C" int td_tile_cut_normalize(const void* image, int elem_size, int bands, int H, int W, const int* tile_win,
                                     const int* tile_net, int n_tiles, const long long* out_off, float* out,
                                     unsigned char* rescale16, void* stream) {
  TD_ARG(n_tiles >= 0);
  if (n_tiles == 0) return TD_OK;
  TD_ARG(image && tile_win && tile_net && out_off && out);
  TD_ARG(bands >= 3 && H > 0 && W > 0 && (elem_size == 1 || elem_size == 2));
  cudaStream_t st = (cudaStream_t)stream;
  std::vector<TileDesc> td(n_tiles);
  std::map<std::pair<int, int>, std::pair<int, int>> tabs;  // (in,out) -> (offset, ksize)
  std::vector<int> tmin, tcnt, tk;
  int max_nw = 0, max_nh = 0, max_rows = 1;
  auto table = [&](int in_size, int out_size, int group) {
    auto key = std::make_pair(in_size, out_size);
    auto it = tabs.find(key);
    std::pair<int, int> res;
    std::vector<int> xm, cn, kk;
    std::vector<double> kd;
    int ks = 0;
    if (it == tabs.end()) {
      pil_coeffs(in_size, out_size, xm, cn, kk, kd, ks);
      res = std::make_pair((int)tmin.size(), ks);
      if (ks <= kMaxK) {
        for (int i = 0; i < out_size; ++i) {
          tmin.push_back(xm[i]); tcnt.push_back(cn[i]);
          for (int q = 0; q < kMaxK; ++q) tk.push_back(q < ks ? kk[(size_t)i * ks + q] : 0);
        }
      }
      tabs[key] = res;
    } else {
      res = it->second;
    }
    if (group > 0 && res.second <= kMaxK) {  // rows of shared memory the vertical pass needs
      for (int i = 0; i < out_size; i += group) {
        const int last = (i + group < out_size ? i + group : out_size) - 1;
        const int span = tmin[res.first + last] + tcnt[res.first + last] - tmin[res.first + i];
        if (span > max_rows) max_rows = span;
      }
    }
    return res;
  };
  for (int t = 0; t < n_tiles; ++t) {
    TileDesc& d = td[t];
    d.c_off = tile_win[4 * t]; d.r_off = tile_win[4 * t + 1]; d.w = tile_win[4 * t + 2]; d.h = tile_win[4 * t + 3];
    d.nh = tile_net[2 * t]; d.nw = tile_net[2 * t + 1];
    d.out_off = out_off[t];
    TD_ARG(d.w > 0 && d.h > 0 && d.nh > 0 && d.nw > 0 && d.c_off >= 0 && d.r_off >= 0 && d.c_off + d.w <= W &&
           d.r_off + d.h <= H);
    if (elem_size == 1) {
      auto tx = table(d.w, d.nw, 0);
      auto ty = table(d.h, d.nh, kBY);
      if (tx.second > kMaxK || ty.second > kMaxK) {
        td_set_error("td_tile_cut_normalize: down-scaling factor needs %d taps (max %d)", tx.second, kMaxK);
        return TD_ERR_UNSUPPORTED;
      }
      d.xtab = tx.first; d.kx = tx.second; d.ytab = ty.first; d.ky = ty.second;
    } else {
      d.xtab = d.ytab = d.kx = d.ky = 0;
    }
    if (d.nw > max_nw) max_nw = d.nw;
    if (d.nh > max_nh) max_nh = d.nh;
  }
  TileDesc* d_td = nullptr;
  int *d_min = nullptr, *d_cnt = nullptr, *d_k = nullptr, *d_max = nullptr;
  TD_CUDA(cudaMallocAsync((void**)&d_td, sizeof(TileDesc) * n_tiles, st));
  TD_CUDA(cudaMemcpyAsync(d_td, td.data(), sizeof(TileDesc) * n_tiles, cudaMemcpyHostToDevice, st));
  int rc = TD_OK;
  if (elem_size == 1) {
    TD_CUDA(cudaMallocAsync((void**)&d_min, sizeof(int) * (tmin.size() + 1), st));
    TD_CUDA(cudaMallocAsync((void**)&d_cnt, sizeof(int) * (tcnt.size() + 1), st));
    TD_CUDA(cudaMallocAsync((void**)&d_k, sizeof(int) * (tk.size() + 1), st));
    TD_CUDA(cudaMemcpyAsync(d_min, tmin.data(), sizeof(int) * tmin.size(), cudaMemcpyHostToDevice, st));
    TD_CUDA(cudaMemcpyAsync(d_cnt, tcnt.data(), sizeof(int) * tcnt.size(), cudaMemcpyHostToDevice, st));
    TD_CUDA(cudaMemcpyAsync(d_k, tk.data(), sizeof(int) * tk.size(), cudaMemcpyHostToDevice, st));
    if (rescale16) TD_CUDA(cudaMemsetAsync(rescale16, 0, n_tiles, st));
    const size_t smem = (size_t)3 * max_rows * kBX;
    if (smem > 48 * 1024)
      cudaFuncSetAttribute(tile_resize_u8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    dim3 grid(td_div_up(max_nw, kBX), td_div_up(max_nh, kBY), n_tiles);
    tile_resize_u8_kernel<<<grid, kThreads, smem, st>>>((const unsigned char*)image, H, W, d_td, d_min, d_cnt, d_k, out,
                                                        max_rows);
  } else {
    TD_CUDA(cudaMallocAsync((void**)&d_max, sizeof(int) * n_tiles, st));
    TD_CUDA(cudaMemsetAsync(d_max, 0, sizeof(int) * n_tiles, st));
    tile_band1_max_kernel<<<n_tiles, 256, 0, st>>>((const unsigned short*)image, H, W, d_td, d_max);
    dim3 grid(td_div_up(max_nw, 128), max_nh, n_tiles);
    tile_resize_u16_kernel<<<grid, 128, 0, st>>>((const unsigned short*)image, H, W, d_td, d_max, out, rescale16);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { td_set_error("td_tile_cut_normalize: %s", cudaGetErrorString(e)); rc = TD_ERR_CUDA; }
  // host vectors were staged by the copies above; device tables die in stream order
  cudaFreeAsync(d_td, st);
  if (d_min) cudaFreeAsync(d_min, st);
  if (d_cnt) cudaFreeAsync(d_cnt, st);
  if (d_k) cudaFreeAsync(d_k, st);
  if (d_max) cudaFreeAsync(d_max, st);
  return rc;
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
