// P9 -- pre-selection (border / overlap-band / height / NDVI rules), the containment
// case analysis and coordinate rounding of process_features
// (TreeDetection/postprocessing.py:571-720), element_is_near_border
// (TreeDetection/helpers.py:501-522) and round_coordinates (utilities.py:146-161).
//
// The reference walks Python lists with O(N) `features.index(...)` lookups per crown;
// the decisions themselves are O(1) per crown once three global facts are known: the
// crown's rank among the pre-selected ones, and the index of the globally FIRST
// contained crown.  Quirks that decide the crown set are reproduced verbatim
// (SURVEY.md §8a-P9): "contains two" is always dropped; "contains one" compares
// against the globally first contained crown, reads var_ndvi at the crown's rank in
// the pre-selected list, may emit the OTHER crown (duplicates possible), and its area
// fallback compares against 0.
#include <cub/device/device_scan.cuh>

#include "chain_internal.cuh"
#include "common.cuh"

namespace {

typedef TdSelectParams SelectParams;

__global__ void preselect_kernel(const double* __restrict__ bounds, const float* __restrict__ max_h,
                                 const float* __restrict__ ndvi_stats, int n, SelectParams P,
                                 const SelectParams* __restrict__ P_dev,
                                 int* __restrict__ pre, int* __restrict__ first_contained,
                                 const unsigned char* __restrict__ is_contained,
                                 const long long* __restrict__ n_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (P_dev) P = *P_dev;     // per-image parameters of the chain (device memory)
  if (n_dev && i >= *n_dev) { pre[i] = 0; return; }   // capacity tail: keeps the rank scan exact
  bool keep = true;
  const double bx0 = bounds[4 * i], by0 = bounds[4 * i + 1], bx1 = bounds[4 * i + 2], by1 = bounds[4 * i + 3];
  if (P.use_overlap) {
    const double eps = 1.0;
    if (bx0 < P.left + eps || bx1 > P.right - eps || by0 < P.bottom + eps || by1 > P.top - eps) keep = false;
    if (keep && !P.is_seam_image) {
      const bool in_top = P.band_top < by0;
      const bool in_bottom = P.band_bottom > by1;
      const bool in_left = P.band_left > bx1;
      const bool in_right = P.band_right < bx0;
      if (in_top || in_bottom || in_left || in_right) keep = false;
    }
  }
  const float h = max_h[i];
  if (keep && h < P.height_threshold && h > -1.0f) keep = false;
  const float mean = ndvi_stats[4 * i + 2], var = ndvi_stats[4 * i + 3];
  if (keep && (mean < P.ndvi_mean_threshold || var > P.ndvi_var_threshold) && mean > -1.0f) keep = false;
  pre[i] = keep ? 1 : 0;
  if (is_contained[i]) atomicMin(first_contained, i);
}

__global__ void decide_kernel(const int* __restrict__ pre, const int* __restrict__ rank,
                              const int* __restrict__ num_contained, const float* __restrict__ ndvi_stats,
                              const double* __restrict__ area, const int* __restrict__ first_contained, int n,
                              int* __restrict__ out_idx, const long long* __restrict__ n_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) n = (int)*n_dev < n ? (int)*n_dev : n;
  if (i >= n) return;
  int out = -1;
  if (pre[i]) {
    const int nc = num_contained[i];
    if (nc >= 3 || nc == 2) {
      out = -1;  // ">= 3": discarded; "== 2": no branch of the reference ever appends
    } else if (nc == 1) {
      const int o = *first_contained;  // exists: somebody is contained by crown i
      if (o >= 0 && o < n) {
        const float mi = ndvi_stats[4 * i + 2], mo = ndvi_stats[4 * o + 2];
        if (fabsf(__fsub_rn(mi, mo)) > 0.05f) {
          const float v_rank = ndvi_stats[4 * rank[i] + 3];  // var_ndvi[i] with i = rank in the pre-selected list
          out = (v_rank < ndvi_stats[4 * o + 3]) ? i : o;
        } else if (area[i] > 0.0) {
          out = i;
        }
      }
    } else {
      out = i;
    }
  }
  out_idx[i] = out;
}

// P9 head (process_geojson, postprocessing.py:739-768): confidence filter, then the area filter on
// simplify(2).area; poly_id = enumeration index after the confidence filter.
__global__ void head_flags_kernel(const double* __restrict__ conf, const double* __restrict__ area, int n,
                                  const long long* __restrict__ n_dev, double conf_thr, double area_min,
                                  double area_max, int* __restrict__ conf_ok, unsigned char* __restrict__ flags) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const bool live = !(n_dev && i >= *n_dev);
  const bool ok = live && conf[i] >= conf_thr;
  conf_ok[i] = ok ? 1 : 0;
  flags[i] = (ok && area[i] >= area_min && area[i] <= area_max) ? 1 : 0;
}

// single-block form of the head (flags + rank scan + poly_id) for small tables (<= 8192 rows): one launch
__global__ void __launch_bounds__(1024)
select_head_block_kernel(const double* __restrict__ conf, const double* __restrict__ area, int n,
                         const long long* __restrict__ n_dev, double conf_thr, double area_min, double area_max,
                         unsigned char* __restrict__ flags, long long* __restrict__ poly_id) {
  __shared__ long long s_warp[32];
  const int live = n_dev ? (int)min((long long)n, *n_dev) : n;
  const int chunk = (n + 1023) / 1024;
  const int lo = min((int)threadIdx.x * chunk, n), hi = min(lo + chunk, n);
  long long mine = 0;
  for (int i = lo; i < hi; ++i) mine += (i < live && conf[i] >= conf_thr) ? 1 : 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    const long long w = s_warp[lane];
    long long wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
  }
  __syncthreads();
  long long run = s_warp[warp] + incl - mine;
  for (int i = lo; i < hi; ++i) {
    const bool ok = i < live && conf[i] >= conf_thr;
    poly_id[i] = run;
    flags[i] = (ok && area[i] >= area_min && area[i] <= area_max) ? 1 : 0;
    run += ok ? 1 : 0;
  }
}

__global__ void widen_kernel(const int* __restrict__ in, int n, long long* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}

__global__ void round_coords_kernel(const double* __restrict__ in, long long n, double* __restrict__ out) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
    out[i] = __ddiv_rn(rint(__dmul_rn(in[i], 1000.0)), 1000.0);
}

}  // namespace

// out_idx[i]: index of the crown emitted at position i of the pre-selected walk, or -1.
// p_host / p_dev: the parameters by value, or (chain) in device memory.
int td_select_crowns_ex(const double* bounds, const float* max_h, const float* ndvi_stats, const double* area,
                        const int* num_contained, const unsigned char* is_contained, int n,
                        const TdSelectParams* p_host, const TdSelectParams* p_dev, int* pre, int* out_idx,
                        const long long* n_dev, cudaStream_t st) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(bounds && max_h && ndvi_stats && area && num_contained && is_contained && (p_host || p_dev) && pre && out_idx);
  SelectParams P = {};
  if (p_host) P = *p_host;
  td_ensure_pool();
  int* first = nullptr;
  int* rank = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  TD_CUDA(td_tmp_alloc((void**)&first, sizeof(int), st));
  TD_CUDA(td_tmp_alloc((void**)&rank, sizeof(int) * n, st));
  TD_CUDA(cudaMemsetAsync(first, 0x7f, sizeof(int), st));   // 0x7f7f7f7f: larger than any index
  const int blocks = td_div_up(n, 256);
  preselect_kernel<<<blocks, 256, 0, st>>>(bounds, max_h, ndvi_stats, n, P, p_dev, pre, first, is_contained, n_dev);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, pre, rank, n, st);
  TD_CUDA(td_tmp_alloc(&tmp, tmp_bytes, st));
  cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, pre, rank, n, st);
  decide_kernel<<<blocks, 256, 0, st>>>(pre, rank, num_contained, ndvi_stats, area, first, n, out_idx, n_dev);
  cudaError_t e = cudaGetLastError();
  td_tmp_free(tmp, st);
  td_tmp_free(rank, st);
  td_tmp_free(first, st);
  if (e != cudaSuccess) { td_set_error("td_select_crowns: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

// params: 14 doubles = [use_overlap, is_seam_image, left, bottom, right, top, band_left,
//   band_right, band_top, band_bottom, height_threshold, ndvi_mean_threshold,
//   ndvi_var_threshold, reserved] (host pointer)
extern "C" int td_select_crowns(const double* bounds, const float* max_h, const float* ndvi_stats,
                                const double* area, const int* num_contained, const unsigned char* is_contained,
                                int n, const double* params, int* pre, int* out_idx, const long long* n_dev,
                                void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(params);
  SelectParams P;
  P.use_overlap = params[0] != 0.0; P.is_seam_image = params[1] != 0.0;
  P.left = params[2]; P.bottom = params[3]; P.right = params[4]; P.top = params[5];
  P.band_left = params[6]; P.band_right = params[7]; P.band_top = params[8]; P.band_bottom = params[9];
  // python scalars compared with float32 array elements: the comparison is in float32
  P.height_threshold = (float)params[10]; P.ndvi_mean_threshold = (float)params[11];
  P.ndvi_var_threshold = (float)params[12];
  return td_select_crowns_ex(bounds, max_h, ndvi_stats, area, num_contained, is_contained, n, &P, nullptr, pre, out_idx,
                             n_dev, (cudaStream_t)stream);
}

// flags (n) u8 = conf >= conf_thr && area_min <= area <= area_max; poly_id (n) i64 = number of crowns
// before i that pass the confidence filter (the id process_geojson gives crown i if it passes).
extern "C" int td_select_head(const double* conf, const double* area, int n, const long long* n_dev,
                              double conf_thr, double area_min, double area_max, unsigned char* flags,
                              long long* poly_id, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(conf && area && flags && poly_id);
  cudaStream_t st = (cudaStream_t)stream;
  if (n <= 8192) {
    select_head_block_kernel<<<1, 1024, 0, st>>>(conf, area, n, n_dev, conf_thr, area_min, area_max, flags, poly_id);
    TD_CHECK_LAUNCH("td_select_head");
    return TD_OK;
  }
  td_ensure_pool();
  int* ok = nullptr;
  int* rank = nullptr;
  void* tmp = nullptr;
  size_t tmp_bytes = 0;
  TD_CUDA(td_tmp_alloc((void**)&ok, sizeof(int) * n, st));
  TD_CUDA(td_tmp_alloc((void**)&rank, sizeof(int) * n, st));
  const int blocks = td_div_up(n, 256);
  head_flags_kernel<<<blocks, 256, 0, st>>>(conf, area, n, n_dev, conf_thr, area_min, area_max, ok, flags);
  cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, ok, rank, n, st);
  cudaError_t e = td_tmp_alloc(&tmp, tmp_bytes ? tmp_bytes : 1, st);
  if (e == cudaSuccess) {
    cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, ok, rank, n, st);
    widen_kernel<<<blocks, 256, 0, st>>>(rank, n, poly_id);
    e = cudaGetLastError();
    td_tmp_free(tmp, st);
  }
  td_tmp_free(rank, st);
  td_tmp_free(ok, st);
  if (e != cudaSuccess) { td_set_error("td_select_head: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

// round(coord * 1000) / 1000 with round-half-even, float64 (utilities.py:146-161)
extern "C" int td_round_coords(const double* in, long long n, double* out, void* stream) {
  TD_ARG(n >= 0);
  if (n == 0) return TD_OK;
  TD_ARG(in && out);
  round_coords_kernel<<<td_num_sms() * 4, 256, 0, (cudaStream_t)stream>>>(in, n, out);
  TD_CHECK_LAUNCH("td_round_coords");
  return TD_OK;
}
