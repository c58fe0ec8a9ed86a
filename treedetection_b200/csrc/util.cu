// Device-side bookkeeping of the sync-free chain: prefix sums with capacity clamping and
// stream compaction whose COUNTS STAY ON THE DEVICE.
//
// The reference sizes every intermediate with a host round trip (len(), cp.where(...).get(),
// Python lists; e.g. TreeDetection/postprocessing.py:389-405, 739-768).  Here the caller
// allocates variable-length buffers with a CAPACITY (remembered from earlier images of the
// same tiling), every kernel reads the live count through a `const long long* n_dev`, and one
// read back at the end of the image returns all counts plus an overflow flag; an overflow
// truncates safely (items from the first one that does not fit on are dropped) and the caller
// re-runs that image through the exact-size path.
#include <cub/device/device_scan.cuh>
#include <cub/device/device_select.cuh>
#include <cub/iterator/counting_input_iterator.cuh>
#include <cub/iterator/transform_input_iterator.cuh>

#include "chain_internal.cuh"
#include "common.cuh"

namespace {

constexpr int kMaxRows = 8;

struct Caps {
  long long cap[kMaxRows];
};

__global__ void first_overflow_kernel(const long long* __restrict__ sizes, const long long* __restrict__ offs, int k,
                                      int n, Caps caps, int* __restrict__ i0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool over = false;
  for (int r = 0; r < k; ++r) {
    const long long s = sizes[(size_t)r * n + i];
    over = over || s < 0 || offs[(size_t)r * (n + 1) + i] + s > caps.cap[r];
  }
  if (over) atomicMin(i0, i);
}

// offs[r][i] <- offs[r][min(i, i0)]: every item from the first overflowing one on becomes empty
__global__ void clamp_offsets_kernel(const long long* __restrict__ sizes, long long* __restrict__ offs, int k, int n,
                                     const int* __restrict__ i0p, long long* __restrict__ totals,
                                     long long* __restrict__ flag, int* __restrict__ win_zero) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n) return;
  const int i0 = *i0p;
  for (int r = 0; r < k; ++r) {
    long long* o = offs + (size_t)r * (n + 1);
    long long v;
    if (i0 < n) v = o[i0];                                        // o[i0] is left unchanged by its own thread
    else v = i < n ? o[i] : o[n - 1] + sizes[(size_t)r * n + (n - 1)];
    if (i >= i0 || i == n) o[i] = v;
    if (i == n) totals[r] = v;
  }
  if (i < n && i >= i0 && win_zero) { win_zero[4 * i + 2] = 0; win_zero[4 * i + 3] = 0; }
  if (i == n && i0 < n) atomicOr((unsigned long long*)flag, 1ull);
}

struct FlagLive {
  const unsigned char* flags;
  const long long* n_dev;
  int want;    // 1: flags[i] != 0 selects, 0: flags[i] == 0 selects
  __host__ __device__ bool operator()(int i) const {
    return (n_dev == nullptr || i < *n_dev) && ((flags[i] != 0) == (want != 0));
  }
};
struct NonNegLive {
  const int* values;
  const long long* n_dev;
  __host__ __device__ bool operator()(int i) const { return (n_dev == nullptr || i < *n_dev) && values[i] >= 0; }
};
struct AsLongLong {
  const int* values;
  __host__ __device__ long long operator()(int i) const { return (long long)values[i]; }
};

__global__ void fill_tail_kernel(long long* __restrict__ out, int n, const long long* __restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && i >= *count) out[i] = 0;
}

__global__ void ring_tail_kernel(long long* __restrict__ ring_off, int* __restrict__ ring_inst, int cap_rings,
                                 const long long* __restrict__ n_rings, const long long* __restrict__ n_verts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > cap_rings) return;
  if (i >= *n_rings) {
    ring_off[i] = *n_verts;
    if (i < cap_rings) ring_inst[i] = 0;
  }
}

// lens[i] = length of ring sel[i]: from the kept-vertex counts of td_simplify_rings, or from ring_off
struct SelLen {
  const long long* sel;
  const int* count;
  const long long* ring_off;
  __host__ __device__ long long operator()(int i) const {
    const long long r = sel[i];
    return count ? (long long)count[r] : ring_off[r + 1] - ring_off[r];
  }
};

constexpr int kMaxGather = 8;
struct GatherArgs {
  const unsigned char* in[kMaxGather];
  unsigned char* out[kMaxGather];
  int row_bytes[kMaxGather];
  int k;
};

// out_a[i] = in_a[sel[i]] for every array a: one thread per (row, array)
__global__ void gather_rows_kernel(GatherArgs A, const long long* __restrict__ sel, int n,
                                   const long long* __restrict__ n_dev) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || (n_dev && i >= *n_dev)) return;
  const int a = blockIdx.y;
  const int rb = A.row_bytes[a];
  const unsigned char* src = A.in[a] + (size_t)sel[i] * rb;
  unsigned char* dst = A.out[a] + (size_t)i * rb;
  if ((rb & 7) == 0) {
    for (int b = 0; b < rb; b += 8) *reinterpret_cast<uint64_t*>(dst + b) = *reinterpret_cast<const uint64_t*>(src + b);
  } else if ((rb & 3) == 0) {
    for (int b = 0; b < rb; b += 4) *reinterpret_cast<uint32_t*>(dst + b) = *reinterpret_cast<const uint32_t*>(src + b);
  } else {
    for (int b = 0; b < rb; ++b) dst[b] = src[b];
  }
}

template <typename InIt, typename FlagIt>
int select_flagged(InIt in, FlagIt flags, long long* out, long long* count, int n, cudaStream_t st) {
  size_t bytes = 0;
  TD_CUDA(cub::DeviceSelect::Flagged(nullptr, bytes, in, flags, out, count, n, st));
  void* tmp = nullptr;
  TD_CUDA(td_tmp_alloc(&tmp, bytes ? bytes : 1, st));
  cudaError_t e = cub::DeviceSelect::Flagged(tmp, bytes, in, flags, out, count, n, st);
  if (e == cudaSuccess) {
    fill_tail_kernel<<<td_div_up(n, 256), 256, 0, st>>>(out, n, count);
    e = cudaGetLastError();
  }
  td_tmp_free(tmp, st);
  if (e != cudaSuccess) { td_set_error("compaction: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

}  // namespace

// sizes (k, n) i64 -> offs (k, n + 1) i64 exclusive scans, truncated at the first item whose end
// exceeds caps[r] in any row r (or whose size is negative): that item and all later ones become
// empty, bit 0 of *flag is set, and (win_zero != null) their window sizes win[4i+2], win[4i+3] are
// zeroed so that the per-instance kernels skip them.  totals[r] = offs[r][n].  caps: HOST.
extern "C" int td_scan_clamp(const long long* sizes, int k, int n, const long long* caps, long long* offs,
                             long long* totals, long long* flag, int* win_zero, void* stream) {
  TD_ARG(k > 0 && k <= kMaxRows && n >= 0 && caps && offs && totals && flag);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) {
    TD_CUDA(cudaMemsetAsync(offs, 0, sizeof(long long) * k, st));
    TD_CUDA(cudaMemsetAsync(totals, 0, sizeof(long long) * k, st));
    return TD_OK;
  }
  TD_ARG(sizes);
  td_ensure_pool();
  size_t bytes = 0;
  TD_CUDA(cub::DeviceScan::ExclusiveSum(nullptr, bytes, sizes, offs, n, st));
  void* tmp = nullptr;
  int* i0 = nullptr;
  TD_CUDA(td_tmp_alloc(&tmp, bytes ? bytes : 1, st));
  TD_CUDA(td_tmp_alloc((void**)&i0, sizeof(int), st));
  cudaError_t e = cudaMemsetAsync(i0, 0x7f, sizeof(int), st);   // 0x7f7f7f7f: larger than any index
  for (int r = 0; r < k && e == cudaSuccess; ++r)
    e = cub::DeviceScan::ExclusiveSum(tmp, bytes, sizes + (size_t)r * n, offs + (size_t)r * (n + 1), n, st);
  if (e == cudaSuccess) {
    Caps c;
    for (int r = 0; r < kMaxRows; ++r) c.cap[r] = r < k ? caps[r] : 0;
    first_overflow_kernel<<<td_div_up(n, 256), 256, 0, st>>>(sizes, offs, k, n, c, i0);
    clamp_offsets_kernel<<<td_div_up(n + 1, 256), 256, 0, st>>>(sizes, offs, k, n, i0, totals, flag, win_zero);
    e = cudaGetLastError();
  }
  td_tmp_free(i0, st);
  td_tmp_free(tmp, st);
  if (e != cudaSuccess) { td_set_error("td_scan_clamp: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

// sel[0..count) = ascending indices i < n (and < *n_dev when given) with flags[i] != 0; sel[count..n) = 0;
// *count stays on the device.
int td_compact_flags_ex(const unsigned char* flags, int want, int n, const long long* n_dev, long long* sel,
                        long long* count, cudaStream_t st) {
  TD_ARG(n >= 0 && count);
  if (n == 0) { TD_CUDA(cudaMemsetAsync(count, 0, sizeof(long long), st)); return TD_OK; }
  TD_ARG(flags && sel);
  td_ensure_pool();
  cub::CountingInputIterator<int> iota(0);
  cub::TransformInputIterator<bool, FlagLive, cub::CountingInputIterator<int>> fl(iota, FlagLive{flags, n_dev, want});
  cub::CountingInputIterator<long long> ids(0);
  return select_flagged(ids, fl, sel, count, n, st);
}

extern "C" int td_compact_flags(const unsigned char* flags, int n, const long long* n_dev, long long* sel,
                                long long* count, void* stream) {
  return td_compact_flags_ex(flags, 1, n, n_dev, sel, count, (cudaStream_t)stream);
}

// out[0..count) = the non-negative values[i] (i < n, i < *n_dev) in order; out[count..n) = 0
// (`final = out_idx[out_idx >= 0]` of the selection step).
extern "C" int td_compact_nonneg(const int* values, int n, const long long* n_dev, long long* out, long long* count,
                                 void* stream) {
  TD_ARG(n >= 0 && count);
  cudaStream_t st = (cudaStream_t)stream;
  if (n == 0) { TD_CUDA(cudaMemsetAsync(count, 0, sizeof(long long), st)); return TD_OK; }
  TD_ARG(values && out);
  td_ensure_pool();
  cub::CountingInputIterator<int> iota(0);
  cub::TransformInputIterator<bool, NonNegLive, cub::CountingInputIterator<int>> fl(iota, NonNegLive{values, n_dev});
  cub::TransformInputIterator<long long, AsLongLong, cub::CountingInputIterator<int>> in(iota, AsLongLong{values});
  return select_flagged(in, fl, out, count, n, st);
}

// closes a capacity-sized ring table: ring_off[i] = *n_verts for *n_rings <= i <= cap_rings
// (the rings past the live count are empty), ring_inst[i] = 0 for *n_rings <= i < cap_rings.
extern "C" int td_ring_tail(long long* ring_off, int* ring_inst, int cap_rings, const long long* n_rings,
                            const long long* n_verts, void* stream) {
  TD_ARG(ring_off && ring_inst && cap_rings >= 0 && n_rings && n_verts);
  ring_tail_kernel<<<td_div_up(cap_rings + 1, 256), 256, 0, (cudaStream_t)stream>>>(ring_off, ring_inst, cap_rings,
                                                                                     n_rings, n_verts);
  TD_CHECK_LAUNCH("td_ring_tail");
  return TD_OK;
}

// dst_off (n + 1) = exclusive offsets of the rings sel[0..n): lengths from `count` (kept vertices of
// td_simplify_rings) when given, else from ring_off.  Rows past *n_dev are harmless garbage.
extern "C" int td_ring_offsets(const long long* ring_off, const int* count, const long long* sel, int n,
                               long long* dst_off, void* stream) {
  TD_ARG(n >= 0 && dst_off && (ring_off || count));
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(dst_off, 0, sizeof(long long), st));
  if (n == 0) return TD_OK;
  TD_ARG(sel);
  td_ensure_pool();
  cub::CountingInputIterator<int> iota(0);
  cub::TransformInputIterator<long long, SelLen, cub::CountingInputIterator<int>> lens(iota, SelLen{sel, count, ring_off});
  size_t bytes = 0;
  TD_CUDA(cub::DeviceScan::InclusiveSum(nullptr, bytes, lens, dst_off + 1, n, st));
  void* tmp = nullptr;
  TD_CUDA(td_tmp_alloc(&tmp, bytes ? bytes : 1, st));
  cudaError_t e = cub::DeviceScan::InclusiveSum(tmp, bytes, lens, dst_off + 1, n, st);
  td_tmp_free(tmp, st);
  if (e != cudaSuccess) { td_set_error("td_ring_offsets: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
  return TD_OK;
}

// k <= 8 row gathers in one launch: out[a][i] = in[a][sel[i]] (rows of row_bytes[a] bytes) for i < n
// (and < *n_dev).  in / out / row_bytes: HOST arrays of k device pointers / sizes.
extern "C" int td_gather_rows(const void* const* in, void* const* out, const int* row_bytes, int k,
                              const long long* sel, int n, const long long* n_dev, void* stream) {
  TD_ARG(k > 0 && k <= kMaxGather && n >= 0 && in && out && row_bytes);
  if (n == 0) return TD_OK;
  TD_ARG(sel);
  GatherArgs A;
  A.k = k;
  for (int a = 0; a < k; ++a) {
    TD_ARG(in[a] && out[a] && row_bytes[a] > 0);
    A.in[a] = (const unsigned char*)in[a]; A.out[a] = (unsigned char*)out[a]; A.row_bytes[a] = row_bytes[a];
  }
  gather_rows_kernel<<<dim3(td_div_up(n, 256), k), 256, 0, (cudaStream_t)stream>>>(A, sel, n, n_dev);
  TD_CHECK_LAUNCH("td_gather_rows");
  return TD_OK;
}
