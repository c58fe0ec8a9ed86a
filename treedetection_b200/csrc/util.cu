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

// ---- single-block forms ------------------------------------------------------------------------------------
// The tables of one image have tens of thousands of rows.  A device-wide CUB scan / select spends three kernels
// (init, sweep, tail) on them; in the chain those are 40 dependent launches whose latency, not their work, is
// what one pays.  Up to kBlockMax rows ONE block of 1024 threads does a whole scan (or compaction) in a single
// launch: every thread sums a contiguous chunk, the 1024 partial sums are scanned in shared memory, every thread
// walks its chunk again.  That pays for the small tables (seam strips, small images: a few microseconds instead
// of three launches); beyond a few thousand rows one SM is slower than CUB's device-wide passes (measured:
// 109 us against ~12 us at 43 k rows), so larger tables keep those.
constexpr int kBlockMax = 8192;
constexpr int kBlockThreads = 1024;

// exclusive prefix of `mine` over the block's threads (all 1024 threads call); returns the block total through *total
__device__ long long block_exclusive(long long mine, long long* total) {
  __shared__ long long s_warp[32];
  __shared__ long long s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long incl = mine;
  for (int o = 1; o < 32; o <<= 1) {
    const long long t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  __syncthreads();                      // s_warp may still be read by a previous call
  if (lane == 31) s_warp[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    long long w = s_warp[lane];
    long long wi = w;
    for (int o = 1; o < 32; o <<= 1) {
      const long long t = __shfl_up_sync(0xffffffffu, wi, o);
      if (lane >= o) wi += t;
    }
    s_warp[lane] = wi - w;
    if (lane == 31) s_total = wi;
  }
  __syncthreads();
  *total = s_total;
  return s_warp[warp] + incl - mine;
}

__global__ void __launch_bounds__(kBlockThreads)
scan_clamp_block_kernel(const long long* __restrict__ sizes, int k, int n, Caps caps, long long* __restrict__ offs,
                        long long* __restrict__ totals, long long* __restrict__ flag, int* __restrict__ win_zero) {
  __shared__ int s_i0;
  __shared__ long long s_at[kMaxRows];
  const int chunk = (n + kBlockThreads - 1) / kBlockThreads;
  const int lo = min(threadIdx.x * chunk, n), hi = min(lo + chunk, n);
  if (threadIdx.x == 0) s_i0 = 0x7f7f7f7f;
  __syncthreads();
  int my_i0 = 0x7f7f7f7f;
  for (int r = 0; r < k; ++r) {
    const long long* sz = sizes + (size_t)r * n;
    long long* o = offs + (size_t)r * (n + 1);
    long long sum = 0;
    for (int i = lo; i < hi; ++i) sum += sz[i];
    long long total;
    long long run = block_exclusive(sum, &total);
    for (int i = lo; i < hi; ++i) {
      const long long v = sz[i];
      o[i] = run;
      if ((v < 0 || run + v > caps.cap[r]) && i < my_i0) my_i0 = i;
      run += v;
    }
    if (threadIdx.x == 0) o[n] = total;
  }
  if (my_i0 != 0x7f7f7f7f) atomicMin(&s_i0, my_i0);
  __syncthreads();
  const int i0 = s_i0;
  if (i0 >= n) {                                   // nothing overflowed
    if (threadIdx.x < k) totals[threadIdx.x] = offs[(size_t)threadIdx.x * (n + 1) + n];
    return;
  }
  if (threadIdx.x < k) s_at[threadIdx.x] = offs[(size_t)threadIdx.x * (n + 1) + i0];
  __syncthreads();
  for (int r = 0; r < k; ++r) {
    long long* o = offs + (size_t)r * (n + 1);
    for (int i = max(lo, i0); i < hi; ++i) o[i] = s_at[r];
    if (threadIdx.x == 0) { o[n] = s_at[r]; totals[r] = s_at[r]; }
  }
  if (win_zero)
    for (int i = max(lo, i0); i < hi; ++i) { win_zero[4 * i + 2] = 0; win_zero[4 * i + 3] = 0; }
  if (threadIdx.x == 0) atomicOr((unsigned long long*)flag, 1ull);
}

// mode 0: flags[i] != 0 (want 1) / == 0 (want 0) selects index i; mode 1: values[i] >= 0 selects values[i]
template <int kMode>
__global__ void __launch_bounds__(kBlockThreads)
compact_block_kernel(const unsigned char* __restrict__ flags, const int* __restrict__ values, int want, int n,
                     const long long* __restrict__ n_dev, long long* __restrict__ out, long long* __restrict__ count) {
  const int live = n_dev ? (int)min((long long)n, *n_dev) : n;
  const int chunk = (n + kBlockThreads - 1) / kBlockThreads;
  const int lo = min(threadIdx.x * chunk, n), hi = min(lo + chunk, n);
  auto sel = [&](int i) {
    if (i >= live) return false;
    return kMode == 0 ? ((flags[i] != 0) == (want != 0)) : (values[i] >= 0);
  };
  long long mine = 0;
  for (int i = lo; i < hi; ++i) mine += sel(i) ? 1 : 0;
  long long total;
  long long run = block_exclusive(mine, &total);
  for (int i = lo; i < hi; ++i)
    if (sel(i)) out[run++] = kMode == 0 ? (long long)i : (long long)values[i];
  __syncthreads();
  for (int i = (int)total + threadIdx.x; i < n; i += kBlockThreads) out[i] = 0;      // zero tail
  if (threadIdx.x == 0) *count = total;
}

__global__ void __launch_bounds__(kBlockThreads)
ring_offsets_block_kernel(SelLen len, int n, long long* __restrict__ dst_off) {
  const int chunk = (n + kBlockThreads - 1) / kBlockThreads;
  const int lo = min(threadIdx.x * chunk, n), hi = min(lo + chunk, n);
  long long mine = 0;
  for (int i = lo; i < hi; ++i) mine += len(i);
  long long total;
  long long run = block_exclusive(mine, &total);
  if (threadIdx.x == 0) dst_off[0] = 0;
  for (int i = lo; i < hi; ++i) { run += len(i); dst_off[i + 1] = run; }
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
  if (n <= kBlockMax) {
    Caps c;
    for (int r = 0; r < kMaxRows; ++r) c.cap[r] = r < k ? caps[r] : 0;
    scan_clamp_block_kernel<<<1, kBlockThreads, 0, st>>>(sizes, k, n, c, offs, totals, flag, win_zero);
    TD_CHECK_LAUNCH("td_scan_clamp");
    return TD_OK;
  }
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
  if (n <= kBlockMax) {
    compact_block_kernel<0><<<1, kBlockThreads, 0, st>>>(flags, nullptr, want, n, n_dev, sel, count);
    TD_CHECK_LAUNCH("td_compact_flags");
    return TD_OK;
  }
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
  if (n <= kBlockMax) {
    compact_block_kernel<1><<<1, kBlockThreads, 0, st>>>(nullptr, values, 1, n, n_dev, out, count);
    TD_CHECK_LAUNCH("td_compact_nonneg");
    return TD_OK;
  }
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
  if (n > 0 && n <= kBlockMax) {
    TD_ARG(sel);
    ring_offsets_block_kernel<<<1, kBlockThreads, 0, st>>>(SelLen{sel, count, ring_off}, n, dst_off);
    TD_CHECK_LAUNCH("td_ring_offsets");
    return TD_OK;
  }
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
