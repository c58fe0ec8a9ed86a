// P3 -- mask -> polygon rings: border following on the packed rasters of P2, the
// >= 4 point filter, ring closure and the pixel -> CRS affine.
//
// Replaces Predictor._process_and_save_single (TreeDetection/prediction.py:197-265:
// identity resize, cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE), contour.size >= 8,
// closing point) and xy_gpu (TreeDetection/utilities.py:182-207: corner convention,
// float64).  The reference moves H*W*4 bytes to the device and back and launches ~8
// kernels per contour; here every lane of a warp walks its own instance window in lock step, the
// windows' bit planes staged in shared memory (the walk is inherently sequential per instance,
// there are ~10^5 independent instances per image).
//
// Two passes because output sizes are data dependent: td_trace_count returns, per
// instance, the number of borders / points / kept rings / ring vertices; after a scan
// (caller side) td_trace_emit re-walks and writes rings in OpenCV's order.
#include <cstdlib>

#include "common.cuh"
#include "contour_core.cuh"
#include "contour_lockstep.cuh"

namespace {

constexpr int kSmemPerWarpDefault = 32 * 1024;   // bit planes of the 32 windows a warp walks

// Shared memory per warp (= per CTA).  The walk is latency bound, so what counts is that ALL warps of
// the launch are resident at once (one wave): the budget shrinks from 32 KB as far as needed for
// ceil(warps / SMs) CTAs to fit the 227 KB of an SM, but not below 16 KB (windows that do not fit
// use the global scratch planes, which costs more than a second wave).
int smem_per_warp(int n_inst, int lanes = 32) {
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("TREEDET_TRACE_SMEM");
    forced = e && atoi(e) > 0 ? atoi(e) : 0;
  }
  if (forced) return forced;
  const int warps = td_div_up(n_inst, lanes);
  const int per_sm = td_div_up(warps, td_num_sms());
  const int full = kSmemPerWarpDefault * lanes / 32, least = 16 * 1024 * lanes / 32;
  int v = full;
  if (per_sm > 0) {
    const int fit = ((227 * 1024) / per_sm - 1024) & ~1023;      // 1 KB per CTA is reserved by the system
    if (fit < v) v = fit < least ? full : fit;
  }
  return v;
}

// Border following is sequential per instance and issue bound, so the 32 lanes of a warp each
// walk their OWN instance window in lock step (contour_lockstep.cuh: a per-lane state machine,
// one bounded micro-step per iteration).  The three bit planes of a window (foreground + the
// two scratch planes) are packed into the warp's shared memory by a warp prefix sum over the
// windows' sizes; windows that do not fit (rare, large boxes) use the global scratch planes.
// Labels (needed only for the parent lookup at a border start) stay in global memory.
template <typename LabelT, typename Mem>
__device__ bool stage_windows(td::RasterT<LabelT, Mem>& R, bool active, const uint32_t* __restrict__ bits,
                              const int* __restrict__ win, const long long* __restrict__ word_off, int i,
                              uint32_t* __restrict__ planes, long long total_words, unsigned char* smem,
                              int smem_bytes) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  R.w = active ? win[4 * i + 2] : 0;
  R.h = active ? win[4 * i + 3] : 0;
  R.wpr = (R.w + 31) >> 5;
  const int nwords = R.wpr * R.h;
  const long long woff = active ? word_off[i] : 0;
  // warp exclusive scan of the bytes each window needs in shared memory
  const int need = 12 * nwords;
  int incl = need;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(full, incl, o);
    if (lane >= o) incl += v;
  }
  const int off = incl - need;
  const bool in_smem = nwords > 0 && incl <= smem_bytes;
  uint32_t* s_fg = reinterpret_cast<uint32_t*>(smem + off);
  // cooperative copy: all lanes copy window j's words (coalesced), then clear its scratch planes
  for (int j = 0; j < 32; ++j) {
    const int nw_j = __shfl_sync(full, nwords, j);
    const int in_j = __shfl_sync(full, (int)in_smem, j);
    const int off_j = __shfl_sync(full, off, j);
    const long long woff_j = __shfl_sync(full, woff, j);
    if (!in_j) continue;
    uint32_t* dst = reinterpret_cast<uint32_t*>(smem + off_j);
    const uint32_t* src = bits + woff_j;
    for (int k = lane; k < nw_j; k += 32) {
      dst[k] = src[k];
      dst[nw_j + k] = 0u;
      dst[2 * nw_j + k] = 0u;
    }
  }
  __syncwarp();
  if (in_smem) {
    R.fg = s_fg; R.visited = s_fg + nwords; R.right = s_fg + 2 * nwords;
  } else {
    R.fg = bits + woff;
    R.visited = planes + woff;
    R.right = planes + total_words + woff;
  }
  R.label = nullptr;
  return in_smem || nwords == 0;
}

__global__ void __launch_bounds__(32)
trace_count_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ win, const long long* __restrict__ word_off,
                   int n, uint32_t* __restrict__ planes, long long total_words, int* __restrict__ counts,
                   long long* __restrict__ sizes_kn, int smem_bytes) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int i = blockIdx.x * 32 + threadIdx.x;
  const bool active = i < n;
  td::LaneState<unsigned short> S;
  stage_windows(S.R, active, bits, win, word_off, i, planes, total_words, smem, smem_bytes);
  td::lane_init(S, nullptr);
  while (__any_sync(0xffffffffu, S.mode != td::kDone)) td::lane_step(S);
  if (active) {
    counts[4 * i + 0] = S.cc.n_contours;
    counts[4 * i + 1] = S.cc.n_points;
    counts[4 * i + 2] = S.cc.n_rings;
    counts[4 * i + 3] = S.cc.n_ring_verts;
    if (sizes_kn) {      // the same four counts as (4, n) int64 rows: what td_scan_clamp consumes
      sizes_kn[i] = S.cc.n_contours;
      sizes_kn[(size_t)n + i] = S.cc.n_points;
      sizes_kn[2 * (size_t)n + i] = S.cc.n_rings;
      sizes_kn[3 * (size_t)n + i] = S.cc.n_ring_verts;
    }
  }
}

struct EmitArgs {
  const uint32_t* bits;
  const int* win;
  const long long* word_off;
  int n;
  uint32_t* planes;
  long long total_words;
  unsigned short* labels;
  const long long* px_off;     // (n+1) label offsets
  const long long* cont_off;   // (n+1)
  const long long* pts_off;    // (n+1)
  const long long* ring_base;  // (n+1) kept rings before instance i
  const long long* vert_base;  // (n+1) ring vertices before instance i
  int* ct_parent;              // per contour tables (sum n_contours)
  int* ct_npts;
  int* ct_ptoff;
  unsigned char* ct_hole;
  int* ct_scratch;             // 3 ints per contour: last_child, prev_sibling, order
  short* pts;                  // 2 shorts per point (sum n_points)
  const int* inst_tile;
  const double* tile_tf;       // (T, 6)
  long long* ring_off;         // (R + 1)  [R written by the caller]
  int* ring_inst;              // (R)
  double* verts;               // (V, 2)
  int smem_bytes;
};

__global__ void __launch_bounds__(32) trace_emit_kernel(EmitArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int i = blockIdx.x * 32 + threadIdx.x;
  const bool active = i < A.n;
  td::LaneState<unsigned short> S;
  stage_windows(S.R, active, A.bits, A.win, A.word_off, i, A.planes, A.total_words, smem, A.smem_bytes);
  const long long c0 = active ? A.cont_off[i] : 0;
  const int nc = active ? (int)(A.cont_off[i + 1] - c0) : 0;
  td::ContourOut out;
  out.parent = A.ct_parent + c0;
  out.npts = A.ct_npts + c0;
  out.pt_off = A.ct_ptoff + c0;
  out.is_hole = A.ct_hole + c0;
  out.pts = A.pts + 2 * (active ? A.pts_off[i] : 0);
  if (active) S.R.label = A.labels + A.px_off[i];
  td::lane_init(S, &out);
  if (nc == 0) S.mode = td::kDone;          // nothing to emit: skip the walk
  while (__any_sync(0xffffffffu, S.mode != td::kDone)) td::lane_step(S);
}

// Second half of the emit pass: contours in OpenCV's order -> closed CRS rings.  One WARP per
// instance (the walk above leaves per-lane contour tables and window-relative int16 points):
// lane 0 orders the (few) contours, all lanes convert and store the vertices, so the float64
// writes are coalesced instead of being a serial tail of the border walk.
__global__ void __launch_bounds__(256) trace_rings_kernel(EmitArgs A) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= A.n) return;
  const long long c0 = A.cont_off[i];
  const int nc = (int)(A.cont_off[i + 1] - c0);
  if (nc == 0) return;
  const int* parent = A.ct_parent + c0;
  const int* npts = A.ct_npts + c0;
  const int* pt_off = A.ct_ptoff + c0;
  const short* pts = A.pts + 2 * A.pts_off[i];
  int* last_child = A.ct_scratch + 3 * c0;
  int* prev_sib = last_child + nc;
  int* order = prev_sib + nc;
  if (lane == 0) td::contour_order(nc, parent, last_child, prev_sib, order);
  __syncwarp();
  const int wx0 = A.win[4 * i + 0], wy0 = A.win[4 * i + 1];
  const double* tf = A.tile_tf + 6 * (size_t)A.inst_tile[i];
  const double ta = tf[0], tb = tf[1], tc = tf[2], td_ = tf[3], te = tf[4], tff = tf[5];
  long long ring = A.ring_base[i];
  long long v = A.vert_base[i];
  for (int k = 0; k < nc; ++k) {
    const int c = order[k];
    const int np = npts[c];
    if (np < 4) continue;
    const short* p = pts + 2 * (size_t)pt_off[c];
    const bool close = (p[0] != p[2 * (np - 1)]) || (p[1] != p[2 * (np - 1) + 1]);
    const int nv = np + (close ? 1 : 0);
    if (lane == 0) {
      A.ring_off[ring] = v;
      A.ring_inst[ring] = i;
    }
    for (int q = lane; q < nv; q += 32) {
      const int qq = q < np ? q : 0;
      const double col = (double)(p[2 * qq] + wx0), row = (double)(p[2 * qq + 1] + wy0);
      // xy_gpu: a * x + b * y + c, every operation rounded (float64)
      A.verts[2 * (v + q)] = __dadd_rn(__dadd_rn(__dmul_rn(ta, col), __dmul_rn(tb, row)), tc);
      A.verts[2 * (v + q) + 1] = __dadd_rn(__dadd_rn(__dmul_rn(td_, col), __dmul_rn(te, row)), tff);
    }
    ++ring;
    v += nv;
  }
}

// ---- single pass into per-instance capacity slots (the sync-free chain) ------------------------------
// The two-pass form walks every border twice because the exact sizes must be known before the
// outputs are allocated.  With capacity buffers the walk can write into a SLOT per instance --
// cap_contours table rows at i * cap_contours, points at pts_off[i] .. pts_off[i + 1] -- and count
// at the same time; an instance that outgrows its slot is counted to the end but not stored and
// raises bit 2 of *flag (the caller then redoes the image through the exact two-pass form).
struct WalkArgs {
  const uint32_t* bits;
  const int* win;
  const long long* word_off;
  int n;
  uint32_t* planes;
  long long total_words;
  unsigned short* labels;
  const long long* px_off;
  const long long* pts_off;    // (n+1) point slots
  int cap_contours;
  int* ct_parent;              // n * cap_contours each
  int* ct_npts;
  int* ct_ptoff;
  unsigned char* ct_hole;
  int* ct_scratch;             // 3 * n * cap_contours
  short* pts;
  int* counts;                 // (n, 4)
  long long* sizes_kn;         // (2, n): kept rings, ring vertices
  long long* flag;
  const long long* ring_base;  // (n+1), second kernel
  const long long* vert_base;
  const int* inst_tile;
  const double* tile_tf;
  long long* ring_off;
  int* ring_inst;
  double* verts;
  int smem_bytes;
};

// kLanes instances per warp (the other lanes idle): the walk is a chain of dependent shared-memory
// reads, so what hides its latency is the number of resident WARPS, and an image only has n / 32 of them
// with full warps (7 per SM at 34 k instances); 16 instances per warp double that at the price of idle
// issue slots nobody was using.  When all windows of the warp were staged into shared memory (nearly
// always) the planes are accessed with shared-space instructions (td::SharedMem).
template <typename Mem>
__device__ __forceinline__ void walk_lanes(td::LaneState<unsigned short, Mem>& S, const WalkArgs& A, int i,
                                           bool active) {
  const size_t c0 = (size_t)(active ? i : 0) * A.cap_contours;
  td::ContourOut out;
  out.parent = A.ct_parent + c0;
  out.npts = A.ct_npts + c0;
  out.pt_off = A.ct_ptoff + c0;
  out.is_hole = A.ct_hole + c0;
  const long long p0 = active ? A.pts_off[i] : 0;
  out.pts = A.pts + 2 * p0;
  out.cap_contours = A.cap_contours;
  out.cap_points = active ? (int)(A.pts_off[i + 1] - p0) : 0;
  if (active) S.R.label = A.labels + A.px_off[i];
  td::lane_init(S, &out);
  // Lock-step walk with a COOPERATIVE raster scan.  A lane that looks for its next border start would, on its
  // own, test one row per micro-step while the lanes that follow a border wait for it (a third of the kernel's
  // instructions ran with one live lane that way).  Instead the whole warp scans for it: lane k tests row
  // y + k of that lane's window (the planes of all windows of the warp are in its shared memory), a ballot
  // picks the first row with a start.  Then every lane that follows a border takes one step.
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  for (;;) {
    unsigned need = __ballot_sync(full, S.mode == td::kScan);
    if (!need && !__any_sync(full, S.mode == td::kFollow)) break;
    while (need) {
      const int src = __ffs(need) - 1;
      need &= need - 1;
      td::RasterT<unsigned short, Mem> R;
      R.fg = reinterpret_cast<const uint32_t*>(__shfl_sync(full, (unsigned long long)S.R.fg, src));
      R.visited = reinterpret_cast<uint32_t*>(__shfl_sync(full, (unsigned long long)S.R.visited, src));
      R.right = reinterpret_cast<uint32_t*>(__shfl_sync(full, (unsigned long long)S.R.right, src));
      R.label = nullptr;
      R.w = __shfl_sync(full, S.R.w, src); R.h = __shfl_sync(full, S.R.h, src); R.wpr = __shfl_sync(full, S.R.wpr, src);
      const int y0 = __shfl_sync(full, S.y, src), wi0 = __shfl_sync(full, S.wi, src);
      const int mo = __shfl_sync(full, S.min_o, src), mh = __shfl_sync(full, S.min_h, src);
      int row = -1, fx = 0, fhole = 0, fwi = 0;
      for (int base = y0; base < R.h && row < 0; base += 32) {
        const int yy = base + lane;
        int x = 0, hole = 0, wi = 0;
        bool f = false;
        if (yy < R.h) f = td::lane_scan_row(R, yy, yy == y0 ? wi0 : 0, yy == y0 ? mo : 0, yy == y0 ? mh : 0, x, hole, wi);
        const unsigned hit = __ballot_sync(full, f);
        if (hit) {
          const int l = __ffs(hit) - 1;
          row = base + l;
          fx = __shfl_sync(full, x, l); fhole = __shfl_sync(full, hole, l); fwi = __shfl_sync(full, wi, l);
        }
      }
      if (lane == src) {
        if (row < 0) { S.y = S.R.h; S.mode = td::kDone; }
        else {
          if (row != S.y) { S.min_o = 0; S.min_h = 0; }     // a fresh row
          S.y = row; S.wi = fwi;
          td::lane_begin_border(S, fx, row, fhole != 0);
        }
      }
    }
    if (S.mode == td::kFollow) td::lane_follow_step(S);
  }
  if (!active) return;
  const bool over = S.cc.n_contours < 0 || S.cc.n_contours > out.cap_contours || S.cc.n_points > out.cap_points;
  if (over) atomicOr((unsigned long long*)A.flag, 4ull);
  A.counts[4 * i + 0] = over ? 0 : S.cc.n_contours;
  A.counts[4 * i + 1] = over ? 0 : S.cc.n_points;
  A.counts[4 * i + 2] = over ? 0 : S.cc.n_rings;
  A.counts[4 * i + 3] = over ? 0 : S.cc.n_ring_verts;
  A.sizes_kn[i] = over ? 0 : S.cc.n_rings;
  A.sizes_kn[(size_t)A.n + i] = over ? 0 : S.cc.n_ring_verts;
}

template <int kLanes, int kMinBlocks = 1>
__global__ void __launch_bounds__(32, kMinBlocks) trace_walk_kernel(WalkArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x;
  const int i = blockIdx.x * kLanes + lane;
  const bool active = lane < kLanes && i < A.n;
  td::LaneState<unsigned short, td::SharedMem> S;
  const bool ok = stage_windows(S.R, active, A.bits, A.win, A.word_off, i, A.planes, A.total_words, smem, A.smem_bytes);
  if (__all_sync(0xffffffffu, ok)) {
    walk_lanes(S, A, i, active);
  } else {      // some window did not fit the warp's shared memory: generic accesses for this warp
    td::LaneState<unsigned short, td::GenericMem> G;
    G.R.fg = S.R.fg; G.R.visited = S.R.visited; G.R.right = S.R.right; G.R.label = nullptr;
    G.R.w = S.R.w; G.R.h = S.R.h; G.R.wpr = S.R.wpr;
    walk_lanes(G, A, i, active);
  }
}

// slot form of trace_rings_kernel: one warp per instance
__global__ void __launch_bounds__(256) trace_rings_slots_kernel(WalkArgs A) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= A.n) return;
  const int nc = A.counts[4 * i];
  if (nc <= 0) return;
  // instances truncated by the capacity clamp of the ring / vertex scans emit nothing
  if (A.ring_base[i + 1] - A.ring_base[i] != A.counts[4 * i + 2] ||
      A.vert_base[i + 1] - A.vert_base[i] != A.counts[4 * i + 3]) return;
  const size_t c0 = (size_t)i * A.cap_contours;
  const int* parent = A.ct_parent + c0;
  const int* npts = A.ct_npts + c0;
  const int* pt_off = A.ct_ptoff + c0;
  const short* pts = A.pts + 2 * A.pts_off[i];
  int* last_child = A.ct_scratch + 3 * c0;
  int* prev_sib = last_child + nc;
  int* order = prev_sib + nc;
  if (lane == 0) td::contour_order(nc, parent, last_child, prev_sib, order);
  __syncwarp();
  const int wx0 = A.win[4 * i + 0], wy0 = A.win[4 * i + 1];
  const double* tf = A.tile_tf + 6 * (size_t)A.inst_tile[i];
  const double ta = tf[0], tb = tf[1], tc = tf[2], td_ = tf[3], te = tf[4], tff = tf[5];
  long long ring = A.ring_base[i];
  long long v = A.vert_base[i];
  for (int k = 0; k < nc; ++k) {
    const int c = order[k];
    const int np = npts[c];
    if (np < 4) continue;
    const short* p = pts + 2 * (size_t)pt_off[c];
    const bool close = (p[0] != p[2 * (np - 1)]) || (p[1] != p[2 * (np - 1) + 1]);
    const int nv = np + (close ? 1 : 0);
    if (lane == 0) {
      A.ring_off[ring] = v;
      A.ring_inst[ring] = i;
    }
    for (int q = lane; q < nv; q += 32) {
      const int qq = q < np ? q : 0;
      const double col = (double)(p[2 * qq] + wx0), row = (double)(p[2 * qq + 1] + wy0);
      A.verts[2 * (v + q)] = __dadd_rn(__dadd_rn(__dmul_rn(ta, col), __dmul_rn(tb, row)), tc);
      A.verts[2 * (v + q) + 1] = __dadd_rn(__dadd_rn(__dmul_rn(td_, col), __dmul_rn(te, row)), tff);
    }
    ++ring;
    v += nv;
  }
}

}  // namespace

// planes: scratch of 2 * total_words uint32, zeroed by this call.
extern "C" int td_trace_count(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                              long long total_words, uint32_t* planes, int* counts, long long* sizes_kn,
                              void* stream) {
  TD_ARG(n_inst >= 0 && total_words >= 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(bits && win && word_off && planes && counts);
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(planes, 0, sizeof(uint32_t) * 2 * (size_t)total_words, st));
  const int smem_bytes = smem_per_warp(n_inst);
  trace_count_kernel<<<td_div_up(n_inst, 32), 32, smem_bytes, st>>>(bits, win, word_off, n_inst, planes, total_words, counts,
                                                                  sizes_kn, smem_bytes);
  TD_CHECK_LAUNCH("td_trace_count");
  return TD_OK;
}

extern "C" int td_trace_emit(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                             long long total_words, uint32_t* planes, unsigned short* labels,
                             const long long* px_off, const long long* cont_off, const long long* pts_off,
                             const long long* ring_base, const long long* vert_base, int* ct_int5,
                             unsigned char* ct_hole, short* pts, long long total_contours, const int* inst_tile,
                             const double* tile_tf, long long* ring_off, int* ring_inst, double* verts,
                             void* stream) {
  TD_ARG(n_inst >= 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(bits && win && word_off && planes && labels && px_off && cont_off && pts_off && ring_base && vert_base);
  TD_ARG(ct_int5 && ct_hole && pts && inst_tile && tile_tf && ring_off && ring_inst && verts);
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(planes, 0, sizeof(uint32_t) * 2 * (size_t)total_words, st));
  EmitArgs A;
  A.bits = bits; A.win = win; A.word_off = word_off; A.n = n_inst; A.planes = planes; A.total_words = total_words;
  A.labels = labels; A.px_off = px_off; A.cont_off = cont_off; A.pts_off = pts_off;
  A.ring_base = ring_base; A.vert_base = vert_base;
  // ct_int5: 6 int arrays of total_contours each (parent, npts, ptoff, 3 x scratch)
  A.ct_parent = ct_int5;
  A.ct_npts = ct_int5 + total_contours;
  A.ct_ptoff = ct_int5 + 2 * total_contours;
  A.ct_scratch = ct_int5 + 3 * total_contours;
  A.ct_hole = ct_hole; A.pts = pts; A.inst_tile = inst_tile; A.tile_tf = tile_tf;
  A.ring_off = ring_off; A.ring_inst = ring_inst; A.verts = verts;
  A.smem_bytes = smem_per_warp(n_inst);
  trace_emit_kernel<<<td_div_up(n_inst, 32), 32, A.smem_bytes, st>>>(A);
  trace_rings_kernel<<<td_div_up((long long)n_inst * 32, 256), 256, 0, st>>>(A);
  TD_CHECK_LAUNCH("td_trace_emit");
  return TD_OK;
}

// Single-pass walk into capacity slots (see trace_walk_kernel).  ct_int6: 6 * n_inst * cap_contours i32
// (parent, npts, ptoff, 3 x scratch), ct_hole: n_inst * cap_contours u8, pts: 2 * pts_off[n_inst] i16,
// counts (N,4) i32 out, sizes_kn (2,N) i64 out = [kept rings, ring vertices], flag: bit 2 raised when an
// instance outgrew its slot.  planes (2 * total_words) is zeroed by the call.
extern "C" int td_trace_walk(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                             long long total_words, uint32_t* planes, unsigned short* labels,
                             const long long* px_off, const long long* pts_off, int cap_contours, int* ct_int6,
                             unsigned char* ct_hole, short* pts, int* counts, long long* sizes_kn, long long* flag,
                             void* stream) {
  TD_ARG(n_inst >= 0 && total_words >= 0 && cap_contours > 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(bits && win && word_off && planes && labels && px_off && pts_off && ct_int6 && ct_hole && pts && counts &&
         sizes_kn && flag);
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(planes, 0, sizeof(uint32_t) * 2 * (size_t)total_words, st));
  WalkArgs A = {};
  const size_t nc = (size_t)n_inst * cap_contours;
  A.bits = bits; A.win = win; A.word_off = word_off; A.n = n_inst; A.planes = planes; A.total_words = total_words;
  A.labels = labels; A.px_off = px_off; A.pts_off = pts_off; A.cap_contours = cap_contours;
  A.ct_parent = ct_int6; A.ct_npts = ct_int6 + nc; A.ct_ptoff = ct_int6 + 2 * nc; A.ct_scratch = ct_int6 + 3 * nc;
  A.ct_hole = ct_hole; A.pts = pts; A.counts = counts; A.sizes_kn = sizes_kn; A.flag = flag;
  static int lanes = 0;
  if (lanes == 0) { const char* e = getenv("TREEDET_TRACE_LANES"); lanes = e && atoi(e) > 0 ? atoi(e) : 8; }
  const int L = lanes >= 32 ? 32 : (lanes >= 16 ? 16 : 8);
  A.smem_bytes = smem_per_warp(n_inst, L);
  if (L == 32) trace_walk_kernel<32><<<td_div_up(n_inst, 32), 32, A.smem_bytes, st>>>(A);
  else if (L == 16) trace_walk_kernel<16><<<td_div_up(n_inst, 16), 32, A.smem_bytes, st>>>(A);
  else {
    // residency of the 8-lane form: 96 registers without a bound (21 one-warp CTAs per SM), 80 at 24 CTAs (default:
    // step 3.561 -> 3.537 ms), 72 at 28 (a few spilled words: 3.554); the walk is a chain of dependent
    // shared-memory reads, resident warps are what hides it
    static int mb = -1;
    if (mb < 0) { const char* e = getenv("TREEDET_TRACE_MINBLOCKS"); mb = e ? atoi(e) : 24; }
    if (mb >= 28) trace_walk_kernel<8, 28><<<td_div_up(n_inst, 8), 32, A.smem_bytes, st>>>(A);
    else if (mb >= 24) trace_walk_kernel<8, 24><<<td_div_up(n_inst, 8), 32, A.smem_bytes, st>>>(A);
    else trace_walk_kernel<8><<<td_div_up(n_inst, 8), 32, A.smem_bytes, st>>>(A);
  }
  TD_CHECK_LAUNCH("td_trace_walk");
  return TD_OK;
}

// Second half of the single-pass form: slots -> closed CRS rings (ring_base / vert_base = scans of
// sizes_kn; instances the scans truncated emit nothing).  Same outputs as td_trace_emit.
extern "C" int td_trace_rings(const int* win, int n_inst, const int* counts, const long long* pts_off,
                              int cap_contours, int* ct_int6, const unsigned char* ct_hole, const short* pts,
                              const long long* ring_base, const long long* vert_base, const int* inst_tile,
                              const double* tile_tf, long long* ring_off, int* ring_inst, double* verts,
                              void* stream) {
  TD_ARG(n_inst >= 0 && cap_contours > 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(win && counts && pts_off && ct_int6 && ct_hole && pts && ring_base && vert_base && inst_tile && tile_tf &&
         ring_off && ring_inst && verts);
  WalkArgs A = {};
  const size_t nc = (size_t)n_inst * cap_contours;
  A.win = win; A.n = n_inst; A.counts = const_cast<int*>(counts); A.pts_off = pts_off; A.cap_contours = cap_contours;
  A.ct_parent = ct_int6; A.ct_npts = ct_int6 + nc; A.ct_ptoff = ct_int6 + 2 * nc; A.ct_scratch = ct_int6 + 3 * nc;
  A.ct_hole = const_cast<unsigned char*>(ct_hole); A.pts = const_cast<short*>(pts);
  A.ring_base = ring_base; A.vert_base = vert_base; A.inst_tile = inst_tile; A.tile_tf = tile_tf;
  A.ring_off = ring_off; A.ring_inst = ring_inst; A.verts = verts;
  trace_rings_slots_kernel<<<td_div_up((long long)n_inst * 32, 256), 256, 0, (cudaStream_t)stream>>>(A);
  TD_CHECK_LAUNCH("td_trace_rings");
  return TD_OK;
}
