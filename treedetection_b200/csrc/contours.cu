// P3 -- mask -> polygon rings: border following on the packed rasters of P2, the
// >= 4 point filter, ring closure and the pixel -> CRS affine.
//
// Replaces Predictor._process_and_save_single (TreeDetection/prediction.py:197-265:
// identity resize, cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE), contour.size >= 8,
// closing point) and xy_gpu (TreeDetection/utilities.py:182-207: corner convention,
// float64).  The reference moves H*W*4 bytes to the device and back and launches ~8
// kernels per contour; here one warp owns one instance window staged in shared memory (the walk
// is inherently sequential per instance, there are ~10^5 independent instances per image).
//
// Two passes because output sizes are data dependent: td_trace_count returns, per
// instance, the number of borders / points / kept rings / ring vertices; after a scan
// (caller side) td_trace_emit re-walks and writes rings in OpenCV's order.
#include "common.cuh"
#include "contour_core.cuh"

namespace {

constexpr int kTraceWarps = 8;                 // instances per CTA (one warp each)
constexpr int kCountSmemPerWarp = 3 * 1024;    // count pass: 3 bit planes (windows up to ~90 x 90 px)
constexpr int kEmitSmemPerWarp = 6 * 1024;     // emit pass: + labels (1 B / px below 255 borders)

// Border following is sequential per instance, and every step depends on the previous
// pixel test: run from global memory a step costs an L2 round trip.  So one WARP owns one
// instance: the lanes copy the window's foreground plane into shared memory (and clear the
// two scratch planes), lane 0 walks the borders at shared-memory latency, and in the emit
// pass all lanes write the ring vertices.  The per-warp budgets are small on purpose: the
// walk is latency bound, so what matters is how many warps an SM can keep resident (64 with
// these budgets).  Windows that do not fit fall back to the global scratch planes.
template <typename LabelT>
__device__ bool stage_window(td::RasterT<LabelT>& R, const uint32_t* bits, const int* win, const long long* word_off,
                             int i, uint32_t* planes, long long total_words, LabelT* labels_global,
                             unsigned char* smem_warp, int budget, bool want_labels) {
  R.w = win[4 * i + 2];
  R.h = win[4 * i + 3];
  R.wpr = (R.w + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int nwords = R.wpr * R.h;
  const uint32_t* fg = bits + word_off[i];
  const size_t need = (size_t)12 * nwords + (want_labels ? sizeof(LabelT) * (size_t)R.w * R.h : 0);
  const bool in_smem = nwords > 0 && need <= (size_t)budget;
  if (in_smem) {
    uint32_t* s_fg = reinterpret_cast<uint32_t*>(smem_warp);
    uint32_t* s_vis = s_fg + nwords;
    uint32_t* s_rgt = s_vis + nwords;
    for (int k = lane; k < nwords; k += 32) { s_fg[k] = fg[k]; s_vis[k] = 0u; s_rgt[k] = 0u; }
    R.fg = s_fg; R.visited = s_vis; R.right = s_rgt;
    R.label = want_labels ? reinterpret_cast<LabelT*>(s_rgt + nwords) : nullptr;
  } else {
    R.fg = fg;
    R.visited = planes + word_off[i];
    R.right = planes + total_words + word_off[i];
    R.label = want_labels ? labels_global : nullptr;
  }
  __syncwarp();
  return in_smem;
}

__global__ void __launch_bounds__(32 * kTraceWarps)
trace_count_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ win, const long long* __restrict__ word_off,
                   int n, uint32_t* __restrict__ planes, long long total_words, int* __restrict__ counts) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kTraceWarps + warp;
  if (i >= n) return;
  td::ContourCounts cc = {0, 0, 0, 0};
  if (win[4 * i + 2] > 0 && win[4 * i + 3] > 0) {
    td::Raster R;
    stage_window<unsigned short>(R, bits, win, word_off, i, planes, total_words, nullptr,
                                 smem + (size_t)warp * kCountSmemPerWarp, kCountSmemPerWarp, false);
    if (lane == 0) cc = td::scan_instance(R, nullptr);
  }
  if (lane == 0) {
    counts[4 * i + 0] = cc.n_contours;
    counts[4 * i + 1] = cc.n_points;
    counts[4 * i + 2] = cc.n_rings;
    counts[4 * i + 3] = cc.n_ring_verts;
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
};

__global__ void __launch_bounds__(32 * kTraceWarps) trace_emit_kernel(EmitArgs A) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int i = blockIdx.x * kTraceWarps + warp;
  if (i >= A.n) return;
  const int wx0 = A.win[4 * i + 0], wy0 = A.win[4 * i + 1];
  if (A.win[4 * i + 2] <= 0 || A.win[4 * i + 3] <= 0) return;
  const long long c0 = A.cont_off[i];
  const int nc = (int)(A.cont_off[i + 1] - c0);
  if (nc == 0) return;
  unsigned char* my_smem = smem + (size_t)warp * kEmitSmemPerWarp;
  td::ContourOut out;
  out.parent = A.ct_parent + c0;
  out.npts = A.ct_npts + c0;
  out.pt_off = A.ct_ptoff + c0;
  out.is_hole = A.ct_hole + c0;
  out.pts = A.pts + 2 * A.pts_off[i];
  int* last_child = A.ct_scratch + 3 * c0;
  int* prev_sib = last_child + nc;
  int* order = prev_sib + nc;
  // labels are border indices: one byte is enough below 255 borders (global fallback keeps u16)
  bool staged8 = false;
  if (nc < 255) {
    td::RasterT<unsigned char> R8;
    R8.w = A.win[4 * i + 2]; R8.h = A.win[4 * i + 3]; R8.wpr = (R8.w + 31) >> 5;
    const size_t need = (size_t)12 * R8.wpr * R8.h + (size_t)R8.w * R8.h;
    if (need <= (size_t)kEmitSmemPerWarp) {
      stage_window<unsigned char>(R8, A.bits, A.win, A.word_off, i, A.planes, A.total_words, nullptr, my_smem,
                                  kEmitSmemPerWarp, true);
      if (lane == 0) td::scan_instance(R8, &out);
      staged8 = true;
    }
  }
  if (!staged8) {
    td::Raster R;
    stage_window<unsigned short>(R, A.bits, A.win, A.word_off, i, A.planes, A.total_words, A.labels + A.px_off[i],
                                 my_smem, kEmitSmemPerWarp, true);
    if (lane == 0) td::scan_instance(R, &out);
  }
  if (lane == 0) td::contour_order(nc, out.parent, last_child, prev_sib, order);
  __syncwarp();   // lane 0's tables and points become visible to the warp
  const double* tf = A.tile_tf + 6 * (size_t)A.inst_tile[i];
  const double ta = tf[0], tb = tf[1], tc = tf[2], td_ = tf[3], te = tf[4], tff = tf[5];
  long long ring = A.ring_base[i];
  long long v = A.vert_base[i];
  for (int k = 0; k < nc; ++k) {
    const int c = order[k];
    const int np = out.npts[c];
    if (np < 4) continue;
    const short* p = out.pts + 2 * (size_t)out.pt_off[c];
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

}  // namespace

// planes: scratch of 2 * total_words uint32, zeroed by this call.
extern "C" int td_trace_count(const uint32_t* bits, const int* win, const long long* word_off, int n_inst,
                              long long total_words, uint32_t* planes, int* counts, void* stream) {
  TD_ARG(n_inst >= 0 && total_words >= 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(bits && win && word_off && planes && counts);
  cudaStream_t st = (cudaStream_t)stream;
  TD_CUDA(cudaMemsetAsync(planes, 0, sizeof(uint32_t) * 2 * (size_t)total_words, st));
  trace_count_kernel<<<td_div_up(n_inst, kTraceWarps), 32 * kTraceWarps, kTraceWarps * kCountSmemPerWarp, st>>>(
      bits, win, word_off, n_inst, planes, total_words, counts);
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
  trace_emit_kernel<<<td_div_up(n_inst, kTraceWarps), 32 * kTraceWarps, kTraceWarps * kEmitSmemPerWarp, st>>>(A);
  TD_CHECK_LAUNCH("td_trace_emit");
  return TD_OK;
}
