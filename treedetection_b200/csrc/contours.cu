// P3 -- mask -> polygon rings: border following on the packed rasters of P2, the
// >= 4 point filter, ring closure and the pixel -> CRS affine.
//
// Replaces Predictor._process_and_save_single (TreeDetection/prediction.py:197-265:
// identity resize, cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE), contour.size >= 8,
// closing point) and xy_gpu (TreeDetection/utilities.py:182-207: corner convention,
// float64).  The reference moves H*W*4 bytes to the device and back and launches ~8
// kernels per contour; here one thread walks one instance window (the algorithm is
// inherently sequential per instance, there are ~10^5 independent instances per image).
//
// Two passes because output sizes are data dependent: td_trace_count returns, per
// instance, the number of borders / points / kept rings / ring vertices; after a scan
// (caller side) td_trace_emit re-walks and writes rings in OpenCV's order.
#include "common.cuh"
#include "contour_core.cuh"

namespace {

__global__ void __launch_bounds__(64)
trace_count_kernel(const uint32_t* __restrict__ bits, const int* __restrict__ win, const long long* __restrict__ word_off,
                   int n, uint32_t* __restrict__ planes, long long total_words, int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  td::Raster R;
  R.w = win[4 * i + 2];
  R.h = win[4 * i + 3];
  R.wpr = (R.w + 31) >> 5;
  td::ContourCounts cc = {0, 0, 0, 0};
  if (R.w > 0 && R.h > 0) {
    R.fg = bits + word_off[i];
    R.visited = planes + word_off[i];
    R.right = planes + total_words + word_off[i];
    R.label = nullptr;
    cc = td::scan_instance(R, nullptr);
  }
  counts[4 * i + 0] = cc.n_contours;
  counts[4 * i + 1] = cc.n_points;
  counts[4 * i + 2] = cc.n_rings;
  counts[4 * i + 3] = cc.n_ring_verts;
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

__global__ void __launch_bounds__(64) trace_emit_kernel(EmitArgs A) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= A.n) return;
  td::Raster R;
  const int wx0 = A.win[4 * i + 0], wy0 = A.win[4 * i + 1];
  R.w = A.win[4 * i + 2];
  R.h = A.win[4 * i + 3];
  R.wpr = (R.w + 31) >> 5;
  if (R.w <= 0 || R.h <= 0) return;
  const long long c0 = A.cont_off[i];
  const int nc = (int)(A.cont_off[i + 1] - c0);
  if (nc == 0) return;
  R.fg = A.bits + A.word_off[i];
  R.visited = A.planes + A.word_off[i];
  R.right = A.planes + A.total_words + A.word_off[i];
  R.label = A.labels + A.px_off[i];
  td::ContourOut out;
  out.parent = A.ct_parent + c0;
  out.npts = A.ct_npts + c0;
  out.pt_off = A.ct_ptoff + c0;
  out.is_hole = A.ct_hole + c0;
  out.pts = A.pts + 2 * A.pts_off[i];
  td::scan_instance(R, &out);
  int* last_child = A.ct_scratch + 3 * c0;
  int* prev_sib = last_child + nc;
  int* order = prev_sib + nc;
  td::contour_order(nc, out.parent, last_child, prev_sib, order);
  const double* tf = A.tile_tf + 6 * (size_t)A.inst_tile[i];
  const double ta = tf[0], tb = tf[1], tc = tf[2], td_ = tf[3], te = tf[4], tff = tf[5];
  long long ring = A.ring_base[i];
  long long v = A.vert_base[i];
  for (int k = 0; k < nc; ++k) {
    const int c = order[k];
    const int np = out.npts[c];
    if (np < 4) continue;
    const short* p = out.pts + 2 * (size_t)out.pt_off[c];
    A.ring_off[ring] = v;
    A.ring_inst[ring] = i;
    ++ring;
    const bool close = (p[0] != p[2 * (np - 1)]) || (p[1] != p[2 * (np - 1) + 1]);
    const int nv = np + (close ? 1 : 0);
    for (int q = 0; q < nv; ++q) {
      const int qq = q < np ? q : 0;
      const double col = (double)(p[2 * qq] + wx0), row = (double)(p[2 * qq + 1] + wy0);
      // xy_gpu: a * x + b * y + c, every operation rounded (float64)
      A.verts[2 * v] = __dadd_rn(__dadd_rn(__dmul_rn(ta, col), __dmul_rn(tb, row)), tc);
      A.verts[2 * v + 1] = __dadd_rn(__dadd_rn(__dmul_rn(td_, col), __dmul_rn(te, row)), tff);
      ++v;
    }
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
  trace_count_kernel<<<td_div_up(n_inst, 64), 64, 0, st>>>(bits, win, word_off, n_inst, planes, total_words, counts);
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
  trace_emit_kernel<<<td_div_up(n_inst, 64), 64, 0, st>>>(A);
  TD_CHECK_LAUNCH("td_trace_emit");
  return TD_OK;
}
