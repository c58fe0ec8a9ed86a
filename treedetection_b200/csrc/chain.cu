// td_chain_* -- the whole post-model chain of ONE image (P2 .. P9) as one host call per stage.
//
// Replaces, for a stream of images with the same tiling, what the reference does per image in
//   Predictor._process_and_save_single + process_and_stitch_predictions
//       (TreeDetection/prediction.py:197-265, TreeDetection/helpers.py:419-600)      -> td_chain_predict
//   process_geojson + process_features (TreeDetection/postprocessing.py:722-809, 478-720) -> td_chain_post
// The reference sizes every intermediate through the host (Python lists, len(), .get()); here every
// variable-length buffer has a CAPACITY inside a caller-provided device workspace, every live count
// stays on the device (16 int64 counters per output slot), and the launch sequence is therefore
// static: it is captured into a CUDA graph the first time a (slot, input pointers) combination is
// seen and replayed with ONE cudaGraphLaunch afterwards, so the host never sits between two kernels
// of the chain.  Per-image scalars (instance count, raster transforms, selection parameters) travel
// through a small device block that is refreshed by a pinned -> device copy in front of the graph.
//
// An overflow of any capacity truncates safely and raises a bit of counter 0; the caller then redoes
// that image through the exact-size entry points.  Results are bit-identical to the exact-size
// composition (tests/test_gpu_chain.py).
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <new>

#include "chain_internal.cuh"
#include "common.cuh"

// C-ABI entry points composed here (include/treedet.h)
extern "C" {
int td_scan_clamp(const long long* sizes, int k, int n, const long long* caps, long long* offs, long long* totals,
                  long long* flag, int* win_zero, void* stream);
int td_paste_threshold_pack(const float* boxes_px, const int* win, const long long* word_off, const float* probs,
                            int n_inst, float threshold, uint32_t* bits, void* stream);
int td_trace_walk(const uint32_t* bits, const int* win, const long long* word_off, int n_inst, long long total_words,
                  uint32_t* planes, unsigned short* labels, const long long* px_off, const long long* pts_off,
                  int cap_contours, int* ct_int6, unsigned char* ct_hole, short* pts, int* counts,
                  long long* sizes_kn, long long* flag, void* stream);
int td_trace_rings(const int* win, int n_inst, const int* counts, const long long* pts_off, int cap_contours,
                   int* ct_int6, const unsigned char* ct_hole, const short* pts, const long long* ring_base,
                   const long long* vert_base, const int* inst_tile, const double* tile_tf, long long* ring_off,
                   int* ring_inst, double* verts, void* stream);
int td_simplify_rings(const double* verts, const long long* ring_off, int n_rings, double tolerance, int* scratch,
                      uint32_t* alive, const double* boxes, const int* ring_box, int* out_count, double* out_bounds,
                      double* out_area, unsigned char* out_keep, int bounds_of_input, const long long* n_dev,
                      void* stream);
int td_ring_tail(long long* ring_off, int* ring_inst, int cap_rings, const long long* n_rings,
                 const long long* n_verts, void* stream);
int td_ring_offsets(const long long* ring_off, const int* count, const long long* sel, int n, long long* dst_off,
                    void* stream);
int td_gather_rows(const void* const* in, void* const* out, const int* row_bytes, int k, const long long* sel, int n,
                   const long long* n_dev, void* stream);
int td_select_head(const double* conf, const double* area, int n, const long long* n_dev, double conf_thr,
                   double area_min, double area_max, unsigned char* flags, long long* poly_id, void* stream);
int td_bbox_nms_ordered_dyn(const double* bounds, const double* conf, const double* area, int n,
                            const long long* n_dev, double iou_threshold, double area_threshold, long long nbr_cap,
                            long long* flag, unsigned char* removed, void* stream);
int td_compact_nonneg(const int* values, int n, const long long* n_dev, long long* out, long long* count,
                      void* stream);
}

namespace {

// counter slots (mirrored by treedetection_b200/pipeline.py)
enum {
  kCtrFlag = 0, kCtrWords = 1, kCtrPx = 2, kCtrSlots = 3, kCtrCont = 4, kCtrPts = 5, kCtrRings = 6, kCtrVerts = 7,
  kCtrNTable = 8, kCtrVTable = 9, kCtrN1 = 10, kCtrN2 = 11, kCtrNFinal = 12, kCtrVFinal = 13, kCtrSize = 16
};

enum { kCfgMask = 0, kCfgTol1, kCfgTol2, kCfgConf, kCfgAreaMin, kCfgAreaMax, kCfgIou, kCfgAreaThr, kCfgCont, kCfgSize = 16 };
enum { kCapInst = 0, kCapWords, kCapPx, kCapPts, kCapRings, kCapVerts, kCapNbrPer, kCapContours, kCapSize = 8 };

constexpr int kMaxSlots = 8;
constexpr int kMaxGraphs = 24;
constexpr int kParamRing = 64;

// per-slot outputs, as offsets into the workspace (the order td_chain_layout reports them in)
enum {
  kOutCounters = 0, kOutTVerts, kOutTOff, kOutTConf, kOutVerts, kOutOff, kOutPid, kOutConf, kOutArea, kOutHeight,
  kOutCentroid, kOutIsContained, kOutNumContained, kOutHxy, kOutNdviStats, kOutCount
};

struct Slot {
  long long* counters;
  double* tverts; long long* toff; double* tconf; double* tarea; double* tbounds;
  double* verts; long long* off; long long* pid; double* conf; double* area; float* height; float* centroid;
  unsigned char* isc; int* num; float* hxy; float* nst;
  size_t out_off[kOutCount];
};

struct PredictIn {
  const float* boxes_net; const float* scores; const float* probs; const int* inst_tile;
  const int* tile_dims; const double* tile_tf; const double* tile_boxes; int n_tiles;
};
struct PostIn {
  const float* ndvi; int nrows, ncols; const float* height; int hrows, hcols; int combined;
};

struct GraphEntry {
  int kind;            // 0 predict, 1 post
  int slot;
  PredictIn pin;
  PostIn qin;
  cudaGraphExec_t exec;
  unsigned long long last_use;
};

struct Chain {
  double cfg[kCfgSize];
  long long cap[kCapSize];
  int n_slots;
  char* ws; size_t ws_bytes;
  // shared intermediates
  float* boxes_px; int* win; long long* sizes1; long long* offs1;
  uint32_t* bits; uint32_t* planes; unsigned short* labels;
  int* ct_int; unsigned char* ct_hole; short* pts; int* counts; long long* sizes2; long long* offs2;
  long long* ring_off; int* ring_inst; double* verts;
  int* scratch; uint32_t* alive; int* ring_box;
  int* count; unsigned char* keep;
  long long* sel;
  unsigned char* flags1; long long* pid_all; long long* sel1;
  double* conf1; double* area1; long long* pid1; double* b1;
  unsigned char* removed; long long* sel2;
  double* conf2; double* areac2; long long* pid2; double* b2; long long* idx2;
  float* cent; float* max_h; float* hxy; float* nst;
  float* ratio; unsigned char* isc; int* num; int* pre; int* out_idx; long long* fin; long long* idxf;
  int* vmax;
  char* arena_mem;               // scratch arena of the composed calls
  TdArena arena;
  TdImageParams* d_params;       // one per slot
  TdImageParams* h_params;       // pinned ring
  unsigned long long seq;
  Slot slots[kMaxSlots];
  GraphEntry graphs[kMaxGraphs];
  int n_graphs;
  unsigned long long tick;
  cudaStream_t cap_stream;       // graphs are captured here, launched on the caller's stream
};

struct Bump {
  size_t off = 0;
  template <typename T>
  size_t take(size_t n) {
    off = (off + 255) & ~(size_t)255;
    const size_t o = off;
    off += sizeof(T) * (n ? n : 1);
    return o;
  }
};

// scratch of the largest composed call (the NMS: float4 boxes, half conf / area, offsets, best / state, sort
// keys and values in double buffers, neighbour slots) plus CUB's temporary storage, with head room
size_t arena_size(const long long* cap) {
  const size_t R = (size_t)cap[kCapRings], Ni = (size_t)cap[kCapInst];
  return (size_t)(96 + 4 * cap[kCapNbrPer]) * R + 64 * Ni + (8u << 20);
}

// lays the workspace out; with c == nullptr only the size is computed
size_t layout(Chain* c, const long long* cap, int n_slots) {
  const size_t Ni = (size_t)cap[kCapInst], W = (size_t)cap[kCapWords], PX = (size_t)cap[kCapPx],
               PS = (size_t)cap[kCapPts], R = (size_t)cap[kCapRings], V = (size_t)cap[kCapVerts],
               CC = (size_t)cap[kCapContours];
  Bump b;
  char* base = c ? c->ws : nullptr;
#define TAKE(field, T, n)                                   \
  do {                                                      \
    const size_t o__ = b.take<T>(n);                        \
    if (c) c->field = reinterpret_cast<T*>(base + o__);     \
  } while (0)
  TAKE(boxes_px, float, 4 * Ni); TAKE(win, int, 4 * Ni); TAKE(sizes1, long long, 3 * Ni);
  TAKE(offs1, long long, 3 * (Ni + 1));
  TAKE(bits, uint32_t, W); TAKE(planes, uint32_t, 2 * W); TAKE(labels, unsigned short, PX);
  TAKE(ct_int, int, 6 * Ni * CC); TAKE(ct_hole, unsigned char, Ni * CC); TAKE(pts, short, 2 * PS);
  TAKE(counts, int, 4 * Ni); TAKE(sizes2, long long, 2 * Ni); TAKE(offs2, long long, 2 * (Ni + 1));
  TAKE(ring_off, long long, R + 2); TAKE(ring_inst, int, R + 1); TAKE(verts, double, 2 * V);
  TAKE(scratch, int, 5 * V); TAKE(alive, uint32_t, V / 32 + R + 2); TAKE(ring_box, int, R);
  TAKE(count, int, R); TAKE(keep, unsigned char, R);
  TAKE(sel, long long, R);
  TAKE(flags1, unsigned char, R); TAKE(pid_all, long long, R); TAKE(sel1, long long, R);
  TAKE(conf1, double, R); TAKE(area1, double, R); TAKE(pid1, long long, R); TAKE(b1, double, 4 * R);
  TAKE(removed, unsigned char, R); TAKE(sel2, long long, R);
  TAKE(conf2, double, R); TAKE(areac2, double, R); TAKE(pid2, long long, R); TAKE(b2, double, 4 * R);
  TAKE(idx2, long long, R);
  TAKE(cent, float, 2 * R); TAKE(max_h, float, R); TAKE(hxy, float, 2 * R); TAKE(nst, float, 4 * R);
  TAKE(ratio, float, R); TAKE(isc, unsigned char, R); TAKE(num, int, R); TAKE(pre, int, R); TAKE(out_idx, int, R);
  TAKE(fin, long long, R); TAKE(idxf, long long, R);
  TAKE(vmax, int, 4);
  const size_t arena_bytes = arena_size(cap);
  TAKE(arena_mem, char, arena_bytes);
  if (c) c->arena = TdArena{c->arena_mem, arena_bytes, 0};
  TAKE(d_params, TdImageParams, kMaxSlots);
#undef TAKE
  for (int s = 0; s < n_slots; ++s) {
    Slot dummy;
    Slot& S = c ? c->slots[s] : dummy;
#define OUT(field, T, n, id)                                \
  do {                                                      \
    const size_t o__ = b.take<T>(n);                        \
    S.field = reinterpret_cast<T*>(base + o__);             \
    if ((id) >= 0) S.out_off[(id) >= 0 ? (id) : 0] = o__;   \
  } while (0)
    OUT(counters, long long, kCtrSize, kOutCounters);
    OUT(tverts, double, 2 * V, kOutTVerts); OUT(toff, long long, R + 1, kOutTOff); OUT(tconf, double, R, kOutTConf);
    OUT(tarea, double, R, -1); OUT(tbounds, double, 4 * R, -1);
    OUT(verts, double, 2 * V, kOutVerts); OUT(off, long long, R + 1, kOutOff); OUT(pid, long long, R, kOutPid);
    OUT(conf, double, R, kOutConf); OUT(area, double, R, kOutArea); OUT(height, float, R, kOutHeight);
    OUT(centroid, float, 2 * R, kOutCentroid); OUT(isc, unsigned char, R, kOutIsContained);
    OUT(num, int, R, kOutNumContained); OUT(hxy, float, 2 * R, kOutHxy); OUT(nst, float, 4 * R, kOutNdviStats);
#undef OUT
  }
  return (b.off + 255) & ~(size_t)255;
}

// ring_box[r] = tile of the instance that produced ring r (the stitch filter box of helpers.py:280-303)
__global__ void ring_tile_kernel(const int* __restrict__ ring_inst, const int* __restrict__ inst_tile, int n,
                                 const long long* __restrict__ n_dev, int* __restrict__ ring_box) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= n) return;
  ring_box[r] = r < *n_dev ? inst_tile[ring_inst[r]] : 0;
}

// table rows of the kept rings: conf = float(score) of the producing instance; thread 0 also records the
// table's vertex count
__global__ void table_rows_kernel(const long long* __restrict__ sel, int n, const long long* __restrict__ n_dev,
                                  const int* __restrict__ ring_inst, const float* __restrict__ scores,
                                  const long long* __restrict__ dst_off, double* __restrict__ tconf,
                                  long long* __restrict__ v_total) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  const long long live = *n_dev < n ? *n_dev : n;
  if (q == 0) *v_total = dst_off[live];
  if (q >= live) return;
  tconf[q] = (double)scores[ring_inst[sel[q]]];
}

__global__ void pick_total_kernel(const long long* __restrict__ off, const long long* __restrict__ n_dev, int cap,
                                  long long* __restrict__ out) {
  const long long live = *n_dev < cap ? *n_dev : cap;
  *out = off[live];
}

// every composed call starts at the beginning of the scratch arena (see TdArena in common.cuh)
#define RC(call)                    \
  do {                              \
    td_arena_rewind();              \
    const int rc__ = (call);        \
    if (rc__ != TD_OK) return rc__; \
  } while (0)

int gather(cudaStream_t st, const long long* sel, int n, const long long* n_dev, int k, const void* const* in,
           void* const* out, const int* rb) {
  return td_gather_rows(in, out, rb, k, sel, n, n_dev, st);
}

// P2 + P3 + P4 (+ the per-ring part of P9's head) of one image into slot s.  n: grid size of the
// per-instance kernels (the live count comes from d_params[s].n_inst).
int enqueue_predict(Chain* c, int s, const PredictIn& in, int n, cudaStream_t st) {
  Slot& S = c->slots[s];
  long long* ctr = S.counters;
  long long* flag = ctr + kCtrFlag;
  const long long* n_inst_dev = &c->d_params[s].n_inst;
  const int R = (int)c->cap[kCapRings];
  const int cc = (int)c->cap[kCapContours];
  TD_CUDA(cudaMemsetAsync(ctr, 0, sizeof(long long) * kCtrSize, st));
  RC(td_paste_plan_ex(in.boxes_net, in.inst_tile, in.tile_dims, n, in.n_tiles, c->boxes_px, c->win, c->sizes1,
                      c->sizes1 + n, n_inst_dev, st));
  const long long caps1[3] = {c->cap[kCapWords], c->cap[kCapPx], c->cap[kCapPts]};
  RC(td_scan_clamp(c->sizes1, 3, n, caps1, c->offs1, ctr + kCtrWords, flag, c->win, st));
  const long long* word_off = c->offs1;
  const long long* px_off = c->offs1 + (n + 1);
  const long long* slot_off = c->offs1 + 2 * (size_t)(n + 1);
  RC(td_paste_threshold_pack(c->boxes_px, c->win, word_off, in.probs, n, (float)c->cfg[kCfgMask], c->bits, st));
  RC(td_trace_walk(c->bits, c->win, word_off, n, c->cap[kCapWords], c->planes, c->labels, px_off, slot_off, cc,
                   c->ct_int, c->ct_hole, c->pts, c->counts, c->sizes2, flag, st));
  const long long caps2[2] = {c->cap[kCapRings], c->cap[kCapVerts]};
  RC(td_scan_clamp(c->sizes2, 2, n, caps2, c->offs2, ctr + kCtrRings, flag, nullptr, st));
  RC(td_trace_rings(c->win, n, c->counts, slot_off, cc, c->ct_int, c->ct_hole, c->pts, c->offs2, c->offs2 + (n + 1),
                    in.inst_tile, in.tile_tf, c->ring_off, c->ring_inst, c->verts, st));
  RC(td_ring_tail(c->ring_off, c->ring_inst, R + 1, ctr + kCtrRings, ctr + kCtrVerts, st));
  // P4: simplify(tol1) + tile box filter of every traced ring (the classic one-warp-per-ring kernel; a fused
  // two-pass form and lane-per-ring / level-by-level forms were measured slower)
  ring_tile_kernel<<<td_div_up(R, 256), 256, 0, st>>>(c->ring_inst, in.inst_tile, R, ctr + kCtrRings, c->ring_box);
  TD_CHECK_LAUNCH("ring_tile");
  RC(td_simplify_rings(c->verts, c->ring_off, R, c->cfg[kCfgTol1], c->scratch, c->alive, in.tile_boxes, c->ring_box,
                       c->count, nullptr, nullptr, c->keep, 0, ctr + kCtrRings, st));
  RC(td_compact_flags_ex(c->keep, 1, R, ctr + kCtrRings, c->sel, ctr + kCtrNTable, st));
  RC(td_ring_offsets(c->ring_off, c->count, c->sel, R, S.toff, st));
  RC(td_take_rings_ex(c->verts, c->ring_off, c->sel, R, c->scratch, S.toff, S.tverts, ctr + kCtrNTable, 0, st));
  // head of P9 on the stitched table: area of simplify(tol2), bounds of the table ring (postprocessing.py:747-754)
  RC(td_simplify_rings(S.tverts, S.toff, R, c->cfg[kCfgTol2], c->scratch, c->alive, nullptr, nullptr, c->count,
                       S.tbounds, S.tarea, c->keep, 1, ctr + kCtrNTable, st));
  table_rows_kernel<<<td_div_up(R, 256), 256, 0, st>>>(c->sel, R, ctr + kCtrNTable, c->ring_inst, in.scores, S.toff,
                                                       S.tconf, ctr + kCtrVTable);
  TD_CHECK_LAUNCH("table_rows");
  return TD_OK;
}

// P9 head, P6, P7, P8, P9 of the table in slot s
int enqueue_post(Chain* c, int s, const PostIn& in, cudaStream_t st) {
  Slot& S = c->slots[s];
  long long* ctr = S.counters;
  long long* flag = ctr + kCtrFlag;
  const int R = (int)c->cap[kCapRings];
  const TdImageParams* P = &c->d_params[s];
  RC(td_select_head(S.tconf, S.tarea, R, ctr + kCtrNTable, c->cfg[kCfgConf], c->cfg[kCfgAreaMin], c->cfg[kCfgAreaMax],
                    c->flags1, c->pid_all, st));
  RC(td_compact_flags_ex(c->flags1, 1, R, nullptr, c->sel1, ctr + kCtrN1, st));
  {
    const void* gi[4] = {S.tconf, S.tarea, c->pid_all, S.tbounds};
    void* go[4] = {c->conf1, c->area1, c->pid1, c->b1};
    const int rb[4] = {8, 8, 8, 32};
    RC(gather(st, c->sel1, R, ctr + kCtrN1, 4, gi, go, rb));
  }
  RC(td_bbox_nms_ordered_dyn(c->b1, c->conf1, c->area1, R, ctr + kCtrN1, c->cfg[kCfgIou], c->cfg[kCfgAreaThr],
                             c->cap[kCapNbrPer] * (long long)R, flag, c->removed, st));
  RC(td_compact_flags_ex(c->removed, 0, R, ctr + kCtrN1, c->sel2, ctr + kCtrN2, st));
  {
    // idx2 = sel1[sel2]: the table ring of every post-NMS crown (no vertex copies before the final one)
    const void* gi[5] = {c->conf1, c->area1, c->pid1, c->b1, c->sel1};
    void* go[5] = {c->conf2, c->areac2, c->pid2, c->b2, c->idx2};
    const int rb[5] = {8, 8, 8, 32, 8};
    RC(gather(st, c->sel2, R, ctr + kCtrN2, 5, gi, go, rb));
  }
  const long long* n2 = ctr + kCtrN2;
  RC(td_centroids_ex(S.tverts, S.toff, c->idx2, R, c->cent, c->vmax, n2, st));
  if (in.combined) {
    RC(td_crown_stats_ex(S.tverts, S.toff, c->idx2, R, in.ndvi, in.height, in.nrows, in.ncols, nullptr, &P->ndvi_tf, 0,
                         c->max_h, c->hxy, c->nst, n2, st));
  } else {
    RC(td_crown_stats_ex(S.tverts, S.toff, c->idx2, R, nullptr, in.height, in.hrows, in.hcols, nullptr, &P->height_tf,
                         1, c->max_h, c->hxy, nullptr, n2, st));
    RC(td_crown_stats_ex(S.tverts, S.toff, c->idx2, R, in.ndvi, nullptr, in.nrows, in.ncols, nullptr, &P->ndvi_tf, 2,
                         nullptr, nullptr, c->nst, n2, st));
  }
  RC(td_containment_ex(c->b2, nullptr, R, c->cfg[kCfgCont], c->ratio, c->isc, c->num, n2, st));
  RC(td_select_crowns_ex(c->b2, c->max_h, c->nst, c->areac2, c->num, c->isc, R, nullptr, &P->sel, c->pre, c->out_idx,
                         n2, st));
  RC(td_compact_nonneg(c->out_idx, R, n2, c->fin, ctr + kCtrNFinal, st));
  const long long* nf = ctr + kCtrNFinal;
  TD_CUDA(cudaMemsetAsync(c->idxf, 0, sizeof(long long) * R, st));   // rows past the live count must stay valid ring indices
  {
    const void* gi[8] = {c->pid2, c->conf2, c->areac2, c->max_h, c->cent, c->isc, c->num, c->idx2};
    void* go[8] = {S.pid, S.conf, S.area, S.height, S.centroid, S.isc, S.num, c->idxf};
    const int rb[8] = {8, 8, 8, 4, 8, 1, 4, 8};
    RC(gather(st, c->fin, R, nf, 8, gi, go, rb));
  }
  {
    const void* gi[2] = {c->hxy, c->nst};
    void* go[2] = {S.hxy, S.nst};
    const int rb[2] = {8, 16};
    RC(gather(st, c->fin, R, nf, 2, gi, go, rb));
  }
  RC(td_ring_offsets(S.toff, nullptr, c->idxf, R, S.off, st));
  RC(td_take_rings_ex(S.tverts, S.toff, c->idxf, R, nullptr, S.off, S.verts, nf, 1, st));
  pick_total_kernel<<<1, 1, 0, st>>>(S.off, nf, R, ctr + kCtrVFinal);
  TD_CHECK_LAUNCH("pick_total");
  return TD_OK;
}

struct ArenaScope {
  explicit ArenaScope(TdArena* a) { td_set_arena(a); }
  ~ArenaScope() { td_set_arena(nullptr); }
};

bool same_predict(const PredictIn& a, const PredictIn& b) { return memcmp(&a, &b, sizeof(a)) == 0; }
bool same_post(const PostIn& a, const PostIn& b) { return memcmp(&a, &b, sizeof(a)) == 0; }

GraphEntry* find_graph(Chain* c, int kind, int slot, const PredictIn* pin, const PostIn* qin) {
  for (int i = 0; i < c->n_graphs; ++i) {
    GraphEntry& g = c->graphs[i];
    if (g.kind != kind || g.slot != slot) continue;
    if (kind == 0 ? same_predict(g.pin, *pin) : same_post(g.qin, *qin)) return &g;
  }
  return nullptr;
}

GraphEntry* new_graph_entry(Chain* c) {
  if (c->n_graphs < kMaxGraphs) return &c->graphs[c->n_graphs++];
  int lru = 0;
  for (int i = 1; i < kMaxGraphs; ++i)
    if (c->graphs[i].last_use < c->graphs[lru].last_use) lru = i;
  cudaGraphExecDestroy(c->graphs[lru].exec);
  return &c->graphs[lru];
}

// body(stream) enqueues the launch sequence.  It is captured on the chain's own stream (the caller's may
// be the legacy default stream, which cannot capture) and the instantiated graph is launched on `st`.
template <typename F>
int run_graph(Chain* c, int kind, int slot, const PredictIn* pin, const PostIn* qin, cudaStream_t st, F body) {
  GraphEntry* g = find_graph(c, kind, slot, pin, qin);
  if (!g) {
    if (!c->cap_stream) TD_CUDA(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    cudaGraph_t graph = nullptr;
    TD_CUDA(cudaStreamBeginCapture(c->cap_stream, cudaStreamCaptureModeThreadLocal));
    const int rc = body(c->cap_stream);
    const cudaError_t e = cudaStreamEndCapture(c->cap_stream, &graph);
    if (rc != TD_OK) { if (graph) cudaGraphDestroy(graph); return rc; }
    if (e != cudaSuccess) { td_set_error("cudaStreamEndCapture: %s", cudaGetErrorString(e)); return TD_ERR_CUDA; }
    cudaGraphExec_t exec = nullptr;
    const cudaError_t e2 = cudaGraphInstantiate(&exec, graph, 0);
    cudaGraphDestroy(graph);
    if (e2 != cudaSuccess) { td_set_error("cudaGraphInstantiate: %s", cudaGetErrorString(e2)); return TD_ERR_CUDA; }
    g = new_graph_entry(c);
    g->kind = kind; g->slot = slot; g->exec = exec;
    memset(&g->pin, 0, sizeof(g->pin)); memset(&g->qin, 0, sizeof(g->qin));
    if (pin) g->pin = *pin;
    if (qin) g->qin = *qin;
  }
  g->last_use = ++c->tick;
  TD_CUDA(cudaGraphLaunch(g->exec, st));
  return TD_OK;
}

}  // namespace

// cfg (HOST, 16 doubles): [mask_threshold, simplify_tolerance, area_simplify_tolerance (2.0),
//   confidence_threshold, area_min, area_max, iou_threshold, area_threshold, containment_threshold, 0...]
// caps (HOST, 8 int64): [instances, packed words, label pixels, point slots, rings, vertices,
//   NMS neighbour slots per crown, contour rows per instance]
extern "C" long long td_chain_workspace_bytes(const long long* caps, int n_slots) {
  if (!caps || n_slots < 1 || n_slots > kMaxSlots) return TD_ERR_ARG;
  for (int i = 0; i < kCapSize; ++i)
    if (caps[i] < 1) return TD_ERR_ARG;
  return (long long)layout(nullptr, caps, n_slots);
}

extern "C" int td_chain_create(const double* cfg, const long long* caps, int n_slots, void* workspace,
                               long long workspace_bytes, void** chain_out) {
  TD_ARG(cfg && caps && workspace && chain_out && n_slots >= 1 && n_slots <= kMaxSlots);
  for (int i = 0; i < kCapSize; ++i) TD_ARG(caps[i] >= 1);
  TD_ARG(caps[kCapRings] < (1ll << 30) && caps[kCapInst] < (1ll << 30));
  TD_ARG(((uintptr_t)workspace & 255) == 0);
  const size_t need = layout(nullptr, caps, n_slots);
  if ((long long)need > workspace_bytes) {
    td_set_error("td_chain_create: workspace of %lld bytes, %zu needed", workspace_bytes, need);
    return TD_ERR_ARG;
  }
  Chain* c = new (std::nothrow) Chain();
  TD_ARG(c != nullptr);
  memcpy(c->cfg, cfg, sizeof(c->cfg));
  memcpy(c->cap, caps, sizeof(c->cap));
  c->n_slots = n_slots;
  c->ws = (char*)workspace; c->ws_bytes = need;
  c->seq = 0; c->n_graphs = 0; c->tick = 0; c->cap_stream = nullptr;
  layout(c, caps, n_slots);
  if (cudaHostAlloc((void**)&c->h_params, sizeof(TdImageParams) * kParamRing, cudaHostAllocDefault) != cudaSuccess) {
    delete c;
    td_set_error("td_chain_create: cudaHostAlloc failed");
    return TD_ERR_CUDA;
  }
  *chain_out = c;
  return TD_OK;
}

extern "C" int td_chain_destroy(void* chain) {
  Chain* c = (Chain*)chain;
  if (!c) return TD_OK;
  for (int i = 0; i < c->n_graphs; ++i) cudaGraphExecDestroy(c->graphs[i].exec);
  if (c->cap_stream) cudaStreamDestroy(c->cap_stream);
  cudaFreeHost(c->h_params);
  delete c;
  return TD_OK;
}

// byte offsets (HOST, 16 int64) of slot `slot`'s outputs inside the workspace, in this order:
//   counters (16 i64) | table verts (V,2) f64 | table ring_off (R+1) i64 | table conf (R) f64 |
//   verts (V,2) f64 | ring_off (R+1) i64 | poly_id (R) i64 | conf (R) f64 | area (R) f64 | tree_height (R) f32 |
//   centroid (R,2) f32 | is_contained (R) u8 | num_contained (R) i32 | height arg-max xy (R,2) f32 |
//   ndvi stats (R,4) f32
extern "C" int td_chain_layout(const void* chain, int slot, long long* offsets16) {
  const Chain* c = (const Chain*)chain;
  TD_ARG(c && offsets16 && slot >= 0 && slot < c->n_slots);
  for (int i = 0; i < 16; ++i) offsets16[i] = i < kOutCount ? (long long)c->slots[slot].out_off[i] : -1;
  return TD_OK;
}

// P2 + P3 + P4 of one image into output slot `slot`.  boxes_net (N,4) f32, scores (N) f32, probs (N,28,28) f32,
// inst_tile (N) i32, tile_dims (T,4) i32, tile_tf (T,6) f64, tile_boxes (T,4) f64 -- all device.
// use_graph != 0: replay (or capture) the CUDA graph of this (slot, pointers) combination.
extern "C" int td_chain_predict(void* chain, int slot, const float* boxes_net, const float* scores, const float* probs,
                                const int* inst_tile, int n_inst, const int* tile_dims, const double* tile_tf,
                                const double* tile_boxes, int n_tiles, int use_graph, void* stream) {
  Chain* c = (Chain*)chain;
  TD_ARG(c && slot >= 0 && slot < c->n_slots && n_inst >= 0 && n_tiles > 0);
  TD_ARG(boxes_net && scores && probs && inst_tile && tile_dims && tile_tf && tile_boxes);
  if (n_inst > c->cap[kCapInst]) { td_set_error("td_chain_predict: %d instances, capacity %lld", n_inst, c->cap[kCapInst]); return TD_ERR_OVERFLOW; }
  cudaStream_t st = (cudaStream_t)stream;
  TdImageParams* hp = &c->h_params[c->seq++ % kParamRing];
  hp->n_inst = n_inst;
  TD_CUDA(cudaMemcpyAsync(&c->d_params[slot].n_inst, &hp->n_inst, sizeof(long long), cudaMemcpyHostToDevice, st));
  PredictIn in = {boxes_net, scores, probs, inst_tile, tile_dims, tile_tf, tile_boxes, n_tiles};
  ArenaScope scope(&c->arena);
  if (!use_graph) return enqueue_predict(c, slot, in, n_inst > 0 ? n_inst : 1, st);
  const int n_cap = (int)c->cap[kCapInst];
  return run_graph(c, 0, slot, &in, nullptr, st, [&](cudaStream_t cs) { return enqueue_predict(c, slot, in, n_cap, cs); });
}

// P9 head + P6 + P7 + P8 + P9 of the table td_chain_predict left in `slot`.  ndvi (nrows, ncols) f32 and
// height (hrows, hcols) f32 device rasters with their HOST 6-float transforms; combined != 0: both
// rasters share grid and bounds (get_metadata_within_polygon), else the split pair; select_params: the
// 14 HOST doubles of td_select_crowns.
extern "C" int td_chain_post(void* chain, int slot, const float* ndvi, int nrows, int ncols, const double* ndvi_tf,
                             const float* height, int hrows, int hcols, const double* height_tf, int combined,
                             const double* select_params, int use_graph, void* stream) {
  Chain* c = (Chain*)chain;
  TD_ARG(c && slot >= 0 && slot < c->n_slots && ndvi && height && ndvi_tf && height_tf && select_params);
  TD_ARG(nrows > 0 && ncols > 0 && hrows > 0 && hcols > 0);
  cudaStream_t st = (cudaStream_t)stream;
  TdImageParams* hp = &c->h_params[c->seq++ % kParamRing];
  hp->ndvi_tf = TdAffine6{ndvi_tf[0], ndvi_tf[1], ndvi_tf[2], ndvi_tf[3], ndvi_tf[4], ndvi_tf[5]};
  hp->height_tf = TdAffine6{height_tf[0], height_tf[1], height_tf[2], height_tf[3], height_tf[4], height_tf[5]};
  const double* p = select_params;
  TdSelectParams& P = hp->sel;
  P.use_overlap = p[0] != 0.0; P.is_seam_image = p[1] != 0.0;
  P.left = p[2]; P.bottom = p[3]; P.right = p[4]; P.top = p[5];
  P.band_left = p[6]; P.band_right = p[7]; P.band_top = p[8]; P.band_bottom = p[9];
  P.height_threshold = (float)p[10]; P.ndvi_mean_threshold = (float)p[11]; P.ndvi_var_threshold = (float)p[12];
  const size_t o = offsetof(TdImageParams, ndvi_tf);
  TD_CUDA(cudaMemcpyAsync((char*)&c->d_params[slot] + o, (const char*)hp + o, sizeof(TdImageParams) - o,
                          cudaMemcpyHostToDevice, st));
  PostIn in = {ndvi, nrows, ncols, height, hrows, hcols, combined != 0};
  ArenaScope scope(&c->arena);
  if (!use_graph) return enqueue_post(c, slot, in, st);
  return run_graph(c, 1, slot, nullptr, &in, st, [&](cudaStream_t cs) { return enqueue_post(c, slot, in, cs); });
}
