// P2 -- mask paste + threshold + bit-pack.
//
// Replaces detectron2's detector_postprocess -> ROIMasks.to_bitmasks ->
// paste_masks_in_image -> _do_paste_mask (external, entered from the reference at
// TreeDetection/prediction.py:181-183) and the identity resize + uint8 cast at
// prediction.py:222-229.  The reference materialises an (N,H,W) bool raster
// (202 500 B per instance on a 450x450 tile); here every instance writes only
// the 1-bit raster of its own paste window (the CPU-path window of
// _do_paste_mask: [floor(x0)-1, ceil(x1)+1) x [floor(y0)-1, ceil(y1)+1), clipped
// to the tile), 32 pixels per word, rows padded to a whole word.
//
// Arithmetic (float32, every operation rounded, contractions written out as
// explicit fmaf so that the result is bit-identical to ATen's CPU grid_sample;
// oracle: oracle/port.py paste_probs_closed_form):
//     g   = ((p + 0.5) - b0) / (b1 - b0) * 2 - 1
//     u   = fma(g + 1, M/2, -0.5)
//     w   = u - floor(u);  e = 1 - w      (x axis; n, s on the y axis)
//     out = fma(se, n*w, fma(sw, n*e, fma(ne, s*w, nw * (s*e))))   zero padded
//     bit = out >= threshold
#include "chain_internal.cuh"
#include "common.cuh"

namespace {

constexpr int kM = 28;              // mask side of Mask R-CNN's ROI head
constexpr int kPad = kM + 2;        // zero border so that taps need no bounds test
constexpr int kPasteThreads = 128;  // 4 warps per instance

// ---------------------------------------------------------------------------
// plan: scale + clip boxes, drop empty ones, compute the paste window
// ---------------------------------------------------------------------------
__global__ void paste_plan_kernel(const float* __restrict__ boxes_net, const int* __restrict__ inst_tile,
                                  const int* __restrict__ tile_dims, int n, float* __restrict__ boxes_px,
                                  int* __restrict__ win, long long* __restrict__ nwords,
                                  long long* __restrict__ npx, const long long* __restrict__ n_dev) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (n_dev && i >= *n_dev) {   // capacity tail (n is then the capacity): an empty window, nothing read
    win[4 * i + 0] = 0; win[4 * i + 1] = 0; win[4 * i + 2] = 0; win[4 * i + 3] = 0;
    nwords[i] = 0;
    if (npx) { npx[i] = 0; npx[(size_t)n + i] = 0; }
    return;
  }
  const int t = inst_tile[i];
  const int out_h = tile_dims[4 * t + 0], out_w = tile_dims[4 * t + 1];
  const int net_h = tile_dims[4 * t + 2], net_w = tile_dims[4 * t + 3];
  // detector_postprocess: python-float scale applied to a float32 tensor
  const float sx = (float)((double)out_w / (double)net_w);
  const float sy = (float)((double)out_h / (double)net_h);
  float x0 = __fmul_rn(boxes_net[4 * i + 0], sx);
  float y0 = __fmul_rn(boxes_net[4 * i + 1], sy);
  float x1 = __fmul_rn(boxes_net[4 * i + 2], sx);
  float y1 = __fmul_rn(boxes_net[4 * i + 3], sy);
  const float fw = (float)out_w, fh = (float)out_h;
  x0 = fminf(fmaxf(x0, 0.f), fw);
  x1 = fminf(fmaxf(x1, 0.f), fw);
  y0 = fminf(fmaxf(y0, 0.f), fh);
  y1 = fminf(fmaxf(y1, 0.f), fh);
  boxes_px[4 * i + 0] = x0;
  boxes_px[4 * i + 1] = y0;
  boxes_px[4 * i + 2] = x1;
  boxes_px[4 * i + 3] = y1;
  const bool keep = (__fsub_rn(x1, x0) > 0.f) && (__fsub_rn(y1, y0) > 0.f);
  int wx0 = 0, wy0 = 0, ww = 0, wh = 0;
  if (keep) {
    wx0 = max((int)floorf(x0) - 1, 0);
    wy0 = max((int)floorf(y0) - 1, 0);
    const int wx1 = min((int)ceilf(x1) + 1, out_w);
    const int wy1 = min((int)ceilf(y1) + 1, out_h);
    ww = max(wx1 - wx0, 0);
    wh = max(wy1 - wy0, 0);
  }
  win[4 * i + 0] = wx0;
  win[4 * i + 1] = wy0;
  win[4 * i + 2] = ww;
  win[4 * i + 3] = wh;
  nwords[i] = (long long)((ww + 31) >> 5) * (long long)wh;
  if (npx) {   // (2, n): label-plane pixels; point slot of the single-pass border walk (td_trace_walk)
    npx[i] = (long long)ww * (long long)wh;
    npx[(size_t)n + i] = ww > 0 && wh > 0 ? 4ll * (ww + wh) + 64 : 0;
  }
}

// one axis of the sampling grid: pixel centre -> tap index and the two weights
struct Tap {
  int i0;    // floor(u) + 1 (index into the zero-padded 30x30 table), clamped
  int i1;    // i0 + 1, clamped
  float w1;  // weight of i1 (u - floor(u))
  float w0;  // weight of i0 (1 - w1)
};

TD_D Tap make_tap(int p, float b0, float b1) {
  const float c = __fadd_rn((float)p, 0.5f);
  float g = __fdiv_rn(__fsub_rn(c, b0), __fsub_rn(b1, b0));
  g = __fsub_rn(__fmul_rn(g, 2.f), 1.f);
  const float u = __fmaf_rn(__fadd_rn(g, 1.f), (float)(kM / 2), -0.5f);
  const float fl = floorf(u);
  Tap t;
  t.w1 = __fsub_rn(u, fl);
  t.w0 = __fsub_rn(1.f, t.w1);
  // clamp in float first: u can be far outside the table for pixels of the
  // window that lie outside the box (and NaN/inf never occur: b1 > b0)
  const float flc = fminf(fmaxf(fl, -2.f), (float)(kM + 1));
  const int i = (int)flc;
  t.i0 = min(max(i + 1, 0), kPad - 1);
  t.i1 = min(max(i + 2, 0), kPad - 1);
  return t;
}

TD_D float sample(const float (*tab)[kPad], const Tap& tx, const Tap& ty) {
  const float nw = tab[ty.i0][tx.i0], ne = tab[ty.i0][tx.i1];
  const float sw = tab[ty.i1][tx.i0], se = tab[ty.i1][tx.i1];
  // weights: "n" = ty.w1 (distance from the north tap), "s" = ty.w0, "w" = tx.w1, "e" = tx.w0
  const float cnw = __fmul_rn(ty.w0, tx.w0);
  const float cne = __fmul_rn(ty.w0, tx.w1);
  const float csw = __fmul_rn(ty.w1, tx.w0);
  const float cse = __fmul_rn(ty.w1, tx.w1);
  float acc = __fmul_rn(nw, cnw);
  acc = __fmaf_rn(ne, cne, acc);
  acc = __fmaf_rn(sw, csw, acc);
  acc = __fmaf_rn(se, cse, acc);
  return acc;
}

// ---------------------------------------------------------------------------
// paste + threshold + pack: one CTA per instance, one warp per 32-pixel word
// ---------------------------------------------------------------------------
// The sampling grid is separable: the tap of a pixel column depends on x only, that of a row on y
// only.  Both tap tables are built once per instance in shared memory (w + h divisions instead of
// 2 * w * h); windows wider / higher than kTapCap (rare: crowns are tens of pixels across) compute
// the missing taps on the fly.  Same operations on the same values -> same bits.
constexpr int kTapCap = 160;

struct PackedTap {
  int idx;    // i0 | i1 << 8
  float w1;
};

struct __align__(16) TapRec {
  int off0, off1;   // byte offsets of the two taps (rows: i * kPad * 4, columns: i * 4)
  float w1, w0;
};

TD_D PackedTap pack_tap(const Tap& t) { return PackedTap{t.i0 | (t.i1 << 8), t.w1}; }
TD_D Tap unpack_tap(const PackedTap& p) {
  Tap t;
  t.i0 = p.idx & 0xff; t.i1 = p.idx >> 8;
  t.w1 = p.w1; t.w0 = __fsub_rn(1.f, p.w1);
  return t;
}

__global__ void __launch_bounds__(kPasteThreads)
paste_pack_kernel(const float* __restrict__ boxes_px, const int* __restrict__ win,
                  const long long* __restrict__ word_off, const float* __restrict__ probs, int n, float thr,
                  uint32_t* __restrict__ bits) {
  __shared__ __align__(16) float tab[kPad][kPad];
  // one tap-table area, two views: byte-offset records (windows up to kTapCap x kTapCap, nearly all) or packed taps
  __shared__ __align__(16) TapRec s_rec[2 * kTapCap];
  TapRec* s_rx = s_rec;
  TapRec* s_ry = s_rec + kTapCap;
  PackedTap* s_tx = reinterpret_cast<PackedTap*>(s_rec);
  PackedTap* s_ty = s_tx + kTapCap;
  const int i = blockIdx.x;
  if (i >= n) return;
  const int ww = win[4 * i + 2], wh = win[4 * i + 3];
  if (ww <= 0 || wh <= 0) return;
  const int wx0 = win[4 * i + 0], wy0 = win[4 * i + 1];
  const float bx0 = boxes_px[4 * i + 0], by0 = boxes_px[4 * i + 1];
  const float bx1 = boxes_px[4 * i + 2], by1 = boxes_px[4 * i + 3];
  for (int k = threadIdx.x; k < kPad * kPad; k += kPasteThreads) {
    const int r = k / kPad, c = k % kPad;
    float v = 0.f;
    if (r >= 1 && r <= kM && c >= 1 && c <= kM) v = probs[(size_t)i * kM * kM + (r - 1) * kM + (c - 1)];
    tab[r][c] = v;
  }
  if (ww > kTapCap || wh > kTapCap) {
    for (int k = threadIdx.x; k < min(ww, kTapCap); k += kPasteThreads) s_tx[k] = pack_tap(make_tap(wx0 + k, bx0, bx1));
    for (int k = threadIdx.x; k < min(wh, kTapCap); k += kPasteThreads) s_ty[k] = pack_tap(make_tap(wy0 + k, by0, by1));
  }
  // the same taps as byte offsets into the table (rows: i * kPad * 4, columns: i * 4) with both weights: the
  // inner loop below then needs one 16-byte broadcast load, four adds and four table loads per row
  const bool fast = ww <= kTapCap && wh <= kTapCap;
  if (fast) {
    for (int k = threadIdx.x; k < ww; k += kPasteThreads) {
      const Tap t = make_tap(wx0 + k, bx0, bx1);
      s_rx[k] = TapRec{t.i0 * 4, t.i1 * 4, t.w1, t.w0};
    }
    for (int k = threadIdx.x; k < wh; k += kPasteThreads) {
      const Tap t = make_tap(wy0 + k, by0, by1);
      s_ry[k] = TapRec{t.i0 * kPad * 4, t.i1 * kPad * 4, t.w1, t.w0};
    }
  }
  __syncthreads();
  const int wpr = (ww + 31) >> 5;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t* out = bits + word_off[i];
  if (fast) {
    // a warp owns a 32-column word column and walks down its rows; lanes past the window sample table entry
    // (0, 0) with zero weights and are masked out of the ballot.  sample(): the same products and fused adds
    // in the same order
    const char* tab0 = reinterpret_cast<const char*>(&tab[0][0]);
    for (int wi = warp; wi < wpr; wi += kPasteThreads / 32) {
      const int x = wi * 32 + lane;
      const bool live = x < ww;
      const TapRec cx = live ? s_rx[x] : TapRec{0, 0, 0.f, 0.f};
      const char* col0 = tab0 + cx.off0;
      const char* col1 = tab0 + cx.off1;
      uint32_t* o = out + wi;
#pragma unroll 2
      for (int y = 0; y < wh; ++y) {
        const TapRec ry = s_ry[y];
        const float nw = *reinterpret_cast<const float*>(col0 + ry.off0);
        const float ne = *reinterpret_cast<const float*>(col1 + ry.off0);
        const float sw = *reinterpret_cast<const float*>(col0 + ry.off1);
        const float se = *reinterpret_cast<const float*>(col1 + ry.off1);
        const float cnw = __fmul_rn(ry.w0, cx.w0);
        const float cne = __fmul_rn(ry.w0, cx.w1);
        const float csw = __fmul_rn(ry.w1, cx.w0);
        const float cse = __fmul_rn(ry.w1, cx.w1);
        float acc = __fmul_rn(nw, cnw);
        acc = __fmaf_rn(ne, cne, acc);
        acc = __fmaf_rn(sw, csw, acc);
        acc = __fmaf_rn(se, cse, acc);
        const uint32_t word = __ballot_sync(0xffffffffu, live && acc >= thr);
        if (lane == 0) o[(size_t)y * wpr] = word;
      }
    }
    return;
  }
  // windows wider / higher than kTapCap (rare): packed taps for the first kTapCap columns / rows, the rest on the fly
  for (int wi = warp; wi < wpr; wi += kPasteThreads / 32) {
    const int x = wi * 32 + lane;
    const bool live = x < ww;
    Tap tx{0, 0, 0.f, 0.f};
    if (live) tx = x < kTapCap ? unpack_tap(s_tx[x]) : make_tap(wx0 + x, bx0, bx1);
    const float* col0 = &tab[0][0] + tx.i0;
    const float* col1 = &tab[0][0] + tx.i1;
    for (int y = 0; y < wh; ++y) {
      const Tap ty = y < kTapCap ? unpack_tap(s_ty[y]) : make_tap(wy0 + y, by0, by1);
      const int r0 = ty.i0 * kPad, r1 = ty.i1 * kPad;
      bool on = false;
      if (live) {
        // sample(): the same products and fused adds in the same order
        const float nw = col0[r0], ne = col1[r0], sw = col0[r1], se = col1[r1];
        const float cnw = __fmul_rn(ty.w0, tx.w0);
        const float cne = __fmul_rn(ty.w0, tx.w1);
        const float csw = __fmul_rn(ty.w1, tx.w0);
        const float cse = __fmul_rn(ty.w1, tx.w1);
        float acc = __fmul_rn(nw, cnw);
        acc = __fmaf_rn(ne, cne, acc);
        acc = __fmaf_rn(sw, csw, acc);
        acc = __fmaf_rn(se, cse, acc);
        on = acc >= thr;
      }
      const uint32_t word = __ballot_sync(0xffffffffu, on);
      if (lane == 0) out[y * wpr + wi] = word;
    }
  }
}

// debug / tolerance-test path: the pasted probabilities themselves (float32)
__global__ void paste_values_kernel(const float* __restrict__ boxes_px, const int* __restrict__ win,
                                    const long long* __restrict__ val_off, const float* __restrict__ probs, int n,
                                    float* __restrict__ vals) {
  __shared__ float tab[kPad][kPad];
  const int i = blockIdx.x;
  if (i >= n) return;
  const int ww = win[4 * i + 2], wh = win[4 * i + 3];
  if (ww <= 0 || wh <= 0) return;
  const int wx0 = win[4 * i + 0], wy0 = win[4 * i + 1];
  for (int k = threadIdx.x; k < kPad * kPad; k += blockDim.x) {
    const int r = k / kPad, c = k % kPad;
    float v = 0.f;
    if (r >= 1 && r <= kM && c >= 1 && c <= kM) v = probs[(size_t)i * kM * kM + (r - 1) * kM + (c - 1)];
    tab[r][c] = v;
  }
  __syncthreads();
  const float bx0 = boxes_px[4 * i + 0], by0 = boxes_px[4 * i + 1];
  const float bx1 = boxes_px[4 * i + 2], by1 = boxes_px[4 * i + 3];
  float* out = vals + val_off[i];
  for (int k = threadIdx.x; k < ww * wh; k += blockDim.x) {
    const int y = k / ww, x = k - y * ww;
    out[k] = sample(tab, make_tap(wx0 + x, bx0, bx1), make_tap(wy0 + y, by0, by1));
  }
}

}  // namespace

int td_paste_plan_ex(const float* boxes_net, const int* inst_tile, const int* tile_dims, int n_inst, int n_tiles,
                     float* boxes_px, int* win, long long* nwords, long long* npx, const long long* n_dev,
                     cudaStream_t st) {
  TD_ARG(n_inst >= 0 && n_tiles >= 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(boxes_net && inst_tile && tile_dims && boxes_px && win && nwords);
  paste_plan_kernel<<<td_div_up(n_inst, 256), 256, 0, st>>>(boxes_net, inst_tile, tile_dims, n_inst, boxes_px, win,
                                                            nwords, npx, n_dev);
  TD_CHECK_LAUNCH("td_paste_plan");
  return TD_OK;
}

extern "C" int td_paste_plan(const float* boxes_net, const int* inst_tile, const int* tile_dims, int n_inst,
                             int n_tiles, float* boxes_px, int* win, long long* nwords, long long* npx,
                             void* stream) {
  return td_paste_plan_ex(boxes_net, inst_tile, tile_dims, n_inst, n_tiles, boxes_px, win, nwords, npx, nullptr,
                          (cudaStream_t)stream);
}

extern "C" int td_paste_threshold_pack(const float* boxes_px, const int* win, const long long* word_off,
                                       const float* probs, int n_inst, float threshold, uint32_t* bits,
                                       void* stream) {
  TD_ARG(n_inst >= 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(boxes_px && win && word_off && probs && bits);
  paste_pack_kernel<<<n_inst, kPasteThreads, 0, (cudaStream_t)stream>>>(boxes_px, win, word_off, probs, n_inst,
                                                                       threshold, bits);
  TD_CHECK_LAUNCH("td_paste_threshold_pack");
  return TD_OK;
}

extern "C" int td_paste_values(const float* boxes_px, const int* win, const long long* val_off, const float* probs,
                               int n_inst, float* vals, void* stream) {
  TD_ARG(n_inst >= 0);
  if (n_inst == 0) return TD_OK;
  TD_ARG(boxes_px && win && val_off && probs && vals);
  paste_values_kernel<<<n_inst, 128, 0, (cudaStream_t)stream>>>(boxes_px, win, val_off, probs, n_inst, vals);
  TD_CHECK_LAUNCH("td_paste_values");
  return TD_OK;
}
