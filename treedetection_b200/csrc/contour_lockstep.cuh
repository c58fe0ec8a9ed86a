// P3 core, lock-step form -- the same border following as contour_core.cuh (Suzuki-Abe with
// OpenCV's conventions and output order; oracle: cv2.findContours(RETR_TREE, CHAIN_APPROX_SIMPLE),
// TreeDetection/prediction.py:232-234), restructured as a per-lane STATE MACHINE so that the 32
// lanes of a warp can each walk their own instance window in lock step:
//
//     while (any lane still working)  lane_step(state)
//
// One call performs one bounded micro-step: either "advance the raster scan to the next border
// start (and begin that border)" or "one step along the current border".  The step along a
// border is branch-light: the 8 neighbours are gathered into a bit mask from three 3-bit row
// reads (one funnel shift per row) and the next direction is a find-first-set on the rotated mask.  contour_core.cuh keeps
// the straightforward sequential form; both are checked against cv2 (tests/test_hostsim_contours.py).
//
// Plain C++ (tests/hostsim runs it with a single lane).
#pragma once
#include "contour_core.cuh"

namespace td {

enum LaneMode { kScan = 0, kFollow = 1, kDone = 2 };
constexpr int kScanRows = 4;

template <typename LabelT, typename Mem = GenericMem>
struct LaneState {
  RasterT<LabelT, Mem> R;
  ContourOut* out;      // null: count only
  ContourCounts cc;
  int mode;
  // raster scan
  int y, wi, min_o, min_h;
  // current border
  int x0, y0, x1, y1, x3, y3, s, prev_s, px, py, npts, hole;
  int first_x, first_y, last_x, last_y, parent;
};

TD_HD inline int dir_dx(int k) { return (int)((0x901Au >> (2 * (k & 7))) & 3u) - 1; }
TD_HD inline int dir_dy(int k) { return (int)((0xA901u >> (2 * (k & 7))) & 3u) - 1; }

TD_HD inline int ffs32(uint32_t v) {   // index of the lowest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

// low 32 bits of the 64-bit value (hi : lo) shifted right by s (0 <= s < 32): one funnel shift on the device
TD_HD inline uint32_t shr64(uint32_t lo, uint32_t hi, int s) {
#if defined(__CUDA_ARCH__)
  return __funnelshift_r(lo, hi, s);
#else
  return (uint32_t)((((uint64_t)hi << 32) | lo) >> (s & 31));
#endif
}

// 8-neighbour foreground mask of (x, y), bit k = direction k (E, NE, N, NW, W, SW, S, SE).
// The step along a border calls this once per pixel, so it is written without a per-row case analysis:
// word index, bit position and the "neighbour word needed" predicates are computed once, each of
// the three rows is one load (plus one predicated load when x sits on a word boundary) and the three bits
// come out of a 64-bit window (previous word's top bit : this word : next word's low bit) with a funnel
// shift.  A word that is not loaded only feeds bits the shift does not select.
template <typename LabelT, typename Mem>
TD_HD inline uint32_t neighbours(const RasterT<LabelT, Mem>& R, int x, int y) {
  const int wi = x >> 5, b = x & 31;
  const bool need_l = b == 0 && wi > 0, need_r = b == 31 && wi + 1 < R.wpr;
  const uint32_t* mid = R.fg + ((size_t)y * R.wpr + wi);
  uint32_t r[3];
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
  for (int d = 0; d < 3; ++d) {
    const int yy = y + d - 1;
    uint32_t v = 0u;
    if ((unsigned)yy < (unsigned)R.h) {
      const uint32_t* p = mid + (d - 1) * R.wpr;
      const uint32_t lo = Mem::ld(p);
      const uint32_t pl = need_l ? Mem::ld(p - 1) : 0u;
      const uint32_t pr = need_r ? Mem::ld(p + 1) : 0u;
      const uint32_t w_lo = (lo << 1) | (pl >> 31);            // bit k = pixel 32 * wi + k - 1
      const uint32_t w_hi = (lo >> 31) | ((pr & 1u) << 1);
      v = shr64(w_lo, w_hi, b) & 7u;                           // pixels x - 1, x, x + 1
    }
    r[d] = v;
  }
  const uint32_t t = r[0], m = r[1], bt = r[2];
  return ((m >> 2) & 1u) | (((t >> 2) & 1u) << 1) | (((t >> 1) & 1u) << 2) | ((t & 1u) << 3) | ((m & 1u) << 4) |
         ((bt & 1u) << 5) | (((bt >> 1) & 1u) << 6) | (((bt >> 2) & 1u) << 7);
}

template <typename LabelT, typename Mem>
TD_HD inline void lane_emit(LaneState<LabelT, Mem>& S, int x, int y) {
  if (S.out && S.cc.n_points + S.npts < S.out->cap_points) {
    short* p = S.out->pts + 2 * ((size_t)S.cc.n_points + S.npts);
    p[0] = (short)x;
    p[1] = (short)y;
  }
  if (S.npts == 0) { S.first_x = x; S.first_y = y; }
  S.last_x = x;
  S.last_y = y;
  ++S.npts;
}

template <typename LabelT, typename Mem>
TD_HD inline void lane_finish_border(LaneState<LabelT, Mem>& S) {
  const int idx = S.cc.n_contours;
  if (S.out && idx < S.out->cap_contours) {
    S.out->parent[idx] = S.parent;
    S.out->npts[idx] = S.npts;
    S.out->pt_off[idx] = S.cc.n_points;
    S.out->is_hole[idx] = S.hole ? 1 : 0;
  }
  S.cc.n_contours += 1;
  S.cc.n_points += S.npts;
  if (S.npts >= 4) {
    S.cc.n_rings += 1;
    S.cc.n_ring_verts += S.npts + ((S.first_x != S.last_x || S.first_y != S.last_y) ? 1 : 0);
  }
  if (S.hole) { S.min_o = S.x0 + 1; S.min_h = S.x0 + 1; }
  else { S.min_o = S.x0 + 1; S.min_h = S.x0; }
  S.mode = kScan;
}

template <typename LabelT, typename Mem>
TD_HD inline void lane_init(LaneState<LabelT, Mem>& S, ContourOut* out) {
  S.out = out;
  S.cc.n_contours = S.cc.n_points = S.cc.n_rings = S.cc.n_ring_verts = 0;
  S.y = 0; S.wi = 0; S.min_o = 0; S.min_h = 0;
  S.mode = (S.R.w > 0 && S.R.h > 0) ? kScan : kDone;
  S.npts = 0; S.hole = 0; S.parent = -1;
  S.x0 = S.y0 = S.x1 = S.y1 = S.x3 = S.y3 = S.s = S.prev_s = S.px = S.py = 0;
  S.first_x = S.first_y = S.last_x = S.last_y = 0;
}

// one step along the current border (mode kFollow)
template <typename LabelT, typename Mem>
TD_HD inline void lane_follow_step(LaneState<LabelT, Mem>& S) {
  RasterT<LabelT, Mem>& R = S.R;
  const int s_end = S.s;
  const uint32_t m = neighbours(R, S.x3, S.y3);
  // first foreground neighbour counter-clockwise after s_end: directions s_end+1 .. s_end+8
  const uint32_t rot = ((m | (m << 8)) >> (s_end + 1)) & 0xffu;
  const int k = s_end + 1 + ffs32(rot | 0x100u);     // rot != 0: we arrived from a neighbour
  const int s = k & 7;
  const int x4 = S.x3 + dir_dx(s), y4 = S.y3 + dir_dy(s);
  R.mark(S.x3, S.y3, (unsigned)(s - 1) < (unsigned)s_end, S.cc.n_contours);
  if (s != S.prev_s) {
    lane_emit(S, S.px, S.py);
    S.prev_s = s;
  }
  S.px += dir_dx(s);
  S.py += dir_dy(s);
  if (x4 == S.x0 && y4 == S.y0 && S.x3 == S.x1 && S.y3 == S.y1) {
    lane_finish_border(S);
    return;
  }
  S.x3 = x4;
  S.y3 = y4;
  S.s = (s + 4) & 7;
}

// Scans row y from word wi0 on for the next border start: an unvisited pixel after a 0 (outer border, allowed
// at x >= min_o) or a pixel without the "right" flag before a 0 (hole border, x >= min_h).  Returns true with
// the start's column, kind and word.  Reads the planes only: any lane may scan any lane's window.
template <typename LabelT, typename Mem>
TD_HD inline bool lane_scan_row(const RasterT<LabelT, Mem>& R, int y, int wi0, int min_o, int min_h, int& x_out,
                                int& hole_out, int& wi_out) {
  for (int wi = wi0; wi < R.wpr; ++wi) {
    const uint32_t F = R.word(R.fg, y, wi);
    if (!F) continue;
    const uint32_t Fl = R.word(R.fg, y, wi - 1), Fr = R.word(R.fg, y, wi + 1);
    const uint32_t V = R.word(R.visited, y, wi), N = R.word(R.right, y, wi);
    const uint32_t prevfg = (F << 1) | (Fl >> 31);
    const uint32_t nextfg = (F >> 1) | (Fr << 31);
    uint32_t O = F & ~V & ~prevfg;   // unvisited pixel after a 0
    uint32_t H = F & ~N & ~nextfg;   // pixel without the right flag before a 0
    const int base = wi * 32;
    if (min_o > base) O &= (min_o - base >= 32) ? 0u : ~((1u << (min_o - base)) - 1u);
    if (min_h > base) H &= (min_h - base >= 32) ? 0u : ~((1u << (min_h - base)) - 1u);
    if (!(O | H)) continue;
    const int a = O ? ffs32(O) : 64, b = H ? ffs32(H) : 64;
    const bool hole = !(a <= b);
    x_out = base + (hole ? b : a);
    hole_out = hole ? 1 : 0;
    wi_out = wi;
    return true;
  }
  return false;
}

// starts the border found at (x, y): parent from the label plane, first neighbour, isolated pixels
template <typename LabelT, typename Mem>
TD_HD inline void lane_begin_border(LaneState<LabelT, Mem>& S, int x, int y, bool hole) {
  RasterT<LabelT, Mem>& R = S.R;
  if (S.cc.n_contours >= 65534) { S.cc.n_contours = -1; S.mode = kDone; return; }
  // parent from the label of the last visited pixel on this row (see contour_core.cuh)
  int parent = -1;
  if (S.out) {
    const int ln = R.lnbd(hole ? x + 1 : x, y);
    if (ln >= 0 && ln < S.out->cap_contours) {
      parent = ln;
      if ((S.out->is_hole[ln] != 0) == hole) parent = S.out->parent[ln];
    }
  }
  S.parent = parent;
  S.hole = hole ? 1 : 0;
  S.x0 = x; S.y0 = y; S.npts = 0;
  // first neighbour clockwise from west (outer) / east (hole)
  const int s_end = hole ? 0 : 4;
  const uint32_t m = neighbours(R, x, y);
  // directions s_end - 1, s_end - 2, ... s_end - 7: bit t of the rotated mask is direction s_end + t,
  // so the first hit clockwise is the HIGHEST set bit among t = 7 .. 1
  const uint32_t rot = ((m | (m << 8)) >> s_end) & 0xfeu;
  const bool found = rot != 0u;
  const int s = found ? ((s_end + highest_bit(rot)) & 7) : s_end;
  if (!found) {   // isolated pixel (the start pixel's own s_end neighbour is background by construction)
    R.mark(x, y, true, S.cc.n_contours);
    lane_emit(S, x, y);
    lane_finish_border(S);
    return;                     // the scan stays on this word: more starts may follow
  }
  S.x1 = x + dir_dx(s); S.y1 = y + dir_dy(s);
  S.x3 = x; S.y3 = y;
  S.s = s; S.prev_s = s ^ 4;
  S.px = x; S.py = y;
  S.mode = kFollow;
}

// one micro-step of one lane: a step along the border, or the raster scan advanced to the next border start
// (up to kScanRows rows per micro-step: rows without a start are the common case)
template <typename LabelT, typename Mem>
TD_HD inline void lane_step(LaneState<LabelT, Mem>& S) {
  if (S.mode == kFollow) { lane_follow_step(S); return; }
  if (S.mode != kScan) return;
  for (int scanned = 0; scanned < kScanRows; ++scanned) {
    int x = 0, hole = 0, wi = 0;
    if (lane_scan_row(S.R, S.y, S.wi, S.min_o, S.min_h, x, hole, wi)) {
      S.wi = wi;
      lane_begin_border(S, x, S.y, hole != 0);
      return;
    }
    // row exhausted
    S.wi = 0; S.min_o = 0; S.min_h = 0;
    if (++S.y >= S.R.h) { S.mode = kDone; return; }
  }
}

}  // namespace td
