// P3 core -- border following on a packed 1-bit raster (one instance window).
//
// Restates what the reference obtains from
//     cv2.findContours(mask_u8, cv2.RETR_TREE, cv2.CHAIN_APPROX_SIMPLE)
// (TreeDetection/prediction.py:232-234): Suzuki-Abe border following on the
// 8-connected foreground with OpenCV's conventions -- raster scan for border
// starts, outer borders start at a 0->1 transition on an unvisited pixel, hole
// borders at a (positive-labelled) 1->0 transition, the first neighbour search
// runs clockwise from west (outer) / east (hole), the walk counter-clockwise from
// the reversed arrival direction, pixels whose east neighbour was examined as 0 get
// the "right" flag, and CHAIN_APPROX_SIMPLE keeps a point whenever the step
// direction changes.  Output ORDER is OpenCV's too: pre-order walk of the border
// tree, siblings in reverse order of discovery.  The oracle is cv2 itself
// (tests/test_hostsim_contours.py, tests/test_gpu_contours.py).
//
// State per pixel: foreground bit (input), "visited" bit, "right" bit (two scratch
// bit-planes with the raster's layout) and a 16-bit label = index of the border
// that wrote it last (the LNBD of Suzuki-Abe), needed only for the parent lookup.
//
// Plain C++ (no CUDA intrinsics) so that tests/hostsim can compile the very same
// function with g++; on the device one thread runs one instance.
#pragma once
#include "common.cuh"

namespace td {

struct ContourOut {
  // pass 1 (count): all null.  pass 2 (fill): arrays sized from pass 1.
  int* parent;        // per contour: parent contour index, -1 = frame
  int* npts;          // per contour
  int* pt_off;        // per contour: offset of its first point (relative to the instance)
  unsigned char* is_hole;
  short* pts;         // (x, y) pairs, window-relative
  // capacities of the tables above (single-pass walk into per-instance slots): contours / points
  // beyond them are counted but not stored
  int cap_contours = 0x7fffffff;
  int cap_points = 0x7fffffff;
};

struct ContourCounts {
  int n_contours;
  int n_points;       // all contours
  int n_rings;        // contours with >= 4 points (prediction.py:236)
  int n_ring_verts;   // their points + the closing point where first != last (:238-239)
};

// How the three bit planes are accessed: through generic pointers (anywhere), or -- when the caller
// KNOWS they sit in shared memory -- with explicit shared-space instructions (a generic access costs a
// trip through the L1TEX pipe even when it lands in shared memory; the walk is a chain of dependent
// plane reads, so that latency is its critical path).
struct GenericMem {
  TD_HD static uint32_t ld(const uint32_t* p) { return *p; }
  TD_HD static void st(uint32_t* p, uint32_t v) { *p = v; }
};
#if defined(__CUDACC__)
struct SharedMem {
  __device__ static uint32_t ld(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"((uint32_t)__cvta_generic_to_shared(p)) : "memory");
    return v;
  }
  __device__ static void st(uint32_t* p, uint32_t v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(p)), "r"(v) : "memory");
  }
};
#endif

TD_HD inline int highest_bit(uint32_t v) {   // index of the highest set bit, v != 0
#if defined(__CUDA_ARCH__)
  return 31 - __clz((int)v);
#else
  return 31 - __builtin_clz(v);
#endif
}

// LabelT: unsigned short in general; unsigned char when the window has < 255 borders (known
// from the count pass), which halves the shared-memory footprint of the emit pass.
template <typename LabelT, typename Mem = GenericMem>
struct RasterT {
  const uint32_t* fg;
  uint32_t* visited;
  uint32_t* right;
  LabelT* label;
  int w, h, wpr;

  TD_HD bool is_fg(int x, int y) const {
    if ((unsigned)x >= (unsigned)w || (unsigned)y >= (unsigned)h) return false;
    return (Mem::ld(fg + (size_t)y * wpr + (x >> 5)) >> (x & 31)) & 1u;
  }
  TD_HD uint32_t word(const uint32_t* plane, int y, int wi) const {
    if (wi < 0 || wi >= wpr) return 0u;
    return Mem::ld(plane + (size_t)y * wpr + wi);
  }
  TD_HD void mark(int x, int y, bool right_flag, int lab) {
    const size_t wi = (size_t)y * wpr + (x >> 5);
    const uint32_t bit = 1u << (x & 31);
    const uint32_t v = Mem::ld(visited + wi);
    // a pixel left through its right side always takes the border's label, any other pixel only on its first
    // visit (one store site for both cases: lanes of a warp that differ in right_flag do not serialise)
    if (right_flag) Mem::st(right + wi, Mem::ld(right + wi) | bit);
    if (right_flag || !(v & bit)) {
      Mem::st(visited + wi, v | bit);
      if (label) label[(size_t)y * w + x] = (LabelT)lab;
    }
  }
  // label of the nearest visited pixel strictly left of x in row y; -1 if none
  TD_HD int lnbd(int x, int y) const {
    int wi = (x - 1) >> 5;
    if (x <= 0) return -1;
    uint32_t m = word(visited, y, wi);
    const int b = (x - 1) & 31;
    if (b < 31) m &= (2u << b) - 1u;
    while (true) {
      if (m) return (int)label[(size_t)y * w + (wi * 32 + highest_bit(m))];
      if (--wi < 0) return -1;
      m = word(visited, y, wi);
    }
  }
};

using Raster = RasterT<unsigned short>;

TD_HD inline int ctz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)v) - 1;
#else
  return __builtin_ctz(v);
#endif
}

// Follows one border from (x0, y0).  Returns the number of CHAIN_APPROX_SIMPLE
// points; writes them when pts != nullptr; reports first / last point.
template <typename RasterType>
TD_HD inline int follow_border(RasterType& R, int x0, int y0, bool hole, int lab, short* pts, int* first_xy,
                               int* last_xy) {
  // 8-neighbourhood in OpenCV's order (E, NE, N, NW, W, SW, S, SE), two bits per entry (d + 1),
  // so that a direction lookup is two ALU ops instead of a local-memory load
  auto dx = [](int k) { return (int)((0x901Au >> (2 * (k & 7))) & 3u) - 1; };
  auto dy = [](int k) { return (int)((0xA901u >> (2 * (k & 7))) & 3u) - 1; };
  int s_end = hole ? 0 : 4, s = s_end;
  int x1 = x0, y1 = y0;
  bool found = false;
  do {
    s = (s - 1) & 7;
    x1 = x0 + dx(s);
    y1 = y0 + dy(s);
    found = R.is_fg(x1, y1);
  } while (!found && s != s_end);
  int n = 0;
  auto emit = [&](int x, int y) {
    if (pts) { pts[2 * n] = (short)x; pts[2 * n + 1] = (short)y; }
    if (n == 0) { first_xy[0] = x; first_xy[1] = y; }
    last_xy[0] = x; last_xy[1] = y;
    ++n;
  };
  if (!found) {  // isolated pixel
    R.mark(x0, y0, true, lab);
    emit(x0, y0);
    return n;
  }
  int x3 = x0, y3 = y0, x4 = x0, y4 = y0;
  int prev_s = s ^ 4;
  int px = x0, py = y0;
  for (;;) {
    s_end = s;
    int k = s;
    while (k < 15) {
      ++k;
      x4 = x3 + dx(k);
      y4 = y3 + dy(k);
      if (R.is_fg(x4, y4)) break;
    }
    s = k & 7;
    // the east neighbour was examined and found empty <=> 1 <= s <= s_end
    R.mark(x3, y3, (unsigned)(s - 1) < (unsigned)s_end, lab);
    if (s != prev_s) {
      emit(px, py);
      prev_s = s;
    }
    px += dx(s);
    py += dy(s);
    if (x4 == x0 && y4 == y0 && x3 == x1 && y3 == y1) break;
    x3 = x4;
    y3 = y4;
    s = (s + 4) & 7;
  }
  return n;
}

// Scans one instance raster.  `out` == nullptr: count only (labels are not needed:
// R.label may be null).  Returns counts; counts.n_contours < 0 when the 16-bit label
// space would overflow.
template <typename RasterType>
TD_HD inline ContourCounts scan_instance(RasterType& R, ContourOut* out) {
  ContourCounts cc = {0, 0, 0, 0};
  const int kMaxLabel = 65534;
  for (int y = 0; y < R.h; ++y) {
    int min_o = 0;  // outer starts allowed at positions >= min_o
    int min_h = 0;  // hole origins allowed at positions >= min_h
    for (int wi = 0; wi < R.wpr; ++wi) {
      for (;;) {
        const uint32_t F = R.word(R.fg, y, wi);
        if (!F) break;
        const uint32_t Fl = R.word(R.fg, y, wi - 1), Fr = R.word(R.fg, y, wi + 1);
        const uint32_t V = R.word(R.visited, y, wi), N = R.word(R.right, y, wi);
        const uint32_t prevfg = (F << 1) | (Fl >> 31);
        const uint32_t nextfg = (F >> 1) | (Fr << 31);
        uint32_t O = F & ~V & ~prevfg;   // unvisited pixel after a 0
        uint32_t H = F & ~N & ~nextfg;   // pixel without the right flag before a 0
        const int base = wi * 32;
        if (min_o > base) O &= (min_o - base >= 32) ? 0u : ~((1u << (min_o - base)) - 1u);
        if (min_h > base) H &= (min_h - base >= 32) ? 0u : ~((1u << (min_h - base)) - 1u);
        if (!(O | H)) break;
        const int a = O ? ctz32(O) : 64, b = H ? ctz32(H) : 64;
        const bool hole = !(a <= b);
        const int x = base + (hole ? b : a);
        if (cc.n_contours >= kMaxLabel) { cc.n_contours = -1; return cc; }
        // ---- parent from the label of the last visited pixel on this row ----
        // transition position is x (outer) or x + 1 (hole): "strictly left of it"
        const int ln = out ? R.lnbd(hole ? x + 1 : x, y) : -1;
        int parent = -1;
        if (out && ln >= 0) {
          parent = ln;
          // same kind -> sibling: take its parent (the frame counts as a hole)
          if ((out->is_hole[ln] != 0) == hole) parent = out->parent[ln];
        }
        const int idx = cc.n_contours;
        int first_xy[2] = {0, 0}, last_xy[2] = {0, 0};
        short* pts = out ? out->pts + 2 * (size_t)cc.n_points : nullptr;
        const int np = follow_border(R, x, y, hole, idx, pts, first_xy, last_xy);
        if (out) {
          out->parent[idx] = parent;
          out->npts[idx] = np;
          out->pt_off[idx] = cc.n_points;
          out->is_hole[idx] = hole ? 1 : 0;
        }
        cc.n_contours += 1;
        cc.n_points += np;
        if (np >= 4) {
          cc.n_rings += 1;
          cc.n_ring_verts += np + ((first_xy[0] != last_xy[0] || first_xy[1] != last_xy[1]) ? 1 : 0);
        }
        if (hole) { min_o = x + 1; min_h = x + 1; }
        else { min_o = x + 1; min_h = x; }
      }
    }
  }
  return cc;
}

// OpenCV's output order: pre-order walk, children in reverse order of discovery.
// order[k] = index of the k-th contour.  scratch: last_child, prev_sibling (n each).
TD_HD inline void contour_order(int n, const int* parent, int* last_child, int* prev_sibling, int* order) {
  int root_last = -1;
  for (int c = 0; c < n; ++c) last_child[c] = -1;
  for (int c = 0; c < n; ++c) {
    const int p = parent[c];
    if (p < 0) { prev_sibling[c] = root_last; root_last = c; }
    else { prev_sibling[c] = last_child[p]; last_child[p] = c; }
  }
  int k = 0;
  int node = root_last;
  while (node >= 0) {
    order[k++] = node;
    if (last_child[node] >= 0) { node = last_child[node]; continue; }
    while (node >= 0 && prev_sibling[node] < 0) node = parent[node];
    if (node < 0) break;
    node = prev_sibling[node];
  }
}

}  // namespace td
