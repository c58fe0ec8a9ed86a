"""Torch-tensor wrappers over the C-ABI (device pointers + current stream in,
caller-allocated outputs).  One function per kernel family of SURVEY.md §8a.

Every function requires CUDA tensors and raises otherwise -- there is no CPU path.
"""
from __future__ import annotations

import torch

from . import _lib

M = 28  # Mask R-CNN ROI mask side


def _stream():
    # raw handle of torch's current stream on the current device (the public
    # torch.cuda.current_stream() costs ~17 us per call, this ~1 us)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _ptr(t):
    if t is None:
        return None
    if not t.is_cuda:
        raise _lib.TreedetError("libtreedet needs CUDA tensors (no CPU fallback)")
    if not t.is_contiguous():
        raise _lib.TreedetError("libtreedet needs contiguous tensors")
    return t.data_ptr()


def _chk(t, dtype, name):
    if t.dtype != dtype:
        raise _lib.TreedetError(f"{name}: expected {dtype}, got {t.dtype}")
    return t


# ----------------------------------------------------------------------------
# P2  paste + threshold + pack
# ----------------------------------------------------------------------------
def paste_plan(boxes_net, inst_tile, tile_dims, sizes=None):
    """boxes_net (N,4) f32 in network-input pixels, inst_tile (N,) i32, tile_dims
    (T,4) i32 = [tile_h, tile_w, net_h, net_w].  Returns boxes_px (N,4) f32, win
    (N,4) i32 = [x0, y0, w, h] (w = h = 0 for dropped instances), nwords (N,) i64.
    ``sizes``: optional (3,N) i64 buffer receiving [nwords, w*h, point slot] (``nwords`` is then its
    first row; the point slot sizes the single-pass border walk, :func:`trace_rings_slots`)."""
    n = boxes_net.shape[0]
    dev = boxes_net.device
    _chk(boxes_net, torch.float32, "boxes_net"); _chk(inst_tile, torch.int32, "inst_tile")
    _chk(tile_dims, torch.int32, "tile_dims")
    boxes_px = torch.empty((n, 4), dtype=torch.float32, device=dev)
    win = torch.empty((n, 4), dtype=torch.int32, device=dev)
    nwords = torch.empty((n,), dtype=torch.int64, device=dev) if sizes is None else sizes[0]
    _lib.call("td_paste_plan", _ptr(boxes_net), _ptr(inst_tile), _ptr(tile_dims), n, tile_dims.shape[0],
              _ptr(boxes_px), _ptr(win), _ptr(nwords), None if sizes is None else _ptr(sizes[1]), _stream())
    return boxes_px, win, nwords


def exclusive_offsets(counts):
    """(N,) i64 counts -> (N+1,) i64 offsets (torch plumbing)."""
    off = torch.zeros(counts.shape[0] + 1, dtype=torch.int64, device=counts.device)
    torch.cumsum(counts, 0, out=off[1:])
    return off


def paste_threshold_pack(boxes_px, win, word_off, probs, threshold=0.5, total_words=None):
    """Returns the packed 1-bit rasters (uint32 words viewed as int32 tensor)."""
    n = boxes_px.shape[0]
    if total_words is None:
        total_words = int(word_off[-1].item())
    bits = torch.empty((max(total_words, 1),), dtype=torch.int32, device=boxes_px.device)
    _chk(probs, torch.float32, "probs")
    _lib.call("td_paste_threshold_pack", _ptr(boxes_px), _ptr(win), _ptr(word_off), _ptr(probs), n,
              float(threshold), _ptr(bits), _stream())
    return bits


def paste_values(boxes_px, win, probs):
    """Pasted probabilities of every instance window (float32), for tolerance tests."""
    n = boxes_px.shape[0]
    npx = (win[:, 2].to(torch.int64) * win[:, 3].to(torch.int64))
    off = exclusive_offsets(npx)
    vals = torch.empty((max(int(off[-1].item()), 1),), dtype=torch.float32, device=boxes_px.device)
    _lib.call("td_paste_values", _ptr(boxes_px), _ptr(win), _ptr(off), _ptr(probs), n, _ptr(vals), _stream())
    return vals, off


# ----------------------------------------------------------------------------
# P3  contours -> rings
# ----------------------------------------------------------------------------
class Rings:
    """Ragged polygon rings on the device: ``verts`` (V,2) f64 CRS coordinates,
    ``ring_off`` (R+1,) i64, ``ring_inst`` (R,) i32 index of the producing instance."""

    def __init__(self, verts, ring_off, ring_inst, n_contours=0, n_points=0):
        self.verts, self.ring_off, self.ring_inst = verts, ring_off, ring_inst
        self.n_contours, self.n_points = n_contours, n_points   # totals of the border walk (capacity planning)

    def __len__(self):
        return self.ring_off.shape[0] - 1


def trace_rings(bits, win, word_off, inst_tile, tile_tf, total_words=None):
    """P3: packed rasters -> closed CRS rings in cv2.findContours order.

    tile_tf (T,6) f64: the tile window transforms of the tiles JSON."""
    n = win.shape[0]
    dev = win.device
    if total_words is None:
        total_words = int(word_off[-1].item())
    _chk(tile_tf, torch.float64, "tile_tf")
    planes = torch.empty((2 * max(total_words, 1),), dtype=torch.int32, device=dev)
    counts = torch.empty((n, 4), dtype=torch.int32, device=dev)
    _lib.call("td_trace_count", _ptr(bits), _ptr(win), _ptr(word_off), n, total_words, _ptr(planes), _ptr(counts),
              None, _stream())
    c64 = counts.to(torch.int64)
    # one exclusive scan over the five per-instance sizes (rows of a (5, n) tensor, scanned along
    # the contiguous dimension), one read back for the totals
    sizes = torch.cat([c64.t(), (win[:, 2].to(torch.int64) * win[:, 3].to(torch.int64))[None, :]], dim=0).contiguous()
    offs = torch.zeros((5, n + 1), dtype=torch.int64, device=dev)
    torch.cumsum(sizes, 1, out=offs[:, 1:])
    cont_off, pts_off, ring_base, vert_base, px_off = offs[0], offs[1], offs[2], offs[3], offs[4]
    totals = torch.cat([offs[:, -1], c64[:, 0].min().reshape(1) if n else offs[:1, -1]]).cpu().tolist()
    tc, tp, tr, tv, tpx, cmin = [int(v) for v in totals]
    if n and cmin < 0:
        raise _lib.TreedetError("td_trace_count: more than 65534 borders in one instance window")
    labels = torch.empty((max(tpx, 1),), dtype=torch.int16, device=dev)
    ct_int = torch.empty((6 * max(tc, 1),), dtype=torch.int32, device=dev)
    ct_hole = torch.empty((max(tc, 1),), dtype=torch.uint8, device=dev)
    pts = torch.empty((2 * max(tp, 1),), dtype=torch.int16, device=dev)
    ring_off = torch.empty((tr + 1,), dtype=torch.int64, device=dev)
    ring_inst = torch.empty((max(tr, 1),), dtype=torch.int32, device=dev)[:tr]
    verts = torch.empty((max(tv, 1), 2), dtype=torch.float64, device=dev)[:tv]
    ring_off[tr] = tv
    _lib.call("td_trace_emit", _ptr(bits), _ptr(win), _ptr(word_off), n, total_words, _ptr(planes), _ptr(labels),
              _ptr(px_off), _ptr(cont_off), _ptr(pts_off), _ptr(ring_base), _ptr(vert_base), _ptr(ct_int),
              _ptr(ct_hole), _ptr(pts), tc, _ptr(inst_tile), _ptr(tile_tf), _ptr(ring_off), ring_inst.data_ptr(),
              verts.data_ptr(), _stream())
    return Rings(verts, ring_off, ring_inst, tc, tp)


def trace_rings_dyn(bits, win, word_off, px_off, inst_tile, tile_tf, caps, flag, totals):
    """Capacity form of :func:`trace_rings` (no synchronisation).  ``caps``: dict with the capacities
    ``words, px, contours, points, rings, verts``; ``totals`` (4,) i64 device slot receiving the live
    [contours, points, rings, vertices].  Returns Rings whose arrays have capacity length: ``ring_off``
    (rings + 2,), ``ring_inst`` (rings + 1,), ``verts`` (verts, 2); rings past the live count are empty."""
    n = win.shape[0]
    dev = win.device
    _chk(tile_tf, torch.float64, "tile_tf")
    cw, cpx, cc, cp, cr, cv = (int(caps[k]) for k in ("words", "px", "contours", "points", "rings", "verts"))
    planes = torch.empty((2 * max(cw, 1),), dtype=torch.int32, device=dev)
    counts = torch.empty((n, 4), dtype=torch.int32, device=dev)
    sizes = torch.empty((4, n), dtype=torch.int64, device=dev)
    _lib.call("td_trace_count", _ptr(bits), _ptr(win), _ptr(word_off), n, cw, _ptr(planes), _ptr(counts), _ptr(sizes),
              _stream())
    offs, _ = scan_clamp(sizes, [cc, cp, cr, cv], flag, totals=totals)
    labels = torch.empty((max(cpx, 1),), dtype=torch.int16, device=dev)
    ct_int = torch.empty((6 * max(cc, 1),), dtype=torch.int32, device=dev)
    ct_hole = torch.empty((max(cc, 1),), dtype=torch.uint8, device=dev)
    pts = torch.empty((2 * max(cp, 1),), dtype=torch.int16, device=dev)
    ring_off = torch.empty((cr + 2,), dtype=torch.int64, device=dev)
    ring_inst = torch.empty((cr + 1,), dtype=torch.int32, device=dev)
    verts = torch.empty((max(cv, 1), 2), dtype=torch.float64, device=dev)
    _lib.call("td_trace_emit", _ptr(bits), _ptr(win), _ptr(word_off), n, cw, _ptr(planes), _ptr(labels),
              _ptr(px_off), _ptr(offs[0]), _ptr(offs[1]), _ptr(offs[2]), _ptr(offs[3]), _ptr(ct_int),
              _ptr(ct_hole), _ptr(pts), cc, _ptr(inst_tile), _ptr(tile_tf), _ptr(ring_off), _ptr(ring_inst),
              _ptr(verts), _stream())
    _lib.call("td_ring_tail", _ptr(ring_off), _ptr(ring_inst), cr + 1, _ptr(totals[2:3]), _ptr(totals[3:4]), _stream())
    return Rings(verts, ring_off, ring_inst)


TRACE_SLOT_CONTOURS = 16     # contour table rows per instance of the single-pass walk


def trace_rings_slots(bits, win, word_off, px_off, slot_off, inst_tile, tile_tf, caps, flag, totals):
    """Single-pass capacity form of :func:`trace_rings`: ONE border walk into per-instance slots
    (td_trace_walk), a scan of the ring / vertex counts, then td_trace_rings.  ``slot_off`` (N+1,):
    point slots; ``caps``: words, px, ptslots, rings, verts; ``totals`` (2,) i64 device slot receiving
    the live [rings, vertices].  An instance that outgrows its slot raises bit 2 of ``flag``."""
    n = win.shape[0]
    dev = win.device
    _chk(tile_tf, torch.float64, "tile_tf")
    cw, cpx, cps, cr, cv = (int(caps[k]) for k in ("words", "px", "ptslots", "rings", "verts"))
    cc = TRACE_SLOT_CONTOURS
    planes = torch.empty((2 * max(cw, 1),), dtype=torch.int32, device=dev)
    labels = torch.empty((max(cpx, 1),), dtype=torch.int16, device=dev)
    ct_int = torch.empty((6 * max(n * cc, 1),), dtype=torch.int32, device=dev)
    ct_hole = torch.empty((max(n * cc, 1),), dtype=torch.uint8, device=dev)
    pts = torch.empty((2 * max(cps, 1),), dtype=torch.int16, device=dev)
    counts = torch.empty((n, 4), dtype=torch.int32, device=dev)
    sizes = torch.empty((2, n), dtype=torch.int64, device=dev)
    _lib.call("td_trace_walk", _ptr(bits), _ptr(win), _ptr(word_off), n, cw, _ptr(planes), _ptr(labels), _ptr(px_off),
              _ptr(slot_off), cc, _ptr(ct_int), _ptr(ct_hole), _ptr(pts), _ptr(counts), _ptr(sizes), _ptr(flag),
              _stream())
    offs, _ = scan_clamp(sizes, [cr, cv], flag, totals=totals)
    ring_off = torch.empty((cr + 2,), dtype=torch.int64, device=dev)
    ring_inst = torch.empty((cr + 1,), dtype=torch.int32, device=dev)
    verts = torch.empty((max(cv, 1), 2), dtype=torch.float64, device=dev)
    _lib.call("td_trace_rings", _ptr(win), n, _ptr(counts), _ptr(slot_off), cc, _ptr(ct_int), _ptr(ct_hole), _ptr(pts),
              _ptr(offs[0]), _ptr(offs[1]), _ptr(inst_tile), _ptr(tile_tf), _ptr(ring_off), _ptr(ring_inst),
              _ptr(verts), _stream())
    _lib.call("td_ring_tail", _ptr(ring_off), _ptr(ring_inst), cr + 1, _ptr(totals[0:1]), _ptr(totals[1:2]), _stream())
    return Rings(verts, ring_off, ring_inst)


# ----------------------------------------------------------------------------
# P5  NDVI + decimation
# ----------------------------------------------------------------------------
def ndvi_decimate(rgbi, out_h, out_w, out=None):
    """rgbi (bands>=4, H, W) uint8 -> NDVI (out_h, out_w) float32 (into ``out`` when given: a buffer kept
    across images keeps the pointers -- and with them the captured graphs of the chain -- stable)."""
    _chk(rgbi, torch.uint8, "rgbi")
    b, h, w = rgbi.shape
    if out is None:
        out = torch.empty((out_h, out_w), dtype=torch.float32, device=rgbi.device)
    elif out.shape != (out_h, out_w) or out.dtype != torch.float32:
        raise _lib.TreedetError("ndvi_decimate: output buffer of the wrong shape / dtype")
    _lib.call("td_ndvi_decimate", _ptr(rgbi), b, h, w, out_h, out_w, _ptr(out), _stream())
    return out


def decimate_f32(band, out_h, out_w, out=None):
    _chk(band, torch.float32, "band")
    h, w = band.shape
    if out is None:
        out = torch.empty((out_h, out_w), dtype=torch.float32, device=band.device)
    elif out.shape != (out_h, out_w) or out.dtype != torch.float32:
        raise _lib.TreedetError("decimate_f32: output buffer of the wrong shape / dtype")
    _lib.call("td_decimate_f32", _ptr(band), h, w, out_h, out_w, _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------
# P6 / P8
# ----------------------------------------------------------------------------
def bbox_nms_ordered(bounds, conf, area, iou_threshold, area_threshold):
    """bounds (N,4) f64, conf (N,) f64, area (N,) f64 -> removed (N,) uint8."""
    n = bounds.shape[0]
    _chk(bounds, torch.float64, "bounds"); _chk(conf, torch.float64, "conf"); _chk(area, torch.float64, "area")
    removed = torch.empty((n,), dtype=torch.uint8, device=bounds.device)
    _lib.call("td_bbox_nms_ordered", _ptr(bounds), _ptr(conf), _ptr(area), n, float(iou_threshold),
              float(area_threshold), _ptr(removed), _stream())
    return removed


def bbox_nms_ordered_dyn(bounds, conf, area, n_dev, iou_threshold, area_threshold, nbr_cap, flag):
    """Capacity form of :func:`bbox_nms_ordered` (no synchronisation): rows >= *n_dev are ignored,
    bit 1 of ``flag`` is raised when ``nbr_cap`` neighbour entries were not enough."""
    n = bounds.shape[0]
    _chk(bounds, torch.float64, "bounds"); _chk(conf, torch.float64, "conf"); _chk(area, torch.float64, "area")
    removed = torch.empty((n,), dtype=torch.uint8, device=bounds.device)
    _lib.call("td_bbox_nms_ordered_dyn", _ptr(bounds), _ptr(conf), _ptr(area), n, _ptr(n_dev), float(iou_threshold),
              float(area_threshold), int(nbr_cap), _ptr(flag), _ptr(removed), _stream())
    return removed


def mask_iou_clean(bits, word_off, win, tile_org, inst_tile, scores, iou_threshold=0.7, confidence=0.2):
    """Opt-in ``iou_mode: mask``: the rule of the reference's (uncalled) clean_crowns on the packed crown
    rasters of P2 with pixel IoU.  tile_org (T,2) i32 = [col_off, row_off] of every tile window.
    Returns (keep u8 (N,), match i32 (N,), best_iou f32 (N,))."""
    n = win.shape[0]
    dev = win.device
    _chk(win, torch.int32, "win"); _chk(tile_org, torch.int32, "tile_org"); _chk(inst_tile, torch.int32, "inst_tile")
    _chk(scores, torch.float32, "scores"); _chk(word_off, torch.int64, "word_off")
    keep = torch.empty((n,), dtype=torch.uint8, device=dev)
    match = torch.empty((n,), dtype=torch.int32, device=dev)
    best = torch.empty((n,), dtype=torch.float32, device=dev)
    _lib.call("td_mask_iou_clean", _ptr(bits), _ptr(word_off), _ptr(win), _ptr(tile_org), _ptr(inst_tile), _ptr(scores),
              n, float(iou_threshold), float(confidence), _ptr(keep), _ptr(match), _ptr(best), _stream())
    return keep, match, best


# ----------------------------------------------------------------------------
# device-side bookkeeping (counts stay on the device)
# ----------------------------------------------------------------------------
def scan_clamp(sizes, caps, flag, win_zero=None, totals=None):
    """sizes (k,n) i64 -> (offs (k,n+1) i64, totals (k,) i64 on the device); see td_scan_clamp.
    ``totals``: optional (k,) i64 device slot to write into."""
    _chk(sizes, torch.int64, "sizes"); _chk(flag, torch.int64, "flag")
    k, n = sizes.shape
    offs = torch.empty((k, n + 1), dtype=torch.int64, device=sizes.device)
    if totals is None:
        totals = torch.empty((k,), dtype=torch.int64, device=sizes.device)
    caps_h = torch.tensor([int(c) for c in caps], dtype=torch.int64)
    _lib.call("td_scan_clamp", _ptr(sizes), k, n, caps_h.data_ptr(), _ptr(offs), _ptr(totals), _ptr(flag),
              _ptr(win_zero), _stream())
    return offs, totals


def compact_flags(flags, n_dev=None, count=None):
    """flags (n,) u8 / bool -> (sel (n,) i64 with a zero tail, count (1,) i64 on the device)."""
    if flags.dtype == torch.bool:
        flags = flags.view(torch.uint8)
    _chk(flags, torch.uint8, "flags")
    n = flags.shape[0]
    sel = torch.empty((n,), dtype=torch.int64, device=flags.device)
    if count is None:
        count = torch.empty((1,), dtype=torch.int64, device=flags.device)
    _lib.call("td_compact_flags", _ptr(flags), n, _ptr(n_dev), _ptr(sel), _ptr(count), _stream())
    return sel, count


def compact_nonneg(values, n_dev=None, count=None):
    """values (n,) i32 -> (the non-negative ones in order (n,) i64 with a zero tail, count (1,) i64)."""
    _chk(values, torch.int32, "values")
    n = values.shape[0]
    out = torch.empty((n,), dtype=torch.int64, device=values.device)
    if count is None:
        count = torch.empty((1,), dtype=torch.int64, device=values.device)
    _lib.call("td_compact_nonneg", _ptr(values), n, _ptr(n_dev), _ptr(out), _ptr(count), _stream())
    return out, count


def containment(bounds32, threshold, n_dev=None):
    """bounds32 (N,4) f32 -> (ratio_max f32, is_contained u8, num_contained i32)."""
    n = bounds32.shape[0]
    _chk(bounds32, torch.float32, "bounds32")
    dev = bounds32.device
    ratio = torch.empty((n,), dtype=torch.float32, device=dev)
    isc = torch.empty((n,), dtype=torch.uint8, device=dev)
    num = torch.empty((n,), dtype=torch.int32, device=dev)
    _lib.call("td_containment", _ptr(bounds32), n, float(threshold), _ptr(ratio), _ptr(isc), _ptr(num), _ptr(n_dev),
              _stream())
    return ratio, isc, num


# ----------------------------------------------------------------------------
# P7
# ----------------------------------------------------------------------------
STATS_COMBINED, STATS_HEIGHT_ONLY, STATS_NDVI_ONLY = 0, 1, 2


def crown_stats(verts, ring_off, ndvi, height, transform6, mode=STATS_COMBINED, n_dev=None):
    """verts (V,2) f64, ring_off (N+1,) i64; rasters (H,W) f32; transform6 = 6 floats
    (a,b,c,d,e,f).  Returns dict of float32 tensors."""
    n = ring_off.shape[0] - 1
    _chk(verts, torch.float64, "verts"); _chk(ring_off, torch.int64, "ring_off")
    dev = verts.device
    ref = ndvi if mode != STATS_HEIGHT_ONLY else height
    rows, cols = ref.shape
    tf = torch.tensor(list(transform6)[:6], dtype=torch.float64)  # host: passed by value through a pointer
    max_h = hxy = stats = None
    if mode != STATS_NDVI_ONLY:
        _chk(height, torch.float32, "height")
        max_h = torch.empty((n,), dtype=torch.float32, device=dev)
        hxy = torch.empty((n, 2), dtype=torch.float32, device=dev)
    if mode != STATS_HEIGHT_ONLY:
        _chk(ndvi, torch.float32, "ndvi")
        stats = torch.empty((n, 4), dtype=torch.float32, device=dev)
    _lib.call("td_crown_stats", _ptr(verts), _ptr(ring_off), n,
              _ptr(ndvi) if mode != STATS_HEIGHT_ONLY else None,
              _ptr(height) if mode != STATS_NDVI_ONLY else None,
              rows, cols, tf.data_ptr(), mode, _ptr(max_h), _ptr(hxy), _ptr(stats), _ptr(n_dev), _stream())
    return {"max_h": max_h, "hxy": hxy, "ndvi": stats}


def crown_height_summary(verts, ring_off, height, transform6, q=95.0, n_dev=None):
    """Optional nDSM summary per crown (the reference computes the maximum only): (N,4) f32 = [min, mean,
    percentile q (numpy "linear"), pixel count] over the pixel set of get_height_within_polygon."""
    n = ring_off.shape[0] - 1
    _chk(verts, torch.float64, "verts"); _chk(ring_off, torch.int64, "ring_off"); _chk(height, torch.float32, "height")
    out = torch.empty((n, 4), dtype=torch.float32, device=verts.device)
    tf = torch.tensor(list(transform6)[:6], dtype=torch.float64)
    _lib.call("td_crown_height_summary", _ptr(verts), _ptr(ring_off), n, _ptr(height), height.shape[0], height.shape[1],
              tf.data_ptr(), float(q), _ptr(out), _ptr(n_dev), _stream())
    return out


def centroids(verts, ring_off, n_dev=None):
    n = ring_off.shape[0] - 1
    out = torch.empty((n, 2), dtype=torch.float32, device=verts.device)
    _lib.call("td_centroids", _ptr(verts), _ptr(ring_off), n, _ptr(out), _ptr(n_dev), _stream())
    return out


# ----------------------------------------------------------------------------
# P4 / P9 geometry
# ----------------------------------------------------------------------------
def simplify_rings(verts, ring_off, tolerance, boxes=None, ring_box=None, want_bounds=True, want_area=False,
                   bounds_of_input=False, n_dev=None):
    """GEOS-semantics simplify(tol, preserve_topology=True) of every ring.

    Returns dict(count i32 (R,), bounds f64 (R,4) of the simplified ring (of the INPUT ring with
    ``bounds_of_input``), area f64 (R,) of the simplified ring,
    keep u8 (R,) = simplified ring within boxes[ring_box] (all ones without boxes),
    scratch = kept-vertex index lists for :func:`take_rings`)."""
    n = ring_off.shape[0] - 1
    dev = verts.device
    nv = verts.shape[0]
    _chk(verts, torch.float64, "verts"); _chk(ring_off, torch.int64, "ring_off")
    scratch = torch.empty((5 * max(nv, 1),), dtype=torch.int32, device=dev)
    alive = torch.empty((nv // 32 + n + 2,), dtype=torch.int32, device=dev)
    count = torch.empty((n,), dtype=torch.int32, device=dev)
    bounds = torch.empty((n, 4), dtype=torch.float64, device=dev) if want_bounds else None
    area = torch.empty((n,), dtype=torch.float64, device=dev) if want_area else None
    keep = torch.empty((n,), dtype=torch.uint8, device=dev)
    if boxes is not None:
        _chk(boxes, torch.float64, "boxes"); _chk(ring_box, torch.int32, "ring_box")
    _lib.call("td_simplify_rings", _ptr(verts), _ptr(ring_off), n, float(tolerance), _ptr(scratch), _ptr(alive),
              _ptr(boxes), _ptr(ring_box), _ptr(count), _ptr(bounds), _ptr(area), _ptr(keep),
              1 if bounds_of_input else 0, _ptr(n_dev), _stream())
    return {"count": count, "bounds": bounds, "area": area, "keep": keep, "scratch": scratch}


def take_rings(verts, ring_off, sel, scratch=None, count=None):
    """Rings ``sel`` (i64 indices) of a ragged ring set -> (verts, ring_off) of the subset.
    With ``scratch``/``count`` from :func:`simplify_rings` only the kept vertices are copied."""
    dev = verts.device
    sel = sel.to(torch.int64).contiguous()
    if count is not None:
        lens = count.to(torch.int64)[sel]
    else:
        lens = (ring_off[1:] - ring_off[:-1])[sel]
    dst_off = exclusive_offsets(lens)
    total = int(dst_off[-1].item())
    out = torch.empty((max(total, 1), 2), dtype=torch.float64, device=dev)[:total]
    _lib.call("td_take_rings", _ptr(verts), _ptr(ring_off), _ptr(sel), sel.shape[0], _ptr(scratch), _ptr(dst_off),
              out.data_ptr(), None, _stream())
    return out, dst_off


def take_rings_dyn(verts, ring_off, sel, n_dev, out_cap, scratch=None, count=None):
    """Capacity form of :func:`take_rings`: ``sel`` has capacity length (tail = 0), ``n_dev`` is the
    live count on the device, ``out_cap`` bounds the output vertices (the caller knows a bound: the
    input vertex count).  Returns (verts (out_cap,2), dst_off (cap+1,)); rows / offsets past the live
    count are undefined.  No synchronisation."""
    dev = verts.device
    dst_off = torch.empty((sel.shape[0] + 1,), dtype=torch.int64, device=dev)
    _lib.call("td_ring_offsets", _ptr(ring_off), _ptr(count), _ptr(sel), sel.shape[0], _ptr(dst_off), _stream())
    out = torch.empty((max(out_cap, 1), 2), dtype=torch.float64, device=dev)
    _lib.call("td_take_rings", _ptr(verts), _ptr(ring_off), _ptr(sel), sel.shape[0], _ptr(scratch), _ptr(dst_off),
              out.data_ptr(), _ptr(n_dev), _stream())
    return out, dst_off


def gather_rows(arrays, sel, n_dev=None):
    """[a[sel] for a in arrays] in ONE launch (td_gather_rows); arrays: contiguous device tensors whose
    first dimension is the row; ``sel`` (n,) i64 with valid indices everywhere (zero tail)."""
    import ctypes as C
    k = len(arrays)
    n = sel.shape[0]
    outs = [torch.empty((n,) + tuple(a.shape[1:]), dtype=a.dtype, device=a.device) for a in arrays]
    rb = [a.element_size() * (a[0].numel() if a.dim() > 1 else 1) for a in arrays]
    ins_p = (C.c_void_p * k)(*[_ptr(a) for a in arrays])
    outs_p = (C.c_void_p * k)(*[o.data_ptr() for o in outs])
    rb_p = (C.c_int * k)(*rb)
    _lib.call("td_gather_rows", ins_p, outs_p, rb_p, k, _ptr(sel), n, _ptr(n_dev), _stream())
    return outs


# ----------------------------------------------------------------------------
# P9 selection
# ----------------------------------------------------------------------------
def select_crowns(bounds, max_h, ndvi_stats, area, num_contained, is_contained, params, n_dev=None):
    """params: the 14 doubles documented at td_select_crowns.  Returns (pre i32 (N,),
    out_idx i32 (N,)): out_idx[i] = crown emitted by pre-selected crown i, or -1."""
    n = bounds.shape[0]
    dev = bounds.device
    _chk(bounds, torch.float64, "bounds"); _chk(max_h, torch.float32, "max_h")
    _chk(ndvi_stats, torch.float32, "ndvi_stats"); _chk(area, torch.float64, "area")
    _chk(num_contained, torch.int32, "num_contained"); _chk(is_contained, torch.uint8, "is_contained")
    pre = torch.empty((n,), dtype=torch.int32, device=dev)
    out_idx = torch.empty((n,), dtype=torch.int32, device=dev)
    p = torch.tensor([float(v) for v in params] + [0.0] * (14 - len(params)), dtype=torch.float64)
    _lib.call("td_select_crowns", _ptr(bounds), _ptr(max_h), _ptr(ndvi_stats), _ptr(area), _ptr(num_contained),
              _ptr(is_contained), n, p.data_ptr(), _ptr(pre), _ptr(out_idx), _ptr(n_dev), _stream())
    return pre, out_idx


def select_head(conf, area, conf_thr, area_min, area_max, n_dev=None):
    """P9 head: (flags u8 (N,), poly_id i64 (N,)); see td_select_head."""
    n = conf.shape[0]
    _chk(conf, torch.float64, "conf"); _chk(area, torch.float64, "area")
    flags = torch.empty((n,), dtype=torch.uint8, device=conf.device)
    pid = torch.empty((n,), dtype=torch.int64, device=conf.device)
    _lib.call("td_select_head", _ptr(conf), _ptr(area), n, _ptr(n_dev), float(conf_thr), float(area_min),
              float(area_max), _ptr(flags), _ptr(pid), _stream())
    return flags, pid


def round_coords(verts):
    out = torch.empty_like(verts)
    _lib.call("td_round_coords", _ptr(verts), verts.numel(), _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------
# P1  tile cut + normalise,  P0a  seam strips
# ----------------------------------------------------------------------------
class TilePlan:
    """Device tables of one tiling (td_tile_plan_create): re-used by every image / step."""

    def __init__(self, tile_win, tile_net, elem_size, H, W):
        import ctypes as C
        if tile_win.is_cuda or tile_net.is_cuda:
            raise _lib.TreedetError("tile tables are host tensors (they come from the tiles JSON)")
        self.win = tile_win.to(torch.int32).contiguous()
        self.net = tile_net.to(torch.int32).contiguous()
        t = self.win.shape[0]
        sizes = 3 * self.net[:, 0].to(torch.int64) * self.net[:, 1].to(torch.int64)
        self.out_off = torch.zeros(t + 1, dtype=torch.int64)
        torch.cumsum(sizes, 0, out=self.out_off[1:])
        self.total = int(self.out_off[-1])
        self.n_tiles, self.elem_size, self.H, self.W = t, elem_size, H, W
        self._h = C.c_void_p()
        _lib.call("td_tile_plan_create", self.win.data_ptr(), self.net.data_ptr(), self.out_off.data_ptr(), t,
                  elem_size, H, W, C.byref(self._h))

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().td_tile_plan_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def run(self, image, out=None):
        b, h, w = image.shape
        if (h, w) != (self.H, self.W) or image.element_size() != self.elem_size:
            raise _lib.TreedetError("tile plan was made for another raster shape / dtype")
        if out is None:
            out = torch.empty((max(self.total, 1),), dtype=torch.float32, device=image.device)
        elif out.numel() < self.total:
            raise _lib.TreedetError("tile_cut_normalize: output buffer too small")
        flag = torch.empty((self.n_tiles,), dtype=torch.uint8, device=image.device)
        _lib.call("td_tile_cut_normalize", self._h, _ptr(image), b, _ptr(out), _ptr(flag), _stream())
        return out, self.out_off, flag


def tile_cut_normalize(image, tile_win, tile_net, out=None):
    """image (bands,H,W) uint8 / int16-viewed-uint16 device tensor; tile_win (T,4) and
    tile_net (T,2) int32 HOST tensors.  Returns (out f32 flat, out_off i64 host, rescale16 u8).
    One-shot convenience; hold a :class:`TilePlan` when the tiling is re-used."""
    elem = image.element_size()
    if elem not in (1, 2):
        raise _lib.TreedetError("tile_cut_normalize: uint8 or uint16 rasters only")
    _, h, w = image.shape
    return TilePlan(tile_win, tile_net, elem, h, w).run(image, out)


def seam_crop(a, b, axis, strip_w, strip_h, out=None):
    """a, b: (bands,H,W) device rasters of the same dtype; axis 0 = b is the right
    neighbour, 1 = b is the lower neighbour.  Returns (bands, strip_h, strip_w) (``out`` when given)."""
    if a.dtype != b.dtype:
        raise _lib.TreedetError("seam_crop: dtype mismatch")
    bands, ha, wa = a.shape
    _, hb, wb = b.shape
    if out is None:
        out = torch.empty((bands, strip_h, strip_w), dtype=a.dtype, device=a.device)
    elif tuple(out.shape) != (bands, strip_h, strip_w) or out.dtype != a.dtype:
        raise _lib.TreedetError("seam_crop: output buffer of the wrong shape / dtype")
    _lib.call("td_seam_crop", _ptr(a), _ptr(b), a.element_size(), bands, ha, wa, hb, wb, axis, strip_w, strip_h,
              _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------
# P10  forest-outline predicates
# ----------------------------------------------------------------------------
def forest_predicates(a_verts, a_off, f_verts, f_off, f_bounds=None, a_filter=None, f_poly_off=None):
    """Query rings (crowns / tile boxes) against the union of the forest polygons.  ``f_poly_off`` (n_poly + 1,)
    i64 groups the forest rings into polygons (first ring = shell, the others = holes); without it every ring
    is a polygon.  ``f_bounds`` (n_poly, 4): bounds of the shells.  Returns (intersects u8 (Na,), within u8
    (Na,)); raises when an edge of a query crosses more forest edges than the kernel's split buffer holds."""
    na = a_off.shape[0] - 1
    n_rings = f_off.shape[0] - 1
    n_poly = n_rings if f_poly_off is None else f_poly_off.shape[0] - 1
    dev = a_verts.device
    if f_bounds is None and n_poly > 0:
        rb = simplify_rings(f_verts, f_off, 0.0, want_bounds=True)["bounds"]
        f_bounds = rb if f_poly_off is None else rb[f_poly_off[:-1]].contiguous()
    if f_poly_off is not None:
        _chk(f_poly_off, torch.int64, "f_poly_off")
    inter = torch.zeros((na,), dtype=torch.uint8, device=dev)
    within = torch.zeros((na,), dtype=torch.uint8, device=dev)
    _lib.call("td_forest_predicates", _ptr(a_verts), _ptr(a_off), na, _ptr(f_verts) if n_poly else None,
              _ptr(f_off) if n_poly else None, _ptr(f_poly_off) if (n_poly and f_poly_off is not None) else None,
              _ptr(f_bounds) if n_poly else None, n_poly, _ptr(a_filter), _ptr(inter), _ptr(within), _stream())
    if na and int(within.max().item()) > 1:
        raise _lib.TreedetError("td_forest_predicates: an edge crosses more than 62 forest edges (split buffer)")
    return inter, within


def rings_are_simple(verts, ring_off):
    """u8 (R,): 1 where the ring is a valid polygon shell (GEOS ``is_valid`` of a single-ring polygon)."""
    n = ring_off.shape[0] - 1
    _chk(verts, torch.float64, "verts"); _chk(ring_off, torch.int64, "ring_off")
    out = torch.empty((n,), dtype=torch.uint8, device=verts.device)
    _lib.call("td_ring_is_simple", _ptr(verts), _ptr(ring_off), n, _ptr(out), _stream())
    return out


# ----------------------------------------------------------------------------
# The whole chain of one image (td_chain_*): P2-P4 and P6-P9 as CUDA-graph replays
# ----------------------------------------------------------------------------
class Chain:
    """One ``td_chain`` object: a device workspace sized by capacities plus the captured graphs.
    ``params``: pipeline.PipelineParams; ``caps``: [instances, words, label pixels, point slots, rings,
    vertices, NMS neighbour slots per crown, contour rows per instance]."""

    _OUT = (("counters", torch.int64, 16, 1), ("tverts", torch.float64, "V", 2), ("toff", torch.int64, "R1", 1),
            ("tconf", torch.float64, "R", 1), ("verts", torch.float64, "V", 2), ("ring_off", torch.int64, "R1", 1),
            ("poly_id", torch.int64, "R", 1), ("conf", torch.float64, "R", 1), ("area", torch.float64, "R", 1),
            ("tree_height", torch.float32, "R", 1), ("centroid", torch.float32, "R", 2),
            ("is_contained", torch.uint8, "R", 1), ("num_contained", torch.int32, "R", 1),
            ("hxy", torch.float32, "R", 2), ("ndvi_stats", torch.float32, "R", 4))

    def __init__(self, params, caps, n_slots, device):
        import ctypes as C
        self.caps = [int(c) for c in caps]
        self.n_slots = int(n_slots)
        self.device = torch.device(device)
        caps_h = (C.c_longlong * 8)(*self.caps)
        nbytes = _lib.lib().td_chain_workspace_bytes(caps_h, self.n_slots)
        if nbytes < 0:
            raise _lib.TreedetError(f"td_chain_workspace_bytes: bad capacities {self.caps}")
        self.workspace = torch.empty((int(nbytes) + 256,), dtype=torch.uint8, device=self.device)
        base = self.workspace.data_ptr()
        self._pad = (-base) % 256
        cfg = (C.c_double * 16)(float(params.mask_threshold), float(params.simplify_tolerance), 2.0,
                                float(params.confidence_threshold), float(params.area_threshold), 1000.0,
                                float(params.iou_threshold), float(params.area_threshold),
                                float(params.containment_threshold), *([0.0] * 7))
        self._h = C.c_void_p()
        _lib.call("td_chain_create", cfg, caps_h, self.n_slots, base + self._pad, int(nbytes), C.byref(self._h))
        self._views = []
        dims = {"R": self.caps[4], "R1": self.caps[4] + 1, "V": self.caps[5]}
        for s in range(self.n_slots):
            offs = (C.c_longlong * 16)()
            _lib.call("td_chain_layout", self._h, s, offs)
            v = {}
            for k, (name, dt, rows, cols) in enumerate(self._OUT):
                r = rows if isinstance(rows, int) else dims[rows]
                nb = r * cols * torch.empty((), dtype=dt).element_size()
                o = self._pad + int(offs[k])
                t = self.workspace[o:o + nb].view(dt)
                v[name] = t.view(r, cols) if cols > 1 else t
            self._views.append(v)

    def __del__(self):
        try:
            if getattr(self, "_h", None) is not None and self._h.value:
                _lib.lib().td_chain_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def predict(self, slot, det, tile_tf, tile_boxes, use_graph=True):
        n = det["boxes_net"].shape[0]
        _chk(det["boxes_net"], torch.float32, "boxes_net"); _chk(det["scores"], torch.float32, "scores")
        _chk(det["probs"], torch.float32, "probs"); _chk(det["inst_tile"], torch.int32, "inst_tile")
        _chk(det["tile_dims"], torch.int32, "tile_dims"); _chk(tile_tf, torch.float64, "tile_tf")
        _chk(tile_boxes, torch.float64, "tile_boxes")
        dummy = self.workspace      # n == 0: the pointers are never dereferenced but must not be null
        ptr = (lambda t: _ptr(t)) if n else (lambda t: _ptr(t) if t.numel() else dummy.data_ptr())
        _lib.call("td_chain_predict", self._h, int(slot), ptr(det["boxes_net"]), ptr(det["scores"]), ptr(det["probs"]),
                  ptr(det["inst_tile"]), int(n), _ptr(det["tile_dims"]), _ptr(tile_tf), _ptr(tile_boxes),
                  int(det["tile_dims"].shape[0]), 1 if use_graph else 0, _stream())

    def post(self, slot, ndvi, ndvi_tf, height, height_tf, combined, select_params, use_graph=True):
        import ctypes as C
        _chk(ndvi, torch.float32, "ndvi"); _chk(height, torch.float32, "height")
        ntf = (C.c_double * 6)(*[float(v) for v in list(ndvi_tf)[:6]])
        htf = (C.c_double * 6)(*[float(v) for v in list(height_tf)[:6]])
        sp = (C.c_double * 14)(*([float(v) for v in select_params] + [0.0] * (14 - len(select_params))))
        _lib.call("td_chain_post", self._h, int(slot), _ptr(ndvi), int(ndvi.shape[0]), int(ndvi.shape[1]), ntf,
                  _ptr(height), int(height.shape[0]), int(height.shape[1]), htf, 1 if combined else 0, sp,
                  1 if use_graph else 0, _stream())

    def counters(self, slot):
        return self._views[slot]["counters"]

    def table(self, slot):
        v = self._views[slot]
        return v["tverts"], v["toff"], v["tconf"]

    def features(self, slot):
        from .pipeline import Features
        v = self._views[slot]
        return Features(v["verts"], v["ring_off"], v["poly_id"], v["conf"], v["area"], v["tree_height"], v["centroid"],
                        v["is_contained"], v["num_contained"], {"hxy": v["hxy"], "ndvi_stats": v["ndvi_stats"]})
