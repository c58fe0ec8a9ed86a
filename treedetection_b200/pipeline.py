"""Device-resident crown pipeline for one image: P2 -> P3 -> P4 (predict side) and
P5 -> P9 (post-processing side), composed from the C-ABI kernels in :mod:`ops`.

Stage boundaries mirror the reference's on-disk artefacts:

* ``predict_stage``   = ``Predictor._process_and_save_single`` + ``process_and_stitch_predictions``
  (TreeDetection/prediction.py:197-265, helpers.py:419-600): raw ROI-head outputs ->
  the crown table of ``geojson_predictions/<image>.gpkg`` (rings + Confidence_score).
* ``postprocess_stage`` = ``process_geojson`` + ``process_features``
  (TreeDetection/postprocessing.py:722-809, 478-720): crown table + rasters ->
  the features of ``processed_<image>.gpkg``.

Everything stays on the GPU between the two; only sizes needed for allocation are read
back.  PyTorch is used for device memory, prefix sums and boolean compaction.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import torch

from . import geo, ops


@dataclass
class PipelineParams:
    """The subset of the reference's config.yml schema the path reads
    (TreeDetection/config.py:182-233, example/config.yml)."""
    tile_width: float = 50
    tile_height: float = 50
    buffer: float = 20
    use_overlap: bool = True
    overlapping_tiles_width: int = 3
    overlapping_tiles_height: int = 3
    confidence_threshold: float = 0.3
    containment_threshold: float = 0.75
    height_threshold: float = 3
    ndvi_mean_threshold: float = 0.1
    ndvi_var_threshold: float = 0.1
    iou_threshold: float = 0.6
    area_threshold: float = 1
    ndvi_scaling_factor: float = 0.2
    height_scaling_factor: float = 1.0
    simplify_tolerance: float = 0.2
    shift: int = 1                       # detection.py:240
    mask_threshold: float = 0.5          # detectron2 ROI_HEADS mask threshold

    @classmethod
    def from_config(cls, config: dict):
        names = cls.__dataclass_fields__.keys()
        return cls(**{k: config[k] for k in names if k in config})


@dataclass
class CrownTable:
    """``geojson_predictions/<image>.gpkg`` on the device: ragged rings + confidence."""
    verts: torch.Tensor      # (V, 2) f64
    ring_off: torch.Tensor   # (R + 1,) i64
    conf: torch.Tensor       # (R,) f64  (float(score) of a float32 score)

    def __len__(self):
        return self.ring_off.shape[0] - 1


@dataclass
class Features:
    """``processed_<image>.gpkg`` on the device (schema postprocessing.py:904-919)."""
    verts: torch.Tensor          # (V, 2) f64, rounded to 3 decimals
    ring_off: torch.Tensor       # (K + 1,) i64
    poly_id: torch.Tensor        # (K,) i64
    conf: torch.Tensor           # (K,) f64
    area: torch.Tensor           # (K,) f64
    tree_height: torch.Tensor    # (K,) f32
    centroid: torch.Tensor       # (K, 2) f32
    is_contained: torch.Tensor   # (K,) u8
    num_contained: torch.Tensor  # (K,) i32
    extras: dict = field(default_factory=dict)

    def __len__(self):
        return self.ring_off.shape[0] - 1


def tile_tables(tiles: dict, device):
    """Device tables of the tiles JSON: window transforms (T,6) f64 and the stitch
    filter boxes (T,4) f64 of helpers.py:265-303 (values parsed from the tile id as
    integers; ``width`` is used for both axes)."""
    tfs, boxes = [], []
    for tid, meta in tiles.items():
        tfs.append(list(meta["transform"][:6]))
        parts = [int(p) for p in tid.split("_")[-5:]]
        minx, miny, width, buf = parts[0], parts[1], parts[2], parts[3]
        boxes.append((minx, miny, width, buf))
    tile_tf = torch.tensor(tfs, dtype=torch.float64, device=device).reshape(-1, 6)
    return tile_tf, boxes


def filter_boxes(boxes_int, shift, device):
    rows = [[minx - buf + shift, miny - buf + shift, minx + width + buf - shift, miny + width + buf - shift]
            for (minx, miny, width, buf) in boxes_int]
    return torch.tensor(rows, dtype=torch.float64, device=device).reshape(-1, 4)


def predict_stage(boxes_net, scores, probs, inst_tile, tile_dims, tile_tf, tile_boxes, p: PipelineParams):
    """P2 + P3 + P4.  All arguments are device tensors (see :mod:`synth` for layouts)."""
    boxes_px, win, nwords = ops.paste_plan(boxes_net, inst_tile, tile_dims)
    word_off = ops.exclusive_offsets(nwords)
    total_words = int(word_off[-1].item())
    bits = ops.paste_threshold_pack(boxes_px, win, word_off, probs, p.mask_threshold, total_words)
    rings = ops.trace_rings(bits, win, word_off, inst_tile, tile_tf, total_words)
    return _stitch(rings, scores, inst_tile, tile_boxes, p)


def _stitch(rings, scores, inst_tile, tile_boxes, p: PipelineParams):
    """P4: simplify + tile box filter of the traced rings -> the stitched crown table."""
    n_rings = len(rings)
    dev = scores.device
    if n_rings == 0:
        return CrownTable(torch.zeros((0, 2), dtype=torch.float64, device=dev),
                          torch.zeros((1,), dtype=torch.int64, device=dev),
                          torch.zeros((0,), dtype=torch.float64, device=dev))
    ring_tile = inst_tile[rings.ring_inst.long()].contiguous()
    simp = ops.simplify_rings(rings.verts, rings.ring_off, p.simplify_tolerance, tile_boxes, ring_tile,
                              want_bounds=False)
    sel = torch.nonzero(simp["keep"]).flatten()
    verts, ring_off = ops.take_rings(rings.verts, rings.ring_off, sel, simp["scratch"], simp["count"])
    conf = scores[rings.ring_inst.long()[sel]].to(torch.float64)
    return CrownTable(verts, ring_off, conf)


def raster_stage(rgbi, rgbi_transform, ndsm, ndsm_transform, p: PipelineParams):
    """P5: decimated reads + NDVI (postprocessing.py:780-800).  rgbi (4,H,W) u8,
    ndsm (h,w) f32 on the device; transforms are 6-tuples."""
    _, H, W = rgbi.shape
    oh, ow = int(H * p.ndvi_scaling_factor), int(W * p.ndvi_scaling_factor)
    ndvi = ops.ndvi_decimate(rgbi, oh, ow)
    ndvi_tf = geo.compose(rgbi_transform, geo.scale(W / ow, H / oh))
    ndvi_bounds = geo.raster_bounds(rgbi_transform, W, H)
    h, w = ndsm.shape
    hh, hw = int(h * p.height_scaling_factor), int(w * p.height_scaling_factor)
    height = ops.decimate_f32(ndsm, hh, hw) if (hh, hw) != (h, w) else ndsm
    height_tf = geo.compose(ndsm_transform, geo.scale(w / hw, h / hh))
    height_bounds = geo.raster_bounds(ndsm_transform, w, h)
    return {"ndvi": ndvi, "ndvi_transform": ndvi_tf, "ndvi_bounds": ndvi_bounds,
            "height": height, "height_transform": height_tf, "height_bounds": height_bounds,
            "pixel_x": abs(rgbi_transform[0]), "pixel_y": abs(rgbi_transform[4])}


def _similar_bounds(a, b, tol=1e-3):
    return all(abs(x - y) < tol for x, y in zip(a, b))


def select_params(p: PipelineParams, ndvi_shape, ndvi_bounds, pixel_x, pixel_y):
    """Overlap-band geometry of process_features (postprocessing.py:580-600)."""
    rows, cols = ndvi_shape
    vmh = ((p.tile_height + 2 * p.buffer) * p.overlapping_tiles_height) * pixel_y
    hmw = ((p.tile_width + 2 * p.buffer) * p.overlapping_tiles_width) * pixel_x
    is_seam = (rows == vmh) or (cols == hmw)
    b = ndvi_bounds
    return [1.0 if p.use_overlap else 0.0, 1.0 if is_seam else 0.0, b.left, b.bottom, b.right, b.top,
            b.left + hmw / 2.0, b.right - hmw / 2.0, b.top - vmh / 2.0, b.bottom + vmh / 2.0,
            float(p.height_threshold), float(p.ndvi_mean_threshold), float(p.ndvi_var_threshold)]


def postprocess_stage(table: CrownTable, rasters: dict, p: PipelineParams, keep_debug=False):
    """P9 head, P6, P7, P8, P9 selection + attributes."""
    dev = table.verts.device
    extras = {}

    def empty():
        z = lambda dt, *s: torch.zeros(s, dtype=dt, device=dev)
        return Features(z(torch.float64, 0, 2), z(torch.int64, 1), z(torch.int64, 0), z(torch.float64, 0),
                        z(torch.float64, 0), z(torch.float32, 0), z(torch.float32, 0, 2), z(torch.uint8, 0),
                        z(torch.int32, 0), extras)

    # 1. + 2. confidence filter, poly_id = enumeration index after it, area of simplify(2) in
    #    [area_threshold, 1000] (postprocessing.py:739-768) -- one kernel pass over all stitched rings
    #    (area of simplify(2) + bounds of the ORIGINAL ring), one compaction
    conf_ok = table.conf >= p.confidence_threshold
    s2 = ops.simplify_rings(table.verts, table.ring_off, 2.0, want_bounds=True, want_area=True, bounds_of_input=True)
    area_all = s2["area"]
    pid_all = torch.cumsum(conf_ok.to(torch.int64), 0) - 1
    sel1 = torch.nonzero(conf_ok & (area_all >= p.area_threshold) & (area_all <= 1000)).flatten()
    if sel1.numel() == 0:
        return empty()
    verts1, off1 = ops.take_rings(table.verts, table.ring_off, sel1)
    conf1, area1, pid1 = table.conf[sel1], area_all[sel1].contiguous(), pid_all[sel1]
    b1 = s2["bounds"][sel1].contiguous()
    area0 = area_all
    # 3. ordered bbox NMS (P6)
    removed = ops.bbox_nms_ordered(b1, conf1.contiguous(), area1, p.iou_threshold, p.area_threshold)
    sel2 = torch.nonzero(removed == 0).flatten()
    verts2, off2 = ops.take_rings(verts1, off1, sel2)
    conf2, area2, pid2, b2 = conf1[sel2], area1[sel2].contiguous(), pid1[sel2], b1[sel2].contiguous()
    n2 = sel2.numel()
    if keep_debug:
        extras.update(area0=area0, removed=removed, pid_after_area=pid1, pid_after_nms=pid2)
    if n2 == 0:
        return empty()
    # 4. statistics (P7) + centroids
    cent = ops.centroids(verts2, off2)
    for key in ("height_ready", "ndvi_ready"):       # rasters produced / copied on another stream
        if rasters.get(key) is not None:
            torch.cuda.current_stream().wait_event(rasters[key])
    combined = geo.almost_equals(rasters["height_transform"], rasters["ndvi_transform"]) and \
        _similar_bounds(rasters["height_bounds"], rasters["ndvi_bounds"])
    if combined:
        st = ops.crown_stats(verts2, off2, rasters["ndvi"], rasters["height"], rasters["ndvi_transform"],
                             ops.STATS_COMBINED)
        max_h, hxy, nst = st["max_h"], st["hxy"], st["ndvi"]
    else:
        sh = ops.crown_stats(verts2, off2, None, rasters["height"], rasters["height_transform"], ops.STATS_HEIGHT_ONLY)
        sn = ops.crown_stats(verts2, off2, rasters["ndvi"], None, rasters["ndvi_transform"], ops.STATS_NDVI_ONLY)
        max_h, hxy, nst = sh["max_h"], sh["hxy"], sn["ndvi"]
    # 5. containment (P8) on float32 bounds of ALL post-NMS crowns
    ratio, isc, num = ops.containment(b2.to(torch.float32).contiguous(), p.containment_threshold)
    # 6. selection (P9)
    sp = select_params(p, tuple(rasters["ndvi"].shape), rasters["ndvi_bounds"], rasters["pixel_x"], rasters["pixel_y"])
    pre, out_idx = ops.select_crowns(b2, max_h, nst, area2, num, isc, sp)
    final = out_idx[out_idx >= 0].long()
    if keep_debug:
        extras.update(max_h=max_h, hxy=hxy, ndvi_stats=nst, centroid=cent, is_contained=isc, num_contained=num,
                      containment_ratio=ratio, pre=pre, out_idx=out_idx, combined=combined)
    if final.numel() == 0:
        return empty()
    vf, of = ops.take_rings(verts2, off2, final)
    vf = ops.round_coords(vf)
    return Features(vf, of, pid2[final], conf2[final], area2[final], max_h[final], cent[final], isc[final], num[final],
                    extras)


# ======================================================================================
# Sync-free form of the two stages.
#
# The exact-size functions above read a size back from the device before every allocation
# (about ten host synchronisations per image).  The functions below allocate by CAPACITY
# (remembered from earlier images with the same tiling, see ChainRunner), keep every live
# count on the device (``n_dev`` arguments of the C-ABI) and record them in one small
# counter tensor, so an image is enqueued without waiting for the GPU; the counters are
# read once, when the results are consumed.  Results are bit-identical to the exact-size
# path; an overflow of any capacity raises bit 0 / 1 of the flag and the image is redone
# with exact sizes.
# ======================================================================================
CTR_FLAG, CTR_WORDS, CTR_PX, CTR_SLOTS, CTR_CONT, CTR_PTS, CTR_RINGS, CTR_VERTS = 0, 1, 2, 3, 4, 5, 6, 7
CTR_NTABLE, CTR_VTABLE, CTR_N1, CTR_N2, CTR_NFINAL, CTR_VFINAL, CTR_SIZE = 8, 9, 10, 11, 12, 13, 16
# the border walk of the capacity form: one pass into per-instance slots (default) or count + emit
TRACE_TWO_PASS = bool(__import__("os").environ.get("TREEDET_TRACE_TWO_PASS"))


@dataclass
class DynTable:
    """Capacity-sized CrownTable: rows / vertices past the live counts are undefined."""
    verts: torch.Tensor      # (cap_v, 2) f64
    ring_off: torch.Tensor   # (cap_r + 1,) i64
    conf: torch.Tensor       # (cap_r,) f64
    n_dev: torch.Tensor      # (1,) i64 live ring count (a view into the counter tensor)


def new_counters(device):
    return torch.zeros((CTR_SIZE,), dtype=torch.int64, device=device)


def predict_stage_dyn(boxes_net, scores, probs, inst_tile, tile_dims, tile_tf, tile_boxes, p: PipelineParams,
                      caps: dict, ctr: torch.Tensor, mark=None, two_pass=False) -> DynTable:
    """P2 + P3 + P4 without host synchronisation (see the section comment).  ``mark("p3")`` is called
    once the border walk is enqueued (schedulers use it to start work that should not share the SMs
    with the shared-memory hungry walk)."""
    flag = ctr[CTR_FLAG:CTR_FLAG + 1]
    sizes1 = torch.empty((3, boxes_net.shape[0]), dtype=torch.int64, device=boxes_net.device)
    boxes_px, win, _ = ops.paste_plan(boxes_net, inst_tile, tile_dims, sizes=sizes1)
    offs1, _ = ops.scan_clamp(sizes1, [caps["words"], caps["px"], caps["ptslots"]], flag, win_zero=win,
                              totals=ctr[CTR_WORDS:CTR_SLOTS + 1])
    word_off, px_off, slot_off = offs1[0], offs1[1], offs1[2]
    bits = ops.paste_threshold_pack(boxes_px, win, word_off, probs, p.mask_threshold, int(caps["words"]))
    if TRACE_TWO_PASS or two_pass:
        rings = ops.trace_rings_dyn(bits, win, word_off, px_off, inst_tile, tile_tf, caps, flag,
                                    ctr[CTR_CONT:CTR_VERTS + 1])
    else:      # one border walk into per-instance slots
        rings = ops.trace_rings_slots(bits, win, word_off, px_off, slot_off, inst_tile, tile_tf, caps, flag,
                                      ctr[CTR_RINGS:CTR_VERTS + 1])
    if mark is not None:
        mark("p3")
    n_rings = ctr[CTR_RINGS:CTR_RINGS + 1]
    cap_r = int(caps["rings"])
    ring_inst = rings.ring_inst[:cap_r].long()
    ring_tile = inst_tile[ring_inst].contiguous()
    ring_off = rings.ring_off[:cap_r + 1]
    simp = ops.simplify_rings(rings.verts, ring_off, p.simplify_tolerance, tile_boxes, ring_tile, want_bounds=False,
                              n_dev=n_rings)
    n_table = ctr[CTR_NTABLE:CTR_NTABLE + 1]
    sel, _ = ops.compact_flags(simp["keep"], n_dev=n_rings, count=n_table)
    verts, dst_off = ops.take_rings_dyn(rings.verts, ring_off, sel, n_table, int(caps["verts"]), simp["scratch"],
                                        simp["count"])
    ctr[CTR_VTABLE:CTR_VTABLE + 1] = dst_off.gather(0, n_table)
    conf = scores[ring_inst[sel]].to(torch.float64)
    return DynTable(verts, dst_off, conf, n_table)


def postprocess_stage_dyn(table: DynTable, rasters: dict, p: PipelineParams, caps: dict, ctr: torch.Tensor):
    """P9 head, P6, P7, P8, P9 without host synchronisation.  Returns a capacity-sized Features whose
    live sizes are ctr[CTR_NFINAL] rings / ctr[CTR_VFINAL] vertices."""
    dev = table.verts.device
    cap = table.conf.shape[0]
    cap_v = table.verts.shape[0]
    flag = ctr[CTR_FLAG:CTR_FLAG + 1]
    n0 = table.n_dev
    s2 = ops.simplify_rings(table.verts, table.ring_off, 2.0, want_bounds=True, want_area=True, bounds_of_input=True,
                            n_dev=n0)
    area_all = s2["area"]
    flags1, pid_all = ops.select_head(table.conf, area_all, p.confidence_threshold, p.area_threshold, 1000.0, n_dev=n0)
    n1 = ctr[CTR_N1:CTR_N1 + 1]
    sel1, _ = ops.compact_flags(flags1, count=n1)
    verts1, off1 = ops.take_rings_dyn(table.verts, table.ring_off, sel1, n1, cap_v)
    conf1, area1, pid1, b1 = ops.gather_rows([table.conf, area_all, pid_all, s2["bounds"]], sel1, n1)
    removed = ops.bbox_nms_ordered_dyn(b1, conf1, area1, n1, p.iou_threshold, p.area_threshold,
                                       int(caps["nbr"]), flag)
    n2 = ctr[CTR_N2:CTR_N2 + 1]
    sel2, _ = ops.compact_flags(removed == 0, n_dev=n1, count=n2)
    verts2, off2 = ops.take_rings_dyn(verts1, off1, sel2, n2, cap_v)
    conf2, area2, pid2, b2 = ops.gather_rows([conf1, area1, pid1, b1], sel2, n2)
    cent = ops.centroids(verts2, off2, n_dev=n2)
    for key in ("height_ready", "ndvi_ready"):       # rasters produced / copied on another stream
        if rasters.get(key) is not None:
            torch.cuda.current_stream().wait_event(rasters[key])
    combined = geo.almost_equals(rasters["height_transform"], rasters["ndvi_transform"]) and \
        _similar_bounds(rasters["height_bounds"], rasters["ndvi_bounds"])
    if combined:
        st = ops.crown_stats(verts2, off2, rasters["ndvi"], rasters["height"], rasters["ndvi_transform"],
                             ops.STATS_COMBINED, n_dev=n2)
        max_h, nst = st["max_h"], st["ndvi"]
    else:
        sh = ops.crown_stats(verts2, off2, None, rasters["height"], rasters["height_transform"], ops.STATS_HEIGHT_ONLY,
                             n_dev=n2)
        sn = ops.crown_stats(verts2, off2, rasters["ndvi"], None, rasters["ndvi_transform"], ops.STATS_NDVI_ONLY,
                             n_dev=n2)
        max_h, nst = sh["max_h"], sn["ndvi"]
    ratio, isc, num = ops.containment(b2.to(torch.float32).contiguous(), p.containment_threshold, n_dev=n2)
    sp = select_params(p, tuple(rasters["ndvi"].shape), rasters["ndvi_bounds"], rasters["pixel_x"], rasters["pixel_y"])
    pre, out_idx = ops.select_crowns(b2, max_h, nst, area2, num, isc, sp, n_dev=n2)
    nf = ctr[CTR_NFINAL:CTR_NFINAL + 1]
    final, _ = ops.compact_nonneg(out_idx, n_dev=n2, count=nf)
    vf, of = ops.take_rings_dyn(verts2, off2, final, nf, cap_v)
    ctr[CTR_VFINAL:CTR_VFINAL + 1] = of.gather(0, nf)
    vf = ops.round_coords(vf)
    pidf, conff, areaf, hf, centf, iscf, numf = ops.gather_rows([pid2, conf2, area2, max_h, cent, isc, num], final, nf)
    return Features(vf, of, pidf, conff, areaf, hf, centf, iscf, numf, {})


def trim_features(f: Features, n: int, v: int) -> Features:
    """Views of a capacity-sized Features cut to the live sizes (read from the counters)."""
    return Features(f.verts[:v], f.ring_off[:n + 1], f.poly_id[:n], f.conf[:n], f.area[:n], f.tree_height[:n],
                    f.centroid[:n], f.is_contained[:n], f.num_contained[:n], f.extras)


class ChainRunner:
    """P2-P9 for a stream of images with the same tiling.  The first image (and any image whose
    sizes outgrow the remembered capacities) runs through the exact-size path; the others are
    enqueued without synchronisation.  ``submit`` returns a ticket, ``collect`` turns it into
    (n_candidates, Features) -- the only point where the host waits for the GPU."""

    GROW = 1.25

    def __init__(self, params: PipelineParams):
        self.p = params
        self.caps = None
        self.nbr_per_crown = 8
        self.fallbacks = 0
        self.two_pass = TRACE_TWO_PASS     # border walk: one pass into slots, or count + emit
        self._pinned = []          # recycled pinned read-back buffers (allocating one costs ~0.1 ms)

    def _learn(self, sizes: dict):
        def cap(v):
            return int(v * self.GROW) + 1024
        new = {k: cap(v) for k, v in sizes.items()}
        if self.caps is None:
            self.caps = new
        else:
            self.caps.update({k: max(self.caps.get(k, 0), new[k]) for k in new})
        self.caps["nbr"] = self.nbr_per_crown * self.caps["rings"]

    def _exact(self, det, tile_tf, tile_boxes, rasters_fn):
        """Exact-size path; also measures the sizes the capacities are learnt from."""
        p = self.p
        boxes_px, win, nwords = ops.paste_plan(det["boxes_net"], det["inst_tile"], det["tile_dims"])
        word_off = ops.exclusive_offsets(nwords)
        total_words = int(word_off[-1].item())
        w64, h64 = win[:, 2].to(torch.int64), win[:, 3].to(torch.int64)
        px, slots = [int(v) for v in torch.stack([(w64 * h64).sum(),
                                                  torch.where(w64 * h64 > 0, 4 * (w64 + h64) + 64, 0).sum()]).tolist()]
        bits = ops.paste_threshold_pack(boxes_px, win, word_off, det["probs"], p.mask_threshold, total_words)
        rings = ops.trace_rings(bits, win, word_off, det["inst_tile"], tile_tf, total_words)
        sizes = {"words": total_words, "px": px, "ptslots": slots, "contours": rings.n_contours,
                 "points": rings.n_points,
                 "rings": len(rings), "verts": int(rings.verts.shape[0])}
        self._learn(sizes)
        table = _stitch(rings, det["scores"], det["inst_tile"], tile_boxes, p)
        feats = postprocess_stage(table, rasters_fn(), p)
        return len(table), feats

    def submit(self, det: dict, tile_tf, tile_boxes, rasters_fn, mark=None):
        """det: dict of device tensors boxes_net, scores, probs, inst_tile, tile_dims;
        rasters_fn(): the P5 rasters dict (called after P2-P4 are enqueued);
        mark(name): optional callback at the stage boundaries ("p4", "p5", "p9"), e.g. to record events."""
        mark = mark or (lambda name: None)
        if self.caps is None:
            out = ("done",) + self._exact(det, tile_tf, tile_boxes, rasters_fn)
            for name in ("p3", "p4", "p5", "p9"):
                mark(name)
            return out
        dev = det["boxes_net"].device
        ctr = new_counters(dev)
        table = predict_stage_dyn(det["boxes_net"], det["scores"], det["probs"], det["inst_tile"], det["tile_dims"],
                                  tile_tf, tile_boxes, self.p, self.caps, ctr, mark=mark, two_pass=self.two_pass)
        mark("p4")
        rasters = rasters_fn()
        mark("p5")
        feats = postprocess_stage_dyn(table, rasters, self.p, self.caps, ctr)
        mark("p9")
        host = self._pinned.pop() if self._pinned else torch.empty((CTR_SIZE,), dtype=torch.int64).pin_memory()
        host.copy_(ctr, non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        return ("dyn", ev, host, feats, (det, tile_tf, tile_boxes, rasters_fn))

    def collect(self, ticket):
        if ticket[0] == "done":
            return ticket[1], ticket[2]
        _, ev, host, feats, again = ticket
        ev.synchronize()
        c = host.tolist()
        self._pinned.append(host)
        if c[CTR_FLAG] != 0:
            # some capacity was too small: redo this image with exact sizes (which also re-learns them)
            self.fallbacks += 1
            if c[CTR_FLAG] & 2:
                self.nbr_per_crown *= 2
            if c[CTR_FLAG] & 4:
                # an instance outgrew its slot of the single-pass walk (ragged mask: many borders or a
                # long outline); the same image would do so again, so this runner walks twice from now on
                self.two_pass = True
            return self._exact(*again)
        learnt = {"words": c[CTR_WORDS], "px": c[CTR_PX], "ptslots": c[CTR_SLOTS], "rings": c[CTR_RINGS],
                  "verts": c[CTR_VERTS]}
        if self.two_pass:
            learnt.update(contours=c[CTR_CONT], points=c[CTR_PTS])
        self._learn(learnt)
        return c[CTR_NTABLE], trim_features(feats, c[CTR_NFINAL], c[CTR_VFINAL])
