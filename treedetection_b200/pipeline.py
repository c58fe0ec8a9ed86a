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

import os
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
    # opt-in (not in the reference's schema): "mask" runs the rule of the reference's uncalled clean_crowns
    # (helpers.py:602-701) with pixel IoU on the packed P2 rasters before the contours are traced; its
    # thresholds are clean_crowns' defaults / the config key the reference reserves for it
    iou_mode: str = "bbox"
    mask_iou_threshold: float = 0.7
    confidence_threshold_stitching: float = 0.3

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


def predict_stage(boxes_net, scores, probs, inst_tile, tile_dims, tile_tf, tile_boxes, p: PipelineParams,
                  tile_org=None):
    """P2 + P3 + P4.  All arguments are device tensors (see :mod:`synth` for layouts).  ``tile_org`` (T,2) i32
    [col_off, row_off] of the tile windows is needed by ``iou_mode: mask`` only."""
    boxes_px, win, nwords = ops.paste_plan(boxes_net, inst_tile, tile_dims)
    word_off = ops.exclusive_offsets(nwords)
    total_words = int(word_off[-1].item())
    bits = ops.paste_threshold_pack(boxes_px, win, word_off, probs, p.mask_threshold, total_words)
    if p.iou_mode == "mask":
        if tile_org is None:
            raise ValueError("iou_mode 'mask' needs the tile window origins (tile_org)")
        keep, match, _ = ops.mask_iou_clean(bits, word_off, win, tile_org, inst_tile, scores, p.mask_iou_threshold,
                                            p.confidence_threshold_stitching)
        # a dropped instance gets an empty window (no contours); a kept one takes the confidence of its match
        # (itself, or an exact duplicate with a higher confidence: clean_crowns appends the MATCH)
        win = torch.where(keep.bool()[:, None], win, torch.zeros_like(win)).contiguous()
        scores = scores[match.clamp(min=0).long()].contiguous()
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


def raster_stage(rgbi, rgbi_transform, ndsm, ndsm_transform, p: PipelineParams, buffers: dict = None):
    """P5: decimated reads + NDVI (postprocessing.py:780-800).  rgbi (4,H,W) u8,
    ndsm (h,w) f32 on the device; transforms are 6-tuples.  ``buffers``: optional dict that keeps the
    output rasters ("ndvi", "height") across images of the same shape."""
    _, H, W = rgbi.shape
    oh, ow = int(H * p.ndvi_scaling_factor), int(W * p.ndvi_scaling_factor)

    def buf(name, shape):
        if buffers is None:
            return None
        b = buffers.get(name)
        if b is None or tuple(b.shape) != shape or b.device != rgbi.device:
            b = buffers[name] = torch.empty(shape, dtype=torch.float32, device=rgbi.device)
        return b
    ndvi = ops.ndvi_decimate(rgbi, oh, ow, out=buf("ndvi", (oh, ow)))
    ndvi_tf = geo.compose(rgbi_transform, geo.scale(W / ow, H / oh))
    ndvi_bounds = geo.raster_bounds(rgbi_transform, W, H)
    h, w = ndsm.shape
    hh, hw = int(h * p.height_scaling_factor), int(w * p.height_scaling_factor)
    height = ops.decimate_f32(ndsm, hh, hw, out=buf("height", (hh, hw))) if (hh, hw) != (h, w) else ndsm
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
# Sync-free form of the two stages: td_chain_* (csrc/chain.cu).
#
# The exact-size functions above read a size back from the device before every allocation
# (about ten host synchronisations per image).  ChainRunner drives the C-ABI chain instead:
# every variable-length buffer lives BY CAPACITY (remembered from earlier images with the
# same tiling) in one device workspace, live counts stay on the device in a 16-slot counter
# block, and the static launch sequence of P2-P4 / P6-P9 is a CUDA graph replayed with one
# launch each -- the host neither waits for the GPU nor sits between two kernels.  The
# counters are read once, when the results are consumed.  Results are bit-identical to the
# exact-size path; an overflow of any capacity raises a bit of the flag counter and the
# image is redone with exact sizes.
# ======================================================================================
CTR_FLAG, CTR_WORDS, CTR_PX, CTR_SLOTS, CTR_CONT, CTR_PTS, CTR_RINGS, CTR_VERTS = 0, 1, 2, 3, 4, 5, 6, 7
CTR_NTABLE, CTR_VTABLE, CTR_N1, CTR_N2, CTR_NFINAL, CTR_VFINAL, CTR_SIZE = 8, 9, 10, 11, 12, 13, 16


def trim_features(f: Features, n: int, v: int) -> Features:
    """Views of a capacity-sized Features cut to the live sizes (read from the counters)."""
    return Features(f.verts[:v], f.ring_off[:n + 1], f.poly_id[:n], f.conf[:n], f.area[:n], f.tree_height[:n],
                    f.centroid[:n], f.is_contained[:n], f.num_contained[:n],
                    {k: t[:n] for k, t in f.extras.items()})


class ChainRunner:
    """P2-P9 for a stream of images with the same tiling.  The first image (and any image whose
    sizes outgrow the remembered capacities) runs through the exact-size path; the others are
    enqueued through ``td_chain_predict`` / ``td_chain_post`` without synchronisation.  ``submit``
    returns a ticket, ``collect`` turns it into (n_candidates, Features) -- the only point where the
    host waits for the GPU.  The Features of a ticket are views into the chain's workspace: they
    stay valid until ``n_slots`` further images have been submitted."""

    GROW = 1.25

    def __init__(self, params: PipelineParams, n_slots: int = 4, use_graph: bool = True):
        self.p = params
        self.caps = None
        self.nbr_per_crown = 8
        self.slot_contours = ops.TRACE_SLOT_CONTOURS
        self.fallbacks = 0
        self.n_slots = n_slots
        self.use_graph = use_graph and not os.environ.get("TREEDET_NO_GRAPH")
        self.chain = None
        self._seq = 0
        self._busy = [None] * n_slots   # ticket id occupying each slot
        self.last_table = None          # stitched table (CrownTable) of the image collected last
        self.last_counters = None
        self._pinned = []               # recycled pinned read-back buffers (allocating one costs ~0.1 ms)

    def _learn(self, sizes: dict):
        def cap(v):
            return int(v * self.GROW) + 1024
        new = {k: cap(v) for k, v in sizes.items()}
        if self.caps is None:
            grown = True
            self.caps = new
        else:
            grown = any(new[k] > self.caps.get(k, 0) for k in new)
            self.caps.update({k: max(self.caps.get(k, 0), new[k]) for k in new})
        if grown and self.chain is not None:
            self.chain = None           # rebuilt (with the larger workspace) by the next submit

    def _chain(self, device):
        if self.chain is None:
            c = self.caps
            self.chain = ops.Chain(self.p, [c["inst"], c["words"], c["px"], c["ptslots"], c["rings"], c["verts"],
                                            self.nbr_per_crown, self.slot_contours], self.n_slots, device)
            self._busy = [None] * self.n_slots
        return self.chain

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
        sizes = {"inst": int(det["boxes_net"].shape[0]), "words": total_words, "px": px, "ptslots": slots,
                 "rings": len(rings), "verts": int(rings.verts.shape[0])}
        self._learn(sizes)
        table = _stitch(rings, det["scores"], det["inst_tile"], tile_boxes, p)
        feats = postprocess_stage(table, rasters_fn(), p)
        self.last_table = table
        return len(table), feats

    def submit(self, det: dict, tile_tf, tile_boxes, rasters_fn, mark=None):
        """det: dict of device tensors boxes_net, scores, probs, inst_tile, tile_dims;
        rasters_fn(): the P5 rasters dict (called after P2-P4 are enqueued);
        mark(name): optional callback at the stage boundaries ("p4", "p5", "p9"), e.g. to record events."""
        mark = mark or (lambda name: None)
        n_inst = int(det["boxes_net"].shape[0])
        if self.p.iou_mode != "bbox":
            raise _lib_error("ChainRunner covers the reference's path (iou_mode 'bbox'); the opt-in mask mode runs "
                             "through pipeline.predict_stage")
        if self.caps is None or n_inst > self.caps["inst"]:
            out = ("done",) + self._exact(det, tile_tf, tile_boxes, rasters_fn)
            for name in ("p3", "p4", "p5", "p9"):
                mark(name)
            return out
        dev = det["boxes_net"].device
        chain = self._chain(dev)
        slot = self._seq % self.n_slots
        if self._busy[slot] is not None:
            raise _lib_error(f"ChainRunner: more than {self.n_slots} images in flight (collect the oldest ticket first)")
        self._seq += 1
        chain.predict(slot, det, tile_tf, tile_boxes, self.use_graph)
        mark("p3")
        mark("p4")
        rasters = rasters_fn()
        mark("p5")
        for key in ("height_ready", "ndvi_ready"):       # rasters produced / copied on another stream
            if rasters.get(key) is not None:
                torch.cuda.current_stream().wait_event(rasters[key])
        combined = geo.almost_equals(rasters["height_transform"], rasters["ndvi_transform"]) and \
            _similar_bounds(rasters["height_bounds"], rasters["ndvi_bounds"])
        sp = select_params(self.p, tuple(rasters["ndvi"].shape), rasters["ndvi_bounds"], rasters["pixel_x"],
                           rasters["pixel_y"])
        chain.post(slot, rasters["ndvi"], rasters["ndvi_transform"], rasters["height"], rasters["height_transform"],
                   combined, sp, self.use_graph)
        mark("p9")
        host = self._pinned.pop() if self._pinned else torch.empty((CTR_SIZE,), dtype=torch.int64).pin_memory()
        host.copy_(chain.counters(slot), non_blocking=True)
        ev = torch.cuda.Event()
        ev.record()
        ticket = ("dyn", ev, host, slot, chain, (det, tile_tf, tile_boxes, rasters_fn))
        self._busy[slot] = id(ticket)
        return ticket

    def collect(self, ticket):
        if ticket[0] == "done":
            return ticket[1], ticket[2]
        _, ev, host, slot, chain, again = ticket
        ev.synchronize()
        c = host.tolist()
        self._pinned.append(host)
        if chain is self.chain and self._busy[slot] == id(ticket):
            self._busy[slot] = None
        if c[CTR_FLAG] != 0:
            # some capacity was too small: redo this image with exact sizes (which also re-learns them)
            self.fallbacks += 1
            if c[CTR_FLAG] & 2:
                self.nbr_per_crown *= 2
                self.chain = None
            if c[CTR_FLAG] & 4 and self.slot_contours < 256:
                # an instance outgrew its contour slot of the single-pass walk (ragged mask: many borders)
                self.slot_contours *= 2
                self.chain = None
            return self._exact(*again)
        self._learn({"words": c[CTR_WORDS], "px": c[CTR_PX], "ptslots": c[CTR_SLOTS], "rings": c[CTR_RINGS],
                     "verts": c[CTR_VERTS]})
        self.last_counters = c
        tv, to, tc = chain.table(slot)
        self.last_table = CrownTable(tv[:c[CTR_VTABLE]], to[:c[CTR_NTABLE] + 1], tc[:c[CTR_NTABLE]])
        return c[CTR_NTABLE], trim_features(chain.features(slot), c[CTR_NFINAL], c[CTR_VFINAL])

    def table(self, ticket_counters, slot):
        """The stitched crown table (``geojson_predictions/<image>.gpkg``) of a collected dyn ticket."""
        c = ticket_counters
        verts, ring_off, conf = self.chain.table(slot)
        return CrownTable(verts[:c[CTR_VTABLE]], ring_off[:c[CTR_NTABLE] + 1], conf[:c[CTR_NTABLE]])


def _lib_error(msg):
    from . import _lib
    return _lib.TreedetError(msg)
