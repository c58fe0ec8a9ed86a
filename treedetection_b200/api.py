"""Host-buffer entry of the crown pipeline for one image -- the call the reference-facing
functions (``detection.predict_tiles`` / ``postprocess_files``) make once rasters and the
predictor's raw outputs are in host memory.

``run_image`` takes HOST arrays (pinned tensors are used as they are), copies them to the
device on the current stream, runs P1..P9 through the C-ABI kernels and copies the final
crown table back.  It is also what ``bench.py`` times for its end-to-end (``e2e``) figure.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _lib, ops, pipeline, tiling


@dataclass
class HostImage:
    """Everything the path consumes for one image, in host memory.  ``rgbi`` / ``ndsm`` may also be DEVICE tensors
    (a raster decoded on the GPU, geotiff.read_device): ``ready[name]`` is then the CUDA event after which the
    tensor is complete."""
    rgbi: torch.Tensor            # (4, H, W) uint8  (pinned for async copies)
    transform: tuple
    ndsm: torch.Tensor            # (h, w) float32
    ndsm_transform: tuple
    tiles: dict                   # tiles JSON (tiling.tile_grid)
    boxes_net: torch.Tensor       # (N, 4) f32   raw ROI-head outputs, tile-major
    scores: torch.Tensor          # (N,) f32
    probs: torch.Tensor           # (N, 28, 28) f32
    inst_tile: torch.Tensor       # (N,) i32
    tile_dims: torch.Tensor       # (T, 4) i32
    ready: dict = None            # name -> torch.cuda.Event for rasters that are already on the device

    @classmethod
    def from_scene(cls, sc, pin=True):
        def t(a):
            x = torch.from_numpy(np.ascontiguousarray(a))
            return x.pin_memory() if pin and _lib.cuda_available() else x
        d = sc.det
        return cls(t(sc.rgbi), sc.transform, t(sc.ndsm), sc.ndsm_transform, sc.tiles, t(d.boxes_net), t(d.scores),
                   t(d.probs), t(d.inst_tile), t(d.tile_dims))

    def h2d_bytes(self):
        return sum(x.numel() * x.element_size() for x in
                   (self.rgbi, self.ndsm, self.boxes_net, self.scores, self.probs, self.inst_tile, self.tile_dims)
                   if not x.is_cuda)


class TileTables:
    """Host + device tables derived once from the tiles JSON."""

    def __init__(self, tiles: dict, device, shift=1):
        self.win = torch.tensor([m["window"] for m in tiles.values()], dtype=torch.int32).reshape(-1, 4)
        self.net = torch.tensor([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in self.win.tolist()],
                                dtype=torch.int32).reshape(-1, 2)
        self.tile_org = self.win[:, :2].contiguous().to(device)       # (T,2) window origins (iou_mode: mask)
        self.tile_tf, boxes_int = pipeline.tile_tables(tiles, device)
        self.tile_boxes = pipeline.filter_boxes(boxes_int, shift, device)
        self.p1_floats = int((3 * self.net[:, 0].long() * self.net[:, 1].long()).sum())
        self._plans = {}

    def retarget(self, tiles: dict, device, shift=1):
        """Another image with the SAME tile grid (windows, network sizes): only the georeferenced tables -- the
        tile transforms and the stitch filter boxes -- are refreshed, in place, so that device pointers (and the
        CUDA graphs keyed on them) and the P1 plans stay valid."""
        win = torch.tensor([m["window"] for m in tiles.values()], dtype=torch.int32).reshape(-1, 4)
        if win.shape != self.win.shape or not torch.equal(win, self.win):
            raise _lib.TreedetError("retarget: the image has another tile grid")
        tile_tf, boxes_int = pipeline.tile_tables(tiles, device)
        self.tile_tf.copy_(tile_tf)
        self.tile_boxes.copy_(pipeline.filter_boxes(boxes_int, shift, device))

    def plan(self, image):
        """P1 plan (device tile tables) for rasters shaped like ``image``; built once."""
        return self.plan_for(image.element_size(), image.shape[1], image.shape[2])

    def plan_for(self, elem_size, height, width):
        key = (int(elem_size), int(height), int(width))
        if key not in self._plans:
            self._plans[key] = ops.TilePlan(self.win, self.net, *key)
        return self._plans[key]


_FIELDS = ("verts", "ring_off", "poly_id", "conf", "area", "tree_height", "centroid", "is_contained", "num_contained")
_pinned = {}


def _pinned_like(name, t):
    """A pinned host buffer of at least t.numel() elements (kept and re-used: allocating pinned
    memory costs more than the copy it serves)."""
    buf = _pinned.get(name)
    if buf is None or buf.dtype != t.dtype or buf.numel() < t.numel():
        buf = torch.empty((max(int(t.numel() * 1.25), 1024),), dtype=t.dtype, pin_memory=True)
        _pinned[name] = buf
    return buf[:t.numel()].view(t.shape)


def features_to_host(f: pipeline.Features):
    """The final crown table in host memory: asynchronous copies into pinned buffers, one wait."""
    host = {}
    for name in _FIELDS:
        t = getattr(f, name)
        h = _pinned_like(name, t)
        h.copy_(t, non_blocking=True)
        host[name] = h
    torch.cuda.current_stream().synchronize()
    return {k: v.numpy().copy() for k, v in host.items()}


_copy_streams = {}
_staging = {}


def _device_buffer(device, name, like: torch.Tensor):
    """A persistent device buffer for the host tensor ``like`` (capacity grows by 1.25x; the view handed out has
    ``like``'s shape).  Keeping the inputs of successive images at the same addresses is what lets td_chain_*
    replay the CUDA graphs it captured for the previous image instead of capturing new ones."""
    key = (torch.device(device).index, name, like.dtype, tuple(like.shape[1:]))
    buf = _staging.get(key)
    rows = like.shape[0] if like.dim() else 1
    if buf is None or buf.shape[0] < rows:
        cap = max(int(rows * 1.25), 1)
        buf = torch.empty((cap,) + tuple(like.shape[1:]), dtype=like.dtype, device=device)
        _staging[key] = buf
    return buf[:rows]


def _copy_stream(device):
    key = torch.device(device).index
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def table_to_host(t: pipeline.CrownTable):
    """the stitched crown table (geojson_predictions/<image>.gpkg) in host memory"""
    out = {}
    for name, x in (("table_verts", t.verts), ("table_ring_off", t.ring_off), ("table_conf", t.conf)):
        h = _pinned_like(name, x)
        h.copy_(x, non_blocking=True)
        out[name] = h
    torch.cuda.current_stream().synchronize()
    return {k: v.numpy().copy() for k, v in out.items()}


def run_image(img: HostImage, params: pipeline.PipelineParams, device, tables: TileTables = None, p1_out=None,
              with_p1=True, runner: pipeline.ChainRunner = None, want_table=False, on_h2d=None):
    """Host buffers in, final crowns (host numpy) out.  ``p1_out``: optional reusable
    device buffer for the normalised tiles (they feed the predictor, not this path).
    ``runner``: a :class:`pipeline.ChainRunner` kept across images of the same tiling -- P2-P9 are
    then enqueued without host synchronisation (capacity buffers); without it every image takes the
    exact-size path.

    The three host->device transfers ride a copy stream in the order the stages need them
    (ROI-head outputs, RGBI, nDSM) and the compute streams wait on one event per group: P2-P4
    overlap the RGBI copy, P1 (own stream), P5 and the statistics-free part of P6-P9 overlap the
    nDSM copy; only the per-crown statistics wait for the nDSM.  ``on_h2d(event)``: called once the copies are
    enqueued, with the event after which the host buffers of ``img`` may be overwritten."""
    if not _lib.cuda_available():
        raise _lib.TreedetError("run_image needs a CUDA device (there is no CPU fallback)")
    tables = tables or TileTables(img.tiles, device, params.shift)
    main = torch.cuda.current_stream(device)
    cs = _copy_stream(device)
    cs.wait_stream(main)
    persistent = runner is not None         # images of one runner share a tiling: stage them at fixed addresses
    def h2d(name, t):
        if t.is_cuda:                       # decoded on the device already
            ev = (img.ready or {}).get(name)
            if ev is not None:
                cs.wait_event(ev)
            return t
        if not persistent:
            return t.to(device, non_blocking=True)
        dst = _device_buffer(device, name, t)
        dst.copy_(t, non_blocking=True)
        return dst
    with torch.cuda.stream(cs):
        det = {k: h2d(k, getattr(img, k)) for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
        ev_det = cs.record_event()
        rgbi = h2d("rgbi", img.rgbi)
        ev_rgbi = cs.record_event()
        ndsm = h2d("ndsm", img.ndsm)
        ev_ndsm = cs.record_event()
    if on_h2d is not None:       # every copy out of the caller's host buffers is enqueued: they are free after ev_ndsm
        on_h2d(ev_ndsm)
    if not persistent:
        for t in (*det.values(), rgbi, ndsm):
            t.record_stream(main)
    main.wait_event(ev_det)
    tiles_out = None

    def rasters_fn():
        nonlocal tiles_out
        main.wait_event(ev_rgbi)
        if with_p1:       # P1 feeds the predictor, nothing downstream here: its own stream
            ps = _p1_stream(device)
            ps.wait_event(ev_rgbi)
            rgbi.record_stream(ps)
            with torch.cuda.stream(ps):
                tiles_out, _, _ = tables.plan(rgbi).run(rgbi, p1_out)
        h, w = ndsm.shape
        decimated = (int(h * params.height_scaling_factor), int(w * params.height_scaling_factor)) != (h, w)
        if decimated:
            main.wait_event(ev_ndsm)
        r = pipeline.raster_stage(rgbi, img.transform, ndsm, img.ndsm_transform, params,
                                  buffers=_raster_buffers(device) if persistent else None)
        r["height_ready"] = None if decimated else ev_ndsm
        return r

    if runner is None:
        table = pipeline.predict_stage(det["boxes_net"], det["scores"], det["probs"], det["inst_tile"],
                                       det["tile_dims"], tables.tile_tf, tables.tile_boxes, params)
        feats = pipeline.postprocess_stage(table, rasters_fn(), params)
        n_cand = len(table)
    else:
        n_cand, feats = runner.collect(runner.submit(det, tables.tile_tf, tables.tile_boxes, rasters_fn))
        table = runner.last_table
    host = features_to_host(feats)
    if want_table:
        host.update(table_to_host(table))
    host["n_candidates"] = n_cand
    if with_p1:
        main.wait_stream(_p1_stream(device))
    return host, tiles_out


_p1_streams = {}
_p5_buffers = {}


def _raster_buffers(device):
    return _p5_buffers.setdefault(torch.device(device).index, {})


def _p1_stream(device):
    key = torch.device(device).index
    if key not in _p1_streams:
        _p1_streams[key] = torch.cuda.Stream(device=device)
    return _p1_streams[key]
