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
    """Everything the path consumes for one image, in host memory."""
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

    @classmethod
    def from_scene(cls, sc, pin=True):
        def t(a):
            x = torch.from_numpy(np.ascontiguousarray(a))
            return x.pin_memory() if pin and torch.cuda.is_available() else x
        d = sc.det
        return cls(t(sc.rgbi), sc.transform, t(sc.ndsm), sc.ndsm_transform, sc.tiles, t(d.boxes_net), t(d.scores),
                   t(d.probs), t(d.inst_tile), t(d.tile_dims))

    def h2d_bytes(self):
        return sum(x.numel() * x.element_size() for x in
                   (self.rgbi, self.ndsm, self.boxes_net, self.scores, self.probs, self.inst_tile, self.tile_dims))


class TileTables:
    """Host + device tables derived once from the tiles JSON."""

    def __init__(self, tiles: dict, device, shift=1):
        self.win = torch.tensor([m["window"] for m in tiles.values()], dtype=torch.int32).reshape(-1, 4)
        self.net = torch.tensor([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in self.win.tolist()],
                                dtype=torch.int32).reshape(-1, 2)
        self.tile_tf, boxes_int = pipeline.tile_tables(tiles, device)
        self.tile_boxes = pipeline.filter_boxes(boxes_int, shift, device)
        self.p1_floats = int((3 * self.net[:, 0].long() * self.net[:, 1].long()).sum())
        self._plans = {}

    def plan(self, image):
        """P1 plan (device tile tables) for rasters shaped like ``image``; built once."""
        key = (image.element_size(), image.shape[1], image.shape[2])
        if key not in self._plans:
            self._plans[key] = ops.TilePlan(self.win, self.net, *key)
        return self._plans[key]


def features_to_host(f: pipeline.Features):
    """One device->host transfer of the final crown table."""
    return {
        "verts": f.verts.cpu().numpy(), "ring_off": f.ring_off.cpu().numpy(), "poly_id": f.poly_id.cpu().numpy(),
        "conf": f.conf.cpu().numpy(), "area": f.area.cpu().numpy(), "tree_height": f.tree_height.cpu().numpy(),
        "centroid": f.centroid.cpu().numpy(), "is_contained": f.is_contained.cpu().numpy(),
        "num_contained": f.num_contained.cpu().numpy(),
    }


_copy_streams = {}


def _copy_stream(device):
    key = torch.device(device).index
    if key not in _copy_streams:
        _copy_streams[key] = torch.cuda.Stream(device=device)
    return _copy_streams[key]


def run_image(img: HostImage, params: pipeline.PipelineParams, device, tables: TileTables = None, p1_out=None,
              with_p1=True):
    """Host buffers in, final crowns (host numpy) out.  ``p1_out``: optional reusable
    device buffer for the normalised tiles (they feed the predictor, not this path).

    The three host->device transfers ride a copy stream in the order the stages need them
    (ROI-head outputs, RGBI, nDSM) and the compute stream waits on one event per group, so
    P2-P4 overlap the RGBI copy and P1 / P5 overlap the nDSM copy."""
    if not torch.cuda.is_available():
        raise _lib.TreedetError("run_image needs a CUDA device (there is no CPU fallback)")
    tables = tables or TileTables(img.tiles, device, params.shift)
    main = torch.cuda.current_stream(device)
    cs = _copy_stream(device)
    cs.wait_stream(main)
    with torch.cuda.stream(cs):
        boxes = img.boxes_net.to(device, non_blocking=True); scores = img.scores.to(device, non_blocking=True)
        probs = img.probs.to(device, non_blocking=True); inst_tile = img.inst_tile.to(device, non_blocking=True)
        tile_dims = img.tile_dims.to(device, non_blocking=True)
        ev_det = cs.record_event()
        rgbi = img.rgbi.to(device, non_blocking=True)
        ev_rgbi = cs.record_event()
        ndsm = img.ndsm.to(device, non_blocking=True)
        ev_ndsm = cs.record_event()
    for t in (boxes, scores, probs, inst_tile, tile_dims, rgbi, ndsm):
        t.record_stream(main)
    main.wait_event(ev_det)
    table = pipeline.predict_stage(boxes, scores, probs, inst_tile, tile_dims, tables.tile_tf, tables.tile_boxes, params)
    main.wait_event(ev_rgbi)
    tiles_out = None
    if with_p1:
        tiles_out, _, _ = tables.plan(rgbi).run(rgbi, p1_out)
    main.wait_event(ev_ndsm)
    rasters = pipeline.raster_stage(rgbi, img.transform, ndsm, img.ndsm_transform, params)
    feats = pipeline.postprocess_stage(table, rasters, params)
    host = features_to_host(feats)
    host["n_candidates"] = len(table)
    return host, tiles_out
