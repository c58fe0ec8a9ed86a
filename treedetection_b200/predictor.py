"""Predictor plug of ``predict_tiles``.

The reference's ``Predictor`` (TreeDetection/prediction.py:18-47) wraps detectron2's
``DefaultPredictor``; the Mask R-CNN forward pass is explicitly NOT part of this build
(BASELINE.json north_star).  What the rest of the path needs
from it are the raw ROI-head outputs per tile -- boxes in network-input pixels, scores and
the 28x28 mask probabilities, BEFORE detectron2's paste -- and those are replayed from
fixtures dumped once.

A "model path" in config.yml (``combined_model`` / ``urban_model`` / ``forrest_model``) may
therefore be a directory of fixtures: ``<model_path>/<image stem>.npz`` with arrays
``boxes_net (N,4) f32, scores (N,) f32, probs (N,28,28) f32, inst_tile (N,) i32`` (tile
index in the order of the tiles JSON) and ``tile_ids`` (the ids, for validation).  Any object
with the same ``raw_outputs`` method can be passed as ``config["predictor"]`` instead (the
adapter for a live detectron2 / torchvision model is the "next" row N3).
"""
from __future__ import annotations

import os

import numpy as np

from .synth import Detections
from .tiling import resize_shortest_edge


class FixturePredictor:
    def __init__(self, model_path, exclude_vars=None):
        if not os.path.isdir(model_path):
            raise FileNotFoundError(
                f"{model_path}: the Mask R-CNN forward pass is outside this build; point the model key of config.yml "
                f"at a directory of ROI-head fixtures (<image stem>.npz) or pass config['predictor']")
        self.model_path = model_path
        self.exclude_vars = exclude_vars or []

    def raw_outputs(self, image_stem, tiles: dict) -> Detections:
        path = os.path.join(self.model_path, image_stem + ".npz")
        tile_ids = list(tiles.keys())
        tile_dims = np.zeros((len(tile_ids), 4), dtype=np.int32)
        for t, tid in enumerate(tile_ids):
            _, _, w, h = tiles[tid]["window"]
            nh, nw = resize_shortest_edge(h, w)
            tile_dims[t] = (h, w, nh, nw)
        if not os.path.exists(path):
            z = np.zeros
            return Detections(z((0, 4), np.float32), z(0, np.float32), z((0, 28, 28), np.float32), z(0, np.int32),
                              tile_dims, tile_ids, tiles)
        with np.load(path, allow_pickle=False) as f:
            fx_ids = [str(s) for s in f["tile_ids"]]
            boxes, scores, probs, inst_tile = f["boxes_net"], f["scores"], f["probs"], f["inst_tile"]
        if fx_ids != tile_ids:   # fixtures were dumped with another tiling: remap by tile id
            pos = {tid: t for t, tid in enumerate(tile_ids)}
            remap = np.array([pos.get(tid, -1) for tid in fx_ids], dtype=np.int64)
            new_tile = remap[inst_tile]
            keep = new_tile >= 0
            order = np.argsort(new_tile[keep], kind="stable")
            boxes, scores, probs = boxes[keep][order], scores[keep][order], probs[keep][order]
            inst_tile = new_tile[keep][order].astype(np.int32)
        # tiles excluded for this model (prediction.py:79-93): drop their instances
        if self.exclude_vars:
            skip = np.array([any(bool(tiles[tid].get(v, False)) for v in self.exclude_vars) for tid in tile_ids])
            keep = ~skip[inst_tile]
            boxes, scores, probs, inst_tile = boxes[keep], scores[keep], probs[keep], inst_tile[keep]
        return Detections(np.ascontiguousarray(boxes, np.float32), np.ascontiguousarray(scores, np.float32),
                          np.ascontiguousarray(probs, np.float32), np.ascontiguousarray(inst_tile, np.int32),
                          tile_dims, tile_ids, tiles)


def dump_fixtures(model_path, image_stem, det: Detections):
    """Write the fixtures of one image (used by tests / bench to stage a 'model')."""
    os.makedirs(model_path, exist_ok=True)
    np.savez(os.path.join(model_path, image_stem + ".npz"), boxes_net=det.boxes_net, scores=det.scores,
             probs=det.probs, inst_tile=det.inst_tile, tile_ids=np.array(det.tile_ids))
