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
with the same ``raw_outputs`` method can be passed as ``config["predictor"]`` instead;
:class:`TorchvisionPredictor` is such an object for a live torchvision Mask R-CNN (row N3).
"""
from __future__ import annotations

import os

import numpy as np

from .synth import Detections
from .tiling import resize_shortest_edge


def _load_npz(path):
    """Arrays of an ``.npz`` as a dict.  Members that are STORED (``np.savez``, what ``dump_fixtures`` writes) become
    zero-copy views of the memory-mapped file -- ``np.load`` would read every member through zipfile (one extra copy
    plus a CRC-32 pass, ~0.1 s for the 100 MB of mask probabilities of one image); deflated members fall back to it."""
    import mmap
    import zipfile
    out = {}
    with open(path, "rb") as f:
        buf = mmap.mmap(f.fileno(), 0, access=mmap.ACCESS_READ)
    with zipfile.ZipFile(path) as z:
        infos = z.infolist()
    slow = []
    for zi in infos:
        name = zi.filename[:-4] if zi.filename.endswith(".npy") else zi.filename
        if zi.compress_type != 0:
            slow.append(name)
            continue
        h = zi.header_offset
        if buf[h:h + 4] != b"PK\x03\x04":
            slow.append(name)
            continue
        n_name, n_extra = int.from_bytes(buf[h + 26:h + 28], "little"), int.from_bytes(buf[h + 28:h + 30], "little")
        start = h + 30 + n_name + n_extra
        import io
        head = io.BytesIO(buf[start:start + 4096])
        try:
            version = np.lib.format.read_magic(head)
            shape, fortran, dtype = (np.lib.format.read_array_header_1_0(head) if version == (1, 0)
                                     else np.lib.format.read_array_header_2_0(head))
        except Exception:
            slow.append(name)
            continue
        if fortran or dtype.hasobject:
            slow.append(name)
            continue
        count = int(np.prod(shape)) if len(shape) else 1
        out[name] = np.frombuffer(buf, dtype=dtype, count=count, offset=start + head.tell()).reshape(shape)
    if slow:
        with np.load(path, allow_pickle=False) as f:
            for name in slow:
                out[name] = f[name]
    return out


class FixturePredictor:
    def __init__(self, model_path, exclude_vars=None, allow_missing=False):
        self.allow_missing = allow_missing
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
            if self.allow_missing:      # explicit opt-in: "no fixture" means "no detections"
                z = np.zeros
                return Detections(z((0, 4), np.float32), z(0, np.float32), z((0, 28, 28), np.float32), z(0, np.int32),
                                  tile_dims, tile_ids, tiles)
            # a mis-staged fixture directory must not produce silently empty crown layers that a resumed
            # run then skips as "already predicted": predict_on_model logs the error and does not mark the file
            raise FileNotFoundError(f"no ROI-head fixture for image {image_stem!r}: {path}")
        f = _load_npz(path)
        fx_ids = [str(s) for s in f["tile_ids"]]
        boxes, scores, probs, inst_tile = f["boxes_net"], f["scores"], f["probs"], f["inst_tile"]
        n = len(scores)
        if boxes.shape != (n, 4) or probs.shape != (n, 28, 28) or inst_tile.shape != (n,):
            raise ValueError(f"{path}: boxes_net {boxes.shape}, scores {scores.shape}, probs {probs.shape}, inst_tile "
                             f"{inst_tile.shape} do not describe the same instances")
        if n and (int(inst_tile.min()) < 0 or int(inst_tile.max()) >= len(fx_ids)):
            raise ValueError(f"{path}: inst_tile outside [0, {len(fx_ids)})")
        if n and np.any(np.diff(inst_tile.astype(np.int64)) < 0):
            raise ValueError(f"{path}: instances are not in tile-major order")
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


class TorchvisionPredictor:
    """N3: a live Mask R-CNN behind the predictor plug.

    The reference runs detectron2's ``DefaultPredictor`` per batch of tiles and lets it paste the
    masks (TreeDetection/prediction.py:178-195).  This adapter consumes the P1 output where it lies
    -- normalised float32 CHW tiles in HBM -- runs backbone, RPN and ROI heads of a torchvision
    ``MaskRCNN`` on the same stream and hands the RAW ROI-head outputs (boxes in network-input pixels,
    scores, 28x28 mask probabilities, i.e. before any paste) to P2 as device tensors: no host round
    trip between P1, the model and P2.  detectron2 is absent offline; torchvision's Mask R-CNN has the
    same ROI-head contract.  With ``model=None`` a random-init ResNet-FPN model is built
    (BASELINE.json: checkpoints are not available offline) -- useful for end-to-end timing and for
    exercising the path on real network outputs, not for detection quality.
    """
    wants_tiles = True

    def __init__(self, model=None, device="cuda", batch_tiles=4, backbone="resnet50", score_thresh=0.3,
                 nms_thresh=0.5, detections_per_img=100, seed=0, exclude_vars=None):
        import torch
        self.device = torch.device(device)
        self.batch_tiles = int(batch_tiles)
        self.exclude_vars = exclude_vars or []
        if model is None:
            from torchvision.models.detection import MaskRCNN
            from torchvision.models.detection.backbone_utils import resnet_fpn_backbone
            torch.manual_seed(seed)
            model = MaskRCNN(resnet_fpn_backbone(backbone_name=backbone, weights=None), num_classes=2,
                             box_score_thresh=score_thresh, box_nms_thresh=nms_thresh,
                             box_detections_per_img=detections_per_img)
        self.model = model.to(self.device).eval()
        # P1 delivers BGR 0..255 (detectron2's input convention, prediction.py:166); torchvision's
        # statistics are RGB 0..1
        mean = torch.tensor(self.model.transform.image_mean, dtype=torch.float32, device=self.device) * 255.0
        std = torch.tensor(self.model.transform.image_std, dtype=torch.float32, device=self.device) * 255.0
        self._mean = mean.flip(0).view(3, 1, 1)
        self._std = std.flip(0).view(3, 1, 1)

    def raw_outputs(self, image_stem, tiles):
        """The tiling only (no instances): the live outputs come from :meth:`forward`."""
        tile_ids = list(tiles.keys())
        tile_dims = np.zeros((len(tile_ids), 4), dtype=np.int32)
        for t, tid in enumerate(tile_ids):
            _, _, w, h = tiles[tid]["window"]
            nh, nw = resize_shortest_edge(h, w)
            tile_dims[t] = (h, w, nh, nw)
        z = np.zeros
        return Detections(z((0, 4), np.float32), z(0, np.float32), z((0, 28, 28), np.float32), z(0, np.int32),
                          tile_dims, tile_ids, tiles)

    def forward(self, image_stem, tiles, tiles_dev, tiles_off, flags=None):
        """tiles_dev: flat float32 device buffer of P1 (tile t = (3, net_h, net_w) at tiles_off[t]).
        Returns Detections whose boxes_net / scores / probs / inst_tile are DEVICE tensors."""
        import torch
        from torchvision.models.detection.image_list import ImageList
        base = self.raw_outputs(image_stem, tiles)
        dims = base.tile_dims
        skip = np.zeros(len(base.tile_ids), dtype=bool)
        if self.exclude_vars:     # tiles excluded for this model (prediction.py:79-93)
            skip = np.array([any(bool(tiles[tid].get(v, False)) for v in self.exclude_vars) for tid in base.tile_ids])
        boxes, scores, probs, inst = [], [], [], []
        order = [t for t in range(len(base.tile_ids)) if not skip[t]]
        # batches of tiles with the same network size (ResizeShortestEdge: 800 x 800 except at the image edge)
        groups = {}
        for t in order:
            groups.setdefault((int(dims[t, 2]), int(dims[t, 3])), []).append(t)
        per_tile = {}
        with torch.no_grad():
            for (nh, nw), ts in groups.items():
                for b0 in range(0, len(ts), self.batch_tiles):
                    tb = ts[b0:b0 + self.batch_tiles]
                    x = torch.stack([tiles_dev[int(tiles_off[t]):int(tiles_off[t]) + 3 * nh * nw].view(3, nh, nw)
                                     for t in tb])
                    x = ((x - self._mean) / self._std).flip(1)          # BGR -> RGB, normalised
                    pad_h, pad_w = (-nh) % 32, (-nw) % 32
                    if pad_h or pad_w:
                        x = torch.nn.functional.pad(x, (0, pad_w, 0, pad_h))
                    images = ImageList(x, [(nh, nw)] * len(tb))
                    feats = self.model.backbone(x)
                    if isinstance(feats, torch.Tensor):
                        feats = {"0": feats}
                    proposals, _ = self.model.rpn(images, feats)
                    dets, _ = self.model.roi_heads(feats, proposals, images.image_sizes)
                    for t, det in zip(tb, dets):
                        per_tile[t] = det
        for t in order:           # tile-major, instance order = model output order (score descending)
            det = per_tile[t]
            n = det["scores"].shape[0]
            if n == 0:
                continue
            boxes.append(det["boxes"].to(torch.float32))
            scores.append(det["scores"].to(torch.float32))
            probs.append(det["masks"].to(torch.float32).reshape(n, 28, 28))
            inst.append(torch.full((n,), t, dtype=torch.int32, device=self.device))
        if boxes:
            out = (torch.cat(boxes).contiguous(), torch.cat(scores).contiguous(), torch.cat(probs).contiguous(),
                   torch.cat(inst).contiguous())
        else:
            z = lambda *s, dt=torch.float32: torch.zeros(s, dtype=dt, device=self.device)
            out = (z(0, 4), z(0), z(0, 28, 28), z(0, dt=torch.int32))
        return Detections(out[0], out[1], out[2], out[3], dims, base.tile_ids, tiles)
