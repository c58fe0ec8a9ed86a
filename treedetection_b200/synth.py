"""Seeded synthetic orthophoto + nDSM mosaics and Mask R-CNN output fixtures
(SURVEY.md §8d).  NumPy only.

There are no datasets or checkpoints offline, and the reference's predictor is not
rewritten: the raw ROI-head outputs it would produce (boxes, scores, 28x28 mask
probabilities per tile) are synthesised here from a seeded tree field and replayed
through the pipeline.  Georeferencing follows the bundled tile: EPSG 25832, north-up,
origin (412000, 5318000) at the bottom-left.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

from . import geo
from .tiling import resize_shortest_edge, tile_grid

EPSG = 25832
ORIGIN_X = 412000.0
ORIGIN_Y = 5318000.0


@dataclass
class TreeField:
    x: np.ndarray      # centre (CRS)
    y: np.ndarray
    r: np.ndarray      # crown radius (m)
    h: np.ndarray      # tree height (m)
    score: np.ndarray  # detection confidence of the tree
    ecc: np.ndarray    # y radius / x radius
    left: float
    bottom: float
    width_m: float
    height_m: float


def tree_field(seed, width_m, height_m, density_per_km2=2500.0, left=ORIGIN_X, bottom=ORIGIN_Y,
               r_range=(1.5, 6.0)):
    rng = np.random.default_rng(seed)
    n = max(int(round(density_per_km2 * width_m * height_m / 1e6)), 1)
    x = left + rng.uniform(0, width_m, n)
    y = bottom + rng.uniform(0, height_m, n)
    r = rng.uniform(r_range[0], r_range[1], n)
    h = rng.uniform(3.0, 35.0, n)
    low = rng.uniform(0, 1, n) < 0.10
    h[low] = rng.uniform(0.5, 3.0, int(low.sum()))     # exercises the height filter
    score = rng.uniform(0.25, 0.999, n)
    ecc = rng.uniform(0.8, 1.25, n)
    return TreeField(x, y, r, h, score, ecc, left, bottom, width_m, height_m)


def image_transform(left, top, px):
    return (px, 0.0, left, 0.0, -px, top)


def make_ndsm(field: TreeField, px, seed=0):
    """float32 (H, W): max over trees of h * max(0, 1 - (d/r)^2) + N(0, 0.1), clipped >= 0."""
    w = int(round(field.width_m / px)); h = int(round(field.height_m / px))
    top = field.bottom + field.height_m
    out = np.zeros((h, w), dtype=np.float32)
    for k in range(len(field.x)):
        rx = field.r[k]; ry = field.r[k] * field.ecc[k]
        c0 = max(int((field.x[k] - rx - field.left) / px), 0); c1 = min(int((field.x[k] + rx - field.left) / px) + 2, w)
        r0 = max(int((top - field.y[k] - ry) / px), 0); r1 = min(int((top - field.y[k] + ry) / px) + 2, h)
        if c1 <= c0 or r1 <= r0:
            continue
        xs = field.left + (np.arange(c0, c1) + 0.5) * px - field.x[k]
        ys = top - (np.arange(r0, r1) + 0.5) * px - field.y[k]
        d2 = (xs[None, :] / rx) ** 2 + (ys[:, None] / ry) ** 2
        v = (field.h[k] * np.maximum(0.0, 1.0 - d2)).astype(np.float32)
        np.maximum(out[r0:r1, c0:c1], v, out=out[r0:r1, c0:c1])
    rng = np.random.default_rng(seed + 7)
    out += rng.normal(0.0, 0.1, out.shape).astype(np.float32)
    np.maximum(out, 0.0, out=out)
    return out


def make_rgbi(field: TreeField, px, seed=0, u16=False):
    """uint8 (4, H, W) RGBI: vegetation NIR~N(180,20) red~N(60,15) inside crowns,
    ground NIR~N(90,25) red~N(110,25).  ``u16`` scales by 257 (16-bit branch of P1)."""
    w = int(round(field.width_m / px)); h = int(round(field.height_m / px))
    top = field.bottom + field.height_m
    veg = np.zeros((h, w), dtype=bool)
    for k in range(len(field.x)):
        rx = field.r[k]; ry = field.r[k] * field.ecc[k]
        c0 = max(int((field.x[k] - rx - field.left) / px), 0); c1 = min(int((field.x[k] + rx - field.left) / px) + 2, w)
        r0 = max(int((top - field.y[k] - ry) / px), 0); r1 = min(int((top - field.y[k] + ry) / px) + 2, h)
        if c1 <= c0 or r1 <= r0:
            continue
        xs = field.left + (np.arange(c0, c1) + 0.5) * px - field.x[k]
        ys = top - (np.arange(r0, r1) + 0.5) * px - field.y[k]
        veg[r0:r1, c0:c1] |= ((xs[None, :] / rx) ** 2 + (ys[:, None] / ry) ** 2) <= 1.0
    rng = np.random.default_rng(seed + 11)
    out = np.empty((4, h, w), dtype=np.uint8)
    # cheap gaussian-ish noise: sum of 4 uniforms (Irwin-Hall), generated band by band
    def noise(shape):
        z = rng.integers(0, 256, size=(4,) + shape, dtype=np.uint8).astype(np.float32).sum(axis=0)
        return (z - 510.0) / 147.8     # ~N(0,1)
    nz = noise((h, w))
    out[0] = np.clip(np.where(veg, 60 + 15 * nz, 110 + 25 * nz), 0, 255).astype(np.uint8)
    nz = noise((h, w))
    out[1] = np.clip(np.where(veg, 95 + 15 * nz, 105 + 25 * nz), 0, 255).astype(np.uint8)
    nz = noise((h, w))
    out[2] = np.clip(np.where(veg, 55 + 12 * nz, 95 + 25 * nz), 0, 255).astype(np.uint8)
    nz = noise((h, w))
    out[3] = np.clip(np.where(veg, 180 + 20 * nz, 90 + 25 * nz), 0, 255).astype(np.uint8)
    if u16:
        return out.astype(np.uint16) * 257
    return out


@dataclass
class Detections:
    """Raw ROI-head outputs of all tiles of one image, flattened tile-major in the
    tile order of the tiles JSON; within a tile sorted by descending score."""
    boxes_net: np.ndarray   # (N, 4) f32, network-input pixels (xyxy)
    scores: np.ndarray      # (N,) f32
    probs: np.ndarray       # (N, 28, 28) f32 sigmoid probabilities
    inst_tile: np.ndarray   # (N,) i32
    tile_dims: np.ndarray   # (T, 4) i32 [tile_h, tile_w, net_h, net_w]
    tile_ids: list
    tiles: dict


def make_detections(field: TreeField, tiles: dict, px, seed=0, cap=100, score_floor=0.3, M=28):
    rng = np.random.default_rng(seed + 23)
    tile_ids = list(tiles.keys())
    T = len(tile_ids)
    tile_dims = np.zeros((T, 4), dtype=np.int32)
    # spatial index of trees on a 50 m grid
    cell = 50.0
    gx = ((field.x - field.left) // cell).astype(np.int64); gy = ((field.y - field.bottom) // cell).astype(np.int64)
    ncx = int(field.width_m // cell) + 1
    key = gy * ncx + gx
    order = np.argsort(key, kind="stable")
    skey = key[order]
    boxes, scores, probs, inst_tile = [], [], [], []
    lin = (np.arange(M, dtype=np.float64) + 0.5) / M * 2.0 - 1.0
    for t, tid in enumerate(tile_ids):
        meta = tiles[tid]
        c_off, r_off, tw, th = meta["window"]
        net_h, net_w = resize_shortest_edge(th, tw)
        tile_dims[t] = (th, tw, net_h, net_w)
        tf = meta["transform"]
        left = tf[2]; top = tf[5]
        right = left + tw * px; bottom = top - th * px
        # candidate trees: cells overlapping the tile (+ max radius)
        cx0 = int(max((left - 8 - field.left) // cell, 0)); cx1 = int(min((right + 8 - field.left) // cell, ncx - 1))
        cy0 = int(max((bottom - 8 - field.bottom) // cell, 0)); cy1 = int((top + 8 - field.bottom) // cell)
        cand = []
        for cy in range(cy0, cy1 + 1):
            lo = np.searchsorted(skey, cy * ncx + cx0, side="left"); hi = np.searchsorted(skey, cy * ncx + cx1, side="right")
            cand.append(order[lo:hi])
        cand = np.concatenate(cand) if cand else np.zeros(0, dtype=np.int64)
        if cand.size == 0:
            continue
        rx = field.r[cand]; ry = field.r[cand] * field.ecc[cand]
        x0 = (field.x[cand] - rx - left) / px; x1 = (field.x[cand] + rx - left) / px
        y0 = (top - field.y[cand] - ry) / px; y1 = (top - field.y[cand] + ry) / px
        hit = (x1 > 0) & (x0 < tw) & (y1 > 0) & (y0 < th)
        cand = cand[hit]; x0 = x0[hit]; x1 = x1[hit]; y0 = y0[hit]; y1 = y1[hit]
        if cand.size == 0:
            continue
        k = cand.size
        jit = rng.normal(0.0, 1.0, (k, 4))
        bx = np.stack([x0 + jit[:, 0], y0 + jit[:, 1], x1 + jit[:, 2], y1 + jit[:, 3]], axis=1)
        bx[:, 0::2] = np.clip(bx[:, 0::2], 0, tw); bx[:, 1::2] = np.clip(bx[:, 1::2], 0, th)
        sc = np.clip(field.score[cand] + rng.normal(0, 0.02, k), 0.0, 0.999)
        ok = (sc >= score_floor) & (bx[:, 2] - bx[:, 0] > 1) & (bx[:, 3] - bx[:, 1] > 1)
        bx = bx[ok]; sc = sc[ok]
        if sc.size == 0:
            continue
        o = np.argsort(-sc, kind="stable")[:cap]
        bx = bx[o]; sc = sc[o]
        k = sc.size
        rho2 = lin[None, :, None] ** 2 + lin[None, None, :] ** 2
        logits = 6.0 * (1.0 - rho2 / 0.85) + rng.normal(0.0, 0.5, (k, M, M))
        pr = 1.0 / (1.0 + np.exp(-logits))
        bx_net = bx.copy()
        bx_net[:, 0::2] *= net_w / tw
        bx_net[:, 1::2] *= net_h / th
        boxes.append(bx_net.astype(np.float32)); scores.append(sc.astype(np.float32))
        probs.append(pr.astype(np.float32)); inst_tile.append(np.full(k, t, dtype=np.int32))
    if boxes:
        boxes = np.concatenate(boxes); scores = np.concatenate(scores)
        probs = np.concatenate(probs); inst_tile = np.concatenate(inst_tile)
    else:
        boxes = np.zeros((0, 4), np.float32); scores = np.zeros(0, np.float32)
        probs = np.zeros((0, M, M), np.float32); inst_tile = np.zeros(0, np.int32)
    return Detections(boxes, scores, probs, inst_tile, tile_dims, tile_ids, tiles)


@dataclass
class Scene:
    """One synthetic image with everything the post-model path consumes."""
    stem: str
    transform: tuple          # RGBI transform (a,b,c,d,e,f)
    px: float
    rgbi: np.ndarray          # (4, H, W) u8
    ndsm: np.ndarray          # (h, w) f32
    ndsm_transform: tuple
    field: TreeField
    tiles: dict
    det: Detections

    @property
    def area_km2(self):
        return self.field.width_m * self.field.height_m / 1e6


def make_scene(seed=1234, size_px=10000, px=0.2, ndsm_px=1.0, density_per_km2=2500.0, cap=100,
               tile_width=50, tile_height=50, buffer=20, stem="FDOP20_000000_rgbi", left=ORIGIN_X, bottom=ORIGIN_Y,
               size_px_y=None, r_range=(1.5, 6.0), with_rasters=True):
    wpx = size_px; hpx = size_px_y or size_px
    width_m = wpx * px; height_m = hpx * px
    field = tree_field(seed, width_m, height_m, density_per_km2, left, bottom, r_range)
    top = bottom + height_m
    tf = image_transform(left, top, px)
    tiles = tile_grid(stem, tf, wpx, hpx, EPSG, tile_width, tile_height, buffer)
    det = make_detections(field, tiles, px, seed, cap)
    rgbi = make_rgbi(field, px, seed) if with_rasters else None
    ndsm = make_ndsm(field, ndsm_px, seed) if with_rasters else None
    return Scene(stem, tf, px, rgbi, ndsm, image_transform(left, top, ndsm_px), field, tiles, det)


def config1_scene(ndsm_npz=None, seed=31):
    """BASELINE config 1: the reference's example tile.  The bundled ``data/nDSM/324125317.tif`` (1000 x 1000
    float32, 1 m, top-left 412000 / 5318000; a lossless copy of its pixels is committed as
    tests/golden/ndsm_324125317.npz) plus a SYNTHETIC 5000 x 5000 RGBI companion at 0.2 m (the bundled RGB is
    absent from the reference: .MISSING_LARGE_BLOBS) and ROI-head outputs replayed from seeded fixtures."""
    import os
    if ndsm_npz is None:
        ndsm_npz = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                                "ndsm_324125317.npz")
    ndsm = np.load(ndsm_npz)["ndsm"]
    left, top = 412000.0, 5318000.0
    sc = make_scene(seed=seed, size_px=5000, px=0.2, ndsm_px=1.0, density_per_km2=2500.0, stem="324125317", left=left,
                    bottom=top - 1000.0)
    sc.ndsm = ndsm
    return sc
