"""Row sharding of a mosaic over the GPUs of one box.

Every image file -- seam strips included -- is an independent unit through P0b..P9 (the
reference never compares crowns of different files, TreeDetection/postprocessing.py:1030-1075),
so rank r simply owns image row r.  The one exchange step of the path is the seam strip between
row r and row r + 1 (TreeDetection/merging.py:81-107): it needs the top ``halo`` pixel rows of
the lower neighbour's rasters.  Rank r + 1 sends them to rank r (NCCL send/recv over NVLink;
``gloo`` in the CPU tests), rank r assembles the strip on its device with ``td_seam_crop``.
No other collective is on the data path.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def strip_height(tile_height, buffer, overlapping_tiles_height):
    """Height of a down-seam strip in PIXELS (merging.py:98-100)."""
    return int((tile_height + 2 * buffer) * overlapping_tiles_height)


def halo_rows(tile_height, buffer, overlapping_tiles_height):
    """Rows of the LOWER neighbour inside the down-seam strip.  crop_image (helpers.py:1053-1085) cuts
    ``sh`` rows starting at ``mh // 2 - sh // 2`` of the two-image mosaic, i.e. ``sh // 2`` rows above the
    seam and ``sh - sh // 2`` below it: for an odd strip height the extra row comes from the lower image."""
    sh = strip_height(tile_height, buffer, overlapping_tiles_height)
    return sh - sh // 2


def exchange_down_halos(tops, rank, world, group=None, recv=None):
    """``tops``: list of contiguous tensors holding the TOP halo rows of this rank's rasters.
    Sends them to rank - 1 and returns the list received from rank + 1 (``None`` on the last
    rank).  One batched isend/irecv per call; tensors keep their device.  ``recv``: optional
    pre-allocated receive buffers (same shapes as ``tops``), re-used across steps."""
    if world == 1:
        return None
    ops_ = []
    if rank + 1 < world:
        if recv is None:
            recv = [torch.empty_like(t) for t in tops]
    else:
        recv = None
    if rank + 1 < world:
        ops_ += [dist.P2POp(dist.irecv, t, rank + 1, group) for t in recv]
    if rank > 0:
        ops_ += [dist.P2POp(dist.isend, t, rank - 1, group) for t in tops]
    for w in dist.batch_isend_irecv(ops_):
        w.wait()
    return recv


def assemble_down_strip(own, halo, sh=None, out=None):
    """own (bands, H, W) device raster of this rank, halo (bands, rows, W) = top rows of the
    lower neighbour, ``sh`` = strip height (default 2 * rows).  Returns the (bands, sh, W) seam strip:
    ``sh // 2`` bottom rows of ``own`` followed by ``sh - sh // 2`` rows of the halo -- the same rows
    merging.py's centre crop of the two-image mosaic selects, gathered without building the mosaic."""
    from . import ops
    rows = halo.shape[1]
    sh = 2 * rows if sh is None else int(sh)
    up = sh // 2
    if sh - up != rows:
        raise ValueError(f"assemble_down_strip: a strip of {sh} rows needs {sh - up} halo rows, got {rows}")
    bottom = own[:, own.shape[1] - up:, :].contiguous()
    # mosaic of (up rows, rows rows): its centre crop of height sh starts at (up + rows) // 2 - sh // 2 = 0
    return ops.seam_crop(bottom, halo, 1, own.shape[2], sh, out=out)
