"""N > 1 plumbing on CPU: world_size 2 over gloo.  The halo exchange must hand rank r the top
rows of rank r + 1's rasters (the only exchange step of the path)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from treedetection_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rows = sharding.halo_rows(50, 20, 3)
        rng = np.random.default_rng(100 + rank)
        rgbi = torch.from_numpy(rng.integers(0, 256, size=(4, 400, 64), dtype=np.uint8))
        ndsm = torch.from_numpy(rng.uniform(0, 30, size=(1, 400, 64)).astype(np.float32))
        got = sharding.exchange_down_halos([rgbi[:, :rows].contiguous(), ndsm[:, :rows].contiguous()], rank, world)
        if rank + 1 < world:
            nb = np.random.default_rng(100 + rank + 1)
            want_rgbi = nb.integers(0, 256, size=(4, 400, 64), dtype=np.uint8)[:, :rows]
            want_ndsm = nb.uniform(0, 30, size=(1, 400, 64)).astype(np.float32)[:, :rows]
            ok = np.array_equal(got[0].numpy(), want_rgbi) and np.array_equal(got[1].numpy(), want_ndsm)
        else:
            ok = got is None
        out[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


def test_halo_rows_matches_reference_strip_height():
    # (50 + 2*20) * 3 = 270 px strip, half from each image
    assert sharding.halo_rows(50, 20, 3) == 135 and sharding.strip_height(50, 20, 3) == 270
    # odd strip (45 + 2*20) * 3 = 255: crop_image starts at mh // 2 - sh // 2, i.e. 127 rows above the seam
    # and 128 below it -- the extra row comes from the lower neighbour (what merging.py cuts on one GPU)
    assert sharding.strip_height(45, 20, 3) == 255 and sharding.halo_rows(45, 20, 3) == 128
    H = 400
    mh, sh = 2 * H, 255
    top = max(mh // 2 - sh // 2, 0)
    assert (H - top, top + sh - H) == (sh // 2, sharding.halo_rows(45, 20, 3))


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world)), dict(out)
