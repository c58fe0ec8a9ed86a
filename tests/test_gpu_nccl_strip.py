"""The one exchange step of the path on hardware: two ranks (two GPUs, NCCL over NVLink), rank 1 sends the top halo
rows of its rasters to rank 0 (``sharding.exchange_down_halos``), rank 0 assembles the down-seam strip on its device
and runs it through the CUDA-graph chain.  The strip's final crowns must equal the CPU oracle's on the strip the
reference would cut from the two-image mosaic (merging.py:81-107).  Skipped on a box with fewer than 2 GPUs."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port_no, out):
    import torch.distributed as dist

    from oracle import port
    from treedetection_b200 import api, geo, pipeline, sharding, synth, tiling
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port_no))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        px, size = 0.2, 1500
        p = pipeline.PipelineParams()
        scenes = [synth.make_scene(seed=51 + r, size_px=size, px=px, ndsm_px=px, density_per_km2=4000.0,
                                   bottom=synth.ORIGIN_Y - r * size * px) for r in range(world)]
        sc = scenes[rank]
        rows = sharding.halo_rows(p.tile_height, p.buffer, p.overlapping_tiles_height)
        rgbi, ndsm = torch.from_numpy(sc.rgbi).to(dev), torch.from_numpy(sc.ndsm).to(dev)
        tops = [rgbi[:, :rows].contiguous(), ndsm[None, :rows].contiguous()]
        recv = sharding.exchange_down_halos(tops, rank, world)
        if rank + 1 == world:
            out[rank] = "ok" if recv is None else "last rank received something"
            return
        low = scenes[rank + 1]
        s_rgbi = sharding.assemble_down_strip(rgbi, recv[0])
        s_ndsm = sharding.assemble_down_strip(ndsm[None], recv[1])[0]
        ref_rgbi = np.ascontiguousarray(port.seam_crop(sc.rgbi, low.rgbi, 1, size, 2 * rows)[0])
        ref_ndsm = np.ascontiguousarray(port.seam_crop(sc.ndsm[None], low.ndsm[None], 1, size, 2 * rows)[0][0])
        assert np.array_equal(s_rgbi.cpu().numpy(), ref_rgbi) and np.array_equal(s_ndsm.cpu().numpy(), ref_ndsm)
        both = synth.TreeField(*[np.concatenate([getattr(sc.field, k), getattr(low.field, k)]) for k in
                                 ("x", "y", "r", "h", "score", "ecc")], sc.field.left, low.field.bottom, sc.field.width_m,
                               2 * sc.field.height_m)
        s_tf = synth.image_transform(sc.field.left, sc.field.bottom + rows * px, px)
        s_tiles = tiling.tile_grid("FDOP20_seam_rgbi", s_tf, size, 2 * rows, synth.EPSG, p.tile_width, p.tile_height, p.buffer)
        s_det = synth.make_detections(both, s_tiles, px, 51)
        tables = api.TileTables(s_tiles, dev, p.shift)
        det = {k: torch.from_numpy(getattr(s_det, k)).to(dev) for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
        runner = pipeline.ChainRunner(p)
        rasters = lambda: pipeline.raster_stage(s_rgbi, s_tf, s_ndsm, s_tf, p)
        res = [runner.collect(runner.submit(det, tables.tile_tf, tables.tile_boxes, rasters)) for _ in range(2)]
        assert runner.fallbacks == 0
        # oracle on the strip the reference would cut
        rings, conf = port.predict_stage(s_det, s_tiles)
        H, W = ref_rgbi.shape[1:]
        oh, ow = int(H * p.ndvi_scaling_factor), int(W * p.ndvi_scaling_factor)
        dec = np.stack([port.decimate_bilinear(ref_rgbi[b], oh, ow) for b in (0, 0, 0, 3)])
        ndvi = port.ndvi_from_rgbi(dec).astype(np.float32)
        cfg = {k: getattr(p, k) for k in p.__dataclass_fields__}
        want, _ = port.post_process(rings, conf, ndvi, geo.compose(s_tf, geo.scale(W / ow, H / oh)),
                                    tuple(geo.raster_bounds(s_tf, W, H)), ref_ndsm, s_tf,
                                    tuple(geo.raster_bounds(s_tf, W, H)), px, px, cfg)
        for n_cand, f in res:
            assert n_cand == len(rings) and len(want) > 10
            assert np.array_equal(f.poly_id.cpu().numpy(), np.array([int(o["poly_id"]) for o in want]))
            assert np.array_equal(f.tree_height.cpu().numpy(), np.array([o["TreeHeight"] for o in want], np.float32))
            assert np.array_equal(f.verts.cpu().numpy(), np.array([q for o in want for q in o["coords"]]).reshape(-1, 2))
        out[rank] = "ok"
    except Exception as e:          # reported to the parent, which asserts
        out[rank] = f"{type(e).__name__}: {e}"
    finally:
        dist.destroy_process_group()


def test_seam_strip_over_nccl_matches_oracle():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (the exchange step of the row-sharded mosaic)")
    import torch.multiprocessing as mp
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert dict(out) == {0: "ok", 1: "ok"}, dict(out)
