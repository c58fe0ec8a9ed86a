"""The sync-free chain (td_chain_*: capacity workspace, device-side counts, CUDA-graph replay, fused
simplify; driven by pipeline.ChainRunner) must return exactly what the exact-size composition of the
separate kernels returns -- same crowns, same order, same bits -- and must fall back to it when a
capacity is too small."""
import numpy as np
import pytest
import torch

from treedetection_b200 import ops, pipeline, synth

pytestmark = pytest.mark.gpu

FIELDS = ("verts", "ring_off", "poly_id", "conf", "area", "tree_height", "centroid", "is_contained", "num_contained")


def _setup(dev, seed, size_px=1500, ndsm_px=0.2, density=4000.0):
    sc = synth.make_scene(seed=seed, size_px=size_px, px=0.2, ndsm_px=ndsm_px, density_per_km2=density)
    d = sc.det
    tile_tf, boxes_int = pipeline.tile_tables(sc.tiles, dev)
    det = dict(boxes_net=torch.from_numpy(d.boxes_net).to(dev), scores=torch.from_numpy(d.scores).to(dev),
               probs=torch.from_numpy(d.probs).to(dev), inst_tile=torch.from_numpy(d.inst_tile).to(dev),
               tile_dims=torch.from_numpy(d.tile_dims).to(dev))
    rgbi = torch.from_numpy(sc.rgbi).to(dev)
    ndsm = torch.from_numpy(sc.ndsm).to(dev)
    p = pipeline.PipelineParams()
    rasters = lambda: pipeline.raster_stage(rgbi, sc.transform, ndsm, sc.ndsm_transform, p)
    return sc, det, tile_tf, pipeline.filter_boxes(boxes_int, 1, dev), rasters, p


def _same(a, b):
    for f in FIELDS:
        x, y = getattr(a, f).cpu().numpy(), getattr(b, f).cpu().numpy()
        assert x.shape == y.shape, f
        np.testing.assert_array_equal(x, y, err_msg=f)


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("ndsm_px", [0.2, 1.0])
def test_dyn_chain_equals_exact_chain(dev, ndsm_px, graph):
    sc, det, tile_tf, tile_boxes, rasters, p = _setup(dev, 11, ndsm_px=ndsm_px)
    table = pipeline.predict_stage(**det, tile_tf=tile_tf, tile_boxes=tile_boxes, p=p)
    ref = pipeline.postprocess_stage(table, rasters(), p, keep_debug=True)
    run = pipeline.ChainRunner(p, use_graph=graph)
    n0, f0 = run.collect(run.submit(det, tile_tf, tile_boxes, rasters))       # exact path, learns capacities
    assert n0 == len(table)
    _same(f0, ref)
    r = rasters()                                                             # same raster buffers -> same graph
    for rep in range(2):
        tickets = [run.submit(det, tile_tf, tile_boxes, lambda: r) for _ in range(3)]   # enqueued back to back
        assert all(t[0] == "dyn" for t in tickets)
        for t in tickets:
            n, f = run.collect(t)
            assert n == len(table) and len(f) == len(ref) and len(f) > 50
            _same(f, ref)
            tab = run.table(run.last_counters, t[3])                          # the stitched table of that image
            np.testing.assert_array_equal(tab.verts.cpu().numpy(), table.verts.cpu().numpy())
            np.testing.assert_array_equal(tab.ring_off.cpu().numpy(), table.ring_off.cpu().numpy())
            np.testing.assert_array_equal(tab.conf.cpu().numpy(), table.conf.cpu().numpy())
    assert run.fallbacks == 0


def test_dyn_chain_too_many_in_flight(dev):
    _, det, tile_tf, tile_boxes, rasters, p = _setup(dev, 12, size_px=1000)
    run = pipeline.ChainRunner(p, n_slots=2)
    run.collect(run.submit(det, tile_tf, tile_boxes, rasters))
    t = [run.submit(det, tile_tf, tile_boxes, rasters) for _ in range(2)]
    with pytest.raises(ops._lib.TreedetError):
        run.submit(det, tile_tf, tile_boxes, rasters)
    for x in t:
        run.collect(x)
    run.collect(run.submit(det, tile_tf, tile_boxes, rasters))


def test_dyn_chain_other_image_same_capacities(dev):
    """capacities learnt on one image serve another one of similar size"""
    _, det_a, tile_tf, tile_boxes, rasters_a, p = _setup(dev, 21)
    run = pipeline.ChainRunner(p)
    run.collect(run.submit(det_a, tile_tf, tile_boxes, rasters_a))
    sc, det_b, tile_tf_b, tile_boxes_b, rasters_b, _ = _setup(dev, 22)
    table = pipeline.predict_stage(**det_b, tile_tf=tile_tf_b, tile_boxes=tile_boxes_b, p=p)
    ref = pipeline.postprocess_stage(table, rasters_b(), p)
    n, f = run.collect(run.submit(det_b, tile_tf_b, tile_boxes_b, rasters_b))
    assert n == len(table)
    _same(f, ref)


@pytest.mark.parametrize("small", ["inst", "words", "px", "ptslots", "rings", "verts", "nbr", "contours"])
def test_capacity_overflow_falls_back(dev, small):
    sc, det, tile_tf, tile_boxes, rasters, p = _setup(dev, 31, size_px=1000)
    table = pipeline.predict_stage(**det, tile_tf=tile_tf, tile_boxes=tile_boxes, p=p)
    ref = pipeline.postprocess_stage(table, rasters(), p)
    run = pipeline.ChainRunner(p)
    run.collect(run.submit(det, tile_tf, tile_boxes, rasters))
    if small == "nbr":                                                    # too small on purpose
        run.nbr_per_crown = 1
    elif small == "contours":
        run.slot_contours = 0
    else:
        run.caps[small] = 3
    run.chain = None
    if small == "contours":
        with pytest.raises(ops._lib.TreedetError):
            run.submit(det, tile_tf, tile_boxes, rasters)                 # a zero capacity is refused outright
        return
    t = run.submit(det, tile_tf, tile_boxes, rasters)
    if small == "inst":                                                   # more instances than the workspace holds
        assert t[0] == "done"
        n, f = run.collect(t)
        assert n == len(table)
        _same(f, ref)
        return
    assert t[0] == "dyn"
    n, f = run.collect(t)
    assert run.fallbacks == 1 and n == len(table)
    _same(f, ref)
    for _ in range(4):        # capacities were re-learnt (the neighbour slots double per fallback: 1 -> 2 -> 4 -> 8)
        before = run.fallbacks
        n, f = run.collect(run.submit(det, tile_tf, tile_boxes, rasters))
        _same(f, ref)
        if run.fallbacks == before:
            break
    assert run.fallbacks == before and (small == "nbr" or run.fallbacks == 1)


def test_bookkeeping_ops(dev):
    rng = np.random.default_rng(0)
    sizes = torch.from_numpy(rng.integers(0, 50, size=(3, 1000))).to(dev)
    flag = torch.zeros(1, dtype=torch.int64, device=dev)
    offs, tot = ops.scan_clamp(sizes, [10**9, 10**9, 10**9], flag)
    ref = torch.zeros((3, 1001), dtype=torch.int64, device=dev); ref[:, 1:] = torch.cumsum(sizes, 1)
    assert torch.equal(offs, ref) and torch.equal(tot, ref[:, -1]) and int(flag) == 0
    cap = int(ref[1, 400])                       # row 1 overflows at item 400 (first item whose end > cap)
    win = torch.ones((1000, 4), dtype=torch.int32, device=dev)
    offs, tot = ops.scan_clamp(sizes, [10**9, cap, 10**9], flag, win_zero=win)
    i0 = int((ref[1, 1:] > cap).nonzero()[0])
    assert int(flag) == 1 and torch.equal(offs[:, :i0 + 1], ref[:, :i0 + 1])
    assert bool((offs[:, i0:] == ref[:, i0:i0 + 1]).all()) and torch.equal(tot, ref[:, i0])
    assert int(win[:i0, 2:].min()) == 1 and int(win[i0:, 2:].max()) == 0
    f = torch.from_numpy(rng.integers(0, 2, size=5000).astype(np.uint8)).to(dev)
    nd = torch.tensor([3000], dtype=torch.int64, device=dev)
    sel, cnt = ops.compact_flags(f, n_dev=nd)
    want = torch.nonzero(f[:3000]).flatten()
    assert int(cnt) == want.numel() and torch.equal(sel[:int(cnt)], want) and int(sel[int(cnt):].abs().max()) == 0
    v = torch.from_numpy(rng.integers(-3, 100, size=5000).astype(np.int32)).to(dev)
    out, cnt = ops.compact_nonneg(v, n_dev=nd)
    want = v[:3000][v[:3000] >= 0].long()
    assert int(cnt) == want.numel() and torch.equal(out[:int(cnt)], want)


def test_fused_bookkeeping_ops(dev):
    """td_ring_offsets, td_gather_rows, td_select_head against their torch compositions"""
    rng = np.random.default_rng(1)
    n_rings = 700
    lens = torch.from_numpy(rng.integers(4, 60, n_rings)).to(dev)
    ring_off = ops.exclusive_offsets(lens)
    count = torch.from_numpy(rng.integers(4, 30, n_rings).astype(np.int32)).to(dev)
    sel = torch.from_numpy(np.sort(rng.choice(n_rings, 300, replace=False))).to(dev)
    for cnt in (None, count):
        dst = torch.empty(sel.shape[0] + 1, dtype=torch.int64, device=dev)
        _ = ops._lib.call("td_ring_offsets", ops._ptr(ring_off), ops._ptr(cnt), ops._ptr(sel), sel.shape[0],
                          ops._ptr(dst), ops._stream())
        want = ops.exclusive_offsets((lens if cnt is None else cnt.long())[sel])
        assert torch.equal(dst, want)
    arrays = [torch.from_numpy(rng.normal(size=n_rings)).to(dev), torch.from_numpy(rng.normal(size=(n_rings, 4))).to(dev),
              torch.from_numpy(rng.integers(0, 255, n_rings).astype(np.uint8)).to(dev),
              torch.from_numpy(rng.normal(size=(n_rings, 2)).astype(np.float32)).to(dev),
              torch.from_numpy(rng.integers(0, 9, n_rings).astype(np.int32)).to(dev)]
    nd = torch.tensor([250], dtype=torch.int64, device=dev)
    outs = ops.gather_rows(arrays, sel, nd)
    for a, o in zip(arrays, outs):
        assert o.dtype == a.dtype and torch.equal(o[:250], a[sel][:250])
    conf = torch.from_numpy(np.round(rng.uniform(0.1, 1.0, n_rings), 3)).to(dev)
    area = torch.from_numpy(rng.uniform(0.0, 1500.0, n_rings)).to(dev)
    nd = torch.tensor([600], dtype=torch.int64, device=dev)
    flags, pid = ops.select_head(conf, area, 0.3, 1.0, 1000.0, n_dev=nd)
    valid = torch.arange(n_rings, device=dev) < 600
    ok = (conf >= 0.3) & valid
    want_flags = ok & (area >= 1.0) & (area <= 1000.0)
    assert torch.equal(flags.bool(), want_flags)
    want_pid = torch.cumsum(ok.long(), 0) - 1
    assert torch.equal(pid[ok], want_pid[ok])
