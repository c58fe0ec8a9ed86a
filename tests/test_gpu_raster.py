"""GPU parity of P1 (tile cut + normalise) against PIL / torch-CPU (the libraries the
reference itself calls) and of P0a (seam strips) against the mosaic + crop restatement."""
import numpy as np
import pytest
import torch

from oracle import port
from treedetection_b200 import ops, synth, tiling

pytestmark = pytest.mark.gpu


def _tables(tiles):
    win = np.array([m["window"] for m in tiles.values()], dtype=np.int32)
    net = np.array([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in win], dtype=np.int32)
    return win, net


def test_tile_cut_normalize_uint8_bit_exact(dev):
    sc = synth.make_scene(seed=2, size_px=1100, px=0.2, ndsm_px=1.0, density_per_km2=3000.0)
    win, net = _tables(sc.tiles)
    out, off, flag = ops.tile_cut_normalize(torch.from_numpy(sc.rgbi).to(dev), torch.from_numpy(win),
                                            torch.from_numpy(net))
    out = out.cpu().numpy(); off = off.numpy()
    assert (flag.cpu().numpy() == 0).all()
    shapes = set()
    for t in range(len(win)):
        ref = port.tile_cut_normalize(sc.rgbi, tuple(int(v) for v in win[t]))
        nh, nw = net[t]
        assert ref.shape == (3, nh, nw)
        shapes.add((int(win[t][3]), int(win[t][2]), int(nh), int(nw)))
        np.testing.assert_array_equal(out[off[t]:off[t + 1]].reshape(3, nh, nw), ref)
    assert len(shapes) >= 3          # corner, edge and interior tiles


def test_tile_cut_normalize_tma_path(dev):
    """Raster width a multiple of 16: the TMA-staged kernel (UTMALDG) runs; windows start at
    arbitrary, mostly 16-byte-unaligned columns."""
    rng = np.random.default_rng(12)
    img = rng.integers(0, 256, size=(4, 1200, 1600), dtype=np.uint8)
    tf = synth.image_transform(synth.ORIGIN_X, synth.ORIGIN_Y + 1200 * 0.25, 0.25)
    tiles = tiling.tile_grid("x", tf, 1600, 1200, 25832, 50, 50, 20)
    win, net = _tables(tiles)
    win = np.concatenate([win, np.array([[3, 5, 357, 211], [1243, 7, 357, 300], [0, 0, 1600, 1200][:4]], dtype=np.int32)])
    win[-1] = [1201, 901, 399, 299]
    net = np.array([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in win], dtype=np.int32)
    out, off, _ = ops.tile_cut_normalize(torch.from_numpy(img).to(dev), torch.from_numpy(win), torch.from_numpy(net))
    out = out.cpu().numpy(); off = off.numpy()
    for t in range(len(win)):
        ref = port.tile_cut_normalize(img, tuple(int(v) for v in win[t]))
        np.testing.assert_array_equal(out[off[t]:off[t + 1]].reshape(ref.shape), ref)


def test_tile_cut_normalize_downscale(dev):
    rng = np.random.default_rng(1)
    img = rng.integers(0, 256, size=(4, 1300, 1700), dtype=np.uint8)
    win = np.array([[0, 0, 1700, 1300], [100, 50, 1000, 1000], [7, 9, 901, 333]], dtype=np.int32)
    net = np.array([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in win], dtype=np.int32)
    out, off, _ = ops.tile_cut_normalize(torch.from_numpy(img).to(dev), torch.from_numpy(win), torch.from_numpy(net))
    out = out.cpu().numpy(); off = off.numpy()
    for t in range(len(win)):
        ref = port.tile_cut_normalize(img, tuple(int(v) for v in win[t]))
        np.testing.assert_array_equal(out[off[t]:off[t + 1]].reshape(ref.shape), ref)


def test_tile_cut_normalize_uint16_branch(dev):
    rng = np.random.default_rng(4)
    img8 = rng.integers(0, 256, size=(4, 600, 700), dtype=np.uint8)
    img = img8.astype(np.uint16) * 257
    img[:, 300:, 350:] = img8[:, 300:, 350:]           # one quadrant stays <= 255: the reference skips that tile
    win = np.array([[0, 0, 350, 300], [350, 300, 350, 300], [100, 100, 450, 400]], dtype=np.int32)
    net = np.array([tiling.resize_shortest_edge(int(w[3]), int(w[2])) for w in win], dtype=np.int32)
    dimg = torch.from_numpy(img.view(np.int16)).to(dev)
    out, off, flag = ops.tile_cut_normalize(dimg, torch.from_numpy(win), torch.from_numpy(net))
    out = out.cpu().numpy(); off = off.numpy(); flag = flag.cpu().numpy()
    for t in range(len(win)):
        ref = port.tile_cut_normalize(img, tuple(int(v) for v in win[t]))
        if ref is None:
            assert flag[t] == 2
            continue
        assert flag[t] == 1
        got = out[off[t]:off[t + 1]].reshape(ref.shape)
        np.testing.assert_allclose(got, ref, atol=1e-5, rtol=0)      # float branch: north_star tolerance
    assert list(flag) == [1, 2, 1]


@pytest.mark.parametrize("dtype", [np.uint8, np.float32, np.uint16])
def test_seam_crop(dev, dtype):
    rng = np.random.default_rng(8)
    bands = 4 if dtype == np.uint8 else 1
    a = (rng.uniform(0, 250, (bands, 500, 640))).astype(dtype)
    b = (rng.uniform(0, 250, (bands, 500, 640))).astype(dtype)

    def to_dev(x):
        return torch.from_numpy(x.view(np.int16) if dtype == np.uint16 else x).to(dev)

    for axis, (sw, sh) in ((0, (270, 500)), (1, (640, 270))):
        ref, _ = port.seam_crop(a, b, axis, sw, sh)
        got = ops.seam_crop(to_dev(a), to_dev(b), axis, sw, sh).cpu().numpy()
        if dtype == np.uint16:
            got = got.view(np.uint16)
        np.testing.assert_array_equal(got, ref)
