"""N1 on the device: td_tiff_lzw_decode_batch + td_tiff_place_chunks (the file crosses PCIe still compressed, one
warp per strip / tile, 32 codes per step) against the host reader -- itself pinned to PIL / libtiff in tests/test_geotiff_codec.py --
on files written by libtiff: strips and tiles, chunky RGBA and single bands, predictor 1 / 2, float32, noise (codes
of every width, table resets), long runs (KwKwK strings, long copies)."""
import numpy as np
import pytest
import torch
from PIL import Image

from treedetection_b200 import geotiff

pytestmark = pytest.mark.gpu


def _save(path, arr, predictor=None, tile=None, rows_per_strip=None):
    img = Image.fromarray(np.ascontiguousarray(arr.transpose(1, 2, 0))) if arr.ndim == 3 else Image.fromarray(arr)
    info = {}
    if predictor:
        info[317] = predictor
    if tile:
        info.update({322: tile, 323: tile})
    if rows_per_strip:
        info[278] = rows_per_strip
    img.save(path, format="TIFF", compression="tiff_lzw", tiffinfo=info)


def _cases():
    rng = np.random.default_rng(0)
    smooth = (np.add.outer(np.arange(700), np.arange(1017)) % 251).astype(np.uint8)
    rgba = np.stack([smooth, smooth[::-1], (smooth // 3), rng.integers(0, 255, smooth.shape).astype(np.uint8)])
    runs = np.zeros((500, 2048), np.uint8)
    runs[100:300, 300:1500] = 200                      # long constant runs: KwKwK codes, strings of hundreds of bytes
    f32 = (np.sin(np.arange(700)[:, None] / 17.0) * 20 + rng.normal(0, 0.1, (700, 1017))).astype(np.float32)
    alt = np.zeros((300, 3000), np.uint8)
    alt[:, ::2] = 9                                    # period-2 pattern: every code names the entry made just before
    ramp = (np.arange(300 * 3000) % 7).astype(np.uint8).reshape(300, 3000)
    return {"rgba": rgba, "band": smooth, "runs": runs, "f32": f32,
            "noise": rng.integers(0, 256, (300, 4000)).astype(np.uint8),
            "few": rng.integers(0, 4, (300, 9000)).astype(np.uint8),      # table fills up and is reset many times
            "const": np.full((400, 2500), 123, np.uint8),                 # one long KwKwK chain
            "alt": alt, "ramp": ramp}


@pytest.mark.parametrize("name,predictor,tile", [("rgba", None, None), ("rgba", 2, None), ("band", None, None),
                                                 ("band", 2, None), ("runs", None, None), ("runs", 2, None),
                                                 ("noise", None, None), ("f32", None, None), ("rgba", None, 128),
                                                 ("rgba", 2, 256), ("band", 2, 128), ("few", None, None),
                                                 ("few", 2, None), ("const", None, None), ("const", 2, None),
                                                 ("alt", None, None), ("alt", 2, 128), ("ramp", None, None)])
def test_device_reader_matches_host_reader(tmp_path, dev, name, predictor, tile):
    arr = _cases()[name]
    path = str(tmp_path / "x.tif")
    try:
        _save(path, arr, predictor, tile)
    except Exception as e:
        pytest.skip(f"PIL cannot write this combination: {e}")
    ref, rinfo = geotiff.read(path)
    np.testing.assert_array_equal(ref, arr if arr.ndim == 3 else arr[None])
    if not geotiff.device_decodable(path):
        pytest.skip("layout outside the device decoder (host reader covers it)")
    got, info, status = geotiff.read_device(path, dev)
    torch.cuda.synchronize()
    assert int(status.item()) == 0
    assert info == rinfo and tuple(got.shape) == ref.shape
    np.testing.assert_array_equal(got.cpu().numpy(), ref)


def test_device_reader_into_caller_buffer_and_limits(tmp_path, dev):
    arr = _cases()["rgba"]
    path = str(tmp_path / "x.tif")
    _save(path, arr, 2)
    out = torch.full(arr.shape, 7, dtype=torch.uint8, device=dev)
    got, _, status = geotiff.read_device(path, dev, out=out, slot=1)
    torch.cuda.synchronize()
    assert got is out and int(status.item()) == 0
    np.testing.assert_array_equal(out.cpu().numpy(), arr)
    with pytest.raises(ValueError):
        geotiff.read_device(path, dev, out=torch.zeros((4, 10, 10), dtype=torch.uint8, device=dev))
    # strips of more than 1 MiB and uncompressed files stay with the host reader
    big = np.zeros((600, 4096), np.uint8)
    _save(str(tmp_path / "big.tif"), big, rows_per_strip=512)
    plain = str(tmp_path / "plain.tif")
    geotiff.write(plain, arr, (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0), epsg=25832)
    assert not geotiff.device_decodable(plain) and geotiff.read_device(plain, dev) is None
    if not geotiff.device_decodable(str(tmp_path / "big.tif")):
        assert geotiff.read_device(str(tmp_path / "big.tif"), dev) is None


def test_corrupt_stream_sets_the_status(tmp_path, dev):
    arr = _cases()["band"]
    path = str(tmp_path / "x.tif")
    _save(path, arr)
    raw = bytearray(open(path, "rb").read())
    import mmap
    tags = geotiff._read_ifd(bytes(raw), "<")
    off, cnt = tags[273][3], tags[279][3]
    raw[off + 2:off + cnt] = b"\xff" * (cnt - 2)      # codes beyond the table
    open(path, "wb").write(raw)
    got, _, status = geotiff.read_device(path, dev)
    torch.cuda.synchronize()
    assert int(status.item()) != 0


def test_plain_reader_streams_uncompressed_strips_to_the_device(tmp_path, dev):
    """geotiff.read_device_plain: file -> two-slot pinned ring -> device tensor, pieces smaller than a band (several
    pieces per run, runs across bands), float32 and uint8; other layouts are declined"""
    cases = _cases()
    tf = (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0)
    for name, piece in (("rgba", 100_000), ("rgba", 32 << 20), ("f32", 70_001), ("band", 4096)):
        arr = cases[name]
        path = str(tmp_path / f"{name}.tif")
        geotiff.write(path, arr, tf, epsg=25832)
        ref, rinfo = geotiff.read(path)
        got, info = geotiff.read_device_plain(path, dev, piece=piece)
        torch.cuda.synchronize()
        assert info == rinfo
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
        out = torch.zeros(ref.shape, dtype=got.dtype, device=dev)
        assert geotiff.read_device_plain(path, dev, out=out, slot=1, piece=piece)[0] is out
        torch.cuda.synchronize()
        np.testing.assert_array_equal(out.cpu().numpy(), ref)
    lzw = str(tmp_path / "lzw.tif")
    geotiff.write(lzw, cases["rgba"], tf, epsg=25832, compression="lzw")
    chunky = str(tmp_path / "chunky.tif")
    Image.fromarray(np.ascontiguousarray(cases["rgba"].transpose(1, 2, 0))).save(chunky, format="TIFF")
    assert geotiff.read_device_plain(lzw, dev) is None and geotiff.read_device_plain(chunky, dev) is None
    assert geotiff.read_device_plain(lzw, dev, probe=True) is None
    assert geotiff.read_device_plain(str(tmp_path / "rgba.tif"), dev, probe=True) is True


def test_bigtiff_on_the_device_paths(tmp_path, dev):
    """BigTIFF files (64-bit offsets) through the device LZW decoder and the streaming reader of uncompressed strips:
    same pixels as the host reader, which tests/test_geotiff_codec.py pins against libtiff"""
    cases = _cases()
    tf = (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0)
    for name, predictor in (("rgba", 2), ("f32", 1), ("band", 1)):
        arr = cases[name]
        path = str(tmp_path / f"{name}_lzw.tif")
        geotiff.write(path, arr, tf, epsg=25832, compression="lzw", predictor=predictor, bigtiff=True)
        ref, rinfo = geotiff.read(path)
        np.testing.assert_array_equal(ref, arr if arr.ndim == 3 else arr[None])
        assert geotiff.device_decodable(path)
        got, info, status = geotiff.read_device(path, dev)
        torch.cuda.synchronize()
        assert int(status.item()) == 0 and info == rinfo
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
        plain = str(tmp_path / f"{name}_plain.tif")
        geotiff.write(plain, arr, tf, epsg=25832, bigtiff=True)
        got, info = geotiff.read_device_plain(plain, dev, piece=100_000)
        torch.cuda.synchronize()
        assert info == rinfo
        np.testing.assert_array_equal(got.cpu().numpy(), ref)
