"""N1: the GeoTIFF reader against PIL / libtiff on compressed files -- LZW and deflate,
predictors 1 / 2 / 3, strips and tiles, uint8 / uint16 / float32 (the rasters the reference reads with
rasterio: prediction.py:61, postprocessing.py:781-800).  No GPU needed: td_tiff_lzw_decode is a host
function of libtreedet."""
import ctypes as C
import os

import numpy as np
import pytest
from PIL import Image

from treedetection_b200 import _lib, geotiff


def _save(path, arr, compression, predictor=None, tile=None):
    if arr.ndim == 3:
        img = Image.fromarray(np.ascontiguousarray(arr.transpose(1, 2, 0)))
    else:
        img = Image.fromarray(arr)
    info = {}
    if predictor:
        info[317] = predictor
    kw = {"compression": compression, "tiffinfo": info}
    if tile:
        kw["tiffinfo"] = {**info, 322: tile, 323: tile}
    img.save(path, format="TIFF", **kw)


def _cases():
    rng = np.random.default_rng(0)
    smooth = (np.add.outer(np.arange(300), np.arange(517)) % 251).astype(np.uint8)
    rgba = np.stack([smooth, smooth[::-1], (smooth // 3), rng.integers(0, 255, smooth.shape).astype(np.uint8)])
    f32 = (np.sin(np.arange(300)[:, None] / 17.0) * 20 + rng.normal(0, 0.1, (300, 517))).astype(np.float32)
    u16 = (np.add.outer(np.arange(300), np.arange(517)) * 37 % 60000).astype(np.uint16)
    return {"rgba": rgba, "f32": f32, "u16": u16, "noise": rng.integers(0, 256, (64, 4000)).astype(np.uint8)}


@pytest.mark.parametrize("name", ["rgba", "f32", "u16", "noise"])
@pytest.mark.parametrize("compression,predictor", [("tiff_lzw", None), ("tiff_lzw", 2), ("tiff_adobe_deflate", None),
                                                   ("tiff_adobe_deflate", 2), ("tiff_lzw", 3)])
def test_reader_matches_pil(tmp_path, name, compression, predictor):
    arr = _cases()[name]
    if predictor == 3 and arr.dtype != np.float32:
        pytest.skip("floating-point predictor is for float samples")
    if predictor == 2 and arr.dtype == np.float32:
        pytest.skip("libtiff refuses horizontal differencing for float samples")
    path = str(tmp_path / "x.tif")
    try:
        _save(path, arr, compression, predictor)
    except Exception as e:       # a libtiff build without this combination
        pytest.skip(f"PIL cannot write {compression}/{predictor}: {e}")
    with Image.open(path) as im:
        ref = np.array(im)
    got, info = geotiff.read(path)
    got = got[0] if arr.ndim == 2 else got.transpose(1, 2, 0)
    assert got.dtype == ref.dtype
    np.testing.assert_array_equal(got, ref)
    np.testing.assert_array_equal(got, arr if arr.ndim == 2 else arr.transpose(1, 2, 0))


def test_lzw_decoder_errors_and_kwkwk():
    lib = _lib.lib()
    # the classic KwKwK case: 'aaaaaaa...' forces code == next
    data = np.full((1, 5000), 97, np.uint8)
    import io
    buf = io.BytesIO()
    Image.fromarray(data).save(buf, format="TIFF", compression="tiff_lzw")
    raw = buf.getvalue()
    import tempfile
    with tempfile.NamedTemporaryFile(suffix=".tif", delete=False) as f:
        f.write(raw)
    got, _ = geotiff.read(f.name)
    os.unlink(f.name)
    np.testing.assert_array_equal(got[0], data)
    dst = C.create_string_buffer(16)
    assert lib.td_tiff_lzw_decode(b"\x80\x00", 2, dst, 0) == 0            # ClearCode then end of input
    bad = bytes([0xFF, 0xFF, 0xFF, 0xFF])
    assert lib.td_tiff_lzw_decode(bad, len(bad), dst, 16) < 0             # code beyond the table
    assert b"corrupt" in lib.td_last_error()


@pytest.mark.parametrize("compression,tile", [(None, None), ("tiff_adobe_deflate", None), ("tiff_adobe_deflate", 128),
                                              ("tiff_lzw", None)])
def test_windowed_read_decodes_only_what_it_needs(tmp_path, compression, tile):
    """rasterio's windowed reads (prediction.py:164, postprocessing.py:781-800): any window of a strip / tile
    file equals the slice of the full read; the destination may be a caller buffer (a pinned staging area)."""
    arr = _cases()["rgba"]
    path = str(tmp_path / "w.tif")
    if compression is None:
        geotiff.write(path, arr, (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0), epsg=25832)
    else:
        try:
            _save(path, arr, compression, None, tile)
        except Exception as e:
            pytest.skip(f"PIL cannot write {compression}/{tile}: {e}")
    full, info = geotiff.read(path)
    np.testing.assert_array_equal(full, arr)
    rng = np.random.default_rng(1)
    for _ in range(12):
        w, h = int(rng.integers(1, 300)), int(rng.integers(1, 200))
        c0, r0 = int(rng.integers(0, 517 - w + 1)), int(rng.integers(0, 300 - h + 1))
        dst = np.full((4, h, w), 77, np.uint8)
        got, winfo = geotiff.read(path, window=(c0, r0, w, h), out=dst)
        assert got is dst
        np.testing.assert_array_equal(got, arr[:, r0:r0 + h, c0:c0 + w])
        assert (winfo.width, winfo.height) == (w, h)
    with pytest.raises(ValueError):
        geotiff.read(path, window=(500, 0, 100, 10))
    with pytest.raises(ValueError):
        geotiff.read(path, out=np.zeros((4, 10, 10), np.uint8))


@pytest.mark.parametrize("name,predictor", [("band", 1), ("band", 2), ("f32", 1), ("noise", 1), ("runs", 2), ("few", 1)])
def test_lzw_writer_is_read_by_libtiff(tmp_path, name, predictor):
    """td_tiff_lzw_encode (geotiff.write(compression="lzw")) pinned against libtiff: PIL decodes the file to the same
    pixels -- smooth data, noise (codes of every width), long runs (KwKwK), a 4-symbol alphabet on long strips (the
    table fills up and is reset many times) -- and so does this package's own reader"""
    rng = np.random.default_rng(4)
    smooth = (np.add.outer(np.arange(300), np.arange(517)) % 251).astype(np.uint8)
    runs = np.zeros((200, 2048), np.uint8)
    runs[50:120, 300:1500] = 200
    arr = {"band": smooth, "f32": _cases()["f32"], "noise": _cases()["noise"], "runs": runs,
           "few": rng.integers(0, 4, (300, 9000)).astype(np.uint8)}[name]
    path = str(tmp_path / "w.tif")
    geotiff.write(path, arr, (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0), epsg=25832, compression="lzw",
                  predictor=predictor)
    with Image.open(path) as im:
        np.testing.assert_array_equal(np.array(im), arr)
    got, info = geotiff.read(path)
    np.testing.assert_array_equal(got[0], arr)
    assert info.epsg == 25832 and info.transform == (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0)


def test_lzw_writer_planar_bands_round_trip(tmp_path):
    arr = _cases()["rgba"]
    path = str(tmp_path / "w.tif")
    for predictor in (1, 2):
        geotiff.write(path, arr, (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0), epsg=25832, compression="lzw",
                      predictor=predictor)
        got, _ = geotiff.read(path)
        np.testing.assert_array_equal(got, arr)
        win, _ = geotiff.read(path, window=(100, 50, 200, 120))
        np.testing.assert_array_equal(win, arr[:, 50:170, 100:300])
    with pytest.raises(ValueError):
        geotiff.write(path, _cases()["f32"], (1, 0, 0, 0, -1, 0), compression="lzw", predictor=2)


def test_lzw_encoder_decoder_agree_at_every_stream_length():
    """The decoder adds a table entry (and may widen the codes, or expect a ClearCode) after the LAST data code as
    well, so the encoder must account for it before EOI (libtiff's LZWPostEncode): streams whose final code lands
    exactly on a width boundary (254, 510, 1022, 2046 codes) or on a full table (3836 codes)"""
    lib = _lib.lib()
    rng = np.random.default_rng(9)
    noise = rng.integers(0, 256, 5000, dtype=np.uint8).tobytes()      # ~one code per byte
    lengths = list(range(1, 40)) + list(range(240, 270)) + list(range(500, 520)) + list(range(1015, 1030)) + \
        list(range(2040, 2052)) + list(range(3825, 3850)) + [4999]
    for n in lengths:
        raw = noise[:n]
        enc = geotiff._lzw_encode(raw)
        dst = C.create_string_buffer(n + 8)
        got = lib.td_tiff_lzw_decode(enc, len(enc), dst, n + 8)
        assert got == n and dst.raw[:n] == raw, n


TF = (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0)


@pytest.mark.parametrize("mode", ["gray", "rgb", "f32"])
def test_bigtiff_reader_matches_pil(tmp_path, mode):
    """BigTIFF (magic 43, 64-bit offsets: what GDAL writes for rasters beyond 4 GB) written by PIL's own writer:
    the reader returns the same pixels -- single band, chunky RGB, float32 -- whole and windowed"""
    rng = np.random.default_rng(9)
    arr = {"gray": rng.integers(0, 255, (211, 333), dtype=np.uint8),
           "rgb": rng.integers(0, 255, (97, 120, 3), dtype=np.uint8),
           "f32": rng.normal(size=(64, 75)).astype(np.float32)}[mode]
    path = str(tmp_path / "b.tif")
    try:
        Image.fromarray(arr).save(path, format="TIFF", big_tiff=True)
    except Exception as e:  # pragma: no cover
        pytest.skip(f"PIL cannot write BigTIFF: {e}")
    with open(path, "rb") as f:
        if f.read(4) != b"II+\x00":
            pytest.skip("this PIL ignores big_tiff")
    want = arr.transpose(2, 0, 1) if arr.ndim == 3 else arr[None]
    got, info = geotiff.read(path)
    np.testing.assert_array_equal(got, want)
    assert (info.width, info.height, info.count) == (want.shape[2], want.shape[1], want.shape[0])
    assert geotiff.read_info(path).dtype == arr.dtype
    win, _ = geotiff.read(path, window=(10, 20, 40, 30))
    np.testing.assert_array_equal(win, want[:, 20:50, 10:50])


@pytest.mark.parametrize("compression,predictor", [(None, 1), ("lzw", 1), ("lzw", 2)])
def test_bigtiff_writer_is_read_by_libtiff_and_round_trips(tmp_path, compression, predictor):
    """geotiff.write(bigtiff=True): libtiff (through PIL) decodes the single-band file to the same pixels; planar
    multi-band rasters, the geo tags and windowed reads round-trip through this package's reader; classic and
    BigTIFF layouts hold the same pixels; a classic file cannot be forced past 4 GB of offsets"""
    smooth = (np.add.outer(np.arange(300), np.arange(517)) % 251).astype(np.uint8)
    path = str(tmp_path / "w.tif")
    geotiff.write(path, smooth, TF, epsg=25832, compression=compression, predictor=predictor, bigtiff=True)
    with open(path, "rb") as f:
        assert f.read(4) == b"II+\x00"
    with Image.open(path) as im:
        np.testing.assert_array_equal(np.array(im), smooth)
    got, info = geotiff.read(path)
    np.testing.assert_array_equal(got[0], smooth)
    assert info.epsg == 25832 and info.transform == TF
    arr = _cases()["rgba"]
    geotiff.write(path, arr, TF, epsg=25832, nodata=0.0, compression=compression, predictor=predictor, bigtiff=True)
    got, info = geotiff.read(path)
    np.testing.assert_array_equal(got, arr)
    assert info.nodata == 0.0 and info.count == arr.shape[0]
    win, _ = geotiff.read(path, window=(100, 50, 200, 120))
    np.testing.assert_array_equal(win, arr[:, 50:170, 100:300])
    classic = str(tmp_path / "c.tif")
    geotiff.write(classic, arr, TF, epsg=25832, compression=compression, predictor=predictor)       # auto: classic
    with open(classic, "rb") as f:
        assert f.read(4) == b"II*\x00"
    np.testing.assert_array_equal(geotiff.read(classic)[0], arr)


def test_not_a_tiff_is_rejected(tmp_path):
    p = tmp_path / "x.tif"
    p.write_bytes(b"II\x2c\x00" + b"\x00" * 32)
    with pytest.raises(ValueError):
        geotiff.read_info(str(p))
    p.write_bytes(b"II+\x00\x04\x00\x00\x00" + b"\x00" * 32)       # BigTIFF magic with a wrong offset size
    with pytest.raises(ValueError):
        geotiff.read(str(p))
