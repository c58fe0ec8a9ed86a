"""TEST INFRASTRUCTURE: the oracle's decimated read (oracle/port.py decimate_bilinear -- a restatement of the
down-sampling bilinear RasterIO of GDAL, which is absent here: "parity unpinned" at that boundary) cross-checked
against an INDEPENDENT implementation of the same filter that is in the image: Pillow's antialiased
Image.resize(BILINEAR) (triangle kernel, support = scale, centre (i + 0.5) * scale, weights normalised per output
pixel).  Pillow accumulates float images in double and uint8 images in 22-bit fixed point, so the agreement is to
rounding, not bit for bit: it pins the FILTER (support, centring, normalisation, pass structure), not the last ulp."""
import numpy as np
import pytest
from PIL import Image

from oracle import port


@pytest.mark.parametrize("shape,out", [((500, 640), (100, 128)), ((333, 211), (111, 70)), ((400, 400), (57, 80)),
                                       ((250, 1000), (125, 100)), ((64, 64), (64, 64))])
def test_float_raster_agrees_with_pillow(shape, out):
    rng = np.random.default_rng(shape[0])
    band = (rng.normal(size=shape) * 10 + 20).astype(np.float32)
    got = port.decimate_bilinear(band, *out)
    want = np.asarray(Image.fromarray(band, mode="F").resize((out[1], out[0]), Image.BILINEAR), dtype=np.float32)
    assert got.shape == want.shape == out
    np.testing.assert_allclose(got, want, rtol=0, atol=2e-4)      # values ~ 20 +- 30: float32 vs double accumulation


@pytest.mark.parametrize("shape,factor", [((1000, 1000), 0.2), ((730, 1210), 0.2), ((611, 977), 1 / 3), ((700, 910), 1 / 7)])
def test_uint8_raster_agrees_with_pillow_to_one_level(shape, factor):
    rng = np.random.default_rng(shape[1])
    low = rng.integers(0, 256, size=(shape[0] // 16 + 2, shape[1] // 16 + 2)).astype(np.uint8)
    band = np.asarray(Image.fromarray(low).resize((shape[1], shape[0]), Image.BICUBIC))      # smooth + full range
    band = (band.astype(np.int16) + rng.integers(-6, 7, size=shape)).clip(0, 255).astype(np.uint8)
    oh, ow = int(shape[0] * factor), int(shape[1] * factor)
    got = port.decimate_bilinear(band, oh, ow).astype(np.int16)
    want = np.asarray(Image.fromarray(band).resize((ow, oh), Image.BILINEAR)).astype(np.int16)
    diff = np.abs(got - want)
    # Pillow rounds the horizontal pass to uint8 before the vertical one (and works in 22-bit fixed point), the
    # restated GDAL path keeps float32 between the passes: one level apart on a minority of the pixels, never two
    assert diff.max() <= 1
    assert (diff == 0).mean() > 0.85


def test_coefficients_are_the_normalised_triangle():
    """the tap table itself: support = scale, centre (i + 0.5) * scale, weights sum to one, symmetric in the interior"""
    bounds, coeffs = port._triangle_coeffs(1000, 200)
    assert len(bounds) == len(coeffs) == 200
    for i in (0, 1, 57, 198, 199):
        x0, n = bounds[i]
        k = np.asarray(coeffs[i][:n], dtype=np.float64)
        assert abs(k.sum() - 1.0) < 1e-6 and (k >= 0).all()
        centre = (i + 0.5) * 5.0
        assert x0 >= max(0, int(centre - 5.0 + 0.5)) and x0 + n <= min(1000, int(centre + 5.0 + 0.5))
    x0, n = bounds[57]
    k = np.asarray(coeffs[57][:n], dtype=np.float64)
    assert n == 10 and x0 == 57 * 5 - 2
    # pixel x has its centre at x + 0.5: the peak sits on pixel 5 i + 2 (centre 5 i + 2.5), the triangle reaches zero
    # five pixels away, i.e. exactly on the last tap of the window
    assert int(np.argmax(k)) == 4 and abs(k[4] - 0.2) < 1e-12 and abs(k[9]) < 1e-12
    np.testing.assert_allclose(k[:9], k[:9][::-1], atol=1e-12)
    np.testing.assert_allclose(k[:5], 0.04 * np.arange(1, 6), atol=1e-12)
