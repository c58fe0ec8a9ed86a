"""Affine / window arithmetic of the raster side of the path (host, float64).

Restates what the reference gets from the ``affine`` and ``rasterio`` packages
(un-vendored third-party code): ``Affine.__mul__``, ``~Affine``, ``Affine.scale``,
``dataset.bounds``, ``dataset.window_transform`` and
``rasterio.features.geometry_window`` (floor / ceil of the inverse-transformed
shape bounds, intersected with the raster) as used at
``TreeDetection/preprocessing.py:99-101`` and ``TreeDetection/prediction.py:164``.
A transform is the 6-tuple ``(a, b, c, d, e, f)``.
"""
from __future__ import annotations

import math
from collections import namedtuple

BoundingBox = namedtuple("BoundingBox", "left bottom right top")
Window = namedtuple("Window", "col_off row_off width height")


def compose(t, o):
    """t * o (apply o first, then t)."""
    sa, sb, sc, sd, se, sf = t[:6]
    oa, ob, oc, od, oe, of = o[:6]
    return (sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc,
            sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf)


def apply(t, x, y):
    a, b, c, d, e, f = t[:6]
    return (x * a + y * b + c, x * d + y * e + f)


def invert(t):
    sa, sb, sc, sd, se, sf = t[:6]
    idet = 1.0 / (sa * se - sb * sd)
    ra = se * idet
    rb = -sb * idet
    rd = -sd * idet
    re = sa * idet
    return (ra, rb, -sc * ra - sf * rb, rd, re, -sc * rd - sf * re)


def scale(sx, sy=None):
    return (float(sx), 0.0, 0.0, 0.0, float(sx if sy is None else sy), 0.0)


def translation(x, y):
    return (1.0, 0.0, float(x), 0.0, 1.0, float(y))


def almost_equals(t, o, precision=1e-5):
    """affine.Affine.almost_equals over all nine coefficients."""
    return all(abs(x - y) < precision for x, y in zip(t[:6], o[:6]))


def raster_bounds(t, width, height):
    """rasterio ``dataset.bounds`` for a north-up transform."""
    a, b, c, d, e, f = t[:6]
    return BoundingBox(c, f + e * height, c + a * width, f)


def window_transform(t, col_off, row_off):
    """``dataset.window_transform(window)`` = transform * translation(col_off, row_off)."""
    return compose(t, translation(col_off, row_off))


def geometry_window(t, width, height, minx, miny, maxx, maxy):
    """``rasterio.features.geometry_window(dataset, [box], pad 0)`` intersected with
    the raster: outermost pixel indices containing the box."""
    inv = invert(t)
    cols, rows = [], []
    for (x, y) in ((minx, miny), (minx, maxy), (maxx, maxy), (maxx, miny)):
        c, r = apply(inv, x, y)
        cols.append(c)
        rows.append(r)
    row_start, row_stop = int(math.floor(min(rows))), int(math.ceil(max(rows)))
    col_start, col_stop = int(math.floor(min(cols))), int(math.ceil(max(cols)))
    c0, r0 = max(col_start, 0), max(row_start, 0)
    c1, r1 = min(col_stop, width), min(row_stop, height)
    if c1 <= c0 or r1 <= r0:
        raise ValueError("Input shapes do not overlap raster, check geometry of incoming Tifs.")
    return Window(c0, r0, c1 - c0, r1 - r0)
