"""P0b -- tile grid (metadata only), mirroring ``TreeDetection/preprocessing.py``.

``tile_grid`` restates the loop of ``tile_single_file`` (preprocessing.py:57-120):
x outer / y inner from the raster's bottom-left corner, ids
``{stem}_{int(minx)}_{int(miny)}_{int(tw)}_{int(buffer)}_{epsg}``, bounds +- buffer,
window = geometry_window, transform = window transform.  Pixels are cut lazily at
predict time (P1), exactly as in the reference.
"""
from __future__ import annotations

import numpy as np

from . import geo


def tile_grid(stem, transform, width, height, epsg, tile_width=50, tile_height=50, buffer=20, forest=None):
    """Returns an ordered dict ``tile_id -> metadata`` (JSON-serialisable, the
    reference's schema) -- plus ``window`` which the reference recomputes at predict
    time and this implementation keeps.

    ``forest``: optional object with ``flags(minx, miny, maxx, maxy, bbox) ->
    (only_forest, only_urban)`` (config 3, see fusion.py)."""
    b = geo.raster_bounds(transform, width, height)
    out = {}
    for minx in np.arange(b.left, b.right, tile_width):
        for miny in np.arange(b.bottom, b.top, tile_height):
            minx_f, miny_f = float(minx), float(miny)
            tile_id = f"{stem}_{int(minx_f)}_{int(miny_f)}_{int(tile_width)}_{int(buffer)}_{epsg}"
            bounds = [minx_f - buffer, miny_f - buffer, minx_f + tile_width + buffer, miny_f + tile_height + buffer]
            only_forest, only_urban = False, False
            win = geo.geometry_window(transform, width, height, *bounds)
            tf = geo.window_transform(transform, win.col_off, win.row_off)
            out[tile_id] = {
                "crs": epsg,
                "transform": [tf[0], tf[1], tf[2], tf[3], tf[4], tf[5], 0.0, 0.0, 1.0],
                "bounds": bounds,
                "only_forest": bool(only_forest),
                "only_urban": bool(only_urban),
                "window": [win.col_off, win.row_off, win.width, win.height],
                "_tile_box": [minx_f, miny_f, minx_f + tile_width, miny_f + tile_height],
            }
    if forest is not None and out:
        only_forest, only_urban = forest.tile_flags([m["_tile_box"] for m in out.values()],
                                                    [m["bounds"] for m in out.values()])
        for k, m in enumerate(out.values()):
            m["only_forest"] = bool(only_forest[k])
            m["only_urban"] = bool(only_urban[k])
    for m in out.values():
        del m["_tile_box"]
    return out


def resize_shortest_edge(h, w, short=800, max_size=1333):
    """detectron2 ``ResizeShortestEdge.get_output_shape`` (the predictor's test-time
    augmentation: MIN_SIZE_TEST 800, MAX_SIZE_TEST 1333)."""
    size = short * 1.0
    scale = size / min(h, w)
    if h < w:
        newh, neww = size, scale * w
    else:
        newh, neww = scale * h, size
    if max(newh, neww) > max_size:
        scale = max_size * 1.0 / max(newh, neww)
        newh = newh * scale
        neww = neww * scale
    return int(newh + 0.5), int(neww + 0.5)
