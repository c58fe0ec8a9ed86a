"""P0a -- seam-strip builder, mirroring ``TreeDetection/merging.py`` and the neighbour
search of ``TreeDetection/helpers.py:984-1021``.

For every image the strip across the seam to its RIGHT and to its LOWER neighbour is
written to ``<dir>/<merged_path>/``; the strips are then tiled / predicted / post-processed
like ordinary images.  The reference mosaics both whole images with ``rasterio.merge`` and
keeps 2.7 % of the result; here ``td_seam_crop`` gathers the strip straight from the two
rasters on the device."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import geo, geotiff, ops


def retrieve_neighboring_image_filenames(filename, other_filenames, meta_info):
    """helpers.py:984-1021 -- neighbours by geotransform origin, eps 1e-3 (note the
    reference uses ``transform.a`` for both axes)."""
    info = meta_info[filename]
    transform, width, height = info.transform, info.width, info.height
    x, y = transform[2], transform[5]
    left = right = up = down = None
    eps = 1e-3
    for other in other_filenames:
        if other == filename:
            continue
        ot = meta_info[other].transform
        if abs(ot[2] - (x - (width * ot[0]))) < eps and abs(ot[5] - y) < eps:
            left = other
        if abs(ot[2] - (x + (width * ot[0]))) < eps and abs(ot[5] - y) < eps:
            right = other
        if abs(ot[5] - (y + (height * ot[0]))) < eps and abs(ot[2] - x) < eps:
            up = other
        if abs(ot[5] - (y - (height * ot[0]))) < eps and abs(ot[2] - x) < eps:
            down = other
    return left, right, up, down


def _to_device(arr, device):
    a = np.ascontiguousarray(arr)
    if a.dtype == np.uint16:
        a = a.view(np.int16)
    return torch.from_numpy(a).to(device)


def _from_device(t, dtype):
    a = t.cpu().numpy()
    return a.view(np.uint16) if dtype == np.uint16 else a


def merge_and_crop_images(config, images_paths, height_paths, device=None):
    """merging.py:10-119.  Appends the new strip paths to both lists in place."""
    logger = config.get("logger")
    merged_directory = config["merged_path"]
    device = device if device is not None else torch.device("cuda", int(config.get("device", 0) or 0))
    cache = {}

    def load(path):
        if path not in cache:
            if len(cache) > 6:
                cache.pop(next(iter(cache)))
            arr, info = geotiff.read(path)
            cache[path] = (_to_device(arr, device), info, arr.dtype)
        return cache[path]

    def save_cropped_images(paths, rgbi=True):
        out_names = []
        meta_info = {f: geotiff.read_info(f) for f in paths}
        for f in list(paths):
            _, right, _, down = retrieve_neighboring_image_filenames(f, paths, meta_info)
            directory = os.path.dirname(f)
            result_directory = f"{directory}/{merged_directory}"
            os.makedirs(result_directory, exist_ok=True)
            base = os.path.basename(f).replace(".tif", "")
            f_basename, f_name_end = base.split("_")[0], base.split("_")[-1]
            fx, fy = meta_info[f].transform[2], meta_info[f].transform[5]
            for nb, axis in ((right, 0), (down, 1)):
                if nb is None:
                    continue
                try:
                    a, ia, dt = load(f)
                    b, ib, _ = load(nb)
                    nx, ny = meta_info[nb].transform[2], meta_info[nb].transform[5]
                    if rgbi:
                        name = f"{f_basename}_{round(fx)}_{round(fy)}_{round(nx)}_{round(ny)}_{f_name_end}.tif"
                    else:
                        name = f"{f_basename}_{round(fx)}{round(fy)}{round(nx)}{round(ny)}_{f_name_end}.tif"
                    mw = ia.width + ib.width if axis == 0 else max(ia.width, ib.width)
                    mh = max(ia.height, ib.height) if axis == 0 else ia.height + ib.height
                    if axis == 0:
                        sw = int((config["tile_width"] + 2 * config["buffer"]) * config["overlapping_tiles_width"])
                        sh = mh
                    else:
                        sw = mw
                        sh = int((config["tile_height"] + 2 * config["buffer"]) * config["overlapping_tiles_height"])
                    sw, sh = min(sw, mw), min(sh, mh)
                    left = max(mw // 2 - sw // 2, 0)
                    top = max(mh // 2 - sh // 2, 0)
                    strip = ops.seam_crop(a, b, axis, sw, sh)
                    tf = geo.window_transform(ia.transform, left, top)
                    nodata = ia.nodata
                    if nodata is None or abs(nodata) > 1e10:
                        nodata = 0.0   # helpers.py:1037-1040
                    path = f"{result_directory}/{name}"
                    geotiff.write(path, _from_device(strip, dt), tf, epsg=ia.epsg, nodata=nodata)
                    out_names.append(path)
                except Exception as e:  # per-file failures are logged and skipped (merging.py:78,106)
                    if logger:
                        logger.error(f"Error merging images {f} and {nb}: {e}")
        return out_names

    try:
        cropped_images = save_cropped_images(images_paths, rgbi=True)
        cropped_heights = save_cropped_images(height_paths, rgbi=False)
        images_paths.extend(cropped_images)
        height_paths.extend(cropped_heights)
    except Exception as e:
        if logger:
            logger.error(f"Error merging and cropping images: {e}")
