"""P10 -- two-model fusion against a forest outline (config 3).

Reference: ``fuse_predictions`` (TreeDetection/helpers.py:703-834), the tile flags
``only_forest`` / ``only_urban`` (TreeDetection/preprocessing.py:67-96) and the per-model
tile exclusion (TreeDetection/prediction.py:79-93, implemented in ``predictor.py``).

The GEOS predicates (``intersects`` / ``within`` / ``contains`` against ``unary_union`` of the
forest polygons) run on the device (``td_forest_predicates``, csrc/forest_core.cuh) without
building the union.  Forest polygons are taken as single closed rings; interior rings (holes)
of the outline are ignored with a warning.  GEOS repair steps (``make_valid``, ``buffer(0)``)
are no-ops for the valid simple rings this path produces and are not reproduced.
"""
from __future__ import annotations

import os
import struct

import numpy as np
import torch
import yaml

from . import gpkg, ops


# ----------------------------------------------------------------------------
# outline readers
# ----------------------------------------------------------------------------
def _ring_signed_area(r):
    x, y = r[:, 0], r[:, 1]
    return 0.5 * float(np.sum(x[:-1] * y[1:] - x[1:] * y[:-1]))


def read_shapefile_polygons(path):
    """ESRI shapefile, shape types 5 / 15 / 25 (Polygon[Z/M]): list of outer rings (N,2) f64.
    Outer rings are clockwise in a shapefile; counter-clockwise rings are holes (skipped)."""
    rings, holes = [], 0
    with open(path, "rb") as f:
        data = f.read()
    pos = 100
    while pos + 8 <= len(data):
        _, clen = struct.unpack(">ii", data[pos:pos + 8])
        rec = data[pos + 8: pos + 8 + 2 * clen]
        pos += 8 + 2 * clen
        if len(rec) < 44:
            continue
        stype = struct.unpack("<i", rec[:4])[0]
        if stype not in (5, 15, 25):
            continue
        nparts, npoints = struct.unpack("<ii", rec[36:44])
        parts = list(struct.unpack("<" + "i" * nparts, rec[44:44 + 4 * nparts])) + [npoints]
        pts = np.frombuffer(rec, dtype="<f8", count=2 * npoints, offset=44 + 4 * nparts).reshape(npoints, 2)
        for a, b in zip(parts[:-1], parts[1:]):
            r = np.array(pts[a:b], dtype=np.float64)
            if len(r) < 4:
                continue
            if _ring_signed_area(r) <= 0:      # clockwise = outer ring
                rings.append(r)
            else:
                holes += 1
    return rings, holes


def read_outline(path, logger=None):
    if path.lower().endswith(".shp"):
        rings, holes = read_shapefile_polygons(path)
        if holes and logger:
            logger.warning(f"{holes} interior rings of the forest outline are ignored (holes are not supported).")
        return rings
    verts, off, _, _ = gpkg.read_layer(path)
    return [np.array(verts[off[i]:off[i + 1]]) for i in range(len(off) - 1) if off[i + 1] - off[i] >= 4]


def _ragged(rings, device):
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    if rings:
        off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.concatenate(rings).astype(np.float64) if rings else np.zeros((0, 2))
    return torch.from_numpy(np.ascontiguousarray(verts)).to(device), torch.from_numpy(off).to(device)


class ForestIndex:
    """Forest outline on the device + the two queries the path makes against it."""

    def __init__(self, rings, device):
        self.rings = rings
        self.device = device
        self.verts, self.off = _ragged(rings, device)
        self.bounds = ops.simplify_rings(self.verts, self.off, 0.0, want_bounds=True)["bounds"] if rings else None

    @classmethod
    def from_file(cls, path, device=None, logger=None):
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        rings = read_outline(path, logger)
        if not rings:
            raise ValueError(f"No valid geometries found in the forest shapefile {path}.")
        return cls(rings, device)

    def predicates(self, verts, ring_off, a_filter=None):
        """(intersects, within) uint8 tensors for the query rings."""
        return ops.forest_predicates(verts, ring_off, self.verts, self.off, self.bounds, a_filter)

    def tile_flags(self, tile_boxes, buffered_boxes):
        """preprocessing.py:67-96 for all tiles at once.  tile_boxes / buffered_boxes: (T,4)
        [minx, miny, maxx, maxy].  Returns (only_forest, only_urban) bool arrays."""
        t = len(tile_boxes)
        bb = np.asarray(buffered_boxes, dtype=np.float64).reshape(t, 4)
        # shapely box(minx, miny, maxx, maxy) ring order
        ring = np.stack([bb[:, [2, 1]], bb[:, [2, 3]], bb[:, [0, 3]], bb[:, [0, 1]], bb[:, [2, 1]]], axis=1)
        verts = torch.from_numpy(np.ascontiguousarray(ring.reshape(-1, 2))).to(self.device)
        off = torch.arange(0, 5 * t + 1, 5, dtype=torch.int64, device=self.device)
        filt = torch.from_numpy(np.ascontiguousarray(np.asarray(tile_boxes, dtype=np.float64).reshape(t, 4))).to(
            self.device)
        inter, within = self.predicates(verts, off, filt)
        inter = inter.cpu().numpy().astype(bool)
        within = within.cpu().numpy().astype(bool)
        return inter & within, ~inter


# ----------------------------------------------------------------------------
# fusion
# ----------------------------------------------------------------------------
def fuse_tables(urban, forest_crowns, forest: ForestIndex):
    """urban / forest_crowns: (verts, ring_off, conf) device tensors.  Returns the fused
    (verts, ring_off, conf): forest-model crowns intersecting the forest union first, then
    urban-model crowns that are not within it (helpers.py:804-811)."""
    uv, uo, uc = urban
    fv, fo, fc = forest_crowns
    f_int, _ = forest.predicates(fv, fo)
    _, u_within = forest.predicates(uv, uo)
    sel_f = torch.nonzero(f_int == 1).flatten()
    sel_u = torch.nonzero(u_within == 0).flatten()
    v1, o1 = ops.take_rings(fv, fo, sel_f)
    v2, o2 = ops.take_rings(uv, uo, sel_u)
    verts = torch.cat([v1, v2])
    off = torch.cat([o1, o2[1:] + o1[-1]])
    conf = torch.cat([fc[sel_f], uc[sel_u]])
    return verts, off, conf


def fuse_predictions(urban_fold, forrest_fold, forrest_path, output_dir, logger=None, device=None):
    for p, what in ((urban_fold, "Urban predictions"), (forrest_fold, "Forest predictions")):
        if not os.path.exists(p) or not os.path.isdir(p):
            raise FileNotFoundError(f"{what} path not found: {p}")
    if not os.path.exists(forrest_path) or not os.path.isfile(forrest_path):
        raise FileNotFoundError(f"Forest boundary path not found: {forrest_path}")
    os.makedirs(output_dir, exist_ok=True)
    device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    rec_file = os.path.join(output_dir, "fusion_recovery.yaml")
    completed = set()
    if os.path.exists(rec_file):
        try:
            completed = set((yaml.safe_load(open(rec_file)) or {}).get("completed_files", []))
        except Exception:
            completed = set()
    files = sorted(f for f in os.listdir(urban_fold) if f.endswith(".geojson") or f.endswith(".gpkg"))
    todo = [f for f in files if os.path.splitext(f)[0] not in completed]
    forest = ForestIndex.from_file(forrest_path, device, logger)
    fused = []
    for name in todo:
        up, fp = os.path.join(urban_fold, name), os.path.join(forrest_fold, name)
        if not os.path.exists(fp):
            if logger:
                logger.error(f"Forest GeoJSON for tile {name} at path {fp} not found. Skipping tile.")
            continue
        out_path = os.path.join(output_dir, os.path.basename(name))
        try:
            uv, uo, ucols, epsg = gpkg.read_layer(up)
            fv, fo, fcols, fepsg = gpkg.read_layer(fp)
            layer = os.path.splitext(name)[0]
            uconf = np.array(ucols.get("Confidence_score", []), dtype=np.float64)
            fconf = np.array(fcols.get("Confidence_score", []), dtype=np.float64)
            if len(uo) - 1 == 0:
                gpkg.write_layer(out_path, layer, fv, fo, {"Confidence_score": fconf}, gpkg.STITCHED_SCHEMA,
                                 epsg=fepsg or epsg or 4326)
            elif len(fo) - 1 == 0:
                gpkg.write_layer(out_path, layer, uv, uo, {"Confidence_score": uconf}, gpkg.STITCHED_SCHEMA,
                                 epsg=epsg or 4326)
            else:
                t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(device)
                verts, off, conf = fuse_tables((t(uv), t(uo), t(uconf)), (t(fv), t(fo), t(fconf)), forest)
                gpkg.write_layer(out_path, layer, verts.cpu().numpy(), off.cpu().numpy(),
                                 {"Confidence_score": conf.cpu().numpy()}, gpkg.STITCHED_SCHEMA, epsg=epsg or 4326)
            if logger:
                logger.debug(f"Fused file saved to {out_path}")
            fused.append(os.path.splitext(name)[0])
        except Exception as e:
            if logger:
                logger.error(f"Failed to process tile {name}: {e}")
    try:
        with open(rec_file, "w") as f:
            yaml.safe_dump({"completed_files": sorted(completed | set(fused))}, f, sort_keys=False)
    except Exception:
        pass
