"""P10 -- two-model fusion against a forest outline (config 3).

Reference: ``fuse_predictions`` (TreeDetection/helpers.py:703-834), the tile flags
``only_forest`` / ``only_urban`` (TreeDetection/preprocessing.py:67-96) and the per-model
tile exclusion (TreeDetection/prediction.py:79-93, implemented in ``predictor.py``).

The GEOS predicates (``intersects`` / ``within`` / ``contains`` against ``unary_union`` of the
forest polygons) run on the device (``td_forest_predicates``, csrc/forest_core.cuh) without
building the union.  Forest polygons keep their interior rings (holes).  The reference repairs
INVALID geometries only (``geom.buffer(0) if not geom.is_valid else geom``, helpers.py:743-750 for
the outline, :816-821 for the fused crowns): valid geometries pass through untouched, so vertex
order and count of the fused layer are those of the inputs.  GEOS itself is absent here and its
repair is not reproduced: invalid rings (a mask outline with a one-pixel spur or a diagonal pinch touches
itself) are COUNTED on the device (``td_ring_is_simple``) and reported with a warning; they are written
unchanged (about 2 % of the crowns of the synthetic two-model scene, tests/test_gpu_two_model.py).  The
single-model path of the reference never repairs geometries.
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


def _point_in_ring(p, r):
    """even-odd rule (host helper that assigns shapefile holes to their shells)"""
    x, y = p
    x0, y0, x1, y1 = r[:-1, 0], r[:-1, 1], r[1:, 0], r[1:, 1]
    cond = (y0 > y) != (y1 > y)
    with np.errstate(divide="ignore", invalid="ignore"):
        xi = x0 + (y - y0) * (x1 - x0) / (y1 - y0)
    return bool(np.count_nonzero(cond & (x < xi)) & 1)


def read_shapefile_polygons(path):
    """ESRI shapefile, shape types 5 / 15 / 25 (Polygon[Z/M]): list of polygons, each a list of rings
    (N,2) f64 -- the shell first, then its holes.  Outer rings are clockwise in a shapefile, holes
    counter-clockwise; a hole belongs to the shell of its record that contains its first vertex."""
    polys = []
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
        shells, holes = [], []
        for a, b in zip(parts[:-1], parts[1:]):
            r = np.array(pts[a:b], dtype=np.float64)
            if len(r) < 4:
                continue
            (shells if _ring_signed_area(r) <= 0 else holes).append(r)
        rec_polys = [[s_] for s_ in shells]
        for h in holes:
            for poly in rec_polys:
                if _point_in_ring(h[0], poly[0]):
                    poly.append(h)
                    break
        polys += rec_polys
    return polys


def read_outline(path, logger=None):
    """forest outline as a list of polygons (each: [shell, hole, ...])"""
    if path.lower().endswith(".shp"):
        return read_shapefile_polygons(path)
    verts, off, _, _ = gpkg.read_layer(path)        # GeoPackage outlines: exterior rings
    return [[np.array(verts[off[i]:off[i + 1]])] for i in range(len(off) - 1) if off[i + 1] - off[i] >= 4]


def ring_is_simple(ring):
    """GEOS validity of a crown ring as a polygon shell: no two non-adjacent segments share a point and adjacent
    ones only their common vertex (host, O(n^2) on tens of vertices; used to COUNT the geometries the
    reference would hand to buffer(0))."""
    r = np.asarray(ring, dtype=np.float64)
    if len(r) < 4 or not np.array_equal(r[0], r[-1]):
        return False
    p, q = r[:-1], r[1:]
    n = len(p)

    def orient(a, b, c):
        return np.sign((b[..., 0] - a[..., 0]) * (c[..., 1] - a[..., 1]) - (b[..., 1] - a[..., 1]) * (c[..., 0] - a[..., 0]))
    for i in range(n):
        js = np.arange(i + 1, n)
        if len(js) == 0:
            break
        a, b = p[i], q[i]
        c, d = p[js], q[js]
        o1, o2 = orient(a, b, c), orient(a, b, d)
        o3, o4 = orient(c, d, a[None]), orient(c, d, b[None])
        env = (np.minimum(c[:, 0], d[:, 0]) <= max(a[0], b[0])) & (np.maximum(c[:, 0], d[:, 0]) >= min(a[0], b[0])) & \
              (np.minimum(c[:, 1], d[:, 1]) <= max(a[1], b[1])) & (np.maximum(c[:, 1], d[:, 1]) >= min(a[1], b[1]))
        touch = env & (o1 * o2 <= 0) & (o3 * o4 <= 0)
        adj = (js == i + 1) | ((i == 0) & (js == n - 1))
        # adjacent segments share one vertex; they are only a problem when they fold back onto each other
        fold = adj & (o1 == 0) & (o2 == 0) & (np.einsum("ij,j->i", d - c, b - a) < 0)
        if np.any(touch & ~adj) or np.any(fold):
            return False
    return True


def _ragged(rings, device):
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    if rings:
        off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.concatenate(rings).astype(np.float64) if rings else np.zeros((0, 2))
    return torch.from_numpy(np.ascontiguousarray(verts)).to(device), torch.from_numpy(off).to(device)


class ForestIndex:
    """Forest outline on the device + the two queries the path makes against it.  ``polygons``: list of
    polygons, each a list of rings (shell first, then holes); a bare (N,2) array is a polygon without holes."""

    def __init__(self, polygons, device):
        polygons = [[p] if isinstance(p, np.ndarray) else list(p) for p in polygons]
        self.polygons = polygons
        self.device = device
        rings = [r for p in polygons for r in p]
        self.verts, self.off = _ragged(rings, device)
        poly_off = np.zeros(len(polygons) + 1, dtype=np.int64)
        poly_off[1:] = np.cumsum([len(p) for p in polygons])
        self.poly_off = torch.from_numpy(poly_off).to(device)
        self.bounds = None
        if rings:
            rb = ops.simplify_rings(self.verts, self.off, 0.0, want_bounds=True)["bounds"]
            self.bounds = rb[self.poly_off[:-1]].contiguous()

    @classmethod
    def from_file(cls, path, device=None, logger=None):
        device = device if device is not None else torch.device("cuda", torch.cuda.current_device())
        polygons = read_outline(path, logger)
        if not polygons:
            raise ValueError(f"No valid geometries found in the forest shapefile {path}.")
        return cls(polygons, device)

    def predicates(self, verts, ring_off, a_filter=None):
        """(intersects, within) uint8 tensors for the query rings."""
        return ops.forest_predicates(verts, ring_off, self.verts, self.off, self.bounds, a_filter, self.poly_off)

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
                # helpers.py:816-821 repairs INVALID geometries only; GEOS is absent, so they are counted, not repaired
                n_bad = int((ops.rings_are_simple(verts.contiguous(), off.contiguous()) == 0).sum().item()) if len(off) > 1 else 0
                if n_bad and logger:
                    logger.warning(f"{n_bad} fused crowns of {name} are not valid polygons; the reference would repair "
                                   f"them with buffer(0) / make_valid, this build writes them unchanged.")
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
