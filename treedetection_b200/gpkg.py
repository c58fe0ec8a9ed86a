"""GeoPackage (SQLite + GeoPackageBinary/WKB) codec for the vector artefacts of the path.

The reference writes ``geojson_predictions/<image>.gpkg`` through geopandas
(``TreeDetection/helpers.py:545-548``: columns ``Confidence_score`` + geometry) and
``processed_<image>.gpkg`` through fiona (``TreeDetection/postprocessing.py:904-939``:
``Confidence_score, poly_id, Area, TreeHeight, Centroid, Diameter, is_contained,
num_contained``).  fiona / geopandas / GDAL are absent here; this module writes the same
layers with the standard library so that QGIS / geopandas users see a drop-in
(GeoPackage 1.2: application_id 'GPKG', user_version 10200, gpkg_spatial_ref_sys,
gpkg_contents, gpkg_geometry_columns, one feature table, single-ring Polygon geometries).
"""
from __future__ import annotations

import os
import sqlite3
import struct
from datetime import datetime, timezone

import numpy as np

_WKT = {
    4326: 'GEOGCS["WGS 84",DATUM["WGS_1984",SPHEROID["WGS 84",6378137,298.257223563]],PRIMEM["Greenwich",0],'
          'UNIT["degree",0.0174532925199433],AUTHORITY["EPSG","4326"]]',
    25832: 'PROJCS["ETRS89 / UTM zone 32N",GEOGCS["ETRS89",DATUM["European_Terrestrial_Reference_System_1989",'
           'SPHEROID["GRS 1980",6378137,298.257222101]],PRIMEM["Greenwich",0],UNIT["degree",0.0174532925199433]],'
           'PROJECTION["Transverse_Mercator"],PARAMETER["latitude_of_origin",0],PARAMETER["central_meridian",9],'
           'PARAMETER["scale_factor",0.9996],PARAMETER["false_easting",500000],PARAMETER["false_northing",0],'
           'UNIT["metre",1],AXIS["Easting",EAST],AXIS["Northing",NORTH],AUTHORITY["EPSG","25832"]]',
}

_SQL_TYPES = {"float": "REAL", "int": "MEDIUMINT", "str": "TEXT"}


def _gpb_polygon(ring: np.ndarray, srs_id: int) -> bytes:
    """GeoPackageBinary header (little endian, xy envelope) + WKB Polygon with one ring."""
    n = ring.shape[0]
    if n == 0:
        return b"GP\x00" + bytes([0x11]) + struct.pack("<i", srs_id) + struct.pack("<BII", 1, 3, 0)
    minx, miny = ring.min(axis=0)
    maxx, maxy = ring.max(axis=0)
    head = b"GP\x00" + bytes([0x03]) + struct.pack("<i", srs_id) + struct.pack("<4d", minx, maxx, miny, maxy)
    return head + struct.pack("<BIII", 1, 3, 1, n) + np.ascontiguousarray(ring, dtype="<f8").tobytes()


def _parse_gpb(blob: bytes) -> np.ndarray:
    flags = blob[3]
    env = {0: 0, 1: 32, 2: 48, 3: 48, 4: 64}[(flags >> 1) & 7]
    wkb = blob[8 + env:]
    bo = "<" if wkb[0] == 1 else ">"
    gtype, nrings = struct.unpack(bo + "II", wkb[1:9])
    if gtype % 1000 == 6:      # MultiPolygon: first polygon (postprocessing.py:493-494)
        wkb = wkb[9:]
        bo = "<" if wkb[0] == 1 else ">"
        gtype, nrings = struct.unpack(bo + "II", wkb[1:9])
    if gtype % 1000 != 3 or nrings == 0:
        return np.zeros((0, 2))
    npts = struct.unpack(bo + "I", wkb[9:13])[0]
    dims = 2 + {0: 0, 1: 1, 2: 1, 3: 2}.get(gtype // 1000, 0)      # ISO WKB: +1000 Z, +2000 M, +3000 ZM
    pts = np.frombuffer(wkb, dtype=bo + "f8", count=npts * dims, offset=13).reshape(npts, dims)
    return np.array(pts[:, :2], dtype=np.float64)


def _append_native(path, layer, epsg, verts, ring_off, names, schema, columns):
    """The feature rows through ``td_gpkg_append`` (csrc/gpkgio.cu; the GIL is released for the whole call).
    Returns False when the library or a column cannot take that route (the Python loop then writes the rows)."""
    import ctypes as C
    try:
        from . import _lib
        fn = _lib.lib().td_gpkg_append
    except Exception:
        return False
    n = len(ring_off) - 1
    keep, types, data, offs = [], [], [], []
    for k in names:
        c = columns[k]
        kind = schema[k]
        if kind == "float":
            if not isinstance(c, np.ndarray) and any(v is None for v in c):
                c = [np.nan if v is None else v for v in c]
            a = np.ascontiguousarray(c, dtype=np.float64)
            o = None
        elif kind == "int":
            if not isinstance(c, np.ndarray) and any(v is None for v in c):
                return False
            a = np.ascontiguousarray(c, dtype=np.int64)
            o = None
        else:
            if any(v is None for v in c):
                return False
            enc = [str(v).encode("utf-8") for v in c]
            a = np.frombuffer(b"".join(enc) or b"\0", dtype=np.uint8)
            o = np.zeros(n + 1, dtype=np.int64)
            np.cumsum([len(e) for e in enc], out=o[1:])
        if o is None and a.shape != (n,):
            return False
        keep.append((a, o))
        types.append({"float": 0, "int": 1, "str": 2}[kind])
        data.append(a.ctypes.data)
        offs.append(o.ctypes.data if o is not None else None)
    m = len(names)
    c_names = (C.c_char_p * max(m, 1))(*[k.encode("utf-8") for k in names])
    c_types = (C.c_int * max(m, 1))(*types)
    c_data = (C.c_void_p * max(m, 1))(*data)
    c_offs = (C.c_void_p * max(m, 1))(*offs)
    v = np.ascontiguousarray(verts, dtype=np.float64)
    ro = np.ascontiguousarray(ring_off, dtype=np.int64)
    rc = fn(os.fsencode(path), layer.encode("utf-8"), int(epsg), v.ctypes.data if v.size else None, ro.ctypes.data, n, m,
            C.cast(c_names, C.c_void_p), C.cast(c_types, C.c_void_p), C.cast(c_data, C.c_void_p),
            C.cast(c_offs, C.c_void_p))
    if rc != 0:
        msg = _lib.lib().td_last_error()
        raise RuntimeError(f"td_gpkg_append failed with code {rc}: {msg.decode() if msg else ''}")
    return True


def write_layer(path, layer, verts, ring_off, columns: dict, schema: dict, epsg=25832, native=None):
    """verts (V,2) f64 + ring_off (R+1): one Polygon per ring.  columns: name -> sequence of
    R values; schema: name -> 'float' | 'int' | 'str' (fiona's names), in column order.
    ``native``: None = the feature rows go through ``td_gpkg_append`` when the library is there (same bytes in
    the table), False = the Python loop, True = the native writer or an error."""
    verts = np.asarray(verts, dtype=np.float64).reshape(-1, 2)
    ring_off = np.asarray(ring_off, dtype=np.int64)
    n = len(ring_off) - 1
    if os.path.exists(path):
        os.remove(path)
    con = sqlite3.connect(path)
    try:
        cur = con.cursor()
        cur.execute("PRAGMA application_id = 1196444487")      # 'GPKG'
        cur.execute("PRAGMA user_version = 10200")
        cur.execute("CREATE TABLE gpkg_spatial_ref_sys (srs_name TEXT NOT NULL, srs_id INTEGER NOT NULL PRIMARY KEY, "
                    "organization TEXT NOT NULL, organization_coordsys_id INTEGER NOT NULL, definition TEXT NOT NULL, "
                    "description TEXT)")
        rows = [("Undefined cartesian SRS", -1, "NONE", -1, "undefined", "undefined cartesian coordinate reference system"),
                ("Undefined geographic SRS", 0, "NONE", 0, "undefined", "undefined geographic coordinate reference system"),
                ("WGS 84 geodetic", 4326, "EPSG", 4326, _WKT[4326], "longitude/latitude coordinates in decimal degrees")]
        if epsg not in (-1, 0, 4326):
            rows.append((f"EPSG:{epsg}", epsg, "EPSG", epsg, _WKT.get(epsg, "undefined"), None))
        cur.executemany("INSERT INTO gpkg_spatial_ref_sys VALUES (?,?,?,?,?,?)", rows)
        cur.execute("CREATE TABLE gpkg_contents (table_name TEXT NOT NULL PRIMARY KEY, data_type TEXT NOT NULL, "
                    "identifier TEXT UNIQUE, description TEXT DEFAULT '', last_change DATETIME NOT NULL, min_x DOUBLE, "
                    "min_y DOUBLE, max_x DOUBLE, max_y DOUBLE, srs_id INTEGER)")
        cur.execute("CREATE TABLE gpkg_geometry_columns (table_name TEXT NOT NULL, column_name TEXT NOT NULL, "
                    "geometry_type_name TEXT NOT NULL, srs_id INTEGER NOT NULL, z TINYINT NOT NULL, m TINYINT NOT NULL, "
                    "CONSTRAINT pk_geom_cols PRIMARY KEY (table_name, column_name))")
        cols_sql = ", ".join(f'"{k}" {_SQL_TYPES[v]}' for k, v in schema.items())
        cur.execute(f'CREATE TABLE "{layer}" (fid INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, geom POLYGON'
                    + (", " + cols_sql if cols_sql else "") + ")")
        xs, ys = (np.ascontiguousarray(verts[:, 0]), np.ascontiguousarray(verts[:, 1]))   # unit-stride reductions
        if len(verts):
            ext = (float(xs.min()), float(ys.min()), float(xs.max()), float(ys.max()))
        else:
            ext = (None, None, None, None)
        now = datetime.now(timezone.utc).strftime("%Y-%m-%dT%H:%M:%S.000Z")
        cur.execute("INSERT INTO gpkg_contents VALUES (?,?,?,?,?,?,?,?,?,?)",
                    (layer, "features", layer, "", now, *ext, epsg))
        cur.execute("INSERT INTO gpkg_geometry_columns VALUES (?,?,?,?,?,?)", (layer, "geom", "POLYGON", epsg, 0, 0))
        names = list(schema.keys())
        if native is not False and n and _append_native is not None:
            con.commit()
            con.close()
            con = None
            if _append_native(path, layer, epsg, verts, ring_off, names, schema, columns):
                return
            if native is True:
                raise RuntimeError("gpkg.write_layer: the native feature writer (td_gpkg_append) is not available")
            con = sqlite3.connect(path)
            cur = con.cursor()
        conv = {"float": float, "int": int, "str": str}
        # geometry blobs: header | envelope | WKB polygon header | coordinates, cut from ONE byte string of all
        # coordinates and vectorised envelopes (a layer has tens of thousands of crowns)
        coords = np.ascontiguousarray(verts, dtype="<f8").tobytes()
        lens = np.diff(ring_off)
        blobs = []
        if n:
            safe = np.minimum(ring_off[:-1], max(len(verts) - 1, 0))
            if len(verts):
                env = np.stack([np.minimum.reduceat(xs, safe), np.maximum.reduceat(xs, safe),
                                np.minimum.reduceat(ys, safe), np.maximum.reduceat(ys, safe)], 1)
            else:
                env = np.zeros((n, 4))
            env_b = np.ascontiguousarray(env, dtype="<f8").tobytes()
            head = b"GP\x00" + bytes([0x03]) + struct.pack("<i", epsg)
            empty = _gpb_polygon(np.zeros((0, 2)), epsg)
            wkb_head = {}
            for i in range(n):
                k = int(lens[i])
                if k == 0:
                    blobs.append(empty)
                    continue
                wh = wkb_head.get(k)
                if wh is None:
                    wh = wkb_head[k] = struct.pack("<BIII", 1, 3, 1, k)
                o = int(ring_off[i])
                blobs.append(b"".join((head, env_b[32 * i:32 * i + 32], wh, coords[16 * o:16 * (o + k)])))
        cols = []
        for k in names:
            c, f = columns[k], conv[schema[k]]
            c = c.tolist() if isinstance(c, np.ndarray) else list(c)
            cols.append([None if v is None else f(v) for v in c])
        data = list(zip(blobs, *cols)) if names else [(b,) for b in blobs]
        ph = ",".join("?" * (1 + len(names)))
        colnames = ", ".join(["geom"] + [f'"{k}"' for k in names])
        cur.executemany(f'INSERT INTO "{layer}" ({colnames}) VALUES ({ph})', data)
        con.commit()
    finally:
        if con is not None:
            con.close()


def read_layer(path, layer=None):
    """Returns (verts (V,2) f64, ring_off (R+1) i64, columns dict name -> list, epsg)."""
    con = sqlite3.connect(path)
    try:
        cur = con.cursor()
        if layer is None:
            row = cur.execute("SELECT table_name FROM gpkg_contents WHERE data_type='features'").fetchone()
            if row is None:
                return np.zeros((0, 2)), np.zeros(1, dtype=np.int64), {}, None
            layer = row[0]
        gcol, epsg = cur.execute("SELECT column_name, srs_id FROM gpkg_geometry_columns WHERE table_name=?",
                                 (layer,)).fetchone()
        info = cur.execute(f'PRAGMA table_info("{layer}")').fetchall()
        pk = [r[1] for r in info if r[5]]
        names = [r[1] for r in info if r[1] != gcol and r[1] not in pk]
        sel = ", ".join([f'"{gcol}"'] + [f'"{k}"' for k in names])
        order = f' ORDER BY "{pk[0]}"' if pk else ""
        rings, cols = [], {k: [] for k in names}
        for rec in cur.execute(f'SELECT {sel} FROM "{layer}"{order}'):
            rings.append(_parse_gpb(rec[0]) if rec[0] is not None else np.zeros((0, 2)))
            for k, v in zip(names, rec[1:]):
                cols[k].append(v)
    finally:
        con.close()
    off = np.zeros(len(rings) + 1, dtype=np.int64)
    if rings:
        off[1:] = np.cumsum([len(r) for r in rings])
    verts = np.concatenate(rings) if rings and off[-1] > 0 else np.zeros((0, 2))
    return verts, off, cols, epsg


PROCESSED_SCHEMA = {"Confidence_score": "float", "poly_id": "str", "Area": "float", "TreeHeight": "float",
                    "Centroid": "str", "Diameter": "float", "is_contained": "str", "num_contained": "int"}
STITCHED_SCHEMA = {"Confidence_score": "float"}
