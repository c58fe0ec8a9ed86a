"""Load the REAL reference modules on CPU (TEST INFRASTRUCTURE, container only).

``/root/reference/TreeDetection/*.py`` import cupy, rasterio, fiona, geopandas,
shapely, affine, detectron2, ... unconditionally (postprocessing.py:7-17,
helpers.py:10-28) and none of those are installed here.  This module installs

* a NumPy-backed ``cupy`` (arrays are an ndarray subclass with ``.get()``),
* ``shapely`` / ``shapely.geometry`` backed by :mod:`oracle.geom`,
* ``rasterio.coords.BoundingBox`` and an ``affine.Affine`` stand-in,
* inert stubs for everything else,

into ``sys.modules`` and then executes the reference source files unmodified,
so ``filter_polygons_by_iou_and_area``, ``process_features``,
``get_metadata_within_polygon``, ``ndvi_array_from_rgbi`` ... are the
reference's own code.  ``/root/reference`` does not exist on the GPU box:
nothing that runs there may import this module (tests skip when it is absent).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types
from collections import namedtuple

import numpy as np

REF_ROOT = os.environ.get("TREEDET_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, "TreeDetection", "postprocessing.py"))


# ----------------------------------------------------------------------------
# cupy on NumPy
# ----------------------------------------------------------------------------
class ShimArray(np.ndarray):
    """ndarray with cupy's ``.get()``; reductions stay 0-d arrays so that
    ``x.max().get()`` works as it does on cupy."""

    def get(self):
        return np.asarray(self)

    def _wrap(self, r):
        return np.asarray(r).view(ShimArray)

    def max(self, *a, **k):
        return self._wrap(np.asarray(self).max(*a, **k))

    def min(self, *a, **k):
        return self._wrap(np.asarray(self).min(*a, **k))

    def sum(self, *a, **k):
        return self._wrap(np.asarray(self).sum(*a, **k))

    def mean(self, *a, **k):
        return self._wrap(np.asarray(self).mean(*a, **k))

    def any(self, *a, **k):
        return self._wrap(np.asarray(self).any(*a, **k))


def _w(x):
    return np.asarray(x).view(ShimArray)


def _make_cupy():
    cp = types.ModuleType("cupy")

    def wrapf(f):
        def g(*a, **k):
            return _w(f(*a, **k))
        g.__name__ = getattr(f, "__name__", "f")
        return g

    for name in ("array", "asarray", "zeros", "ones", "full", "empty", "isnan", "sqrt", "argmax", "argmin",
                 "mean", "var", "append", "abs", "maximum", "minimum", "arange", "sum", "any", "nanmean",
                 "stack", "floor", "column_stack", "concatenate", "max", "min"):
        setattr(cp, name, wrapf(getattr(np, name)))

    def where(*a, **k):
        r = np.where(*a, **k)
        return tuple(_w(x) for x in r) if isinstance(r, tuple) else _w(r)

    cp.where = where
    cp.asnumpy = lambda x: np.asarray(x)
    cp.nan = np.nan
    cp.ndarray = ShimArray
    for t in ("float16", "float32", "float64", "int32", "int64", "bool_", "uint8"):
        setattr(cp, t, getattr(np, t))

    class _Device:
        def __init__(self, *a, **k):
            pass

        def __enter__(self):
            return self

        def __exit__(self, *a):
            return False

        def use(self):
            pass

    cuda = types.ModuleType("cupy.cuda")
    cuda.Device = _Device
    runtime = types.ModuleType("cupy.cuda.runtime")
    runtime.CUDARuntimeError = RuntimeError
    cuda.runtime = runtime
    cp.cuda = cuda
    return cp, cuda, runtime


# ----------------------------------------------------------------------------
# affine / rasterio stand-ins
# ----------------------------------------------------------------------------
class Affine(tuple):
    """affine.Affine restated: 9-tuple (a,b,c,d,e,f,0,0,1); ``*`` composes,
    ``~`` inverts, ``scale`` / ``translation`` constructors, ``almost_equals``."""

    def __new__(cls, a, b, c, d, e, f, *rest):
        return tuple.__new__(cls, (float(a), float(b), float(c), float(d), float(e), float(f), 0.0, 0.0, 1.0))

    a = property(lambda s: s[0]); b = property(lambda s: s[1]); c = property(lambda s: s[2])
    d = property(lambda s: s[3]); e = property(lambda s: s[4]); f = property(lambda s: s[5])

    @classmethod
    def scale(cls, *scaling):
        sx = scaling[0]
        sy = scaling[1] if len(scaling) > 1 else sx
        return cls(sx, 0.0, 0.0, 0.0, sy, 0.0)

    @classmethod
    def translation(cls, xoff, yoff):
        return cls(1.0, 0.0, xoff, 0.0, 1.0, yoff)

    def __mul__(self, other):
        sa, sb, sc, sd, se, sf = self[:6]
        if isinstance(other, Affine):
            oa, ob, oc, od, oe, of = other[:6]
            return Affine(sa * oa + sb * od, sa * ob + sb * oe, sa * oc + sb * of + sc,
                          sd * oa + se * od, sd * ob + se * oe, sd * oc + se * of + sf)
        vx, vy = other
        return (vx * sa + vy * sb + sc, vx * sd + vy * se + sf)

    def __invert__(self):
        sa, sb, sc, sd, se, sf = self[:6]
        idet = 1.0 / (sa * se - sb * sd)
        ra = se * idet; rb = -sb * idet; rd = -sd * idet; re = sa * idet
        return Affine(ra, rb, -sc * ra - sf * rb, rd, re, -sc * rd - sf * re)

    def almost_equals(self, other, precision=1e-5):
        return all(abs(x - y) < precision for x, y in zip(self, other))


BoundingBox = namedtuple("BoundingBox", "left bottom right top")


class _Stub(types.ModuleType):
    """Inert module: any attribute is another stub / a no-op callable class."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        v = type(name, (), {"__init__": lambda self, *a, **k: None, "__call__": lambda self, *a, **k: None})
        setattr(self, name, v)
        return v


_loaded = None


def load():
    """Return a namespace with the reference modules ``utilities``, ``helpers``,
    ``postprocessing``, ``config`` loaded from REF_ROOT (cached)."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference not present under {REF_ROOT}")
    from . import geom

    cp, cuda, runtime = _make_cupy()
    mods = {"cupy": cp, "cupy.cuda": cuda, "cupy.cuda.runtime": runtime}

    shp = types.ModuleType("shapely")
    shp_geom = types.ModuleType("shapely.geometry")
    shp_err = types.ModuleType("shapely.errors")
    for m in (shp, shp_geom):
        m.Polygon = geom.Polygon; m.MultiPolygon = geom.MultiPolygon
        m.shape = geom.shape; m.box = geom.box
    shp_err.ShapelyError = Exception
    shp.geometry = shp_geom; shp.errors = shp_err
    mods.update({"shapely": shp, "shapely.geometry": shp_geom, "shapely.errors": shp_err})

    aff = types.ModuleType("affine"); aff.Affine = Affine
    mods["affine"] = aff

    rio = _Stub("rasterio")
    rio_coords = types.ModuleType("rasterio.coords"); rio_coords.BoundingBox = BoundingBox
    rio.coords = rio_coords
    mods["rasterio"] = rio
    mods["rasterio.coords"] = rio_coords
    for sub in ("enums", "mask", "merge", "transform", "windows", "crs"):
        s = _Stub(f"rasterio.{sub}"); setattr(rio, sub, s); mods[f"rasterio.{sub}"] = s

    for name in ("fiona", "fiona.model", "geopandas", "aiofiles", "detectron2", "detectron2.engine",
                 "detectron2.config", "detectron2.model_zoo", "pycocotools", "pycocotools.mask",
                 "matplotlib", "matplotlib.pyplot", "matplotlib.colors"):
        mods[name] = _Stub(name)
    mods["fiona"].model = mods["fiona.model"]
    mods["detectron2"].model_zoo = mods["detectron2.model_zoo"]
    mods["pycocotools"].mask = mods["pycocotools.mask"]
    mods["matplotlib"].pyplot = mods["matplotlib.pyplot"]
    mods["matplotlib"].colors = mods["matplotlib.colors"]

    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        pkg = types.ModuleType("TreeDetection")
        pkg.__path__ = [os.path.join(REF_ROOT, "TreeDetection")]
        sys.modules["TreeDetection"] = pkg
        ns = types.SimpleNamespace(cupy=cp, Affine=Affine, BoundingBox=BoundingBox, geom=geom)
        for name in ("config", "utilities", "recoveries", "helpers", "postprocessing"):
            spec = importlib.util.spec_from_file_location(
                f"TreeDetection.{name}", os.path.join(REF_ROOT, "TreeDetection", f"{name}.py"))
            m = importlib.util.module_from_spec(spec)
            sys.modules[f"TreeDetection.{name}"] = m
            spec.loader.exec_module(m)
            setattr(pkg, name, m)
            setattr(ns, name, m)
    finally:
        # leave the TreeDetection.* modules importable but restore the world
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    _loaded = ns
    return ns


class _NullLogger:
    def debug(self, *a, **k): pass
    info = warning = warn = error = debug


def set_config(ns, **kw):
    """Populate the reference's Config singleton (config.py:12-23) the way
    ``get_config`` would, without touching the filesystem."""
    cfg = dict(tile_width=50, tile_height=50, buffer=20, use_overlap=True,
               overlapping_tiles_width=3, overlapping_tiles_height=3,
               confidence_threshold=0.3, containment_threshold=0.75, height_threshold=3,
               ndvi_mean_threshold=0.1, ndvi_var_threshold=0.1, iou_threshold=0.6,
               area_threshold=1, ndvi_scaling_factor=0.2, height_scaling_factor=1.0,
               logger=_NullLogger())
    cfg.update(kw)
    ns.config.Config()._load_into_config(cfg)
    return cfg
