"""Host logic on the CPU against golden values produced by the REFERENCE'S OWN functions
(tests/golden/make_golden_host.py through oracle.refshim): get_config keys / defaults / assertion
messages, the tile-id parser behind the stitch filter boxes, coordinate rounding; plus the tile grid
(preprocessing.py:57-120 restated in tiling.tile_grid) against its closed form, ResizeShortestEdge, the
GPKG and GeoTIFF codecs' round trips and the capacity planning of the sync-free chain."""
import json
import os

import numpy as np
import pytest
import yaml

from treedetection_b200 import config as tconfig, geo, geotiff, gpkg, pipeline, synth, tiling

G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "host_logic.json")))
SKIP = {"logger", "image_directory", "height_data_path", "combined_model", "urban_model", "forrest_model",
        "forrest_outline", "output_directory", "tiles_path", "continue"}


def _get_config(tmp_path, extra, drop=()):
    img, h, model = (tmp_path / n for n in ("rgb", "ndsm", "model"))
    for d in (img, h, model):
        d.mkdir(exist_ok=True)
    cfg = {"image_directory": str(img), "height_data_path": str(h), "combined_model": str(model),
           "output_directory": str(tmp_path / "out"), "tiles_path": str(tmp_path / "tiles")}
    cfg.update(extra)
    for k in drop:
        cfg.pop(k, None)
    path = tmp_path / "c.yml"
    path.write_text(yaml.safe_dump(cfg))
    return tconfig.get_config(str(path))


@pytest.mark.parametrize("name,extra", [("config_minimal", {}),
                                        ("config_overrides", {"tile_width": 64, "buffer": 8, "iou_threshold": 0.6,
                                                              "exclude_files": ["a.gpkg"], "device": "cuda:1",
                                                              "ndvi_scaling_factor": 0.2, "debug": True})])
def test_get_config_matches_reference(tmp_path, name, extra):
    import torch
    if torch.cuda.is_available():
        pytest.skip("the golden was produced without CUDA (device falls back to 'cpu')")
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        cfg, obj = _get_config(tmp_path, extra)
    got = {k: v for k, v in cfg.items() if k not in SKIP}
    assert got == G[name]
    assert obj.tile_width == cfg["tile_width"] and tconfig.Config().buffer == cfg["buffer"]     # singleton
    assert cfg["continue"] == os.path.join(cfg["output_directory"], "continue.yml")
    assert os.path.isdir(os.path.join(cfg["output_directory"], "logs"))


@pytest.mark.parametrize("name,drop", [("config_no_images", "image_directory"), ("config_no_height", "height_data_path"),
                                       ("config_no_model", "combined_model")])
def test_get_config_assertions_match_reference(tmp_path, name, drop):
    with pytest.raises(AssertionError) as e:
        _get_config(tmp_path, {}, drop=(drop,))
    assert str(e.value) == G[name]["assertion"]


def test_tile_id_parser_matches_filename_geoinfo():
    """pipeline.tile_tables parses the last five '_' fields of a tile id, as helpers.filename_geoinfo does"""
    for tid, want in G["filename_geoinfo"].items():
        parts = [int(p) for p in tid.split("_")[-5:]]
        assert parts == want


def test_round_coordinates_rule():
    """round(coord * 1000) / 1000 with Python's round-half-even (utilities.py:146-161): the rule
    td_round_coords implements with rint; checked here on the host against the reference's outputs"""
    vals = np.array(G["round_coordinates"]["in"])
    want = np.array(G["round_coordinates"]["out"])
    got = np.stack([np.rint(vals * 1000.0) / 1000.0, np.rint(-vals * 1000.0) / 1000.0], 1)
    np.testing.assert_array_equal(got, want)


def test_tile_grid_closed_form():
    px, W, H = 0.2, 1500, 1000
    tf = synth.image_transform(412000.0, 5318200.0, px)          # top-left origin
    tiles = tiling.tile_grid("img", tf, W, H, 25832, 50, 50, 20)
    ids = list(tiles)
    # x outer, y inner, from the BOTTOM-left corner (preprocessing.py:57-58)
    assert ids[0] == "img_412000_5318000_50_20_25832" and ids[1] == "img_412000_5318050_50_20_25832"
    assert len(ids) == 6 * 4
    for tid, m in tiles.items():
        minx, miny = [int(v) for v in tid.split("_")[1:3]]
        assert m["bounds"] == [minx - 20, miny - 20, minx + 70, miny + 70]
        c0, r0, w, h = m["window"]
        # geometry_window: floor / ceil of the pixel bounds, clipped to the raster
        assert c0 == max(int(np.floor((minx - 20 - 412000.0) / px)), 0)
        assert c0 + w == min(int(np.ceil((minx + 70 - 412000.0) / px)), W)
        assert r0 == max(int(np.floor((5318200.0 - (miny + 70)) / px)), 0)
        assert r0 + h == min(int(np.ceil((5318200.0 - (miny - 20)) / px)), H)
        a, b, c, d, e, f = m["transform"][:6]
        assert (a, b, d, e) == (px, 0.0, 0.0, -px) and c == 412000.0 + c0 * px and f == 5318200.0 - r0 * px
        assert m["transform"][6:] == [0.0, 0.0, 1.0] and m["crs"] == 25832


@pytest.mark.parametrize("hw,want", [((450, 450), (800, 800)), ((450, 350), (1029, 800)), ((350, 450), (800, 1029)),
                                     ((100, 400), (333, 1333)), ((1000, 1000), (800, 800))])
def test_resize_shortest_edge(hw, want):
    """detectron2 ResizeShortestEdge(800, max 1333): scale the short side to 800, cap the long side,
    round half up (int(x + 0.5))"""
    assert tiling.resize_shortest_edge(*hw) == want


def test_gpkg_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    rings = [np.concatenate([r, r[:1]]) for r in (412000 + rng.uniform(0, 100, (k, 2)) for k in (4, 9, 33))]
    verts = np.concatenate(rings)
    off = np.zeros(4, dtype=np.int64); off[1:] = np.cumsum([len(r) for r in rings])
    cols = {"Confidence_score": np.array([0.3, 0.75, 0.999])}
    path = str(tmp_path / "x.gpkg")
    gpkg.write_layer(path, "x", verts, off, cols, gpkg.STITCHED_SCHEMA, epsg=25832)
    v2, o2, c2 = gpkg.read_layer(path)[:3]
    np.testing.assert_array_equal(v2, verts)
    np.testing.assert_array_equal(o2, off)
    np.testing.assert_array_equal(np.asarray(c2["Confidence_score"], dtype=np.float64), cols["Confidence_score"])


def test_gpkg_native_writer_writes_the_same_rows(tmp_path):
    """td_gpkg_append (csrc/gpkgio.cu) against the Python row loop: every column of every row of the processed
    schema, an empty ring, a NaN (NULL) and non-ASCII text"""
    import sqlite3
    rng = np.random.default_rng(3)
    n = 500
    lens = rng.integers(4, 40, n)
    lens[7] = 0
    off = np.zeros(n + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    verts = 412000 + rng.uniform(0, 1000, (int(off[-1]), 2))
    conf = rng.random(n)
    conf[3] = np.nan
    cols = {"Confidence_score": conf, "poly_id": [str(i) for i in range(n)], "Area": rng.random(n),
            "TreeHeight": rng.random(n).astype(np.float32), "Centroid": [f'{{"x": {i}.5, "y": "\u00e4"}}' for i in range(n)],
            "Diameter": [float(i) for i in range(n)], "is_contained": [str(bool(i & 1)) for i in range(n)],
            "num_contained": rng.integers(0, 9, n).astype(np.int32)}
    a, b = str(tmp_path / "a.gpkg"), str(tmp_path / "b.gpkg")
    gpkg.write_layer(a, "x", verts, off, cols, gpkg.PROCESSED_SCHEMA, epsg=25832, native=True)
    gpkg.write_layer(b, "x", verts, off, cols, gpkg.PROCESSED_SCHEMA, epsg=25832, native=False)

    def rows(p):
        con = sqlite3.connect(p)
        try:
            return (con.execute('SELECT * FROM "x" ORDER BY fid').fetchall(),
                    con.execute("SELECT table_name, min_x, min_y, max_x, max_y, srs_id FROM gpkg_contents").fetchall(),
                    con.execute("SELECT * FROM gpkg_geometry_columns").fetchall())
        finally:
            con.close()
    ra, rb = rows(a), rows(b)
    assert len(ra[0]) == n and ra == rb
    assert ra[0][3][2] is None            # NaN -> NULL on both routes
    v2, o2, c2 = gpkg.read_layer(a)[:3]
    np.testing.assert_array_equal(v2, verts)
    np.testing.assert_array_equal(o2, off)
    assert c2["Centroid"] == cols["Centroid"]


def test_geotiff_round_trip_with_geo_tags(tmp_path):
    rng = np.random.default_rng(1)
    arr = rng.integers(0, 255, (4, 37, 53), dtype=np.uint8)
    tf = (0.2, 0.0, 412000.0, 0.0, -0.2, 5318000.0)
    path = str(tmp_path / "a.tif")
    geotiff.write(path, arr, tf, epsg=25832)
    got, info = geotiff.read(path)
    np.testing.assert_array_equal(got, arr)
    assert info.transform == tf and info.epsg == 25832 and (info.width, info.height, info.count) == (53, 37, 4)
    win, winfo = geotiff.read(path, window=(5, 7, 20, 11))
    np.testing.assert_array_equal(win, arr[:, 7:18, 5:25])
    assert winfo.transform == geo.window_transform(tf, 5, 7)
    f32 = rng.normal(size=(29, 31)).astype(np.float32)
    geotiff.write(str(tmp_path / "h.tif"), f32, (1.0, 0.0, 412000.0, 0.0, -1.0, 5318000.0), epsg=25832,
                  nodata=-3.4028234663852886e38)
    got, info = geotiff.read(str(tmp_path / "h.tif"))
    np.testing.assert_array_equal(got[0], f32)
    assert info.nodata == pytest.approx(-3.4028234663852886e38)


def test_chain_runner_capacity_planning():
    """capacities grow, never shrink; a grown capacity drops the chain object so that it is rebuilt"""
    run = pipeline.ChainRunner(pipeline.PipelineParams())
    run._learn({"inst": 900, "words": 1000, "px": 50000, "ptslots": 9000, "rings": 35, "verts": 600})
    first = dict(run.caps)
    assert first["words"] == int(1000 * 1.25) + 1024 and first["inst"] == int(900 * 1.25) + 1024
    run.chain = "sentinel"
    run._learn({"words": 10, "px": 10, "ptslots": 10, "rings": 3, "verts": 5})
    assert run.caps == first and run.chain == "sentinel"
    run._learn({"words": 5000, "px": 10, "ptslots": 10, "rings": 300, "verts": 5})
    assert run.caps["words"] == int(5000 * 1.25) + 1024 and run.caps["px"] == first["px"]
    assert run.caps["rings"] == int(300 * 1.25) + 1024 and run.chain is None


def test_prediction_and_stitching_ledgers_follow_the_reference(tmp_path):
    """recoveries.py:5-144: same file formats (``files: {image: [tile ids]}``, ``completed_files``), the tile-count
    validation incl. excluded tiles, and the model-path check"""
    import logging
    from treedetection_b200 import detection
    log = logging.getLogger("ledger-test")
    out, tiles_dir, stitched = tmp_path / "predictions", tmp_path / "tiles", tmp_path / "geojson_predictions"
    for d in (out, tiles_dir, stitched):
        d.mkdir()
    tiles = {f"img1_{k}_0_50_20_25832": {"only_forest": k == 2, "only_urban": False} for k in range(4)}
    for stem in ("img1", "img2", "img3"):
        (tiles_dir / f"{stem}.json").write_text(json.dumps(tiles))
        (stitched / f"{stem}.gpkg").write_bytes(b"")
    files = [str(tmp_path / "rgb" / f"{s}.tif") for s in ("img1", "img2", "img3")]
    rec = str(out / "prediction_recovery.yaml")
    detection._save_prediction_ledger(rec, str(tiles_dir), "modelA", files, log)
    detection._save_stitching_ledger(str(stitched), files[:2], log)
    data = yaml.safe_load(open(rec))
    assert data["model_path"] == "modelA" and data["files"][files[0]] == list(tiles)          # the reference's layout
    assert yaml.safe_load(open(stitched / "stitching_recovery.yaml")) == {"completed_files": ["img1", "img2"]}
    load = lambda model="modelA", excl=None: detection._load_prediction_ledger(rec, str(out), str(tiles_dir), model,
                                                                                 str(stitched), excl, log)
    assert load() == set(files[:2])                      # img3 is not in the stitching ledger, no tile files either
    assert load("modelB") == set()                       # another model: no recovery
    # per-tile prediction files present: the reference's count rule decides
    (out / "img3").mkdir()
    for k in range(3):
        (out / "img3" / f"Prediction_{k}.json").write_text("[]")
    assert files[2] not in load()                        # 3 files, 4 tiles
    assert files[2] in load(excl=["only_forest"])        # ... but 3 tiles once the forest-only tile is excluded
    (out / "img3" / "Prediction_3.json").write_text("[]")
    assert files[2] in load()
    os.remove(stitched / "img1.gpkg")
    assert files[0] not in load()                        # the stitched layer itself is gone


def test_fixture_predictor_fails_loudly(tmp_path):
    """a missing fixture must not turn into a silently empty crown layer (and a ledger entry that makes resumed runs
    skip the image); malformed fixtures are refused before anything reaches the device"""
    from treedetection_b200 import predictor
    tf = synth.image_transform(412000.0, 5318100.0, 0.2)
    tiles = tiling.tile_grid("img", tf, 500, 500, 25832, 50, 50, 20)
    model = tmp_path / "model"
    model.mkdir()
    fp = predictor.FixturePredictor(str(model))
    with pytest.raises(FileNotFoundError):
        fp.raw_outputs("img", tiles)
    det = predictor.FixturePredictor(str(model), allow_missing=True).raw_outputs("img", tiles)       # explicit opt-in
    assert len(det.scores) == 0 and det.tile_dims.shape == (len(tiles), 4)
    n, ids = 5, np.array(list(tiles))
    good = dict(boxes_net=np.zeros((n, 4), np.float32), scores=np.full(n, 0.5, np.float32),
                probs=np.zeros((n, 28, 28), np.float32), inst_tile=np.array([0, 0, 1, 2, 2], np.int32), tile_ids=ids)
    np.savez(model / "img.npz", **good)
    assert len(fp.raw_outputs("img", tiles).scores) == n
    for bad in (dict(inst_tile=np.array([0, 0, 1, 2, len(ids)], np.int32)),       # tile index out of range
                dict(inst_tile=np.array([0, 0, -1, 2, 2], np.int32)),
                dict(inst_tile=np.array([0, 2, 1, 2, 2], np.int32)),              # not tile-major
                dict(scores=np.zeros(n - 1, np.float32)),                        # lengths disagree
                dict(probs=np.zeros((n, 14, 14), np.float32))):
        np.savez(model / "img.npz", **{**good, **bad})
        with pytest.raises(ValueError):
            fp.raw_outputs("img", tiles)


def test_gpkg_conforms_to_the_geopackage_core_requirements(tmp_path):
    """No OGR / fiona exists offline to open the file, so the layer is checked against what the OGC GeoPackage
    encoding standard (12-128r15, core + features) requires and OGR / QGIS rely on: SQLite header fields,
    gpkg_spatial_ref_sys / gpkg_contents / gpkg_geometry_columns rows, an integer primary key, and the
    GeoPackageBinary header + WKB of every geometry (written by both the Python and the native row writer)."""
    import sqlite3
    import struct
    rng = np.random.default_rng(3)
    rings = [np.concatenate([r, r[:1]]) for r in (412000 + rng.uniform(0, 100, (k, 2)) for k in (4, 9, 33, 5))]
    verts = np.concatenate(rings)
    off = np.zeros(len(rings) + 1, dtype=np.int64); off[1:] = np.cumsum([len(r) for r in rings])
    n = len(rings)
    cols = {"Confidence_score": rng.uniform(0, 1, n), "poly_id": [str(k) for k in range(n)], "Area": rng.uniform(1, 50, n),
            "TreeHeight": rng.uniform(3, 30, n), "Centroid": [f"POINT ({k} {k})" for k in range(n)],
            "Diameter": rng.uniform(1, 9, n), "is_contained": ["False"] * n, "num_contained": np.arange(n)}
    for native in (False, True):
        path = str(tmp_path / f"layer_{native}.gpkg")
        gpkg.write_layer(path, "processed_x", verts, off, cols, gpkg.PROCESSED_SCHEMA, epsg=25832, native=native)
        raw = open(path, "rb").read(100)
        assert raw[:16] == b"SQLite format 3\x00"                               # R1
        assert struct.unpack(">I", raw[68:72])[0] == 0x47504B47                   # R2: application_id 'GPKG'
        assert struct.unpack(">I", raw[60:64])[0] >= 10200                        # user_version
        con = sqlite3.connect(path)
        try:
            assert con.execute("PRAGMA integrity_check").fetchone()[0] == "ok"   # R6
            assert con.execute("PRAGMA foreign_key_check").fetchall() == []      # R7
            srs = {r[0]: r for r in con.execute("SELECT srs_id, organization, organization_coordsys_id, definition "
                                                "FROM gpkg_spatial_ref_sys")}
            assert {-1, 0, 4326, 25832} <= set(srs)                               # R11 + the layer's own
            assert srs[25832][1].upper() == "EPSG" and srs[25832][2] == 25832 and "UTM zone 32N" in srs[25832][3]
            tname, dtype, ident, minx, miny, maxx, maxy, sid = con.execute(
                "SELECT table_name, data_type, identifier, min_x, min_y, max_x, max_y, srs_id FROM gpkg_contents").fetchone()
            assert (tname, dtype, sid) == ("processed_x", "features", 25832)      # R13 / R18
            assert minx <= verts[:, 0].min() and maxx >= verts[:, 0].max() and miny <= verts[:, 1].min() \
                and maxy >= verts[:, 1].max()
            gt, gc, gtype, gsrs, z, m = con.execute(
                "SELECT table_name, column_name, geometry_type_name, srs_id, z, m FROM gpkg_geometry_columns").fetchone()
            assert (gt, gtype, gsrs, z, m) == ("processed_x", "POLYGON", 25832, 0, 0)     # R22 - R28
            info = con.execute(f"PRAGMA table_info({tname})").fetchall()
            pk = [c for c in info if c[5] == 1]
            assert len(pk) == 1 and pk[0][2].upper() == "INTEGER"                 # R29: integer primary key
            names = [c[1] for c in info]
            assert gc in names and all(k in names for k in cols)
            blobs = [r[0] for r in con.execute(f"SELECT {gc} FROM {tname} ORDER BY {pk[0][1]}")]
        finally:
            con.close()
        assert len(blobs) == n
        for k, b in enumerate(blobs):                                              # R19: GeoPackageBinary
            assert b[:2] == b"GP" and b[2] == 0
            flags = b[3]
            assert flags & 1 == 1 and (flags >> 5) & 1 == 0 and (flags >> 4) & 1 == 0      # little endian, standard, non-empty
            assert (flags >> 1) & 7 == 1                                          # envelope: minx, maxx, miny, maxy
            assert struct.unpack("<i", b[4:8])[0] == 25832
            env = struct.unpack("<4d", b[8:40])
            ring = rings[k]
            assert env == (ring[:, 0].min(), ring[:, 0].max(), ring[:, 1].min(), ring[:, 1].max())
            order, wtype, nrings, npts = struct.unpack("<BIII", b[40:53])
            assert (order, wtype, nrings, npts) == (1, 3, 1, len(ring))           # WKB Polygon, one closed ring
            xy = np.frombuffer(b, dtype="<f8", offset=53).reshape(-1, 2)
            np.testing.assert_array_equal(xy, ring)
            assert (xy[0] == xy[-1]).all() and len(b) == 53 + 16 * len(ring)


def test_prediction_json_files_become_one_ragged_ring_set(tmp_path):
    """detection._rings_from_prediction_files: the reference's per-tile ``Prediction_*.json`` wire format
    (prediction.py:253-263) in tiles-JSON order; open rings are closed as shapely's Polygon() does, a tile with a
    degenerate ring or an RLE entry is dropped as a whole (helpers.py:419-476), missing tiles are skipped"""
    import logging

    from treedetection_b200 import detection
    tiles = {f"img_4120{k}0_5318000_50_20_25832": {"crs": 25832} for k in range(5)}
    ids = list(tiles)
    sq = lambda x: [[x, 0.0], [x + 1.0, 0.0], [x + 1.0, 1.0], [x, 1.0]]
    files = {
        ids[0]: [{"image_id": "a", "category_id": 0, "score": 0.9, "polygon_coords": [sq(0.0) + [[0.0, 0.0]]]},
                 {"image_id": "a", "category_id": 0, "score": 0.8, "polygon_coords": [sq(5.0)]}],            # open ring
        ids[1]: [{"image_id": "a", "category_id": 0, "score": 0.7, "polygon_coords": [[[0.0, 0.0], [1.0, 1.0]]]}],  # degenerate
        ids[2]: [],
        ids[3]: [{"image_id": "a", "category_id": 0, "score": 0.6, "segmentation": {"counts": "x", "size": [1, 1]}}],
        # ids[4]: no file
    }
    d = tmp_path / "img"
    d.mkdir()
    for tid, ev in files.items():
        (d / f"Prediction_{tid}.json").write_text(json.dumps(ev))
    (d / f"Prediction_{ids[4]}_extra.json").write_text("[]")
    records = []
    logger = logging.getLogger("stitch-test")
    logger.addHandler(type("H", (logging.Handler,), {"emit": lambda self, r: records.append(r.getMessage())})())
    v, off, ring_tile, conf = detection._rings_from_prediction_files(str(d), tiles, logger)
    np.testing.assert_array_equal(off, [0, 5, 10])
    np.testing.assert_array_equal(ring_tile, [0, 0])
    np.testing.assert_array_equal(conf, [0.9, 0.8])
    np.testing.assert_array_equal(v[:5], np.array(sq(0.0) + [[0.0, 0.0]]))
    np.testing.assert_array_equal(v[5:], np.array(sq(5.0) + [[5.0, 0.0]]))
    assert v.dtype == np.float64 and ring_tile.dtype == np.int32
    assert len(records) == 2 and ids[1] in records[0] and ids[3] in records[1]
    v, off, ring_tile, conf = detection._rings_from_prediction_files(str(tmp_path / "nothing"), tiles)
    assert v.shape == (0, 2) and list(off) == [0] and len(conf) == 0
    with pytest.raises(FileNotFoundError):
        detection.process_and_stitch_predictions(str(tmp_path / "no_tiles"), str(d), str(tmp_path / "out"))


def test_gpkg_reader_accepts_the_geometry_encodings_ogr_writes(tmp_path):
    """``geojson_predictions/*.gpkg`` may come from the reference (geopandas / fiona / OGR): feature table with
    ``fid`` + ``geom``, GeoPackageBinary blobs with any envelope type and byte order, Polygon / PolygonZ / PolygonZM,
    polygons with holes (the exterior ring is the crown), MultiPolygon (first polygon, postprocessing.py:493-494),
    NULL and empty geometries"""
    import sqlite3
    import struct
    sq = np.array([[412000.0, 5318000.0], [412004.0, 5318000.0], [412004.0, 5318003.0], [412000.0, 5318003.0],
                   [412000.0, 5318000.0]])
    hole = sq[::-1] * 1.0
    hole[:, 0] = 412001.0 + (hole[:, 0] - 412000.0) * 0.25
    hole[:, 1] = 5318001.0 + (hole[:, 1] - 5318000.0) * 0.25

    def wkb_poly(rings, bo="<", z=0):
        code = 3 + 1000 * z
        out = struct.pack(bo + "BII", 1 if bo == "<" else 0, code, len(rings))
        for r in rings:
            out += struct.pack(bo + "I", len(r))
            for x, y in r:
                out += struct.pack(bo + "dd" + "d" * {0: 0, 1: 1, 3: 2}[z], x, y, *([7.0] * {0: 0, 1: 1, 3: 2}[z]))
        return out

    def gpb(wkb, env=1, bo="<", srs=25832, empty=False):
        flags = (1 if bo == "<" else 0) | (env << 1) | (0x10 if empty else 0)
        nenv = {0: 0, 1: 4, 2: 6, 3: 6, 4: 8}[env]
        return b"GP\x00" + bytes([flags]) + struct.pack(bo + "i", srs) + struct.pack(bo + "d" * nenv, *([0.0] * nenv)) + wkb

    multi = struct.pack("<BII", 1, 6, 2) + wkb_poly([sq + 10.0]) + wkb_poly([sq + 50.0])
    blobs = [
        ("ogr default", gpb(wkb_poly([sq])), sq),
        ("no envelope", gpb(wkb_poly([sq + 1.0]), env=0), sq + 1.0),
        ("xyz envelope, PolygonZ", gpb(wkb_poly([sq + 2.0], z=1), env=2), sq + 2.0),
        ("xyzm envelope, PolygonZM", gpb(wkb_poly([sq + 3.0], z=3), env=4), sq + 3.0),
        ("big endian", gpb(wkb_poly([sq + 4.0], bo=">"), bo=">"), sq + 4.0),
        ("hole", gpb(wkb_poly([sq, hole])), sq),
        ("multipolygon", gpb(multi), sq + 10.0),
        ("empty", gpb(struct.pack("<BII", 1, 3, 0), env=0, empty=True), np.zeros((0, 2))),
        ("null", None, np.zeros((0, 2))),
    ]
    path = str(tmp_path / "ogr.gpkg")
    gpkg.write_layer(path, "crowns", np.zeros((0, 2)), np.zeros(1, dtype=np.int64), {"Confidence_score": np.zeros(0)},
                     gpkg.STITCHED_SCHEMA, epsg=25832, native=False)           # metadata tables of a valid file
    con = sqlite3.connect(path)
    con.execute('DROP TABLE "crowns"')
    con.execute('CREATE TABLE "crowns" (fid INTEGER PRIMARY KEY AUTOINCREMENT NOT NULL, geom POLYGON, '
                '"Confidence_score" REAL, note TEXT)')
    con.execute("UPDATE gpkg_geometry_columns SET column_name = 'geom'")
    for k, (name, blob, _) in enumerate(blobs):
        con.execute('INSERT INTO "crowns" (geom, "Confidence_score", note) VALUES (?, ?, ?)', (blob, 0.1 * k, name))
    con.commit()
    con.close()
    verts, off, cols, epsg = gpkg.read_layer(path)
    assert epsg == 25832 and cols["note"] == [b[0] for b in blobs]
    np.testing.assert_allclose(cols["Confidence_score"], [0.1 * k for k in range(len(blobs))])
    assert len(off) == len(blobs) + 1
    for k, (name, _, want) in enumerate(blobs):
        np.testing.assert_array_equal(verts[off[k]:off[k + 1]], want, err_msg=name)
