"""Drop-in API on the B200: get_config -> preprocess_files -> predict_tiles ->
postprocess_files on a two-image synthetic mosaic (seam strip included), artefacts in the
reference's directory layout, final crowns compared with the CPU oracle."""
import json
import os

import numpy as np
import pytest
import yaml

from oracle import port
from treedetection_b200 import detection, geo, geotiff, gpkg, predictor, synth

pytestmark = pytest.mark.gpu

PX = 0.2


def _project(tmp_path, compression=None):
    left, bottom = synth.ORIGIN_X, synth.ORIGIN_Y
    field = synth.tree_field(77, 400.0, 200.0, 5000.0, left, bottom)
    rgbi = synth.make_rgbi(field, PX, 77)            # (4, 1000, 2000)
    ndsm = synth.make_ndsm(field, 1.0, 77)           # (200, 400)
    img_dir, h_dir = tmp_path / "rgb", tmp_path / "ndsm"
    img_dir.mkdir(); h_dir.mkdir()
    top = bottom + 200.0
    for k, name in enumerate(("000001", "000002")):
        x0 = left + 200.0 * k
        geotiff.write(str(img_dir / f"FDOP20_{name}_rgbi.tif"), rgbi[:, :, 1000 * k:1000 * (k + 1)],
                      (PX, 0.0, x0, 0.0, -PX, top), epsg=25832, compression=compression,
                      predictor=2 if compression else 1)
        geotiff.write(str(h_dir / f"nDSM_{name}_1km.tif"), ndsm[:, 200 * k:200 * (k + 1)],
                      (1.0, 0.0, x0, 0.0, -1.0, top), epsg=25832, nodata=-3.4028234663852886e38, compression=compression)
    model = tmp_path / "model_fixtures"
    model.mkdir()
    cfg = {
        "image_directory": str(img_dir), "height_data_path": str(h_dir),
        "image_regex": "FDOP20_(\\d+)_rgbi\\.tif", "height_data_regex": "nDSM_(\\d+)_1km\\.tif",
        "combined_model": str(model), "output_directory": str(tmp_path / "output"), "tiles_path": str(tmp_path / "tiles"),
        "use_overlap": True, "merged_path": "merged", "overlapping_tiles_width": 3, "overlapping_tiles_height": 3,
        "image_merged_regex": "FDOP20_(\\d+)_(\\d+)_(\\d+)_(\\d+)_rgbi\\.tif",
        "height_data_merged_regex": "nDSM_(\\d+)(\\d+)_1km\\.tif",
        "tile_width": 50, "tile_height": 50, "buffer": 20, "batch_size": 10,
        "ndvi_scaling_factor": 0.2, "height_scaling_factor": 1.0, "confidence_threshold": 0.3,
        "containment_threshold": 0.75, "height_threshold": 3, "ndvi_mean_threshold": 0.1, "ndvi_var_threshold": 0.1,
        "iou_threshold": 0.6, "confidence_threshold_stitching": 0.3, "area_threshold": 1, "parallel": True,
        "num_workers": 5, "keep_intermediate": True, "device": "0",
    }
    path = tmp_path / "config.yml"
    path.write_text(yaml.safe_dump(cfg))
    return str(path), field, model


def test_process_files_end_to_end(tmp_path):
    cfg_path, field, model = _project(tmp_path)
    config, cfg_obj = detection.get_config(cfg_path)
    assert cfg_obj.tile_width == 50 and config["simplify_tolerance"] == 0.2 and config["device"] == "0"
    images = detection.preprocess_files(config)
    # two images + the right-seam strip (270 px wide, full height)
    assert len(images) == 3
    strip = [p for p in images if "merged" in p]
    assert len(strip) == 1 and os.path.basename(strip[0]) == "FDOP20_412000_5318200_412200_5318200_rgbi.tif"
    sinfo = geotiff.read_info(strip[0])
    assert (sinfo.width, sinfo.height) == (270, 1000)
    assert os.path.exists(os.path.join(config["height_data_path"], "merged", "nDSM_41200053182004122005318200_1km.tif"))
    assert os.path.exists(os.path.join(config["tiles_path"], "recovery.yaml"))
    # stage the "model": ROI-head fixtures for every image, from the same tree field
    dets = {}
    for p in images:
        stem = os.path.splitext(os.path.basename(p))[0]
        tiles = json.load(open(os.path.join(config["tiles_path"], stem + ".json")))
        dets[stem] = synth.make_detections(field, tiles, PX, seed=5)
        predictor.dump_fixtures(str(model), stem, dets[stem])
    detection.predict_tiles(config)
    detection.postprocess_files(config)
    out = config["output_directory"]
    for p in images:
        stem = os.path.splitext(os.path.basename(p))[0]
        assert os.path.exists(os.path.join(out, "geojson_predictions", stem + ".gpkg"))
        assert os.path.exists(os.path.join(out, "geojson_predictions", f"processed_{stem}.gpkg"))
        assert os.path.exists(os.path.join(out, stem + ".gpkg"))
        assert os.path.isdir(os.path.join(out, "predictions", stem))
    # ---- parity of one plain image and of the seam strip with the oracle ----
    for p in (images[0], strip[0]):
        stem = os.path.splitext(os.path.basename(p))[0]
        tiles = json.load(open(os.path.join(config["tiles_path"], stem + ".json")))
        rings, conf = port.predict_stage(dets[stem], tiles)
        v, o, cols, epsg = gpkg.read_layer(os.path.join(out, "geojson_predictions", stem + ".gpkg"))
        assert epsg == 25832 and len(o) - 1 == len(rings)
        np.testing.assert_array_equal(v, np.array([q for r in rings for q in r]).reshape(-1, 2))
        np.testing.assert_array_equal(np.array(cols["Confidence_score"]), np.array(conf))
        rgbi, rinfo = geotiff.read(p)
        if "merged" in p:
            hp = os.path.join(config["height_data_path"], "merged", "nDSM_41200053182004122005318200_1km.tif")
        else:
            hp = os.path.join(config["height_data_path"], "nDSM_000001_1km.tif")
        ndsm, hinfo = geotiff.read(hp)
        H, W = rgbi.shape[1:]
        oh, ow = int(H * 0.2), int(W * 0.2)
        dec = np.stack([port.decimate_bilinear(rgbi[b], oh, ow) for b in (0, 0, 0, 3)])
        ndvi = port.ndvi_from_rgbi(dec).astype(np.float32)
        ndvi_tf = geo.compose(rinfo.transform, geo.scale(W / ow, H / oh))
        want, dbg = port.post_process(rings, conf, ndvi, ndvi_tf, tuple(geo.raster_bounds(rinfo.transform, W, H)),
                                      ndsm[0], hinfo.transform,
                                      tuple(geo.raster_bounds(hinfo.transform, hinfo.width, hinfo.height)), PX, PX,
                                      config)
        v, o, cols, _ = gpkg.read_layer(os.path.join(out, stem + ".gpkg"))
        assert list(cols.keys()) == list(gpkg.PROCESSED_SCHEMA.keys())
        assert cols["poly_id"] == [w["poly_id"] for w in want]
        np.testing.assert_array_equal(np.array(cols["Area"]), np.array([w["Area"] for w in want]))
        np.testing.assert_array_equal(np.array(cols["TreeHeight"], dtype=np.float32),
                                      np.array([w["TreeHeight"] for w in want], dtype=np.float32))
        assert cols["is_contained"] == [str(w["is_contained"]) for w in want]
        assert cols["Centroid"] == [json.dumps({"x": w["Centroid"][0], "y": w["Centroid"][1]}) for w in want]
        np.testing.assert_array_equal(v, np.array([q for w in want for q in w["coords"]]).reshape(-1, 2))
        if "merged" not in p:
            assert len(want) > 10
    # cleanup removes intermediates unless asked to keep them
    config["keep_intermediate"] = False
    detection.cleanup_files(config)
    assert not os.path.exists(config["tiles_path"])
    assert sorted(os.listdir(out)) == sorted(["logs"] + [os.path.splitext(os.path.basename(p))[0] + ".gpkg"
                                                        for p in images])


def test_config_errors(tmp_path):
    p = tmp_path / "c.yml"
    p.write_text(yaml.safe_dump({"image_directory": str(tmp_path / "missing"), "height_data_path": str(tmp_path)}))
    with pytest.raises(AssertionError):
        detection.get_config(str(p))


class _FieldPredictor:
    """predictor plug that synthesises the ROI-head outputs of any tiling from one tree field"""

    def __init__(self, field):
        self.field = field

    def raw_outputs(self, stem, tiles):
        return synth.make_detections(self.field, tiles, PX, seed=5)


def test_process_files_fast_path_writes_the_same_files(tmp_path):
    """``process_files`` (one session: every image decoded once into pinned memory, P1 + the CUDA-graph chain
    while the next image is decoded, both artefacts written from the device results) against the same stages
    called one by one (each stage re-reads the previous stage's artefacts): identical layers on disk."""
    layers = {}
    for mode in ("session", "staged"):
        root = tmp_path / mode
        root.mkdir()
        cfg_path, field, _ = _project(root)
        config, _ = detection.get_config(cfg_path)
        config["predictor"] = _FieldPredictor(field)
        if mode == "session":
            detection.process_files(config)
            stats = config["_last_session_stats"]
            assert stats["images"] == 3 and stats["fallback_images"] == 0, stats
            assert set(stats["stage_s"]) >= {"preprocess", "predict", "postprocess", "decode", "device", "write"}
        else:
            detection.preprocess_files(config)
            detection.predict_tiles(config)
            detection.postprocess_files(config)
        out = config["output_directory"]
        got = {}
        for sub in ("geojson_predictions", "."):
            d = os.path.join(out, sub)
            for name in sorted(os.listdir(d)):
                if name.endswith(".gpkg"):
                    got[os.path.join(sub, name)] = gpkg.read_layer(os.path.join(d, name))
        layers[mode] = got
    a, b = layers["session"], layers["staged"]
    assert sorted(a) == sorted(b) and len(a) == 9          # 3 x (stitched, processed, final)
    for name in a:
        (va, oa, ca, ea), (vb, ob, cb, eb) = a[name], b[name]
        np.testing.assert_array_equal(va, vb, err_msg=name)
        np.testing.assert_array_equal(oa, ob, err_msg=name)
        assert list(ca) == list(cb) and ea == eb, name
        for col in ca:
            assert list(ca[col]) == list(cb[col]), (name, col)
    assert any(len(a[n][1]) - 1 > 10 for n in a if n.startswith("./"))


def test_process_files_decodes_lzw_rasters_on_the_device(tmp_path):
    """LZW-compressed inputs (predictor 2 imagery, float32 nDSM): the session decodes them on the GPU
    (geotiff.read_device: compressed bytes over PCIe, one warp per strip) and writes the same layers as for the
    uncompressed files; ``host_decode: true`` keeps the host reader."""
    layers, stats = {}, {}
    for mode, compression, host_decode in (("plain", None, False), ("lzw", "lzw", False), ("lzw_host", "lzw", True)):
        root = tmp_path / mode
        root.mkdir()
        cfg_path, field, _ = _project(root, compression)
        config, _ = detection.get_config(cfg_path)
        config["predictor"] = _FieldPredictor(field)
        config["host_decode"] = host_decode
        detection.process_files(config)
        stats[mode] = config["_last_session_stats"]
        out = config["output_directory"]
        layers[mode] = {n: gpkg.read_layer(os.path.join(out, n)) for n in sorted(os.listdir(out)) if n.endswith(".gpkg")}
    assert stats["plain"]["device_decoded_rasters"] == 0 and stats["lzw_host"]["device_decoded_rasters"] == 0
    # the two images' rasters (the merged seam strip is written uncompressed by the merge step)
    assert stats["lzw"]["device_decoded_rasters"] == 4 and stats["lzw"]["fallback_images"] == 0, stats["lzw"]
    for mode in ("lzw", "lzw_host"):
        assert sorted(layers[mode]) == sorted(layers["plain"]) and len(layers[mode]) == 3
        for name, (v, o, c, e) in layers["plain"].items():
            v2, o2, c2, e2 = layers[mode][name]
            np.testing.assert_array_equal(v2, v, err_msg=name)
            np.testing.assert_array_equal(o2, o, err_msg=name)
            assert c2 == c and e2 == e, name


def test_exclude_files_remove_crowns_within_the_outline(tmp_path, dev):
    """helpers.exclude_outlines (helpers.py:33-69): crowns of the existing processed_* layers that lie within the
    union of an exclusion outline disappear; the others (incl. those only touching it) stay, columns intact."""
    import torch
    from tests.test_gpu_two_model import rect, write_shapefile
    cfg_path, field, _ = _project(tmp_path)
    config, _ = detection.get_config(cfg_path)
    config["predictor"] = _FieldPredictor(field)
    detection.process_files(config)                       # keep_intermediate: the processed_* layers stay
    pred = os.path.join(config["output_directory"], "geojson_predictions")
    name = "processed_FDOP20_000001_rgbi.gpkg"
    v, o, cols, _ = gpkg.read_layer(os.path.join(pred, name))
    rings = [[(float(x), float(y)) for x, y in v[o[i]:o[i + 1]]] for i in range(len(o) - 1)]
    left, bottom = synth.ORIGIN_X, synth.ORIGIN_Y
    outline = [[rect(left + 30.0, bottom + 30.0, left + 120.0, bottom + 150.0), rect(left + 60.0, bottom + 60.0, left + 90.0, bottom + 100.0)]]
    write_shapefile(str(tmp_path / "water.shp"), outline)
    _, within = port.forest_predicates(rings, outline)
    assert 0 < within.sum() < len(rings)
    config["exclude_files"] = [str(tmp_path / "water.shp"), str(tmp_path / "missing.shp")]     # a missing file is logged, not raised
    with torch.cuda.device(dev):
        detection.exclude_outlines(config, config["logger"])
    v2, o2, cols2, _ = gpkg.read_layer(os.path.join(pred, name))
    keep = np.nonzero(~within)[0]
    assert len(o2) - 1 == len(keep) and list(cols2) == list(cols)
    np.testing.assert_array_equal(v2, np.concatenate([v[o[k]:o[k + 1]] for k in keep]))
    assert cols2["poly_id"] == [cols["poly_id"][k] for k in keep]
    assert cols2["TreeHeight"] == [cols["TreeHeight"][k] for k in keep]


def test_non_zero_device_index(tmp_path):
    """``device: "1"`` in config.yml: every entry point makes that GPU current, so tensors, streams and the library's
    scratch all live there.  Needs two GPUs; the layers must equal those of a run on device 0."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    layers = {}
    for dev_index in ("0", "1"):
        root = tmp_path / f"dev{dev_index}"
        root.mkdir()
        cfg_path, field, _ = _project(root)
        config, _ = detection.get_config(cfg_path)
        config["device"] = dev_index
        config["predictor"] = _FieldPredictor(field)
        detection.process_files(config)
        out = config["output_directory"]
        layers[dev_index] = {n: gpkg.read_layer(os.path.join(out, n)) for n in sorted(os.listdir(out)) if n.endswith(".gpkg")}
    assert torch.cuda.current_device() == 0                  # the guard restores the caller's device
    a, b = layers["0"], layers["1"]
    assert sorted(a) == sorted(b) and len(a) == 3
    for name in a:
        np.testing.assert_array_equal(a[name][0], b[name][0], err_msg=name)
        np.testing.assert_array_equal(a[name][1], b[name][1], err_msg=name)
        assert a[name][2] == b[name][2], name


def test_stitching_from_prediction_json_files(tmp_path):
    """detection.process_and_stitch_predictions (helpers.py:556-600): predictions that exist as per-tile
    ``Prediction_*.json`` files -- the reference's wire format, written here by ``keep_intermediate`` -- give the same
    stitched layers as the device path that produced them; a second call is answered from the stitching ledger"""
    cfg_path, field, model = _project(tmp_path)
    config, _ = detection.get_config(cfg_path)
    images = detection.preprocess_files(config)
    for p in images:
        stem = os.path.splitext(os.path.basename(p))[0]
        tiles = json.load(open(os.path.join(config["tiles_path"], stem + ".json")))
        predictor.dump_fixtures(str(model), stem, synth.make_detections(field, tiles, PX, seed=5))
    detection.predict_tiles(config)
    out = config["output_directory"]
    again = str(tmp_path / "stitched_again")
    got = detection.process_and_stitch_predictions(config["tiles_path"], os.path.join(out, "predictions"), again,
                                                   shift=1, simplify_tolerance=config["simplify_tolerance"])
    assert got == again
    n_rows = 0
    for p in images:
        stem = os.path.splitext(os.path.basename(p))[0]
        v0, o0, c0, e0 = gpkg.read_layer(os.path.join(out, "geojson_predictions", stem + ".gpkg"))
        v1, o1, c1, e1 = gpkg.read_layer(os.path.join(again, stem + ".gpkg"))
        assert e0 == e1 == 25832
        np.testing.assert_array_equal(o1, o0)
        np.testing.assert_array_equal(v1, v0)
        np.testing.assert_array_equal(np.array(c1["Confidence_score"]), np.array(c0["Confidence_score"]))
        n_rows += len(o1) - 1
    assert n_rows > 50
    ledger = yaml.safe_load(open(os.path.join(again, "stitching_recovery.yaml")))
    assert len(ledger["completed_files"]) == len(images)
    stamp = {f: os.path.getmtime(os.path.join(again, f)) for f in os.listdir(again) if f.endswith(".gpkg")}
    detection.process_and_stitch_predictions(config["tiles_path"], os.path.join(out, "predictions"), again)
    assert stamp == {f: os.path.getmtime(os.path.join(again, f)) for f in stamp}
