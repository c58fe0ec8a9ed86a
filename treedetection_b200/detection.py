"""Public entry points -- same names, arguments, side effects (directories, file names) and
error behaviour as ``TreeDetection/detection.py``:

    preprocess_files(config) -> list[str]      detection.py:256-339
    predict_tiles(config)    -> None           detection.py:134-253
    postprocess_files(config)-> None           detection.py:23-59
    process_files(config)    -> None           detection.py:342-373
    cleanup_files(config)                      detection.py:375-399

Between the artefacts the work is done by the device pipeline (``pipeline.py`` over the
C-ABI kernels).  Called stage by stage (as the reference's API allows) every stage reads its
inputs from the previous stage's artefacts.  ``process_files`` runs the stages inside a
:class:`_Session`: while ``predict_tiles`` has an image's rasters and crown table on the device
it also runs that image's post-processing (``api.run_image`` / ``pipeline.ChainRunner``: one read
of each raster into pinned memory, one CUDA-graph pass, the next image's rasters decoded meanwhile)
and writes BOTH artefacts; ``postprocess_files`` then finds its outputs in place and only does its
bookkeeping.  The files on disk are the same either way (tests/test_gpu_api.py).
The Mask R-CNN forward is replaced by the predictor plug (``predictor.py``)."""
from __future__ import annotations

import datetime
import json
import os
import re
import shutil
import time
from pathlib import Path

import numpy as np
import torch
import yaml

from . import api, geo, geotiff, gpkg, ops, pipeline, tiling
from .config import Config, get_config  # noqa: F401  (re-exported like the reference)
from .merging import merge_and_crop_images
from .predictor import FixturePredictor


def _yaml_dump(data, f):
    """``yaml.safe_dump`` through libyaml when PyYAML was built with it (same bytes, ~7x faster: a prediction ledger
    lists every tile id of every image)"""
    dumper = getattr(yaml, "CSafeDumper", yaml.SafeDumper)
    yaml.dump(data, f, Dumper=dumper, sort_keys=False)


def _yaml_load(f):
    return yaml.load(f, Loader=getattr(yaml, "CSafeLoader", yaml.SafeLoader))


class _Session:
    """Device-side state shared by the stages of ONE ``process_files`` call (the fast path)."""

    def __init__(self):
        self.processed = {}        # stitched .gpkg path -> processed_*.gpkg written while predicting
        self.runners = {}          # tiling key -> pipeline.ChainRunner (capacities + CUDA graphs per tiling)
        self.tables = {}           # tiling key -> (api.TileTables, reusable P1 output buffer)
        self.pinned = {}           # (name, shape, dtype, parity) -> pinned staging tensor
        self.loader = None         # background decoder of the NEXT image's rasters and fixtures
        self.writer = None         # background writer of the PREVIOUS image's crown layers
        self._loader_stream = None # CUDA stream of the decoder thread (device-side raster decode)
        self.timeline = []         # per image: seconds waited for the decoder, on the device, ... (bench.py e2e_files)
        self.stage_s = {}          # wall seconds per stage (read by bench.py's e2e_files)
        self.images = 0
        self.fallback_images = 0   # images that went through the stage-by-stage path
        self.device_decoded = 0    # rasters decoded on the GPU (LZW strips / tiles, geotiff.read_device)

    def pinned_array(self, name, shape, dtype, parity):
        key = (name, tuple(shape), np.dtype(dtype).str, parity)
        t = self.pinned.get(key)
        if t is None:
            tdt = {"|u1": torch.uint8, "<f4": torch.float32, "<u2": torch.int16, "<i2": torch.int16}[np.dtype(dtype).str]
            t = torch.empty(tuple(shape), dtype=tdt, pin_memory=True)
            self.pinned[key] = t
        return t

    def device_array(self, name, shape, dtype, parity, device):
        """persistent device buffer of a raster decoded on the GPU (two parities: decode k + 1 while k runs)"""
        key = ("dev", name, tuple(shape), np.dtype(dtype).str, parity)
        t = self.pinned.get(key)
        if t is None:
            tdt = {"|u1": torch.uint8, "<f4": torch.float32}[np.dtype(dtype).str]
            t = torch.empty(tuple(shape), dtype=tdt, device=device)
            self.pinned[key] = t
        return t

    def loader_stream(self, device):
        if self._loader_stream is None:
            self._loader_stream = torch.cuda.Stream(device=device)
        return self._loader_stream

    def pinned_copy(self, name, array, parity):
        """``array`` (host numpy) copied into a pinned staging buffer that grows by capacity"""
        a = np.ascontiguousarray(array)
        key = ("fx", name, a.dtype.str, parity)
        t = self.pinned.get(key)
        if t is None or t.numel() < a.size:
            tdt = {"<f4": torch.float32, "<i4": torch.int32}[a.dtype.str]
            t = torch.empty((max(int(a.size * 1.25), 16),), dtype=tdt, pin_memory=True)
            self.pinned[key] = t
        v = t[:a.size].view(a.shape)
        if a.nbytes >= (32 << 20):
            # NumPy's copy loop releases the GIL: a few threads move the mask probabilities (~100 MB per image)
            from concurrent.futures import ThreadPoolExecutor
            src, dst = a.reshape(-1), v.numpy().reshape(-1)
            cuts = np.linspace(0, a.size, 5).astype(np.int64)
            with ThreadPoolExecutor(max_workers=4) as ex:
                list(ex.map(lambda k: np.copyto(dst[cuts[k]:cuts[k + 1]], src[cuts[k]:cuts[k + 1]]), range(4)))
        elif a.size:
            np.copyto(v.numpy(), a)
        return v

    def close(self):
        for name in ("loader", "writer"):
            pool = getattr(self, name)
            if pool is not None:
                pool.shutdown(wait=True)
                setattr(self, name, None)
        self.runners.clear(); self.tables.clear(); self.pinned.clear()


def _device(config):
    dev = config.get("device", "0")
    if dev == "cpu" or not ops._lib.cuda_available():
        raise RuntimeError("treedetection_b200 needs a CUDA device: there is no CPU implementation of this path")
    return torch.device("cuda", int(dev))


# --------------------------------------------------------------------------------------
# tiling (P0b)
# --------------------------------------------------------------------------------------
def tile_data(file_list, out_dir, buffer=30, tile_width=200, tile_height=200, parallel=False, max_workers=4,
              forest_shapefile=None, logger=None, device=None):
    """preprocessing.py:125-224: one ``<stem>.json`` per image + ``recovery.yaml``."""
    os.makedirs(out_dir, exist_ok=True)
    recovery_file = os.path.join(out_dir, "recovery.yaml")
    recovered = set()
    if os.path.exists(recovery_file):
        try:
            with open(recovery_file) as f:
                rec = yaml.safe_load(f) or {}
            if rec.get("buffer") == buffer and rec.get("tile_width") == tile_width and rec.get("tile_height") == tile_height:
                recovered = {f for f in rec.get("processed_files", [])
                             if os.path.exists(Path(out_dir) / f"{Path(f).stem}.json")}
        except Exception as e:
            if logger:
                logger.warning(f"Could not load recovery file: {e}")
    todo = [f for f in file_list if f not in recovered]
    if not todo:
        if logger:
            logger.info("All files have already been processed. Exiting Tiling.")
        return
    forest = None
    if forest_shapefile:
        from .fusion import ForestIndex
        forest = ForestIndex.from_file(forest_shapefile, device=device, logger=logger)
    total = len(todo)
    for i, data_path in enumerate(todo):
        try:
            if not os.path.isfile(data_path):
                raise FileNotFoundError(f"File not found: {data_path}")
            info = geotiff.read_info(data_path)
            tiles = tiling.tile_grid(Path(data_path).stem, info.transform, info.width, info.height, info.epsg,
                                     tile_width, tile_height, buffer, forest)
            with open(Path(out_dir) / f"{Path(data_path).stem}.json", "w") as f:
                f.write(json.dumps(tiles))
            cur, prev = int(100 * (i + 1) / total), int(100 * i / total)
            if logger and ((cur // 5) != (prev // 5) or cur == 100 or i == 0):
                logger.info(f"Tiling file {i + 1}/{total} ({cur}%)")
        except Exception as e:
            if logger:
                logger.error(f"Error processing file: {e}")
    try:
        with open(recovery_file, "w") as f:
            yaml.safe_dump({"buffer": buffer, "tile_width": tile_width, "tile_height": tile_height,
                            "file_list": todo + list(recovered), "processed_files": todo + list(recovered)}, f,
                           sort_keys=False)
    except Exception as e:
        if logger:
            logger.warning(f"Failed to save recovery file: {e}")


def preprocess_files(config):
    config_obj = Config()
    config_obj._load_into_config(config)
    logger = config["logger"]

    images_directory = config["image_directory"]
    height_data_directory = config["height_data_path"]
    if not os.path.exists(images_directory):
        raise FileNotFoundError(f"Image directory not found: {images_directory}")
    if not os.path.isdir(images_directory):
        raise NotADirectoryError(f"Image directory is not a directory: {images_directory}")
    if not os.path.exists(height_data_directory):
        raise FileNotFoundError(f"Height directory not found: {height_data_directory}")
    if not os.path.isdir(height_data_directory):
        raise NotADirectoryError(f"Height directory is not a directory: {height_data_directory}")

    images_paths = [os.path.join(images_directory, f) for f in sorted(os.listdir(images_directory)) if f.endswith(".tif")]
    height_paths = [os.path.join(height_data_directory, f) for f in sorted(os.listdir(height_data_directory))
                    if f.endswith(".tif")]
    if os.path.exists(config["continue"]):
        with open(config["continue"], "r") as f:
            continue_files = f.read().splitlines()
        images_paths = [f for f in images_paths if f not in continue_files]

    image_regex_pattern = re.compile(config["image_regex"])
    height_data_regex_pattern = re.compile(config["height_data_regex"])
    images_paths = [f for f in images_paths if image_regex_pattern.search(os.path.basename(f))]
    height_data_paths = [f for f in height_paths if height_data_regex_pattern.search(os.path.basename(f))]
    image_identifiers = {}
    for f in images_paths:
        match = image_regex_pattern.search(os.path.basename(f))
        if match:
            image_identifiers["".join(match.groups())] = f
    height_data_identifiers = {}
    for f in height_data_paths:
        match = height_data_regex_pattern.search(os.path.basename(f))
        if match:
            height_data_identifiers["".join(match.groups())] = f

    # every kernel of libtreedet runs on the CURRENT device's stream and scratch pool: make the configured
    # device current for the duration of the call (config["device"] may name any GPU of the box)
    dev = _device(config)
    if config["use_overlap"]:
        logger.info("Using overlapping tiles for processing, do merging right now ...")
        with torch.cuda.device(dev):
            merge_and_crop_images(config, images_paths, height_paths, dev)

    for identifier, image_path in image_identifiers.items():
        if identifier not in height_data_identifiers:
            logger.warning(f"No corresponding height data found for image file {image_path}")

    if not images_paths:
        raise FileNotFoundError(
            f"No image TIF-files matching the pattern found in the directory: {images_directory} or all files have "
            f"already been processed.")

    logger.info(f"Found {len(images_paths)} images for processing. Starting tiling...")
    with torch.cuda.device(dev):
        tile_data(images_paths, config["tiles_path"], config["buffer"], config["tile_width"], config["tile_height"],
                  parallel=config["parallel"], max_workers=config["num_workers"], logger=config["logger"],
                  forest_shapefile=config.get("forrest_outline", None), device=dev)
    return images_paths


# --------------------------------------------------------------------------------------
# prediction side (P1-P4)
# --------------------------------------------------------------------------------------
def _write_tile_predictions(pred_subdir, tifpath, tiles, rings, inst_tile, scores):
    """``Prediction_<tile_id>.json`` per tile (prediction.py:253-263), only when
    intermediates are kept -- these are the un-simplified, un-filtered rings of P3."""
    os.makedirs(pred_subdir, exist_ok=True)
    verts = rings.verts.cpu().numpy()
    off = rings.ring_off.cpu().numpy()
    rinst = rings.ring_inst.cpu().numpy()
    per_tile = {tid: [] for tid in tiles}
    tile_ids = list(tiles.keys())
    for r in range(len(off) - 1):
        i = int(rinst[r])
        poly = [[float(x), float(y)] for x, y in verts[off[r]:off[r + 1]]]
        per_tile[tile_ids[int(inst_tile[i])]].append(
            {"image_id": tifpath, "category_id": 0, "score": float(scores[i]), "polygon_coords": [poly]})
    for tid, ev in per_tile.items():
        with open(os.path.join(pred_subdir, f"Prediction_{os.path.basename(tid)}.json"), "w") as f:
            f.write(json.dumps(ev))


class _FastPath:
    """One image of ``process_files`` end to end while its data is on the device, as a three-stage pipeline over the
    images: a LOADER thread decodes the next image's rasters and ROI-head fixtures into pinned staging buffers, the
    calling thread runs ``api.run_image`` (P1 + the CUDA-graph chain P2-P9, H2D copies overlapped with the stages),
    and a WRITER thread writes both artefacts of the previous image -- ``geojson_predictions/<stem>.gpkg`` and
    ``processed_<stem>.gpkg``.  ``run`` returns False for anything it does not cover (no matching nDSM, 16-bit
    imagery, a predictor that consumes the tiles): the caller then takes the stage-by-stage path."""

    MAX_PENDING_WRITES = 3     # images whose layers may be queued for the writer threads

    def __init__(self, config, session, dev, params, predictor, stitched_path, output_path, logger, n_images=1):
        from concurrent.futures import ThreadPoolExecutor
        self.config, self.s, self.dev, self.p = config, session, dev, params
        self.predictor, self.stitched_path, self.output_path, self.logger = predictor, stitched_path, output_path, logger
        if session.loader is None:
            session.loader = ThreadPoolExecutor(max_workers=1)
        if session.writer is None:
            session.writer = ThreadPoolExecutor(max_workers=2)    # the two layers of an image are independent files
        self.pending = {}          # image path -> Future of _load
        self.writes = []           # (image path, Future of _write), oldest first
        self.failed = set()        # images whose layers could not be written
        self.parity = 0
        # ONE set of pinned host staging buffers (pinning memory costs ~0.5 s / GB and stalls other CUDA calls while
        # it happens): the decoder thread may overwrite them as soon as the previous image's host->device copies
        # are done -- run_image reports that through `on_h2d` a few milliseconds after it starts
        # A long file list amortises a second set (the decoder then never waits for the copies: ~12 ms per image).
        import threading
        self.two_sets = n_images >= 32
        self.h2d_gate = threading.Event()
        self.h2d_gate.set()
        self.h2d_event = None
        image_pattern = re.compile(config.get("image_regex") or "(\\d+)\\.tif")
        height_pattern = re.compile(config.get("height_data_regex") or "(\\d+)\\.tif")
        self.patterns = (image_pattern, height_pattern, re.compile(config["image_merged_regex"]),
                         re.compile(config["height_data_merged_regex"]))
        self.height_index = _build_file_index(config["height_data_path"], height_pattern)
        self.image_index = _build_file_index(config["image_directory"], image_pattern)

    def _load(self, fp, tiles_path, parity):
        """decode one image's rasters and fixtures into pinned staging buffers (runs on the loader thread)"""
        t0 = time.time()
        torch.cuda.set_device(self.dev)              # the pinned staging buffers belong to this device's context
        stem = Path(fp).stem
        with open(os.path.join(tiles_path, stem + ".json")) as f:
            tiles = json.load(f)
        hpath, ipath = _match_rasters(stem, self.patterns, self.config["height_data_path"],
                                      self.config["image_directory"], self.height_index, self.image_index)
        if hpath is None or ipath is None:
            return None
        rinfo, hinfo = geotiff.read_info(fp), geotiff.read_info(hpath)
        if rinfo.dtype != np.uint8 or rinfo.count < 4 or hinfo.dtype != np.float32:
            return None
        if not self.two_sets:                         # the previous image's copies out of the staging buffers
            self.h2d_gate.wait()
            if self.h2d_event is not None:
                self.h2d_event.synchronize()
        ready, status = {}, []
        rgbi = self._read_raster(fp, "rgbi", rinfo, np.uint8, parity, ready, status)
        ndsm = self._read_raster(hpath, "ndsm", hinfo, np.float32, parity, ready, status)
        t1 = time.time()
        det = None
        if not getattr(self.predictor, "wants_tiles", False):
            raw = self.predictor.raw_outputs(stem, tiles)
            det = {k: self.s.pinned_copy(k, getattr(raw, k), parity if self.two_sets else 0)
                   for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
        return {"tiles": tiles, "rgbi": rgbi, "rinfo": rinfo, "ndsm": ndsm[0], "hinfo": hinfo, "det": det,
                "ready": ready, "status": status, "decode_s": t1 - t0, "fixtures_s": time.time() - t1}

    def _read_raster(self, path, name, info, dtype, parity, ready, status):
        """One raster into the staging buffers of ``parity``: LZW-compressed files are decoded ON THE DEVICE
        (geotiff.read_device: the compressed bytes cross PCIe, one warp per strip / tile), everything else goes
        through the host reader into pinned memory."""
        shape = (info.count, info.height, info.width)
        if not self.config.get("host_decode", False) and geotiff.device_decodable(path):
            out = self.s.device_array(name, shape, dtype, parity, self.dev)
            stream = self.s.loader_stream(self.dev)
            with torch.cuda.stream(stream):
                # staging buffers per (parity, raster): the copies of one raster are still in flight while the
                # next one is read
                res = geotiff.read_device(path, self.dev, out=out, slot=2 * parity + (name != "rgbi"))
                if res is not None:
                    self.s.device_decoded += 1
                    ready[name] = stream.record_event()
                    status.append((path, res[2]))
                    return out
        if self.config.get("streaming_read", True) and not self.config.get("host_decode", False) and \
                geotiff.read_device_plain(path, self.dev, probe=True):
            # uncompressed planar strips through a small pinned ring (2 x 32 MiB) straight into the device raster:
            # no raster-sized staging buffer to pin (0.5 - 2.5 s per GB measured, and other CUDA calls stall
            # meanwhile), the same 0.08 s per image afterwards; `streaming_read: false` keeps the staging buffer
            out = self.s.device_array(name, shape, dtype, parity, self.dev)
            stream = self.s.loader_stream(self.dev)
            with torch.cuda.stream(stream):
                res = geotiff.read_device_plain(path, self.dev, out=out, slot=2 * parity + (name != "rgbi"))
                if res is not None:
                    ready[name] = stream.record_event()
                    return out
        host = self.s.pinned_array(name, shape, dtype, parity if self.two_sets else 0)
        geotiff.read(path, out=host.numpy())
        return host

    def warm(self, fp, tiles_path):
        """Cold-start work that needs only the headers of the first image -- tile tables, the P1 plan and its
        output buffer -- done on the calling thread while the decoder thread reads that image's pixels.  Purely an
        optimisation: any failure is left to ``run``."""
        try:
            stem = Path(fp).stem
            with open(os.path.join(tiles_path, stem + ".json")) as f:
                tiles = json.load(f)
            hpath, ipath = _match_rasters(stem, self.patterns, self.config["height_data_path"],
                                          self.config["image_directory"], self.height_index, self.image_index)
            if hpath is None or ipath is None:
                return
            rinfo, hinfo = geotiff.read_info(fp), geotiff.read_info(hpath)
            key = (rinfo.height, rinfo.width, len(tiles), hinfo.height, hinfo.width)
            if key not in self.s.tables:
                tables = api.TileTables(tiles, self.dev, self.p.shift)
                self.s.tables[key] = (tables, torch.empty((tables.p1_floats,), dtype=torch.float32, device=self.dev))
                self.s.runners[key] = pipeline.ChainRunner(self.p)
                if rinfo.dtype == np.uint8:
                    tables.plan_for(1, rinfo.height, rinfo.width)
        except Exception:
            pass

    def prefetch(self, fp, tiles_path):
        if fp is not None and fp not in self.pending:
            self.pending[fp] = self.s.loader.submit(self._load, fp, tiles_path, self.parity)
            self.parity ^= 1

    def _write_stitched(self, stem, host, epsg):
        """``geojson_predictions/<stem>.gpkg`` (runs on a writer thread)"""
        t0 = time.time()
        stitched = os.path.join(self.stitched_path, stem + ".gpkg")
        gpkg.write_layer(stitched, stem, host["table_verts"], host["table_ring_off"],
                         {"Confidence_score": host["table_conf"]}, gpkg.STITCHED_SCHEMA, epsg=epsg)
        return time.time() - t0

    def _write_final(self, stem, host, epsg):
        """``processed_<stem>.gpkg`` (runs on a writer thread)"""
        t0 = time.time()
        stitched = os.path.join(self.stitched_path, stem + ".gpkg")
        processed = os.path.join(self.stitched_path, f"processed_{stem}.gpkg")
        self.logger.info(f"Processing file {stitched} with {len(host['table_conf'])} features.")
        _write_processed(host, processed, epsg, self.logger)
        self.s.processed[stitched] = processed
        return time.time() - t0

    def _reap(self, keep):
        """wait for the oldest images' writes until at most ``keep`` images are pending; a failed write un-marks its
        image"""
        while len(self.writes) > keep:
            fp, futs = self.writes.pop(0)
            for fut in futs:
                try:
                    self.s.stage_s["write"] = self.s.stage_s.get("write", 0.0) + fut.result()
                except Exception as e:
                    self.logger.error(f"Error writing the crown layers of {fp}: {e}")
                    self.failed.add(fp)

    def finish(self):
        """all layers on disk; returns the images whose layers could not be written"""
        self._reap(0)
        return self.failed

    def run(self, fp, tiles_path, next_fp):
        t_in = time.time()
        self.prefetch(fp, tiles_path)
        data = self.pending.pop(fp).result()
        t_got = time.time()
        if data is None or data["det"] is None:
            self.prefetch(next_fp, tiles_path)
            return False
        stem = Path(fp).stem
        tiles, rinfo, hinfo, det = data["tiles"], data["rinfo"], data["hinfo"], data["det"]
        # images of one shape share a tile grid: tables, the P1 output buffer, capacities and graphs are re-used
        key = (rinfo.height, rinfo.width, len(tiles), hinfo.height, hinfo.width)
        if key in self.s.tables:
            try:
                self.s.tables[key][0].retarget(tiles, self.dev, self.p.shift)   # same grid, this image's georeference
            except ops._lib.TreedetError:
                del self.s.tables[key]
        if key not in self.s.tables:
            tables = api.TileTables(tiles, self.dev, self.p.shift)
            self.s.tables[key] = (tables, torch.empty((tables.p1_floats,), dtype=torch.float32, device=self.dev))
            self.s.runners[key] = pipeline.ChainRunner(self.p)
        tables, p1_out = self.s.tables[key]
        img = api.HostImage(data["rgbi"], rinfo.transform, data["ndsm"], hinfo.transform, tiles, det["boxes_net"],
                            det["scores"], det["probs"], det["inst_tile"], det["tile_dims"], ready=data["ready"])
        t0 = time.time()
        # the decoder thread starts on the next image now and reads its pixels as soon as this image's host->device
        # copies are through (rasters decoded on the device alternate between two device buffers instead)
        self.h2d_gate.clear()

        def copies_enqueued(ev):
            self.h2d_event = ev
            self.h2d_gate.set()
        self.prefetch(next_fp, tiles_path)
        try:
            host, _ = api.run_image(img, self.p, self.dev, tables, p1_out, runner=self.s.runners[key], want_table=True,
                                    on_h2d=copies_enqueued)
        finally:
            self.h2d_gate.set()
        for path, st in data["status"]:          # run_image has synchronised: the decoder's verdict is in
            if int(st.item()) != 0:
                raise ValueError(f"{path}: corrupt LZW stream (device decoder status {int(st.item())})")
        t1 = time.time()
        if self.config.get("keep_intermediate", False) and int(det["scores"].numel()):
            raw = synth_detections(det, tiles)
            _write_predictions_json(self.output_path, stem, fp, tiles, raw, tables, self.p, self.dev)
        self._reap(self.MAX_PENDING_WRITES - 1)
        epsg = rinfo.epsg or 4326
        self.writes.append((fp, [self.s.writer.submit(self._write_stitched, stem, host, epsg),
                                 self.s.writer.submit(self._write_final, stem, host, epsg)]))
        self.s.images += 1
        st = self.s.stage_s
        for name, v in (("wait_decode", t_got - t_in), ("decode", data["decode_s"]), ("fixtures", data["fixtures_s"]),
                        ("tables", t0 - t_got), ("device", t1 - t0)):
            st[name] = st.get(name, 0.0) + v
        self.s.timeline.append({"image": stem, "t_in": t_in, "wait_decode_s": t_got - t_in, "decode_s": data["decode_s"],
                                "fixtures_s": data["fixtures_s"], "tables_s": t0 - t_got, "device_s": t1 - t0,
                                "t_out": time.time()})
        return True


def synth_detections(det, tiles):
    """Detections (host arrays) from the pinned fixture tensors of the fast path"""
    from .synth import Detections
    return Detections(det["boxes_net"].numpy(), det["scores"].numpy(), det["probs"].numpy(), det["inst_tile"].numpy(),
                      det["tile_dims"].numpy(), list(tiles.keys()), tiles)


def _write_predictions_json(output_path, stem, fp, tiles, det, tables, params, dev):
    """the per-tile ``Prediction_*.json`` wire format (prediction.py:253-263), kept intermediates only"""
    t = lambda a: a.to(dev) if torch.is_tensor(a) else torch.from_numpy(np.require(a, requirements=["C", "W"])).to(dev)
    host = lambda a: a.cpu().numpy() if torch.is_tensor(a) else a
    boxes, probs, inst_tile, tile_dims = t(det.boxes_net), t(det.probs), t(det.inst_tile), t(det.tile_dims)
    bpx, win, nwords = ops.paste_plan(boxes, inst_tile, tile_dims)
    woff = ops.exclusive_offsets(nwords)
    bits = ops.paste_threshold_pack(bpx, win, woff, probs, params.mask_threshold)
    rings = ops.trace_rings(bits, win, woff, inst_tile, tables.tile_tf)
    _write_tile_predictions(os.path.join(output_path, stem), fp, tiles, rings, host(det.inst_tile), host(det.scores))


def _load_prediction_ledger(rec_file, output_path, tiles_path, model_path, stitched_path, exclude, logger):
    """recoveries.load_prediction_recovery_data (recoveries.py:5-73), same file format
    ``{model_path, files: {image path: [tile ids]}}``.  An image counts as predicted when the ledger was written
    for the same model, its tiles JSON still exists and -- the reference's rule -- the number of per-tile
    ``Prediction_*.json`` files equals the number of (non-excluded) tiles; this build writes those files only
    with ``keep_intermediate``, so without them the stitched layer and the stitching ledger
    (recoveries.py:111-144, ``stitching_recovery.yaml``) are the evidence."""
    processed = set()
    if not os.path.exists(rec_file):
        return processed
    try:
        rec = _yaml_load(open(rec_file)) or {}
    except Exception as e:
        if logger:
            logger.warning(f"Could not load prediction recovery file: {e}")
        return processed
    if rec.get("model_path") != model_path:
        if logger:
            logger.warning("Model path does not match the one stored in the recovery file. Skipping recovery.")
        return processed
    stitched = _load_stitching_ledger(stitched_path, logger)
    files = rec.get("files")
    if files is None:                         # ledger of an older build: a flat list
        files = {f: None for f in rec.get("processed_files", [])}
    for file_path, keys in files.items():
        stem = Path(file_path).stem
        json_path = os.path.join(tiles_path, stem + ".json")
        if not os.path.exists(json_path):
            if logger:
                logger.debug(f"Missing JSON metadata for {file_path}. Skipping.")
            continue
        folder = os.path.join(output_path, stem)
        if os.path.isdir(folder) and keys is not None:
            n_files = len(os.listdir(folder))
            ok = n_files == len(keys)
            if not ok:
                with open(json_path) as jf:
                    tiles = json.load(jf)
                valid = [k for k, v in tiles.items() if not any(v.get(flag, False) for flag in (exclude or []))]
                ok = n_files == len(valid)
            if not ok:
                if logger:
                    logger.debug(f"Mismatch between output folder and JSON (after excludes) for {file_path}.")
                continue
        elif stem not in stitched:
            continue
        if os.path.exists(os.path.join(stitched_path, stem + ".gpkg")):
            processed.add(file_path)
    if processed and logger:
        logger.info(f"Skipped {len(processed)} files that were already processed.")
    return processed


def _save_prediction_ledger(rec_file, tiles_path, model_path, files, logger):
    """recoveries.save_prediction_recovery_data (recoveries.py:75-108)"""
    try:
        data = {"model_path": model_path, "files": {}}
        for file_path in files:
            json_path = os.path.join(tiles_path, Path(file_path).stem + ".json")
            if os.path.exists(json_path):
                with open(json_path) as jf:
                    data["files"][file_path] = list(json.load(jf).keys())
        with open(rec_file, "w") as f:
            _yaml_dump(data, f)
    except Exception as e:
        if logger:
            logger.warning(f"Failed to save prediction recovery file: {e}")


def _load_stitching_ledger(stitched_path, logger=None):
    """recoveries.load_stitching_recovery (recoveries.py:111-127): base names of the stitched layers"""
    rec_file = os.path.join(stitched_path, "stitching_recovery.yaml")
    done = set()
    if os.path.exists(rec_file):
        try:
            rec = _yaml_load(open(rec_file)) or {}
            done = {os.path.splitext(os.path.basename(p))[0] for p in rec.get("completed_files", [])}
        except Exception as e:
            if logger:
                logger.warning(f"Failed to load stitching recovery: {e}")
    return done


def _save_stitching_ledger(stitched_path, files, logger=None):
    """recoveries.save_stitching_recovery (recoveries.py:129-144)"""
    try:
        done = _load_stitching_ledger(stitched_path) | {Path(f).stem for f in files}
        with open(os.path.join(stitched_path, "stitching_recovery.yaml"), "w") as f:
            _yaml_dump({"completed_files": sorted(done)}, f)
    except Exception as e:
        if logger:
            logger.warning(f"Failed to save stitching recovery: {e}")


def _rings_from_prediction_files(pred_dir, tiles, logger=None):
    """The ``Prediction_<tile id>.json`` files of one image (prediction.py:253-263: ``[{image_id, category_id,
    score, polygon_coords: [ring]}]`` per tile) as one ragged ring set in the order of the tiles JSON: vertices
    (V, 2) f64, ring offsets, the tile of every ring, its confidence.  What shapely's ``Polygon(coords)`` does to
    a ring is reproduced: an open ring is closed; a tile with a ring of fewer than three points fails as a whole
    (helpers.py:419-476 logs the error and drops the file)."""
    verts, lens, ring_tile, conf = [], [], [], []
    for t, tid in enumerate(tiles.keys()):
        path = os.path.join(pred_dir, f"Prediction_{os.path.basename(tid)}.json")
        if not os.path.exists(path):
            continue
        try:
            with open(path) as f:
                data = json.load(f)
            tile_rings, tile_conf = [], []
            for crown in data:
                if "polygon_coords" not in crown:
                    raise ValueError("RLE-encoded predictions (pycocotools) are not supported")
                xy = np.asarray(crown["polygon_coords"], dtype=np.float64).reshape(-1, 2)
                if len(xy) and (xy[0] != xy[-1]).any():
                    xy = np.concatenate([xy, xy[:1]])
                if len(xy) < 4:
                    raise ValueError("A LinearRing must have at least 3 coordinate tuples")
                tile_rings.append(xy)
                tile_conf.append(float(crown["score"]))
        except Exception as e:
            if logger:
                logger.warning(f"Error processing file {path}: {e}")
            continue
        verts += tile_rings
        lens += [len(r) for r in tile_rings]
        ring_tile += [t] * len(tile_rings)
        conf += tile_conf
    off = np.zeros(len(lens) + 1, dtype=np.int64)
    off[1:] = np.cumsum(lens)
    v = np.concatenate(verts) if verts else np.zeros((0, 2), dtype=np.float64)
    return v, off, np.asarray(ring_tile, dtype=np.int32), np.asarray(conf, dtype=np.float64)


def process_and_stitch_predictions(tiles_path, pred_fold, output_path, max_workers=50, shift=1, simplify_tolerance=0.2,
                                   logger=None, device="0"):
    """helpers.process_and_stitch_predictions (helpers.py:556-600) for predictions that already exist as
    ``<pred_fold>/<image>/Prediction_<tile id>.json`` -- e.g. written by the reference's own detectron2 run, or by
    this package with ``keep_intermediate``: every ring is simplified (``simplify_tolerance``, GEOS semantics) and
    kept when it lies within its tile's box shrunk by ``shift`` (P4 on the device, ``td_simplify_rings``), the
    survivors of an image go to ``<output_path>/<image>.gpkg`` (``Confidence_score`` + geometry).  Rows are in the
    order of the tiles JSON (the reference concatenates in directory-listing order).  Images listed in
    ``stitching_recovery.yaml`` are skipped and the ledger is updated, as in the reference."""
    for pth, what in ((tiles_path, "tiles path"), (pred_fold, "prediction folder")):
        if not os.path.isdir(pth):
            raise FileNotFoundError(f"The {what} '{pth}' does not exist.")
    os.makedirs(output_path, exist_ok=True)
    done = _load_stitching_ledger(output_path, logger)
    folders = sorted(f for f in os.listdir(tiles_path) if f.endswith(".json") and os.path.isfile(os.path.join(tiles_path, f)))
    todo = [f for f in folders if os.path.splitext(f)[0] not in done]
    if logger and len(todo) < len(folders):
        logger.info(f"Skipping stiching {len(folders) - len(todo)} of {len(folders)} folders that have already been processed.")
    dev = _device({"device": device})
    results = []
    with torch.cuda.device(dev):
        for k, folder in enumerate(todo):
            stem = os.path.splitext(folder)[0]
            try:
                with open(os.path.join(tiles_path, folder)) as f:
                    tiles = json.load(f)
                v, off, ring_tile, conf = _rings_from_prediction_files(os.path.join(pred_fold, stem), tiles, logger)
                epsg = next(iter(tiles.values()))["crs"] if tiles else 4326
                if len(off) > 1:
                    _, boxes_int = pipeline.tile_tables(tiles, dev)
                    d_v, d_off = torch.from_numpy(v).to(dev), torch.from_numpy(off).to(dev)
                    simp = ops.simplify_rings(d_v, d_off, float(simplify_tolerance), pipeline.filter_boxes(boxes_int, shift, dev),
                                              torch.from_numpy(ring_tile).to(dev), want_bounds=False)
                    sel = torch.nonzero(simp["keep"]).flatten()
                    kv, koff = ops.take_rings(d_v, d_off, sel, simp["scratch"], simp["count"])
                    v, off, conf = kv.cpu().numpy(), koff.cpu().numpy(), conf[sel.cpu().numpy()]
                digits = re.search(r"(\d+)$", str(epsg))
                gpkg.write_layer(os.path.join(output_path, stem + ".gpkg"), stem, v, off, {"Confidence_score": conf},
                                 gpkg.STITCHED_SCHEMA, epsg=int(digits.group(1)) if digits else 4326)
                results.append(stem)
            except Exception as e:
                if logger:
                    logger.error(f"Error processing folder {folder}: {e}")
            if logger and todo and (k == 0 or (100 * (k + 1) // len(todo)) // 5 != (100 * k // len(todo)) // 5):
                logger.info(f"Stitching file {k + 1}/{len(todo)} ({100 * (k + 1) // len(todo)}%)")
    _save_stitching_ledger(output_path, sorted(done | set(results)), logger)
    return output_path


def predict_on_model(config, model_path, tiles_path, output_path, batch_size=10, exclude_vars=None,
                     stitched_path=None):
    """detection.py:62-132 + helpers.process_and_stitch_predictions (helpers.py:556-600) in one
    device pass per image: raw ROI-head outputs -> ``<stitched_path>/<stem>.gpkg``."""
    with torch.cuda.device(_device(config)):
        return _predict_on_model(config, model_path, tiles_path, output_path, batch_size, exclude_vars, stitched_path)


def _predict_on_model(config, model_path, tiles_path, output_path, batch_size, exclude_vars, stitched_path):
    logger = config.get("logger", None)
    for path, name in [(model_path, "Model file"), (tiles_path, "Tiles directory")]:
        if not os.path.exists(path):
            raise FileNotFoundError(f"{name} not found: {path}")
        if name == "Tiles directory" and not os.path.isdir(path):
            raise NotADirectoryError(f"{name} is not a directory: {path}")
    os.makedirs(output_path, exist_ok=True)
    stitched_path = stitched_path or os.path.join(config["output_directory"], "geojson_predictions")
    os.makedirs(stitched_path, exist_ok=True)
    predictor = config.get("predictor") or FixturePredictor(model_path, exclude_vars)
    dev = _device(config)
    params = pipeline.PipelineParams.from_config(config)

    images_directory = Path(config["image_directory"])
    images_paths = sorted(str(f) for f in images_directory.glob("*.tif"))
    merged_directory = Path(f"{images_directory}/{config['merged_path']}")
    images_paths.extend(sorted(str(f) for f in merged_directory.glob("*.tif")))
    if not images_paths:
        logger.warning("No TIF files found for prediction.")
        return

    rec_file = os.path.join(output_path, "prediction_recovery.yaml")
    processed = _load_prediction_ledger(rec_file, output_path, tiles_path, model_path, stitched_path, exclude_vars,
                                        logger)
    images_paths = [f for f in images_paths if f not in processed]
    if not images_paths:
        logger.info("All files have already been predicted. Exiting Prediction.")
        return

    total = len(images_paths)
    done = []
    session = config.get("_session")
    # the fast path (see _Session) applies to a single-model run whose crown layers go straight to
    # post-processing; with two models the fusion sits between the stages
    fast = session is not None and exclude_vars is None and \
        os.path.abspath(stitched_path) == os.path.abspath(os.path.join(config["output_directory"], "geojson_predictions"))
    if fast:
        fast_ctx = _FastPath(config, session, dev, params, predictor, stitched_path, output_path, logger,
                             n_images=total)
        fast_ctx.prefetch(images_paths[0], tiles_path)
        fast_ctx.warm(images_paths[0], tiles_path)
    for i, fp in enumerate(images_paths):
        cur, prev = int(100 * (i + 1) / total), int(100 * i / total)
        if logger and ((cur // 5) != (prev // 5) or cur == 100 or i == 0):
            logger.info(f"Predicting file {i + 1}/{total} ({cur}%)")
        try:
            stem = Path(fp).stem
            if fast:
                nxt = images_paths[i + 1] if i + 1 < total else None
                if fast_ctx.run(fp, tiles_path, nxt):
                    done.append(fp)
                    continue
                session.fallback_images += 1
            tile_json = os.path.join(tiles_path, stem + ".json")
            with open(tile_json) as f:
                tiles = json.load(f)
            tables = api.TileTables(tiles, dev, params.shift)
            if getattr(predictor, "wants_tiles", False):
                # P1: normalised tiles for a live model adapter, consumed on the device
                img, info = geotiff.read(fp)
                d_img = torch.from_numpy(img.view(np.int16) if img.dtype == np.uint16 else img).to(dev)
                tiles_dev, tiles_off, flags = tables.plan(d_img).run(d_img)
                det = predictor.forward(stem, tiles, tiles_dev, tiles_off, flags)
                del d_img, img
            else:       # fixtures replay: the raster itself is not needed here, only its CRS
                info = geotiff.read_info(fp)
                det = predictor.raw_outputs(stem, tiles)
            t = lambda a: a.to(dev) if torch.is_tensor(a) else torch.from_numpy(np.require(a, requirements=["C", "W"])).to(dev)
            boxes, scores, probs, inst_tile, tile_dims = (t(det.boxes_net), t(det.scores), t(det.probs),
                                                          t(det.inst_tile), t(det.tile_dims))
            host = lambda a: a.cpu().numpy() if torch.is_tensor(a) else a
            if config.get("keep_intermediate", False) and len(det.scores):
                _write_predictions_json(output_path, stem, fp, tiles, det, tables, params, dev)
            table = pipeline.predict_stage(boxes, scores, probs, inst_tile, tile_dims, tables.tile_tf,
                                           tables.tile_boxes, params)
            gpkg.write_layer(os.path.join(stitched_path, stem + ".gpkg"), stem, table.verts.cpu().numpy(),
                             table.ring_off.cpu().numpy(), {"Confidence_score": table.conf.cpu().numpy()},
                             gpkg.STITCHED_SCHEMA, epsg=info.epsg or 4326)
            done.append(fp)
        except Exception as e:   # per-file failures are logged and skipped (detection.py:117-120)
            logger.error(f"Error processing {fp}: {e}")
    if fast:
        failed = fast_ctx.finish()
        done = [f for f in done if f not in failed]
    logger.info(f"Completed prediction for {len(images_paths)} images.")
    _save_prediction_ledger(rec_file, tiles_path, model_path, sorted(processed | set(done)), logger)
    _save_stitching_ledger(stitched_path, sorted(processed | set(done)), logger)


def predict_tiles(config):
    config_obj = Config()
    config_obj._load_into_config(config)
    logger = config["logger"]
    out = config["output_directory"]
    if "urban_model" in config and "forrest_model" in config and "forrest_outline" in config and \
            config["urban_model"] and os.path.exists(config["urban_model"]) and \
            config["forrest_model"] and os.path.exists(config["forrest_model"]) and \
            config["forrest_outline"] and os.path.exists(config["forrest_outline"]):
        from .fusion import fuse_predictions
        logger.info("Urban, forrest models and forrest outline are available. Starting prediction...")
        urban_fold = os.path.join(out, "urban_geojson")
        forrest_fold = os.path.join(out, "forrest_geojson")
        t0 = time.time()
        predict_on_model(config, config["urban_model"], config["tiles_path"], os.path.join(out, "urban_predictions"),
                         batch_size=config["batch_size"], exclude_vars=["only_forest"], stitched_path=urban_fold)
        t1 = time.time()
        predict_on_model(config, config["forrest_model"], config["tiles_path"], os.path.join(out, "forrest_predictions"),
                         batch_size=config["batch_size"], exclude_vars=["only_urban"], stitched_path=forrest_fold)
        t2 = time.time()
        logger.info("Predictions have been processed and stitched. Begin fusing the predictions.")
        with torch.cuda.device(_device(config)):
            fuse_predictions(urban_fold, forrest_fold, config["forrest_outline"], os.path.join(out, "geojson_predictions"),
                             logger=logger, device=_device(config))
        t3 = time.time()
        logger.info("Fusion based on forest outline has been completed.")
        logger.debug(f"predict + stitch for urban took {t1 - t0} seconds")
        logger.debug(f"predict + stitch for forrest took {t2 - t1} seconds")
        logger.debug(f"fuse prediction took {t3 - t2} seconds")
    elif "combined_model" in config and config["combined_model"] and os.path.exists(config["combined_model"]):
        logger.info("Only Combined Model is given. Starting prediction...")
        t0 = time.time()
        predict_on_model(config, config["combined_model"], config["tiles_path"], os.path.join(out, "predictions"),
                         batch_size=config["batch_size"], stitched_path=os.path.join(out, "geojson_predictions"))
        logger.info("Predictions have been processed and stitched. Begin fusing the predictions.")
        logger.debug(f"Prediction + stitching took {time.time() - t0} seconds")
    else:
        raise FileNotFoundError("No model available for prediction. Either urban model or forrest model + outline or "
                                "combined model must be available.")


# --------------------------------------------------------------------------------------
# post-processing side (P5-P9)
# --------------------------------------------------------------------------------------
def _build_file_index(directory, pattern, ending=".tif"):
    idx = {}
    for f in sorted(os.listdir(directory)):
        if f.endswith(ending):
            m = pattern.search(os.path.basename(f))
            if m:
                idx["".join(m.groups())] = os.path.join(directory, f)
    return idx


def _find_matching_file(base_name, geojson_pattern, search_pattern, directory, index=None):
    """postprocessing.py:1001-1018."""
    m = geojson_pattern.match(base_name + ".tif")
    if m:
        groups = m.groups()
        concat = "".join(groups)
        if index is not None and concat in index:
            return index[concat]
        for root, _, files in os.walk(directory):
            for file in sorted(files):
                sm = search_pattern.match(file)
                if sm and "".join(sm.groups()[:len(groups)]) == concat:
                    return os.path.join(root, file)
    return None


def _match_rasters(base_name, patterns, height_directory, image_directory, height_index=None, image_index=None):
    """the nDSM / RGBI rasters of a crown layer (postprocessing.py:1030-1049): plain images by the configured
    regexes, seam strips by the merged ones"""
    image_pattern, height_pattern, image_merged_pattern, height_merged_pattern = patterns
    hpath = _find_matching_file(base_name, image_pattern, height_pattern, height_directory, height_index)
    ipath = _find_matching_file(base_name, image_pattern, image_pattern, image_directory, image_index)
    if hpath is None or ipath is None:
        hpath = _find_matching_file(base_name, image_merged_pattern, height_merged_pattern, height_directory)
        ipath = _find_matching_file(base_name, image_merged_pattern, image_merged_pattern, image_directory)
    return hpath, ipath


_RECOVERY_KEYS = ("confidence_threshold", "containment_threshold", "height_threshold", "ndvi_mean_threshold",
                  "ndvi_var_threshold", "iou_threshold", "area_threshold", "ndvi_scaling_factor",
                  "height_scaling_factor", "use_overlap")


def _write_processed(h, processed_file_path, epsg, logger=None):
    """``processed_<image>.gpkg`` (schema postprocessing.py:904-919) from the host crown layer"""
    columns = {
        "Confidence_score": h["conf"], "poly_id": [str(int(v)) for v in h["poly_id"]], "Area": h["area"],
        "TreeHeight": h["tree_height"],
        "Centroid": [json.dumps({"x": float(c[0]), "y": float(c[1])}) for c in h["centroid"]],
        "Diameter": [2 * (float(a) / np.pi) ** 0.5 for a in h["area"]],
        "is_contained": [str(bool(v)) for v in h["is_contained"]], "num_contained": h["num_contained"],
    }
    layer = Path(processed_file_path).stem
    gpkg.write_layer(processed_file_path, layer, h["verts"], h["ring_off"], columns, gpkg.PROCESSED_SCHEMA, epsg=epsg)
    if logger:
        logger.debug(f" File {os.path.basename(processed_file_path)}, # crowns {len(h['poly_id'])} ")


def process_single_file(file_path, processed_file_path, height_data_path, rgbi_data_path, config, dev):
    """postprocessing.py:876-943 with process_geojson / process_features on the device."""
    logger = config["logger"]
    try:
        params = pipeline.PipelineParams.from_config(config)
        verts, off, cols, epsg = gpkg.read_layer(file_path)
        logger.info(f"Processing file {file_path} with {len(off) - 1} features.")
        conf = np.array([np.nan if c is None else float(c) for c in cols.get("Confidence_score", [])], dtype=np.float64)
        conf = np.where(np.isnan(conf), -np.inf, conf)      # "Confidence_score is not None" (postprocessing.py:741)
        table = pipeline.CrownTable(torch.from_numpy(np.ascontiguousarray(verts)).to(dev),
                                    torch.from_numpy(off).to(dev), torch.from_numpy(conf).to(dev))
        ndsm, hinfo = geotiff.read(height_data_path)
        rgbi, rinfo = geotiff.read(rgbi_data_path)
        if rgbi.dtype != np.uint8:
            raise ValueError("NDVI path expects 8-bit RGBI rasters")
        rasters = pipeline.raster_stage(torch.from_numpy(rgbi).to(dev), rinfo.transform,
                                        torch.from_numpy(np.ascontiguousarray(ndsm[0], dtype=np.float32)).to(dev),
                                        hinfo.transform, params)
        feats = pipeline.postprocess_stage(table, rasters, params)
        h = api.features_to_host(feats)
        _write_processed(h, processed_file_path, epsg or rinfo.epsg or 4326, logger)
        return file_path
    except Exception as e:
        print(f"Error postprocessing file {file_path}: {e}")
        return None


def process_files_in_directory(directory, height_directory, image_directory, config, parallel=True,
                               filename_pattern=None):
    """postprocessing.py:945-1076 (files are independent; one device pass each)."""
    with torch.cuda.device(_device(config)):
        return _process_files_in_directory(directory, height_directory, image_directory, config, filename_pattern)


def _process_files_in_directory(directory, height_directory, image_directory, config, filename_pattern):
    logger = config["logger"]
    dev = _device(config)
    rec_file = os.path.join(directory, "recovery.yaml")
    params_now = {k: config.get(k) for k in _RECOVERY_KEYS}
    processed_files = set()
    if os.path.exists(rec_file):
        try:
            rec = _yaml_load(open(rec_file)) or {}
            if rec.get("params") == params_now:
                processed_files = set(rec.get("processed_files", []))
        except Exception:
            processed_files = set()
    files = sorted(f for f in os.listdir(directory) if f.endswith(".gpkg") and not f.startswith("processed_"))
    files = [f for f in files if os.path.join(directory, f) not in processed_files]
    image_pattern, height_pattern = filename_pattern or ("(\\d+)\\.tif", "(\\d+)\\.tif")
    image_pattern = re.compile(image_pattern or "(\\d+)\\.tif")
    height_pattern = re.compile(height_pattern or "(\\d+)\\.tif")
    image_merged_pattern = re.compile(config["image_merged_regex"])
    height_merged_pattern = re.compile(config["height_data_merged_regex"])
    image_index = _build_file_index(image_directory, image_pattern)
    height_index = _build_file_index(height_directory, height_pattern)
    session = config.get("_session")
    for filename in files:
        file_path = os.path.join(directory, filename)
        base_name = os.path.splitext(os.path.basename(filename))[0]
        done = session.processed.get(file_path) if session is not None else None
        if done and os.path.exists(done):        # post-processed while its rasters were on the device (fast path)
            logger.info(f"Processing file {file_path}: already post-processed in the prediction pass.")
            processed_files.add(file_path)
            continue
        hpath, ipath = _match_rasters(base_name, (image_pattern, height_pattern, image_merged_pattern,
                                                  height_merged_pattern), height_directory, image_directory,
                                      height_index, image_index)
        if hpath and ipath:
            res = process_single_file(file_path, os.path.join(directory, f"processed_{filename}"), hpath, ipath, config,
                                      dev)
            if res is not None:
                processed_files.add(res)
        else:
            logger.warning(f"Height data file not found for: {filename}, searched pattern for base name: {base_name}")
    try:
        with open(rec_file, "w") as f:
            yaml.safe_dump({"params": params_now, "processed_files": sorted(processed_files)}, f, sort_keys=False)
    except Exception as e:
        logger.warning(f"Failed to save recovery file: {e}")


def exclude_outlines(config, logger=None):
    """helpers.py:33-69: every crown of the already existing ``processed_*`` layers that lies WITHIN the union of
    an exclusion outline (water, buildings) is removed and the layer is rewritten.  The reference clips the
    outline to the layer's bounds before the test (no effect on the predicate) and calls it before the
    post-processing of the current run; the ``within`` test runs on the device (td_forest_predicates)."""
    from .fusion import ForestIndex
    pred_dir = os.path.join(config["output_directory"], "geojson_predictions")
    for outline in config.get("exclude_files", []) or []:
        try:
            index = ForestIndex.from_file(outline, _device(config), logger)
        except Exception as e:
            msg = f"Failed to read exclude file '{outline}': {e}"
            logger.error(msg) if logger else print(msg)
            continue
        if not os.path.isdir(pred_dir):
            continue
        for file in sorted(os.listdir(pred_dir)):
            if not (file.endswith(".geojson") or file.endswith(".gpkg")) or not file.startswith("processed_"):
                continue
            file_path = os.path.join(pred_dir, file)
            try:
                verts, off, cols, epsg = gpkg.read_layer(file_path)
                if len(off) <= 1:
                    continue
                _, within = index.predicates(torch.from_numpy(np.ascontiguousarray(verts)).to(index.device),
                                             torch.from_numpy(off).to(index.device))
                keep = np.nonzero(within.cpu().numpy() == 0)[0]
                lens = np.diff(off)[keep]
                new_off = np.zeros(len(keep) + 1, dtype=np.int64)
                new_off[1:] = np.cumsum(lens)
                new_verts = np.concatenate([verts[off[k]:off[k + 1]] for k in keep]) if len(keep) else np.zeros((0, 2))
                new_cols = {k: [v[i] for i in keep] for k, v in cols.items()}
                schema = {k: gpkg.PROCESSED_SCHEMA.get(k, "str") for k in cols}
                gpkg.write_layer(file_path, Path(file_path).stem, new_verts, new_off, new_cols, schema, epsg=epsg or 4326)
            except Exception as e:
                msg = f"Error processing file '{file_path}': {e}."
                logger.error(msg) if logger else print(msg)


def postprocess_files(config):
    config_obj = Config()
    config_obj._load_into_config(config)
    logger = config["logger"]
    logger.info("Postprocessing the predictions.")
    filename_pattern = (config.get("image_regex", "(\\d+)\\.tif"), config.get("height_data_regex", "(\\d+)\\.tif"))
    logger.info("Excluding Outlines.")
    if config.get("exclude_files"):
        with torch.cuda.device(_device(config)):
            exclude_outlines(config, logger)
    pred_dir = os.path.join(config["output_directory"], "geojson_predictions")
    process_files_in_directory(pred_dir, config["height_data_path"], config["image_directory"], config,
                               parallel=config["parallel"], filename_pattern=filename_pattern)
    timestamp = datetime.datetime.now().strftime("%Y-%m-%d_%H-%M-%S")
    for file in sorted(os.listdir(pred_dir)):
        if not (file.endswith(".geojson") or file.endswith(".gpkg")) or not file.startswith("processed_"):
            continue
        name = file.replace("processed_", "")
        src = os.path.join(pred_dir, file)
        if config["timestamped_output_directory"]:
            ts_dir = f"{config['output_directory']}/{timestamp}"
            os.makedirs(ts_dir, exist_ok=True)
            shutil.copyfile(src, os.path.join(ts_dir, name))
        shutil.copyfile(src, os.path.join(config["output_directory"], name))


def process_files(config):
    logger = config["logger"]
    config_obj = Config()
    config_obj._load_into_config(config)
    # the stages share device state for the duration of this call (see _Session); a two-model run needs the
    # fusion between prediction and post-processing and takes the stage-by-stage path
    session = config.get("_session")
    own_session = session is None and not config.get("no_session", False)
    if own_session:
        session = config["_session"] = _Session()
    try:
        t0 = time.time()
        preprocess_files(config)
        t1 = time.time()
        predict_tiles(config)
        t2 = time.time()
        postprocess_files(config)
        t3 = time.time()
        cleanup_files(config)
        if session is not None:
            session.stage_s.update(preprocess=t1 - t0, predict=t2 - t1, postprocess=t3 - t2, cleanup=time.time() - t3)
            config["_last_session_stats"] = {"stage_s": dict(session.stage_s), "images": session.images,
                                             "fallback_images": session.fallback_images,
                                             "device_decoded_rasters": session.device_decoded,
                                             "timeline": list(session.timeline), "t0": t0}
    finally:
        if own_session:
            session.close()
            config.pop("_session", None)
    logger.debug(f"preprocess step took {t1 - t0} seconds. ")
    logger.debug(f"predict step took {t2 - t1} seconds. ")
    logger.debug(f"postprocess step took {t3 - t2} seconds. ")


def cleanup_files(config):
    if not config.get("keep_intermediate", False):
        for d in (config["tiles_path"], config["image_directory"] + "/" + config["merged_path"],
                  config["height_data_path"] + "/" + config["merged_path"]):
            try:
                shutil.rmtree(d)
            except FileNotFoundError:
                pass
        for directory in (config["image_directory"], config["height_data_path"]):
            for file in os.listdir(directory):
                if "__" in file:
                    os.remove(os.path.join(directory, file))
    for folder in os.listdir(config["output_directory"]):
        folder = os.path.join(config["output_directory"], folder)
        if os.path.isdir(folder) and os.path.basename(folder) not in ["logs"] and not config.get("keep_intermediate",
                                                                                                   False):
            shutil.rmtree(folder)
