#!/usr/bin/env python
"""bench.py -- post-model crown pipeline throughput (km^2/s) on synthetic orthophoto + nDSM
mosaics (BASELINE.json: configs[1], "synthetic 10k x 10k px RGB + nDSM orthophoto, single
model, tile size/overlap from example/config.yml, 1 B200").

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--size PX]

A step = one pass of the hot path (P1 tile cut/normalise, P2 paste+pack, P3 contours, P4
stitch, P5 NDVI/decimation, P6 NMS, P7 crown stats, P8 containment, P9 selection) over one
image per GPU.  `value` has inputs resident in HBM; `e2e` goes through
treedetection_b200.api.run_image with pinned HOST buffers (H2D + D2H inside the timed
region).  Under torchrun every rank owns its own image (row-sharded mosaic, weak scaling).
`--impl reference` times the CPU restatement of the reference (oracle/port.py) on the
host cores.
"""
from __future__ import annotations

import argparse
import json

import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "post-model crown pipeline throughput"
UNIT = "km^2/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size", type=int, default=10000, help="image side in pixels (0.2 m)")
    ap.add_argument("--ndsm-px", type=float, default=0.2, help="nDSM pixel size (0.2: split stats path, 1.0: combined)")
    ap.add_argument("--cpu-sample", type=int, default=1500, help="side (px) of the sub-scene of the literal CPU port")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-clocks", action="store_true", help="do not poll nvidia-smi during the timed region")
    ap.add_argument("--no-merged", action="store_true", help="skip the secondary crowns-merged/s measurement")
    ap.add_argument("--images-per-rank", type=int, default=1,
                    help="images per GPU and step (replicas of the rank's image with their right-seam strips between "
                         "them and, N > 1, one down-seam strip each towards the next rank); 8 at --gpus 8 is the full "
                         "BASELINE config 5 (64 images, 56 right + 56 down strips)")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3],
                    help="BASELINE.json config: 2 = single model 10k x 10k (the headline), 3 = two models + forest outline "
                         "on a 20k x 20k mosaic (one extra line, 1 GPU)")
    ap.add_argument("--no-files", action="store_true", help="skip e2e_files (process_files on GeoTIFFs on tmpfs)")
    ap.add_argument("--files-only", action="store_true", help="diagnostic: run e2e_files alone and print its entry")
    ap.add_argument("--files-images", type=int, default=6, help="images of the workload in e2e_files")
    ap.add_argument("--chains", type=int, default=1,
                    help="independent chain contexts (workspace + stream + graphs each) that consecutive images "
                         "alternate between: the P2-P9 chain of an image is a latency-bound sequence of dependent "
                         "launches, two images' chains side by side fill the gaps of each other")
    ap.add_argument("--p1-priority", type=int, default=0, help="stream priority of P1 (0 or -1)")
    ap.add_argument("--chain-priority", type=int, default=-1, help="stream priority of the chain contexts (0 or -1)")
    ap.add_argument("--join-steps", action="store_true",
                    help="join all streams after every image (default: the streams run free between the two ends of "
                         "the timed region; images are independent)")
    ap.add_argument("--no-alone", action="store_true",
                    help="do not time the roofline kernel alone after the timed region (profiling runs: keeps the "
                         "launch list to whole steps); the roofline entry then uses the in-step time")
    ap.add_argument("--exact", action="store_true",
                    help="exact-size chain (a host synchronisation before every allocation) instead of the "
                         "sync-free capacity-buffer chain")
    ap.add_argument("--serial", action="store_true",
                    help="one stream: P1 and the P2-P9 chain back to back (default: P1 on its own stream, "
                         "overlapping the latency-bound chain)")
    return ap.parse_args()


# --------------------------------------------------------------------------------------
# CPU arm: the oracle restatement of the reference, on a bounded sample of the workload
# --------------------------------------------------------------------------------------
_SCENES = {}


def _cpu_scene(seed, size_px, ndsm_px):
    key = (seed, size_px, ndsm_px)
    if key not in _SCENES:
        from treedetection_b200 import synth
        _SCENES[key] = synth.make_scene(seed=seed, size_px=size_px, px=0.2, ndsm_px=ndsm_px, density_per_km2=2500.0)
    return _SCENES[key]


def workload_string(size, ndsm_px):
    """config.workload of BOTH arms (identical strings: the driver compares them)"""
    return (f"synthetic {size}x{size} px RGBI + nDSM ({ndsm_px} m) orthophoto per GPU, single model, "
            f"tile 50 m / buffer 20 m, 2500 trees/km^2, ROI-head outputs replayed from fixtures")


def infeasible_note(n_candidates, n_crowns, raster_px):
    """why the reference's literal loops cannot run the full workload"""
    return (f"the reference's literal post-processing is O(N^2) + O(N x P): at the full workload its NMS builds "
            f"{n_candidates}^2 float32 matrices ({n_candidates ** 2 * 4 / 1e9:.1f} GB each, ~8 temporaries) and its "
            f"statistics touch {n_crowns} crowns x {raster_px:.1e} pixels x ~40 passes "
            f"({n_crowns * raster_px * 4 * 40 / 1e12:.0f} TB of traffic): infeasible, hence bounded sub-scenes "
            f"for the literal form and the windowed / sparse forms (same results, tests/test_oracle_windowed.py) "
            f"for the full workload")


def _cpu_sample_once(args):
    """One pass of the reference's path (restated, oracle/port.py) over one sample scene.  The scene
    is synthesised once per process and re-used; only the path is timed.  args = (seed, size_px, ndsm_px
    [, large]): ``large`` selects the windowed statistics / sparse NMS / chunked containment (identical
    results, tests/test_oracle_windowed.py) instead of the reference's literal N x P / N x N loops."""
    seed, size_px, ndsm_px = args[:3]
    large = bool(args[3]) if len(args) > 3 else False
    import numpy as np
    from oracle import port
    from treedetection_b200 import geo, pipeline
    sc = _cpu_scene(seed, size_px, ndsm_px)
    p = pipeline.PipelineParams()
    cfg = {k: getattr(p, k) for k in p.__dataclass_fields__}
    t0 = time.perf_counter()
    for meta in sc.tiles.values():                                   # P1
        port.tile_cut_normalize(sc.rgbi, tuple(meta["window"]))
    rings, conf = port.predict_stage(sc.det, sc.tiles, paste="torch")   # P2-P4
    H, W = sc.rgbi.shape[1:]
    oh, ow = int(H * p.ndvi_scaling_factor), int(W * p.ndvi_scaling_factor)
    dec = np.stack([port.decimate_bilinear(sc.rgbi[b], oh, ow) for b in (0, 3)])        # P5
    ndvi = port.ndvi_from_rgbi(np.stack([dec[0], dec[0], dec[0], dec[1]])).astype(np.float32)
    ndvi_tf = geo.compose(sc.transform, geo.scale(W / ow, H / oh))
    h, w = sc.ndsm.shape
    out, _ = port.post_process(rings, conf, ndvi, ndvi_tf, tuple(geo.raster_bounds(sc.transform, W, H)), sc.ndsm,
                               sc.ndsm_transform, tuple(geo.raster_bounds(sc.ndsm_transform, w, h)), 0.2, 0.2, cfg,
                               large=large)
    dt = time.perf_counter() - t0
    return dt, sc.area_km2, len(rings), len(out)


def _summarise(res, workers):
    area = sum(r[1] for r in res)
    wall = max(r[0] for r in res)        # concurrent workers finish together: the slowest one is the wall
    return {"wall_s": wall, "area_km2": area, "km2_per_s_wall": area / wall, "rings": res[0][2], "crowns": res[0][3]}


def cpu_rate(sample_px, ndsm_px, workers, repeats=1, large=False):
    """km^2/s of the CPU restatement on one sample scene, one process (the cpu_baseline of the b200 arm)."""
    r = _summarise([_cpu_sample_once((1234, sample_px, ndsm_px, large))], 1)
    return r


def run_reference(a):
    """The reference's CPU path (its restatement oracle/port.py: the reference itself needs CuPy,
    detectron2, rasterio and shapely, none of which exist offline) on all host cores.  A step = every
    worker process runs the path once over its own copy of a bounded sample scene of the workload
    (independent images are how the reference parallelises: ThreadPoolExecutor over files).  The sample
    side is chosen from a first calibration pass so that warmup + steps end within ~2.5 minutes; the
    reference's statistics are O(crowns x pixels), so its km^2/s falls with the sample size -- the line
    states the sample it was measured on."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp
    cores = max(1, min(os.cpu_count() or 1, 32))
    total = a.warmup + a.steps
    budget_s = 150.0
    with mp.get_context("spawn").Pool(cores) as pool:
        def one_step(side):
            return _summarise(pool.map(_cpu_sample_once, [(1234, side, a.ndsm_px)] * cores, chunksize=1), cores)
        # calibration on a small sample (also pays the imports and page faults of every worker)
        cal_side = 600
        one_step(cal_side)
        t_cal = one_step(cal_side)["wall_s"]
        side = cal_side
        for cand in (800, 1000, 1250, 1500, 2000, 2500, 3000):
            est = t_cal * (cand / cal_side) ** 3.5      # measured growth between 600 and 3000 px
            if cand <= a.size and est * total <= budget_s:
                side = cand
        times, area, r = [], 0.0, None
        for step in range(total):
            r = one_step(side)
            if step >= a.warmup:
                times.append(r["wall_s"])
                area += r["area_km2"]
        # the reference's own post-processing concurrency: ThreadPoolExecutor(max_workers=5) over files
        # (postprocessing.py:1051) -- one extra step with 5 concurrent workers
        five = _summarise(pool.map(_cpu_sample_once, [(1234, side, a.ndsm_px)] * min(5, cores), chunksize=1), 5)
    steps = len(times)
    value = area / max(sum(times), 1e-9)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "mixed: u8/f32/f64 (CPU)", "data": "synthetic",
        "config": {"workload": workload_string(a.size, a.ndsm_px),
                   "sample": f"{cores} x ({side}x{side} px sub-scene) per step",
                   "five_workers": {"value": five["km2_per_s_wall"], "unit": UNIT,
                                    "note": "5 concurrent workers, the reference's ThreadPoolExecutor(max_workers=5)"},
                   "full_workload": infeasible_note(31200, 9513, float(a.size) ** 2) if a.size == 10000 else None},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{cores} processes x one {side}x{side} px scene ({r['rings']} candidate rings each) "
                                   f"per step, P1-P9 restated in oracle/port.py; sample side chosen so that "
                                   f"{total} steps fit {budget_s:.0f} s (the reference's statistics are "
                                   f"O(crowns x pixels): km^2/s depends on the sample size)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------
# e2e_files: the reference-facing API on GeoTIFFs (process_files), per-stage wall clock
# --------------------------------------------------------------------------------------
def e2e_files(sc, n_images, label, compression=None):
    """``process_files(config)`` -- the call a user of the reference makes -- over ``n_images`` GeoTIFF pairs on
    tmpfs: RGBI + nDSM rasters written uncompressed, ROI-head fixtures staged as the "model", config.yml as in
    example/config.yml.  The images are copies of the scene at origins 10 km apart (no neighbours, so no
    seam strips; use_overlap stays on).  Returns km^2/s over the whole call and the per-stage seconds."""
    import shutil
    import tempfile

    import numpy as np
    import yaml

    from treedetection_b200 import detection, geotiff, predictor, synth, tiling
    base = "/dev/shm" if os.path.isdir("/dev/shm") else None
    root = tempfile.mkdtemp(prefix="treedet_e2e_", dir=base)
    try:
        img_dir, h_dir, model = (os.path.join(root, d) for d in ("rgb", "ndsm", "model"))
        for d in (img_dir, h_dir, model):
            os.makedirs(d)
        H, W = sc.rgbi.shape[1:]
        h, w = sc.ndsm.shape
        px, npx = sc.px, abs(sc.ndsm_transform[0])
        t_w0 = time.perf_counter()
        for k in range(n_images):
            stem = f"FDOP20_{k:06d}_rgbi"
            left = synth.ORIGIN_X + 10000.0 * k
            top = synth.ORIGIN_Y + H * px
            tf = synth.image_transform(left, top, px)
            geotiff.write(os.path.join(img_dir, stem + ".tif"), sc.rgbi, tf, epsg=synth.EPSG, compression=compression,
                          predictor=2 if compression else 1)
            geotiff.write(os.path.join(h_dir, f"nDSM_{k:06d}_1km.tif"), sc.ndsm, synth.image_transform(left, top, npx),
                          epsg=synth.EPSG, nodata=-3.4028234663852886e38, compression=compression)
            tiles = tiling.tile_grid(stem, tf, W, H, synth.EPSG, 50, 50, 20)
            d = sc.det
            predictor.dump_fixtures(model, stem, synth.Detections(d.boxes_net, d.scores, d.probs, d.inst_tile, d.tile_dims,
                                                                   list(tiles.keys()), tiles))
        write_s = time.perf_counter() - t_w0
        file_bytes = sum(os.path.getsize(os.path.join(d, f)) for d in (img_dir, h_dir) for f in os.listdir(d))
        cfg = {
            "image_directory": img_dir, "height_data_path": h_dir, "image_regex": "FDOP20_(\\d+)_rgbi\\.tif",
            "height_data_regex": "nDSM_(\\d+)_1km\\.tif", "combined_model": model,
            "output_directory": os.path.join(root, "output"), "tiles_path": os.path.join(root, "tiles"),
            "use_overlap": True, "merged_path": "merged", "tile_width": 50, "tile_height": 50, "buffer": 20,
            "ndvi_scaling_factor": 0.2, "height_scaling_factor": 1.0, "keep_intermediate": False, "device": "0",
            # example/config.yml
            "confidence_threshold": 0.3, "containment_threshold": 0.75, "height_threshold": 3, "ndvi_mean_threshold": 0.1,
            "ndvi_var_threshold": 0.1, "iou_threshold": 0.6, "area_threshold": 1,
            "image_merged_regex": "FDOP20_(\\d+)_(\\d+)_(\\d+)_(\\d+)_rgbi\\.tif",
            "height_data_merged_regex": "nDSM_(\\d+)(\\d+)_1km\\.tif",
        }
        path = os.path.join(root, "config.yml")
        with open(path, "w") as f:
            yaml.safe_dump(cfg, f)
        config, _ = detection.get_config(path)
        config["logger"].setLevel("ERROR")
        t0 = time.perf_counter()
        detection.process_files(config)
        wall = time.perf_counter() - t0
        stats = config.get("_last_session_stats", {})
        outs = [f for f in os.listdir(config["output_directory"]) if f.endswith(".gpkg")]
        from treedetection_b200 import gpkg
        layers = [gpkg.read_layer(os.path.join(config["output_directory"], f)) for f in sorted(outs)]
        n_crowns = [len(l[1]) - 1 for l in layers]
        parity = None
        from treedetection_b200 import golden_check
        if golden_check.golden_matches_workload(W, 1234, 2500) and H == W and npx in (0.2, 1.0):
            # the file of image 0 (the golden's georeference; the others are the same pixels 10 km further east, whose
            # float64 coordinates round differently) against the CPU oracle's golden: ids, areas, heights, vertices
            for v, o, cols, _ in layers[:1]:
                golden_check.check_layer({"poly_id": np.array([int(x) for x in cols["poly_id"]]), "area": np.array(cols["Area"]),
                                          "tree_height": np.array(cols["TreeHeight"], dtype=np.float32),
                                          "centroid": np.array([[json.loads(c)["x"], json.loads(c)["y"]] for c in cols["Centroid"]],
                                                               dtype=np.float32),
                                          "is_contained": np.array([c == "True" for c in cols["is_contained"]]),
                                          "num_contained": np.array(cols["num_contained"], dtype=np.int32),
                                          "ring_off": o, "verts": v}, "split" if npx == 0.2 else "combined")
            assert len(set(n_crowns)) == 1
            parity = (f"the output layer of image 0 equals the CPU oracle's golden (ids, areas, heights, centroids, "
                      f"containment columns, vertices); the {len(layers) - 1} shifted copies have the same crown count")
        area = n_images * H * W * px * px / 1e6
        tl = stats.get("timeline", [])
        steady = None
        if len(tl) >= 3:
            # images after the first (which allocates the staging buffers, builds the tile tables and learns the
            # capacities through the exact-size path): what a long file list converges to
            # (the second image's decode hides behind the first image's cold start: counted from the second on)
            k0 = 1 if len(tl) >= 4 else 0
            per_image = (tl[-1]["t_out"] - tl[k0]["t_out"]) / (len(tl) - 1 - k0)
            steady = {"s_per_image": round(per_image, 4), "value": (area / n_images) / per_image, "unit": UNIT,
                      "first_image_s": round(tl[0]["t_out"] - stats.get("t0", tl[0]["t_in"]), 3),
                      "first_image_breakdown_s": {k: round(tl[0][k], 3) for k in
                                                  ("wait_decode_s", "decode_s", "fixtures_s", "tables_s", "device_s")},
                      "per_image_s": {k: round(statistics.mean(t[k] for t in tl[k0 + 1:]), 4)
                                      for k in ("wait_decode_s", "decode_s", "fixtures_s", "tables_s", "device_s")}}
        return {"workload": label, "images": n_images, "value": area / wall, "unit": UNIT, "wall_s": wall,
                "steady_state": steady,
                "stage_s": {k: round(v, 3) for k, v in stats.get("stage_s", {}).items()},
                "fast_path_images": stats.get("images"), "fallback_images": stats.get("fallback_images"),
                "crowns_per_image": n_crowns, "parity": parity,
                "input_bytes": int(n_images * (sc.rgbi.nbytes + sc.ndsm.nbytes)), "file_bytes": int(file_bytes),
                "device_decoded_rasters": stats.get("device_decoded_rasters"),
                "note": f"GeoTIFFs {'LZW-compressed (predictor 2 imagery; strips decoded on the GPU, one warp each)' if compression else 'uncompressed'} on {'tmpfs (/dev/shm)' if base else 'the default temp dir'} (written in "
                        f"{write_s:.1f} s, not timed); timed: get tiles -> read + decode rasters and fixtures -> H2D -> P1 "
                        f"+ P2-P9 -> D2H -> stitched, processed and final .gpkg written (decoder thread | device | writer thread); the first image of a tiling "
                        f"learns the capacities (exact-size path), the others replay the CUDA graphs"}
    finally:
        shutil.rmtree(root, ignore_errors=True)


# --------------------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------
# B200 arm
# --------------------------------------------------------------------------------------
def run_b200(a):
    import numpy as np
    import torch
    import torch.distributed as dist

    from treedetection_b200 import _lib, api, golden_check, ops, pipeline, synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (no CPU fallback on the product path)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _lib.lib()
    p = pipeline.PipelineParams()

    # row-sharded mosaic: rank r owns the image of row r (its own seed / georeference)
    sc = synth.make_scene(seed=1234 + rank, size_px=a.size, px=0.2, ndsm_px=a.ndsm_px, density_per_km2=2500.0,
                          bottom=synth.ORIGIN_Y - rank * a.size * 0.2, stem=f"FDOP20_{rank:06d}_rgbi")
    if a.files_only:
        print(json.dumps(e2e_files(sc, a.files_images, workload_string(a.size, a.ndsm_px))))
        print(json.dumps(e2e_files(sc, a.files_images, workload_string(a.size, a.ndsm_px) + ", LZW-compressed GeoTIFFs",
                                   compression="lzw")))
        return
    host = api.HostImage.from_scene(sc)
    tables = api.TileTables(sc.tiles, dev, p.shift)
    p1_out = torch.empty((tables.p1_floats,), dtype=torch.float32, device=dev)
    # device-resident copies for the kernel-only figure
    d = {k: getattr(host, k).to(dev) for k in ("rgbi", "ndsm", "boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
    n_inst = int(host.scores.numel())
    n_tiles = len(sc.tiles)
    p1_bytes = int(sum(3 * int(w[2]) * int(w[3]) for w in tables.win.tolist())) + 4 * tables.p1_floats

    # ---- N > 1: the down-seam strip between image row r and r + 1 (the one exchange step) ----
    strip = None
    if world > 1:
        from treedetection_b200 import sharding, tiling
        rows = sharding.halo_rows(p.tile_height, p.buffer, p.overlapping_tiles_height)
        tops = [d["rgbi"][:, :rows].contiguous(), d["ndsm"][None, :rows].contiguous()]
        if rank + 1 < world:
            px = 0.2
            nb = synth.tree_field(1234 + rank + 1, a.size * px, a.size * px, 2500.0, synth.ORIGIN_X,
                                  synth.ORIGIN_Y - (rank + 1) * a.size * px)
            own = sc.field
            both = synth.TreeField(*[np.concatenate([getattr(own, k), getattr(nb, k)]) for k in
                                     ("x", "y", "r", "h", "score", "ecc")], own.left, nb.bottom, own.width_m,
                                   2 * own.height_m)
            top = own.bottom + rows * px                       # strip = bottom rows of own + top rows of neighbour
            s_tf = synth.image_transform(own.left, top, px)
            s_tiles = tiling.tile_grid(f"FDOP20_seam{rank}_rgbi", s_tf, a.size, 2 * rows, synth.EPSG, p.tile_width,
                                       p.tile_height, p.buffer)
            s_det = synth.make_detections(both, s_tiles, px, 1234 + rank)
            s_tables = api.TileTables(s_tiles, dev, p.shift)
            n_rows = rows if a.ndsm_px == 0.2 else rows
            strip = {
                "tf": s_tf, "ndsm_tf": synth.image_transform(own.left, own.bottom + n_rows * a.ndsm_px, a.ndsm_px),
                "tables": s_tables, "p1": torch.empty((s_tables.p1_floats,), dtype=torch.float32, device=dev),
                "det": {k: torch.from_numpy(getattr(s_det, k)).to(dev) for k in
                        ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")},
            }

    # ---- more than one image per rank: the right-seam strip between horizontally adjacent images (same rank) ----
    M = max(1, a.images_per_rank)
    rstrip = None
    if M > 1:
        from treedetection_b200 import tiling
        assert a.ndsm_px == 0.2, "--images-per-rank > 1 needs the 0.2 m nDSM (strips are 270 PIXELS wide in both rasters)"
        px = 0.2
        sw = int((p.tile_width + 2 * p.buffer) * p.overlapping_tiles_width)           # 270 px (merging.py:60-62)
        own = sc.field
        right = synth.TreeField(own.x + own.width_m, own.y, own.r, own.h, own.score, own.ecc, own.left + own.width_m,
                                own.bottom, own.width_m, own.height_m)                # the replica to the right
        both = synth.TreeField(*[np.concatenate([getattr(own, k), getattr(right, k)]) for k in
                                 ("x", "y", "r", "h", "score", "ecc")], own.left, own.bottom, 2 * own.width_m, own.height_m)
        r_left = own.left + own.width_m - (sw // 2) * px
        r_tf = synth.image_transform(r_left, own.bottom + own.height_m, px)
        r_tiles = tiling.tile_grid(f"FDOP20_rseam{rank}_rgbi", r_tf, sw, a.size, synth.EPSG, p.tile_width, p.tile_height,
                                   p.buffer)
        r_det = synth.make_detections(both, r_tiles, px, 4321 + rank)
        r_tables = api.TileTables(r_tiles, dev, p.shift)
        rstrip = {"tf": r_tf, "tables": r_tables, "sw": sw,
                  "p1": torch.empty((r_tables.p1_floats,), dtype=torch.float32, device=dev),
                  "rgbi": torch.empty((d["rgbi"].shape[0], a.size, sw), dtype=d["rgbi"].dtype, device=dev),
                  "ndsm": torch.empty((1, a.size, sw), dtype=d["ndsm"].dtype, device=dev), "p5": {},
                  "det": {k: torch.from_numpy(getattr(r_det, k)).to(dev) for k in
                          ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}}

    det_keys = ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")
    # P2-P9 without host synchronisation (capacity workspace); consecutive images alternate between the contexts
    runners = [pipeline.ChainRunner(p) for _ in range(max(1, a.chains))]
    runner = runners[0]
    img_seq = [0]
    strip_runner = pipeline.ChainRunner(p)
    rstrip_runner = pipeline.ChainRunner(p)
    pending = []                               # tickets of enqueued images / strips, oldest first
    MAX_IN_FLIGHT = 3                          # per runner (a ChainRunner has 4 output slots)

    def make_room(r):
        """collect the oldest tickets until runner ``r`` has fewer than MAX_IN_FLIGHT images in flight"""
        while sum(1 for q, _ in pending if q is r) >= MAX_IN_FLIGHT:
            q, t = pending.pop(0)
            n_c, f = q.collect(t)
            if q in runners:
                results.append((n_c, len(f)))

    def step_rstrip():
        """the right-seam strip between this image and its right neighbour (a replica held by the same rank)"""
        ops.seam_crop(d["rgbi"], d["rgbi"], 0, rstrip["sw"], a.size, out=rstrip["rgbi"])
        ops.seam_crop(d["ndsm"][None], d["ndsm"][None], 0, rstrip["sw"], a.size, out=rstrip["ndsm"])
        rstrip["tables"].plan(rstrip["rgbi"]).run(rstrip["rgbi"], rstrip["p1"])
        sd = rstrip["det"]
        return rstrip_runner.submit({k: sd[k] for k in det_keys}, rstrip["tables"].tile_tf, rstrip["tables"].tile_boxes,
                                    lambda: pipeline.raster_stage(rstrip["rgbi"], rstrip["tf"], rstrip["ndsm"][0],
                                                                  rstrip["tf"], p, buffers=rstrip["p5"]))

    # halo receive buffers, strip rasters and P5 outputs are allocated once: nothing is allocated inside a step,
    # and the strip's chain replays the same CUDA graphs every step (their keys are the buffer addresses)
    halo_recv = [torch.empty_like(t) for t in tops] if (world > 1 and rank + 1 < world) else None
    strip_bufs = {}

    def step_strip():
        """halo exchange (NCCL send/recv) + the seam strip through the same path"""
        recv = sharding.exchange_down_halos(tops, rank, world, recv=halo_recv)
        if recv is None or strip is None:
            return None
        if "rgbi" not in strip_bufs:
            sh = 2 * recv[0].shape[1]
            strip_bufs["rgbi"] = torch.empty((d["rgbi"].shape[0], sh, d["rgbi"].shape[2]), dtype=d["rgbi"].dtype, device=dev)
            strip_bufs["ndsm"] = torch.empty((1, 2 * recv[1].shape[1], d["ndsm"].shape[1]), dtype=d["ndsm"].dtype, device=dev)
            strip_bufs["p5"] = {}
        s_rgbi = sharding.assemble_down_strip(d["rgbi"], recv[0], out=strip_bufs["rgbi"])
        s_ndsm = sharding.assemble_down_strip(d["ndsm"][None], recv[1], out=strip_bufs["ndsm"])[0]
        strip["tables"].plan(s_rgbi).run(s_rgbi, strip["p1"])
        sd = strip["det"]
        if a.exact:
            table = pipeline.predict_stage(sd["boxes_net"], sd["scores"], sd["probs"], sd["inst_tile"],
                                           sd["tile_dims"], strip["tables"].tile_tf, strip["tables"].tile_boxes, p)
            rasters = pipeline.raster_stage(s_rgbi, strip["tf"], s_ndsm, strip["ndsm_tf"], p)
            pipeline.postprocess_stage(table, rasters, p)
            return None
        return strip_runner.submit({k: sd[k] for k in det_keys}, strip["tables"].tile_tf, strip["tables"].tile_boxes,
                                   lambda: pipeline.raster_stage(s_rgbi, strip["tf"], s_ndsm, strip["ndsm_tf"], p,
                                                                 buffers=strip_bufs["p5"]))

    ev = lambda: torch.cuda.Event(enable_timing=True)
    p1_ev = []

    stage_ev = []
    results = []

    # P1 (HBM bound, feeds the predictor) and the P2-P9 chain (latency bound, consumes the
    # predictor's outputs) are independent: P1 rides its own stream, the chain a high-priority one
    p1_stream = torch.cuda.Stream(device=dev, priority=a.p1_priority)
    chain_streams = [torch.cuda.Stream(device=dev, priority=a.chain_priority) for _ in runners]
    chain_stream = chain_streams[0]
    strip_stream = torch.cuda.Stream(device=dev, priority=-1)   # N > 1: the seam strip, next to the image
    p5_stream = torch.cuda.Stream(device=dev)
    pre_rasters = []
    p5_ev = []
    # NDVI / height output rasters, rotated with the (chain context, output slot) an image lands in, so that every
    # context replays ONE graph per slot (a graph's key includes the raster addresses)
    p5_bufs = [{} for _ in range(4 * len(runners))]
    p5_seq = [0]

    def chain(e):
        """P2-P9 of the resident image on the current stream; e[2..5] bracket the stages"""
        e[2].record()
        if a.exact:
            table = pipeline.predict_stage(d["boxes_net"], d["scores"], d["probs"], d["inst_tile"], d["tile_dims"],
                                           tables.tile_tf, tables.tile_boxes, p)
            e[3].record()
            rasters = pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p)
            e[4].record()
            feats = pipeline.postprocess_stage(table, rasters, p)
            e[5].record()
            results.append((len(table), len(feats)))
            return None
        marks = {"p4": e[3], "p5": e[4], "p9": e[5]}

        def mark(name):
            if name in marks:
                marks[name].record()
        pre = pre_rasters.pop() if pre_rasters else None
        return runners[img_seq[0] % len(runners)].submit({k: d[k] for k in det_keys}, tables.tile_tf, tables.tile_boxes,
                             (lambda: pre) if pre is not None else
                             (lambda: pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p)),
                             mark=mark)

    side_streams = (p1_stream, *chain_streams, strip_stream, p5_stream)
    p5_guard = [None] * len(p5_bufs)           # chain event after which a P5 output buffer may be overwritten

    def region_begin():
        main = torch.cuda.current_stream()
        for st in side_streams:
            st.wait_stream(main)

    def region_end():
        main = torch.cuda.current_stream()
        for st in side_streams:
            main.wait_stream(st)

    def step_resident():
        """One step = M images per GPU (default 1), the right-seam strips between them and, N > 1, the down-seam
        strip below each of them.  The streams run FREE between the two ends of a timed region: images are
        independent, so P1 of the next image may start while the chain of this one finishes -- there is no join
        per image (``--join-steps`` restores one).  The host collects the counters of the oldest image whenever a
        runner has three images in flight."""
        for m in range(M):
            e = [ev() for _ in range(6)]
            t = ts = tr = None
            img_seq[0] += 1
            cur, cur_stream = runners[img_seq[0] % len(runners)], chain_streams[img_seq[0] % len(runners)]
            make_room(cur)
            if a.serial:
                e[0].record()
                tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
                e[1].record()
                t = chain(e)
                if world > 1:
                    make_room(strip_runner)
                    ts = step_strip()
                if rstrip is not None and m + 1 < M:
                    make_room(rstrip_runner)
                    tr = step_rstrip()
            else:
                if a.join_steps:
                    region_begin()
                with torch.cuda.stream(p1_stream):
                    e[0].record()
                    tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
                    e[1].record()
                if not a.exact:
                    # P5 (issue bound, needed only by the statistics) on its own stream, next to P2-P4
                    with torch.cuda.stream(p5_stream):
                        p5_seq[0] += 1
                        b = p5_seq[0] % len(p5_bufs)
                        if p5_guard[b] is not None:
                            p5_stream.wait_event(p5_guard[b])        # the chain that read this buffer is done
                        q0, q1 = ev(), ev()
                        q0.record()
                        r = pipeline.raster_stage(d["rgbi"], host.transform, d["ndsm"], host.ndsm_transform, p,
                                                  buffers=p5_bufs[b])
                        q1.record()
                        p5_ev.append((q0, q1))
                        r["ndvi_ready"] = p5_stream.record_event()
                    pre_rasters.append(r)
                # the strips' small dependent launches go in first: they run under P1 while the host is still enqueuing
                if world > 1:
                    make_room(strip_runner)
                    with torch.cuda.stream(strip_stream):
                        ts = step_strip()
                if rstrip is not None and m + 1 < M:
                    make_room(rstrip_runner)
                    with torch.cuda.stream(strip_stream):
                        tr = step_rstrip()
                with torch.cuda.stream(cur_stream):
                    t = chain(e)
                    if not a.exact:
                        p5_guard[p5_seq[0] % len(p5_bufs)] = cur_stream.record_event()
                if a.join_steps:
                    region_end()
            p1_ev.append((e[0], e[1]))
            stage_ev.append(e)
            for q, tk in ((cur, t), (strip_runner, ts), (rstrip_runner, tr)):
                if tk is not None:
                    pending.append((q, tk))

    last_feats = []

    def drain():
        while pending:
            r, t = pending.pop(0)
            n_c, f = r.collect(t)
            if r in runners:
                results.append((n_c, len(f)))
                last_feats[:] = [f]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, free_running=False):
        barrier()
        s, e = ev(), ev()
        s.record()
        if free_running:
            region_begin()
        out = None
        for _ in range(steps):
            out = fn()
        if free_running:
            region_end()
        e.record()
        barrier()
        ms = s.elapsed_time(e)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, out

    # setup, untimed: every chain context learns its capacities from one exact-size image and captures the two
    # graphs of each of its 4 output slots; then the W warm-up steps
    priming = 0 if a.exact else -(-5 * len(runners) // M)
    region_begin()
    for _ in range(priming + a.warmup):
        step_resident()
    region_end()
    drain()
    p1_ev.clear()
    stage_ev.clear()
    results.clear()
    sampler = ClockSampler(local)
    if rank == 0 and not a.no_clocks:
        sampler.start()
    l0 = _lib.launch_count
    ms, _ = timed(step_resident, a.steps, free_running=not a.serial)
    drain()
    launches = _lib.launch_count - l0
    assert len(results) == a.steps * M and len(set(results)) == 1, "steps disagree on the crown counts"
    n_cand, n_final = results[-1]
    # parity of THIS run's crowns: the last step's final layer against the golden the CPU oracle produced for the same
    # scene (tests/golden/config2.npz; rank 0's image is the golden's scene); outside the timed region
    parity = None
    if rank == 0 and last_feats and golden_check.golden_matches_workload(a.size, 1234, 2500) and a.ndsm_px in (0.2, 1.0):
        parity = golden_check.check_layer(api.features_to_host(last_feats[-1]), "split" if a.ndsm_px == 0.2 else "combined",
                                          n_candidates=n_cand)
    # ---- the other nDSM resolution of BASELINE config 2 (1 m: the reference's COMBINED statistics path,
    # get_metadata_within_polygon), same image, a short run of its own ----
    combined = None
    if rank == 0 and world == 1 and a.ndsm_px == 0.2 and not a.no_merged and not a.serial and not a.exact:
        ndsm1 = torch.from_numpy(synth.make_ndsm(sc.field, 1.0, 1234)).to(dev)
        tf1 = synth.image_transform(sc.field.left, sc.field.bottom + sc.field.height_m, 1.0)
        run1, bufs1, ticks = pipeline.ChainRunner(p), {}, []

        def step1():
            with torch.cuda.stream(p1_stream):
                tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
            with torch.cuda.stream(chain_stream):
                while len(ticks) >= 2:
                    run1.collect(ticks.pop(0))
                ticks.append(run1.submit({k: d[k] for k in det_keys}, tables.tile_tf, tables.tile_boxes,
                                         lambda: pipeline.raster_stage(d["rgbi"], host.transform, ndsm1, tf1, p, buffers=bufs1)))
        region_begin()
        for _ in range(8):       # untimed: capacity learning + the graph capture of each of the runner's 4 output slots
            step1()
        region_end()
        torch.cuda.synchronize()
        ms1, _ = timed(step1, max(5, a.steps // 2), free_running=True)
        f1 = None
        while ticks:
            n1c, f1 = run1.collect(ticks.pop(0))
        par1 = golden_check.check_layer(api.features_to_host(f1), "combined", n_candidates=n1c) \
            if golden_check.golden_matches_workload(a.size, 1234, 2500) else None
        combined = {"workload": workload_string(a.size, 1.0), "value": sc.area_km2 * max(5, a.steps // 2) / (ms1 / 1e3),
                    "unit": UNIT, "ms_per_step": ms1 / max(5, a.steps // 2), "crowns": len(f1), "parity": par1,
                    "exact_size_fallbacks_incl_warmup": run1.fallbacks}
        del ndsm1
    p1_ms_in_step = statistics.mean(x.elapsed_time(y) for x, y in p1_ev)
    # the roofline kernel timed alone (CUDA events on its stream, after the timed region): inside the
    # step it shares the GPU with the P2-P9 chain, which says nothing about the kernel itself
    p1_alone = []
    for _ in range(0 if a.no_alone else 3):
        tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
    torch.cuda.synchronize()
    for _ in range(0 if a.no_alone else 20):
        s0, s1 = ev(), ev()
        s0.record()
        tables.plan(d["rgbi"]).run(d["rgbi"], p1_out)
        s1.record()
        p1_alone.append((s0, s1))
    torch.cuda.synchronize()
    p1_ms = statistics.mean(x.elapsed_time(y) for x, y in p1_alone) if p1_alone else p1_ms_in_step
    names = ["P1 tile cut/normalise", "P2-P4 paste/contours/stitch", "P5 NDVI/decimation", "P6-P9 NMS/stats/select"]
    pairs = [(0, 1), (2, 3), (3, 4), (4, 5)]
    chain_mode = "exact sizes (host sync before every allocation)" if a.exact else \
        f"td_chain (capacity workspace, device-side counts, {'CUDA-graph replay' if runner.use_graph else 'direct launches'}; " \
        f"{len(runners)} chain contexts that consecutive images alternate between; exact-size fallbacks incl. warm-up: " \
        f"image {sum(r.fallbacks for r in runners)}, seam strip {strip_runner.fallbacks}; {priming} untimed priming steps " \
        f"(capacity learning + graph capture) before the warm-up)"
    stage_ms = {n: statistics.mean(e[i].elapsed_time(e[j]) for e in stage_ev) for (i, j), n in zip(pairs, names)}
    p1_ms_in_step = stage_ms[names[0]]
    if p5_ev:      # P5 ran on its own stream
        stage_ms[names[2]] = statistics.mean(x.elapsed_time(y) for x, y in p5_ev[-a.steps:])
    area = sc.area_km2
    value = world * M * area * a.steps / (ms / 1e3)

    # end to end through the host-buffer API
    e2e_runner = pipeline.ChainRunner(p)
    def step_e2e():
        ts = step_strip() if world > 1 else None     # enqueued first: runs under the image's H2D copies
        out, _ = api.run_image(host, p, dev, tables, p1_out, runner=None if a.exact else e2e_runner)
        if ts is not None:
            strip_runner.collect(ts)
        return out
    for _ in range(max(a.warmup, 5)):      # untimed: capacity learning + graph capture of the runner's 4 output slots
        step_e2e()
    e2e_steps = max(1, min(a.steps, 20))
    ms_e2e, out = timed(step_e2e, e2e_steps)
    e2e_value = world * area * e2e_steps / (ms_e2e / 1e3)
    # the ceiling of e2e: the same pinned buffers copied to the device and nothing else, all ranks at once
    # (one NUMA node feeds every GPU of the box: the host side, not the GPUs, bounds e2e at N > 1)
    h2d_dst = {k: torch.empty_like(getattr(host, k), device=dev) for k in ("rgbi", "ndsm", "probs")}
    def step_h2d():
        for k, t in h2d_dst.items():
            t.copy_(getattr(host, k), non_blocking=True)
    step_h2d()
    ms_h2d, _ = timed(step_h2d, 5)
    h2d_bytes = sum(t.numel() * t.element_size() for t in h2d_dst.values())
    h2d_gbs = h2d_bytes * 5 / (ms_h2d / 1e3) / 1e9           # per GPU, slowest rank
    del h2d_dst
    clocks = sampler.stop() if rank == 0 else None     # sampled over both timed regions
    d2h = int(sum(v.nbytes for v in out.values() if hasattr(v, "nbytes")))

    # ---- secondary metric of BASELINE.json: crowns merged/s (config 4, dense-forest stress) ----
    merged = None
    if rank == 0 and not a.no_merged:
        rng = np.random.default_rng(4)
        n_c = 800_000                                   # ~2,000 crowns per 50 m tile over 400 tiles (1 km^2)
        cx = synth.ORIGIN_X + rng.uniform(0, 1000.0, n_c); cy = synth.ORIGIN_Y + rng.uniform(0, 1000.0, n_c)
        rx = rng.uniform(0.5, 2.0, n_c); ry = rx * rng.uniform(0.8, 1.25, n_c)
        bounds = torch.from_numpy(np.stack([cx - rx, cy - ry, cx + rx, cy + ry], 1)).to(dev)
        conf = torch.from_numpy(np.round(rng.uniform(0.3, 1.0, n_c), 3)).to(dev)
        area_c = torch.from_numpy(np.pi * rx * ry).to(dev)
        b32 = bounds.to(torch.float32).contiguous()

        def merge_once():
            rem = ops.bbox_nms_ordered(bounds, conf, area_c, p.iou_threshold, p.area_threshold)
            ops.containment(b32, p.containment_threshold)
            return rem
        for _ in range(3):
            merge_once()
        torch.cuda.synchronize()
        s0, s1 = ev(), ev()
        s0.record()
        reps = 10
        for _ in range(reps):
            rem = merge_once()
        s1.record()
        torch.cuda.synchronize()
        ms_m = s0.elapsed_time(s1) / reps
        merged = {"metric": "crowns merged/s", "value": n_c / (ms_m / 1e3), "unit": "crowns/s", "ms": ms_m,
                  "config": f"dense-forest stress: {n_c} candidate crowns on 1 km^2 (~2,000 per 50 m tile, 25 % tile "
                            f"overlap equivalent), P6 ordered bbox NMS + P8 containment, "
                            f"{int(rem.sum().item())} suppressed"}

    files = None
    if rank == 0 and world == 1 and not a.no_files:
        torch.cuda.synchronize()
        del d, p1_out                   # e2e_files brings its own device buffers
        torch.cuda.empty_cache()
        files = [e2e_files(sc, a.files_images, workload_string(a.size, a.ndsm_px)),
                 e2e_files(sc, a.files_images, workload_string(a.size, a.ndsm_px) + ", LZW-compressed GeoTIFFs",
                           compression="lzw")]
        try:      # BASELINE config 1: the reference's example tile (bundled nDSM 1000^2 @ 1 m + synthetic RGBI 5000^2)
            files.append(e2e_files(synth.config1_scene(), 2, "BASELINE config 1: example/config.yml on the bundled nDSM tile "
                                                       "324125317 (1 m) + synthetic 5000x5000 px RGBI"))
        except Exception as e:          # the fixture of the bundled tile is test data; never fail the bench on it
            files.append({"workload": "BASELINE config 1", "error": str(e)})
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = p1_bytes / (p1_ms / 1e3) / 1e9
        traffic = traffic_src = None
        try:      # dram bytes per launch from the committed ncu capture of this kernel (profiles/ncu_traffic.py)
            tj = json.load(open(os.path.join(ROOT, "profiles", "p1_traffic.json")))
            if a.size == 10000:
                traffic, traffic_src = tj["traffic_bytes_per_launch"], tj["source"]
        except Exception:
            pass
        # whole path: SURVEY.md section 8d per-unit algorithmic bytes x the units of this run
        hh, ww = host.rgbi.shape[1:]
        ndvi_px = int(hh * p.ndvi_scaling_factor) * int(ww * p.ndvi_scaling_factor)
        path_bytes = {
            "P1": p1_bytes,
            "P2-P4": n_inst * (3136 + 20 + 320) + n_inst * (320 + 512) + n_cand * 1024,
            # bands 0 and 3 in, NDVI out; the nDSM is read by P5 only when it is decimated (height_scaling_factor != 1:
            # otherwise the statistics of P7 read it per crown window, counted under P6-P9)
            "P5": 2 * hh * ww + 4 * ndvi_px + (4 * int(host.ndsm.numel()) * (1 + p.height_scaling_factor ** 2)
                                                 if p.height_scaling_factor != 1.0 else 0),
            "P6-P9": 37 * n_cand + n_final * ((10752 if a.ndsm_px == 0.2 else 717) + 205),
        }
        path_total = sum(path_bytes.values())
        path_gbs = path_total / (ms / a.steps / 1e3) / 1e9
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8/f32 rasters, f32/f16 NMS, f64 geometry", "data": "synthetic",
            "config": {"workload": workload_string(a.size, a.ndsm_px),
                       "counts": f"{n_tiles} tiles, {n_inst} ROI-head instances -> {n_cand} candidate crowns -> "
                                 f"{n_final} crowns per image",
                       "area_km2_per_gpu": area * M, "images_per_gpu_per_step": M,
                       "strips_per_step": {"right_seam": world * (M - 1) if rstrip is not None else 0,
                                           "down_seam": (world - 1) * M,
                                           "note": "seam strips re-process imagery the images already cover: their area is "
                                                   "NOT counted in value"}, "stages": "P1+P2+P3+P4+P5+P6+P7+P8+P9",
                       "chain": chain_mode,
                       "streams": ("one stream, stages back to back" if a.serial else
                                   "P1 and P5 on their own streams concurrent with the P2-P4 / P6-P9 chain (high-priority stream); "
                                   + ("all streams joined after every image; " if a.join_steps else
                                      "streams run free between the two ends of the timed region (images are independent; the "
                                      "host stays <= 3 images ahead per chain context); ") +
                                   "stage_ms are per-stream CUDA-event times and overlap"),
                       "stage_ms": {k: round(v, 3) for k, v in stage_ms.items()},
                       "cache": "inputs (rasters 0.8 GB, P1 output 12 GB) exceed the 126 MB L2; no flush needed",
                       "parallelism": (f"image-row sharding x{world}; per step every rank r < N-1 also receives the "
                                       f"135-row RGBI + nDSM halo of rank r+1 (NCCL send/recv) and runs the down-seam "
                                       f"strip through the same path" if world > 1 else
                                       "1 GPU; image-row sharding for N > 1")},
            "roofline": {"bound": "hbm", "kernel": "tile_resize_u8_up_warp_kernel (P1, td_tile_cut_normalize)", "achieved": achieved, "peak": peak,
                         "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                         "write_only_peak": "a pure fill of 12 GiB runs at 7.45 TB/s on this part (scripts/hbm_probe.py); "
                                            "the kernel writes 12.6 GB and reads 0.4 GB, so frac can approach 1.1",
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650 GB/s",
                         "algorithmic_bytes_per_launch": p1_bytes, "ms_per_launch": p1_ms,
                         "timed": ("in the step" if a.no_alone else
                                   "alone: 20 launches after the timed region, CUDA events on the launching stream "
                                   "(burst peak applies)"),
                         "ms_per_launch_in_step": p1_ms_in_step,
                         "in_step_note": ("serial step: the kernel runs alone inside the step" if a.serial else
                                          "inside the step the kernel shares the GPU with the concurrent P2-P9 chain"),
                         "share_of_step": p1_ms / (ms / a.steps),
                         "launches": "one td_tile_cut_normalize call = aligned-rows kernel + unaligned-rows kernel "
                                     "(side stream) + a 1.6 KB memset"},
            "path_roofline": {"bound": "hbm", "algorithmic_bytes_per_step": path_total, "by_stage": path_bytes,
                              "achieved": path_gbs, "peak": peak, "unit": "GB/s", "frac": path_gbs / peak,
                              "note": "all stages of one step against the same HBM peak (SURVEY 8d per-unit bytes)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": host.h2d_bytes(), "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / e2e_steps,
                    "h2d_gbs_per_gpu_in_e2e": host.h2d_bytes() / (ms_e2e / e2e_steps / 1e3) / 1e9,
                    "h2d_ceiling_gbs_per_gpu": h2d_gbs, "h2d_ceiling_gbs_aggregate": h2d_gbs * world,
                    "ceiling_note": "pure pinned->device copies of the same buffers on all ranks at once: what the host "
                                    "side (PCIe + the one NUMA node that feeds every GPU of the box) delivers; e2e cannot "
                                    "exceed area / (bytes / ceiling)",
                    "value_at_ceiling": world * area / (host.h2d_bytes() / (h2d_gbs * 1e9))},
            "gpu_launches": launches,
            "e2e_files": files,
            "combined_path": combined,
            "parity": parity or "not checked (no golden for this workload size)",
            "clocks": clocks,
            "crowns_merged": merged,
        }
        if not a.no_cpu_baseline and world == 1:
            # (1) the FULL workload on one core through the oracle's full-size forms (windowed statistics, sparse NMS,
            # chunked containment: same results as the literal loops, tests/test_oracle_windowed.py); (2) the literal
            # restatement of the reference's N x P / N x N loops on a bounded sub-scene
            full = cpu_rate(a.size, a.ndsm_px, 1, large=True)
            lit = cpu_rate(a.cpu_sample, a.ndsm_px, 1)
            line["cpu_baseline"] = {"value": full["km2_per_s_wall"], "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"the full workload ({a.size}x{a.size} px, {full['rings']} candidate rings -> "
                                              f"{full['crowns']} crowns), P1-P9 restated in oracle/port.py with windowed "
                                              f"statistics / sparse NMS, {full['wall_s']:.1f} s on one core of "
                                              f"{os.cpu_count()}",
                                    "literal": {"value": lit["km2_per_s_wall"], "unit": UNIT,
                                                "sample": f"one {a.cpu_sample}x{a.cpu_sample} px sub-scene "
                                                          f"({lit['rings']} candidate rings), the reference's literal "
                                                          f"N x P / N x N loops, {lit['wall_s']:.1f} s on one core"},
                                    "note": infeasible_note(n_cand, full["crowns"], float(a.size) ** 2)}
            assert full["rings"] == n_cand and full["crowns"] == n_final, "CPU port and CUDA path disagree on the counts"
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------------------
# BASELINE config 3: two models + forest outline on a 20 000 x 20 000 px mosaic (1 GPU)
# --------------------------------------------------------------------------------------
def run_config3(a):
    """P1 once, P2-P4 for the urban and the forest model (each skipping the tiles the outline assigns to the other,
    prediction.py:79-93), P10 fusion against the forest outline (helpers.py:795-811), P5, P6-P9 on the fused layer.
    Exact-size composition (the two tables meet in the fusion, which the single-model CUDA-graph chain does not
    cover).  Parity of this path: tests/test_gpu_two_model.py (small scene, against the oracle)."""
    import numpy as np
    import torch

    from treedetection_b200 import api, fusion, pipeline, synth, tiling
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    size = a.size if a.size != 10000 else 20000
    p = pipeline.PipelineParams()
    t0 = time.perf_counter()
    field = synth.tree_field(1234, size * 0.2, size * 0.2, 2500.0)
    rgbi, ndsm = synth.make_rgbi(field, 0.2, 1234), synth.make_ndsm(field, a.ndsm_px, 1234)
    top = synth.ORIGIN_Y + size * 0.2
    tf, ntf = synth.image_transform(synth.ORIGIN_X, top, 0.2), synth.image_transform(synth.ORIGIN_X, top, a.ndsm_px)
    # forest outline: ~40 % cover, random convex patches and rectangles, a third of them with a clearing (hole)
    rng = np.random.default_rng(3)
    ext = size * 0.2
    polys, cover = [], 0.0
    while cover < 0.4 * ext * ext:
        cx, cy = synth.ORIGIN_X + rng.uniform(0, ext), synth.ORIGIN_Y + rng.uniform(0, ext)
        r = rng.uniform(0.04, 0.12) * ext
        if rng.uniform() < 0.4:
            w, h = rng.uniform(0.8, 2.0, 2) * r
            ring = np.array([(cx + w, cy - h), (cx + w, cy + h), (cx - w, cy + h), (cx - w, cy - h), (cx + w, cy - h)])
            cover += 4 * w * h
        else:
            ang = np.sort(rng.uniform(0, 2 * np.pi, int(rng.integers(6, 14))))
            ring = np.stack([cx + r * np.cos(ang), cy + r * np.sin(ang)], 1)
            ring = np.concatenate([ring, ring[:1]])
            cover += 2.6 * r * r
        poly = [ring]
        if rng.uniform() < 0.33:
            ha = np.sort(rng.uniform(0, 2 * np.pi, 8))[::-1]
            hole = np.stack([cx + 0.3 * r * np.cos(ha), cy + 0.3 * r * np.sin(ha)], 1)
            poly.append(np.concatenate([hole, hole[:1]]))
        polys.append(poly)
    forest = fusion.ForestIndex(polys, dev)
    tiles = tiling.tile_grid("FDOP20_000000_rgbi", tf, size, size, synth.EPSG, p.tile_width, p.tile_height, p.buffer, forest)
    flags = np.array([[m["only_forest"], m["only_urban"]] for m in tiles.values()])
    dets = {}
    for name, seed, skip_col in (("urban", 5, 0), ("forest", 6, 1)):
        d = synth.make_detections(field, tiles, 0.2, seed)
        keep = ~flags[d.inst_tile, skip_col]
        dets[name] = {k: torch.from_numpy(np.ascontiguousarray(getattr(d, k)[keep] if k != "tile_dims" else d.tile_dims)).to(dev)
                      for k in ("boxes_net", "scores", "probs", "inst_tile", "tile_dims")}
    tables = api.TileTables(tiles, dev, p.shift)
    d_rgbi, d_ndsm = torch.from_numpy(rgbi).to(dev), torch.from_numpy(ndsm).to(dev)
    p1_out = torch.empty((tables.p1_floats,), dtype=torch.float32, device=dev)
    setup_s = time.perf_counter() - t0
    p1_stream = torch.cuda.Stream(device=dev)
    counts = {}

    def step():
        main = torch.cuda.current_stream()
        p1_stream.wait_stream(main)
        with torch.cuda.stream(p1_stream):
            tables.plan(d_rgbi).run(d_rgbi, p1_out)
        tabs = {n: pipeline.predict_stage(**dets[n], tile_tf=tables.tile_tf, tile_boxes=tables.tile_boxes, p=p)
                for n in ("urban", "forest")}
        u, f = tabs["urban"], tabs["forest"]
        verts, off, conf = fusion.fuse_tables((u.verts, u.ring_off, u.conf), (f.verts, f.ring_off, f.conf), forest)
        rasters = pipeline.raster_stage(d_rgbi, tf, d_ndsm, ntf, p)
        feats = pipeline.postprocess_stage(pipeline.CrownTable(verts, off, conf), rasters, p)
        main.wait_stream(p1_stream)
        counts.update(urban=len(u), forest=len(f), fused=int(off.shape[0] - 1), crowns=len(feats))
        return feats

    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()
    s, e = ev(), ev()
    steps = max(3, min(a.steps, 10))
    s.record()
    for _ in range(steps):
        feats = step()
    e.record()
    torch.cuda.synchronize()
    ms = s.elapsed_time(e) / steps
    area = (size * 0.2 / 1000.0) ** 2
    v = feats.verts.cpu().numpy()
    assert len(feats) > 1000 and np.array_equal(v, np.round(v, 3))
    print(json.dumps({
        "metric": METRIC, "value": area / (ms / 1e3), "unit": UNIT, "n_gpus": 1, "steps": steps, "warmup": max(a.warmup, 3),
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u8/f32 rasters, f32/f16 NMS, f64 geometry", "data": "synthetic",
        "config": {"workload": f"BASELINE config 3: two-model run with a forest outline on a synthetic {size}x{size} px mosaic "
                               f"({len(tiles)} tiles, {int(flags[:, 0].sum())} forest-only / {int(flags[:, 1].sum())} urban-only; "
                               f"{len(polys)} outline polygons, {sum(len(q) - 1 for q in polys)} holes)",
                   "counts": counts, "instances": {k: int(v_["scores"].numel()) for k, v_ in dets.items()},
                   "chain": "exact-size composition: P1 || 2 x (P2-P4) -> P10 fusion -> P5 -> P6-P9",
                   "setup_s": round(setup_s, 1)},
        "gpu_launches": None, "e2e": None}))


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == 3:
        run_config3(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
