"""Golden values of the reference's HOST logic (no device work), produced by the REFERENCE'S OWN
functions loaded through oracle.refshim:

* get_config (TreeDetection/config.py:144-238): keys and defaults for a minimal YAML, for a YAML that
  overrides some keys, and the assertion messages of incomplete configurations;
* filename_geoinfo / box_make / box_filter parameters (TreeDetection/helpers.py:265-319) on tile ids;
* round_coordinates (TreeDetection/utilities.py:146-161), check_similarity_bounds, element_is_near_border.

    python tests/golden/make_golden_host.py        (build container only) -> tests/golden/host_logic.json
"""
import json
import os
import sys
import tempfile

import yaml

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import refshim  # noqa: E402

SKIP = {"logger", "image_directory", "height_data_path", "combined_model", "urban_model", "forrest_model",
        "forrest_outline", "output_directory", "tiles_path", "continue"}


def run_get_config(ns, tmp, extra, drop=()):
    img, h, model = (os.path.join(tmp, n) for n in ("rgb", "ndsm", "model"))
    for d in (img, h, model):
        os.makedirs(d, exist_ok=True)
    cfg = {"image_directory": img, "height_data_path": h, "combined_model": model,
           "output_directory": os.path.join(tmp, "out"), "tiles_path": os.path.join(tmp, "tiles")}
    cfg.update(extra)
    for k in drop:
        cfg.pop(k, None)
    path = os.path.join(tmp, "c.yml")
    with open(path, "w") as f:
        yaml.safe_dump(cfg, f)
    try:
        out, _ = ns.config.get_config(path)
    except AssertionError as e:
        return {"assertion": str(e)}
    return {k: v for k, v in out.items() if k not in SKIP}


def main():
    ns = refshim.load()
    g = {}
    with tempfile.TemporaryDirectory() as tmp:
        g["config_minimal"] = run_get_config(ns, tmp, {})
        g["config_overrides"] = run_get_config(ns, tmp, {"tile_width": 64, "buffer": 8, "iou_threshold": 0.6,
                                                         "exclude_files": ["a.gpkg"], "device": "cuda:1",
                                                         "ndvi_scaling_factor": 0.2, "debug": True})
        g["config_no_images"] = run_get_config(ns, tmp, {}, drop=("image_directory",))
        g["config_no_height"] = run_get_config(ns, tmp, {}, drop=("height_data_path",))
        g["config_no_model"] = run_get_config(ns, tmp, {}, drop=("combined_model",))
    ids = ["FDOP20_000001_rgbi_412000_5318000_50_20_25832", "324125317_412950_5318950_50_20_25832",
           "x_y_z_7_-3_200_30_4326"]
    g["filename_geoinfo"] = {}
    for tid in ids:
        try:
            g["filename_geoinfo"][tid] = list(ns.helpers.filename_geoinfo("Prediction_" + tid + ".json"))
        except Exception as e:      # the reference's parser rejects it
            g["filename_geoinfo"][tid] = {"error": type(e).__name__}
    vals = [412000.0004999, 412000.0005, 412000.0015, 5318000.123456789, -0.0005, 2.5e-4, 1e-12, 123456.7895]
    # round_coordinates takes a polygon = list of rings of (x, y) points
    g["round_coordinates"] = {"in": vals, "out": [list(map(float, p)) for p in
                                                   ns.utilities.round_coordinates([[[v, -v] for v in vals]])[0]]}
    json.dump(g, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "host_logic.json"), "w"), indent=1,
              sort_keys=True)
    print(json.dumps(g, indent=1, sort_keys=True)[:1500])


if __name__ == "__main__":
    main()
