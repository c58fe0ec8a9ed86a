"""Configuration -- same YAML schema, defaults, assertions and ``Config`` singleton as
``TreeDetection/config.py`` (load_config :68-79, setup_logging :81-110,
set_device_configuration :112-142, get_config :144-238).

The detectron2 part (``setup_model_cfg``, config.py:25-66) is out of scope: the Mask R-CNN
is not rewritten, its raw outputs are replayed from fixtures (``predictor.py``)."""
from __future__ import annotations

import logging
import os
import warnings
from datetime import datetime

import yaml


class Config:
    """Process-wide singleton mirroring the config dict as attributes (config.py:12-23)."""
    _instance = None

    def __new__(cls, *args, **kwargs):
        if cls._instance is None:
            cls._instance = super().__new__(cls)
        return cls._instance

    @classmethod
    def _load_into_config(cls, config):
        for key, value in config.items():
            setattr(cls, key, value)


def load_config(config_path: str):
    with open(config_path, "r") as file:
        return yaml.safe_load(file)


def setup_logging(log_path: str, debug: bool):
    os.makedirs(log_path, exist_ok=True)
    log_filename = f"logs_{datetime.now().strftime('%Y%m%d_%H%M%S')}.log"
    log_file_path = os.path.join(log_path, log_filename)
    logging.basicConfig(filename=log_file_path, format="%(asctime)s - %(levelname)s - %(message)s",
                        level=logging.DEBUG if debug else logging.INFO, datefmt="%Y-%m-%d %H:%M:%S")
    logger = logging.getLogger(__name__)
    if not any(isinstance(h, logging.StreamHandler) and not isinstance(h, logging.FileHandler) for h in logger.handlers):
        console_handler = logging.StreamHandler()
        console_handler.setFormatter(logging.Formatter("%(asctime)s - %(levelname)s - %(message)s"))
        logger.addHandler(console_handler)
    return logger


def _gpu_index(raw_device):
    """Device spec of the YAML (int, "2", "cuda", "cuda:2", anything else -> 0) as a GPU index string."""
    if isinstance(raw_device, int):
        return raw_device
    if isinstance(raw_device, str):
        spec = raw_device[5:] if raw_device.startswith("cuda:") else raw_device
        if spec.isdigit():
            return spec
    return "0"


def set_device_configuration(config, raw_device):
    """config["device"]: GPU index string, or "cpu" when CUDA is absent (reference config.py:112-142:
    same accepted spellings, same messages).  With "cpu" this implementation refuses to run the
    kernels -- there is no CPU path."""
    import torch
    if not torch.cuda.is_available():
        if isinstance(raw_device, str) and raw_device.startswith("cuda"):
            warnings.warn(f"CUDA device '{raw_device}' requested but CUDA is not available. Falling back to CPU.")
        config["device"] = "cpu"
        return
    index = "0" if raw_device is None else _gpu_index(raw_device)
    if raw_device is not None:
        try:
            assert torch.cuda.device_count() > int(index), f"GPU index {int(index)} is out of range."
        except (IndexError, ValueError):
            raise ValueError(f"Invalid CUDA device specification: {raw_device}")
    config["device"] = str(index)


# (key, message) -- paths that must be present and exist (reference config.py:173-180)
_REQUIRED = [
    ("image_directory", "Input path is missing from the configuration or path is incorrect."),
    ("height_data_path", "nDOM path is missing from the configuration or path is incorrect."),
]
_REQUIRED_WITHOUT_COMBINED_MODEL = [
    ("urban_model", "Urban model path is missing from the configuration or path is incorrect."),
    ("forrest_model", "Forrest model path is missing from the configuration."),
    ("forrest_outline", "Forrest outline path is missing from the configuration."),
]

# the reference's defaults (config.py:182-233), in the order it fills them in
_DEFAULTS = [
    ("output_directory", "./output"), ("tiles_path", "./tiles"),
    ("tile_width", 50), ("tile_height", 50), ("buffer", 20), ("batch_size", 10),
    ("use_overlap", True), ("overlapping_tiles_width", 3), ("overlapping_tiles_height", 3), ("merged_path", "merged"),
    ("image_merged_regex", "FDOP20_(\\d+)_(\\d+)_(\\d+)_(\\d+)_(\\d+)\\.tif"),
    ("height_data_merged_regex", "FDOP20_(\\d+)_(\\d+)\\.tif"),
    ("iou_threshold", 0.5), ("confidence_threshold_stitching", 0.3), ("area_threshold", 1),
    ("exclude_files", []), ("confidence_threshold", 0.3), ("containment_threshold", 0.9), ("height_threshold", 3),
    ("parallel", True), ("num_workers", None), ("verbose", False), ("debug", False),
    ("keep_intermediate", False), ("timestamped_output_directory", False), ("simplify_tolerance", 0.2),
    ("building_shapes", None),
]


def _present(config, key):
    return bool(config.get(key)) and os.path.exists(config.get(key))


def get_config(config_path: str):
    """YAML -> (dict with defaults filled in, Config singleton).  Same keys, defaults and
    assertion messages as the reference (TreeDetection/config.py:144-238); ``ndvi_scaling_factor``,
    ``height_scaling_factor``, ``ndvi_mean_threshold`` and ``ndvi_var_threshold`` have no defaults
    there either."""
    config = load_config(config_path)
    for key, message in _REQUIRED:
        assert _present(config, key), message
    if not _present(config, "combined_model"):
        for key, message in _REQUIRED_WITHOUT_COMBINED_MODEL:
            assert _present(config, key), message
    for key, default in _DEFAULTS:
        config.setdefault(key, list(default) if isinstance(default, list) else default)
    for key in ("output_directory", "tiles_path"):
        if not config[key]:                      # as the reference: an empty path fails here
            os.makedirs(config[key], exist_ok=True)
    config.setdefault("continue", os.path.join(config["output_directory"], "continue.yml"))
    set_device_configuration(config, config.get("device", None))
    config["logger"] = setup_logging(os.path.join(config["output_directory"], "logs"), config["debug"])
    cfg = Config()
    cfg._load_into_config(config)
    return config, cfg
