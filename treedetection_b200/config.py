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


def set_device_configuration(config, raw_device):
    """config["device"]: GPU index string, or "cpu" when CUDA is absent (config.py:112-142).
    NOTE: with "cpu" this implementation refuses to run the kernels -- there is no CPU path."""
    import torch
    if torch.cuda.is_available():
        if raw_device is not None:
            device_str = "0"
            if isinstance(raw_device, int):
                device_str = raw_device
            elif isinstance(raw_device, str) and raw_device.startswith("cuda"):
                if raw_device.replace("cuda:", "").isdigit():
                    device_str = raw_device.replace("cuda:", "")
                elif raw_device == "cuda":
                    device_str = "0"
            elif isinstance(raw_device, str) and raw_device.isdigit():
                device_str = raw_device
            try:
                gpu_index = int(device_str)
                assert torch.cuda.device_count() > gpu_index, f"GPU index {gpu_index} is out of range."
            except (IndexError, ValueError):
                raise ValueError(f"Invalid CUDA device specification: {raw_device}")
            config["device"] = str(device_str)
        else:
            config["device"] = "0"
    else:
        if isinstance(raw_device, str) and raw_device.startswith("cuda"):
            warnings.warn(f"CUDA device '{raw_device}' requested but CUDA is not available. Falling back to CPU.")
        config["device"] = "cpu"


def get_config(config_path: str):
    """YAML -> (dict with defaults filled in, Config singleton).  Same keys, defaults and
    assertion messages as the reference; ``ndvi_scaling_factor``, ``height_scaling_factor``,
    ``ndvi_mean_threshold`` and ``ndvi_var_threshold`` have no defaults there either."""
    config = load_config(config_path)

    assert config.get("image_directory") and os.path.exists(config.get("image_directory")), \
        "Input path is missing from the configuration or path is incorrect."
    assert config.get("height_data_path") and os.path.exists(config.get("height_data_path")), \
        "nDOM path is missing from the configuration or path is incorrect."

    if not config.get("combined_model") or not os.path.exists(config.get("combined_model")):
        assert config.get("urban_model") and os.path.exists(config.get("urban_model")), \
            "Urban model path is missing from the configuration or path is incorrect."
        assert config.get("forrest_model") and os.path.exists(config.get("forrest_model")), \
            "Forrest model path is missing from the configuration."
        assert config.get("forrest_outline") and os.path.exists(config.get("forrest_outline")), \
            "Forrest outline path is missing from the configuration."

    config["output_directory"] = config.get("output_directory", "./output")
    if not config["output_directory"]:
        os.makedirs(config["output_directory"], exist_ok=True)
    config["tiles_path"] = config.get("tiles_path", "./tiles")
    if not config["tiles_path"]:
        os.makedirs(config["tiles_path"], exist_ok=True)
    config["continue"] = config.get("continue", os.path.join(config["output_directory"], "continue.yml"))

    config["tile_width"] = config.get("tile_width", 50)
    config["tile_height"] = config.get("tile_height", 50)
    config["buffer"] = config.get("buffer", 20)
    config["batch_size"] = config.get("batch_size", 10)

    config["use_overlap"] = config.get("use_overlap", True)
    config["overlapping_tiles_width"] = config.get("overlapping_tiles_width", 3)
    config["overlapping_tiles_height"] = config.get("overlapping_tiles_height", 3)
    config["merged_path"] = config.get("merged_path", "merged")
    config["image_merged_regex"] = config.get("image_merged_regex", "FDOP20_(\\d+)_(\\d+)_(\\d+)_(\\d+)_(\\d+)\\.tif")
    config["height_data_merged_regex"] = config.get("height_data_merged_regex", "FDOP20_(\\d+)_(\\d+)\\.tif")

    config["iou_threshold"] = config.get("iou_threshold", 0.5)
    config["confidence_threshold_stitching"] = config.get("confidence_threshold_stitching", 0.3)
    config["area_threshold"] = config.get("area_threshold", 1)

    config["exclude_files"] = config.get("exclude_files", [])
    config["confidence_threshold"] = config.get("confidence_threshold", 0.3)
    config["containment_threshold"] = config.get("containment_threshold", 0.9)
    config["height_threshold"] = config.get("height_threshold", 3)

    raw_device = config.get("device", None)
    set_device_configuration(config, raw_device)

    config["parallel"] = config.get("parallel", True)
    config["num_workers"] = config.get("num_workers", None)
    config["verbose"] = config.get("verbose", False)
    config["debug"] = config.get("debug", False)
    config["logger"] = setup_logging(os.path.join(config["output_directory"], "logs"), config["debug"])
    config["keep_intermediate"] = config.get("keep_intermediate", False)
    config["timestamped_output_directory"] = config.get("timestamped_output_directory", False)
    config["simplify_tolerance"] = config.get("simplify_tolerance", 0.2)

    config["building_shapes"] = config.get("building_shapes", None)

    config_obj = Config()
    config_obj._load_into_config(config)
    return config, config_obj
