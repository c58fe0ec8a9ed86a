"""treedetection_b200 -- B200-native post-model crown pipeline of Jonetz/TreeDetection.

Public API mirrors the reference (``TreeDetection/detection.py``, ``config.py``):
``get_config``, ``preprocess_files``, ``predict_tiles``, ``postprocess_files``,
``process_files``.  The compute path is hand-written sm_100a CUDA behind the C-ABI of
``include/treedet.h`` (``csrc/libtreedet.so``); there is no CPU fallback.
"""
__version__ = "0.1.0"
