"""ctypes binding of ``libtreedet.so`` (the C-ABI declared in ``include/treedet.h``).

There is no CPU fallback: if the shared object is missing or a call fails, the
error is raised.  PyTorch tensors are only the interchange -- the library sees raw
device pointers, sizes and the current CUDA stream handle.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libtreedet.so")

_p = C.c_void_p
_i = C.c_int
_ll = C.c_longlong
_f = C.c_float
_d = C.c_double

# name -> (restype, argtypes).  Mirrors include/treedet.h one to one
# (tests/test_abi.py checks both directions).
SIGNATURES = {
    "td_version": (_i, []),
    "td_last_error": (C.c_char_p, []),
    "td_device_sms": (_i, []),
    # P2
    "td_paste_plan": (_i, [_p, _p, _p, _i, _i, _p, _p, _p, _p, _p]),
    "td_paste_threshold_pack": (_i, [_p, _p, _p, _p, _i, _f, _p, _p]),
    "td_paste_values": (_i, [_p, _p, _p, _p, _i, _p, _p]),
    # P3
    "td_trace_count": (_i, [_p, _p, _p, _i, _ll, _p, _p, _p, _p]),
    "td_trace_emit": (_i, [_p, _p, _p, _i, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _ll, _p, _p, _p, _p, _p,
                           _p]),
    "td_trace_walk": (_i, [_p, _p, _p, _i, _ll, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p]),
    "td_trace_rings": (_i, [_p, _i, _p, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    # P4 / P9 geometry
    "td_simplify_rings": (_i, [_p, _p, _i, _d, _p, _p, _p, _p, _p, _p, _p, _p, _i, _p, _p]),
    "td_take_rings": (_i, [_p, _p, _p, _i, _p, _p, _p, _p, _p]),
    # P5
    "td_ndvi_decimate": (_i, [_p, _i, _i, _i, _i, _i, _p, _p]),
    "td_decimate_f32": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    # P6 / P8
    "td_bbox_nms_ordered": (_i, [_p, _p, _p, _i, _d, _d, _p, _p]),
    "td_bbox_nms_ordered_dyn": (_i, [_p, _p, _p, _i, _p, _d, _d, _ll, _p, _p, _p]),
    "td_containment": (_i, [_p, _i, _d, _p, _p, _p, _p, _p]),
    "td_mask_iou_clean": (_i, [_p, _p, _p, _p, _p, _p, _i, _f, _f, _p, _p, _p, _p]),
    # P7
    "td_crown_stats": (_i, [_p, _p, _i, _p, _p, _i, _i, _p, _i, _p, _p, _p, _p, _p]),
    "td_centroids": (_i, [_p, _p, _i, _p, _p, _p]),
    "td_crown_height_summary": (_i, [_p, _p, _i, _p, _i, _i, _p, _d, _p, _p, _p]),
    # P9
    "td_select_crowns": (_i, [_p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p, _p]),
    "td_round_coords": (_i, [_p, _ll, _p, _p]),
    "td_select_head": (_i, [_p, _p, _i, _p, _d, _d, _d, _p, _p, _p]),
    # P1
    "td_tile_plan_create": (_i, [_p, _p, _p, _i, _i, _i, _i, _p]),
    "td_tile_plan_destroy": (_i, [_p]),
    "td_tile_cut_normalize": (_i, [_p, _p, _i, _p, _p, _p]),
    # P10
    "td_forest_predicates": (_i, [_p, _p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _p]),
    "td_ring_is_simple": (_i, [_p, _p, _i, _p, _p]),
    # device-side bookkeeping
    "td_scan_clamp": (_i, [_p, _i, _i, _p, _p, _p, _p, _p, _p]),
    "td_compact_flags": (_i, [_p, _i, _p, _p, _p, _p]),
    "td_compact_nonneg": (_i, [_p, _i, _p, _p, _p, _p]),
    "td_ring_tail": (_i, [_p, _p, _i, _p, _p, _p]),
    "td_ring_offsets": (_i, [_p, _p, _p, _i, _p, _p]),
    "td_gather_rows": (_i, [_p, _p, _p, _i, _p, _i, _p, _p]),
    # whole chain (CUDA-graph replay)
    "td_chain_workspace_bytes": (_ll, [_p, _i]),
    "td_chain_create": (_i, [_p, _p, _i, _p, _ll, _p]),
    "td_chain_destroy": (_i, [_p]),
    "td_chain_layout": (_i, [_p, _i, _p]),
    "td_chain_predict": (_i, [_p, _i, _p, _p, _p, _p, _i, _p, _p, _p, _i, _i, _p]),
    "td_chain_post": (_i, [_p, _i, _p, _i, _i, _p, _p, _i, _i, _p, _i, _p, _i, _p]),
    # N1 (host function)
    "td_tiff_lzw_decode": (_ll, [_p, _ll, _p, _ll]),
    "td_tiff_lzw_encode": (_ll, [_p, _ll, _p, _ll]),
    "td_tiff_lzw_decode_batch": (_i, [_p, _p, _p, _i, _p, _ll, _p, _p, _p, _p]),
    "td_tiff_place_chunks": (_i, [_p, _ll, _i, _p, _i, _i, _i, _i, _i, _i, _i, _i, _p]),
    # N2 (host function)
    "td_gpkg_append": (_i, [C.c_char_p, C.c_char_p, _i, _p, _p, _ll, _i, _p, _p, _p, _p]),
    # P0a
    "td_seam_crop": (_i, [_p, _p, _i, _i, _i, _i, _i, _i, _i, _i, _i, _p, _p]),
}

# hand-written kernels launched per call (library kernels -- CUB sort / scan -- not counted);
# bench.py reports the sum over the timed region as ``gpu_launches``
OWN_KERNELS = {
    "td_paste_plan": 1, "td_paste_threshold_pack": 1, "td_paste_values": 1, "td_trace_count": 1, "td_trace_emit": 2, "td_trace_walk": 1, "td_trace_rings": 1,
    "td_simplify_rings": 1, "td_take_rings": 1, "td_ndvi_decimate": 1, "td_decimate_f32": 1,
    "td_bbox_nms_ordered": 7, "td_bbox_nms_ordered_dyn": 7, "td_scan_clamp": 1, "td_compact_flags": 1, "td_compact_nonneg": 1,
    "td_ring_tail": 1, "td_ring_offsets": 1, "td_gather_rows": 1, "td_containment": 4, "td_mask_iou_clean": 5, "td_crown_stats": 1, "td_centroids": 2, "td_crown_height_summary": 1, "td_select_crowns": 2,
    "td_round_coords": 1, "td_select_head": 1, "td_chain_predict": 14, "td_chain_post": 28, "td_tile_cut_normalize": 1, "td_seam_crop": 1, "td_tile_plan_create": 0, "td_forest_predicates": 1, "td_ring_is_simple": 1,
    "td_tiff_lzw_encode": (_ll, [_p, _ll, _p, _ll]),
    "td_tiff_lzw_decode_batch": 1, "td_tiff_place_chunks": 1,
}
launch_count = 0

_lib = None


class TreedetError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load the library (once).  Raises when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TreedetError(
            f"{LIB_PATH} is missing: run `python -m treedetection_b200.build` (there is no CPU fallback)")
    handle = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        try:
            fn = getattr(handle, name)
        except AttributeError:
            continue  # reported by exported_symbols() / test_abi
        fn.restype = res
        fn.argtypes = args
    _lib = handle
    return handle


def exported_symbols():
    """Names from SIGNATURES the shared object really exports."""
    h = lib()
    out = []
    for name in SIGNATURES:
        try:
            getattr(h, name)
            out.append(name)
        except AttributeError:
            pass
    return out


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().td_last_error()
        raise TreedetError(f"{what} failed with code {rc}: {msg.decode() if msg else ''}")


def call(name: str, *args):
    global launch_count
    fn = getattr(lib(), name)
    check(fn(*args), name)
    launch_count += OWN_KERNELS.get(name, 0)


_cuda_ok = None


def cuda_available():
    """``torch.cuda.is_available()`` asked once per process (every call costs a driver round trip of ~20 ms)"""
    global _cuda_ok
    if _cuda_ok is None:
        import torch
        _cuda_ok = bool(torch.cuda.is_available())
    return _cuda_ok
