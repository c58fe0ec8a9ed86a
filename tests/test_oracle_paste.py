"""P2 oracle: the closed-form float32 restatement of detectron2's ``_do_paste_mask`` + ATen ``grid_sample``
(``port.paste_probs_closed_form`` -- the operation order the CUDA kernel reproduces) against the same composition
run through torch's own CPU ``grid_sample`` (``port.paste_probs``).  CPU only; the kernel is compared with both in
tests/test_gpu_kernels.py::test_paste_bits_and_values."""
import numpy as np
import pytest

from oracle import port
from treedetection_b200 import synth


@pytest.fixture(scope="module")
def scene():
    return synth.make_scene(seed=7, size_px=1500, px=0.2, ndsm_px=1.0, density_per_km2=4000.0)


def test_closed_form_equals_torch_grid_sample(scene):
    d = scene.det
    n = d.boxes_net.shape[0]
    assert n > 200
    worst, differing = 0.0, 0
    for i in range(0, n, 2):
        th, tw, nh, nw = d.tile_dims[d.inst_tile[i]]
        box, keep = port.scale_clip_boxes(d.boxes_net[i:i + 1], (nh, nw), (th, tw))
        assert keep[0]
        v_t, win_t = port.paste_probs(box[0], d.probs[i], th, tw)
        v_c, win_c = port.paste_probs_closed_form(box[0], d.probs[i], th, tw)
        assert win_t == win_c and v_t.shape == v_c.shape
        # the thresholded mask -- what the path consumes -- is identical
        np.testing.assert_array_equal(v_c >= np.float32(0.5), v_t >= np.float32(0.5))
        worst = max(worst, float(np.abs(v_c - v_t).max()))
        differing += int((v_c != v_t).sum())
    # float32 values: ATen's vectorised path may contract differently on some CPUs; 1e-6 absolute at most
    assert worst <= 1e-6, (worst, differing)


def test_window_and_empty_boxes():
    """paste_window: integer window of the box clipped to the tile (floor(x0 - 1) .. ceil(x1 + 1));
    boxes that detector_postprocess drops as empty"""
    assert port.paste_window(np.array([10.2, 20.7, 30.1, 40.9], np.float32), 450, 450) == (9, 19, 32, 42)
    assert port.paste_window(np.array([-5.0, -5.0, 500.0, 500.0], np.float32), 450, 450) == (0, 0, 450, 450)
    b, keep = port.scale_clip_boxes(np.array([[10, 10, 10, 50], [5, 5, 60, 70]], np.float32), (800, 800), (450, 450))
    assert list(keep) == [False, True]
