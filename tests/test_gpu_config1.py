"""BASELINE config 1 on the B200: the reference's example configuration over the bundled nDSM
tile (pixels in tests/golden/ndsm_324125317.npz) + synthetic RGBI, checked against the outputs
of the REFERENCE'S OWN process_features (tests/golden/make_golden_config1.py)."""
import os

import numpy as np
import pytest
import torch

from treedetection_b200 import pipeline, synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_config1_bundled_tile(dev):
    g = np.load(os.path.join(G, "config1.npz"))
    ndsm = np.load(os.path.join(G, "ndsm_324125317.npz"))["ndsm"]
    sc = synth.make_scene(seed=int(g["seed"][0]), size_px=5000, px=0.2, ndsm_px=1.0, density_per_km2=2500.0,
                          stem="324125317", left=412000.0, bottom=5318000.0 - 1000.0)
    assert len(sc.det.scores) == int(g["n_instances"][0]), "synthetic fixtures are not reproducible on this box"
    p = pipeline.PipelineParams(containment_threshold=0.75, iou_threshold=0.6)
    d = sc.det
    tile_tf, boxes_int = pipeline.tile_tables(sc.tiles, dev)
    table = pipeline.predict_stage(
        torch.from_numpy(d.boxes_net).to(dev), torch.from_numpy(d.scores).to(dev), torch.from_numpy(d.probs).to(dev),
        torch.from_numpy(d.inst_tile).to(dev), torch.from_numpy(d.tile_dims).to(dev), tile_tf,
        pipeline.filter_boxes(boxes_int, 1, dev), p)
    np.testing.assert_array_equal(table.ring_off.cpu().numpy(), g["stitched_off"])
    np.testing.assert_array_equal(table.conf.cpu().numpy(), g["stitched_conf"])
    np.testing.assert_array_equal(table.verts.cpu().numpy()[:64], g["stitched_first"])
    rasters = pipeline.raster_stage(torch.from_numpy(sc.rgbi).to(dev), sc.transform, torch.from_numpy(ndsm).to(dev),
                                    (1.0, 0.0, 412000.0, 0.0, -1.0, 5318000.0), p)
    f = pipeline.postprocess_stage(table, rasters, p, keep_debug=True)
    assert f.extras["combined"]        # 1 m nDSM + 0.2-decimated NDVI share one grid: the combined statistics path
    np.testing.assert_array_equal(f.extras["pid_after_nms"].cpu().numpy(), g["ids_after_nms"])
    np.testing.assert_array_equal(f.poly_id.cpu().numpy(), g["out_poly_id"])
    np.testing.assert_array_equal(f.area.cpu().numpy(), g["out_area"])
    np.testing.assert_array_equal(f.tree_height.cpu().numpy(), g["out_height"])
    np.testing.assert_array_equal(f.centroid.cpu().numpy().astype(np.float64), g["out_centroid"])
    np.testing.assert_array_equal(f.is_contained.cpu().numpy().astype(bool), g["out_is_contained"])
    np.testing.assert_array_equal(f.num_contained.cpu().numpy(), g["out_num_contained"])
    np.testing.assert_array_equal(f.ring_off.cpu().numpy(), g["out_off"])
    np.testing.assert_array_equal(f.verts.cpu().numpy(), g["out_verts"])
    assert len(f) > 1000
