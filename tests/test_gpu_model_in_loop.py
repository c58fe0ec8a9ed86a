"""N3: a live random-init Mask R-CNN between P1 and P2 on the device.  P1 tiles feed the
model where they lie, its RAW ROI-head outputs (boxes in network pixels, scores, 28x28 probabilities)
go to P2-P4 as device tensors; the crown table must equal what the oracle restatement makes of the
very same network outputs (dumped to the host as fixtures, as BASELINE.json config 1 prescribes)."""
import numpy as np
import pytest
import torch

from oracle import port
from treedetection_b200 import api, pipeline, predictor, synth

pytestmark = pytest.mark.gpu


def test_torchvision_model_in_the_loop(dev):
    pytest.importorskip("torchvision")
    sc = synth.make_scene(seed=3, size_px=500, px=0.2, ndsm_px=1.0, density_per_km2=4000.0)   # 2 x 2 tiles
    p = pipeline.PipelineParams()
    tables = api.TileTables(sc.tiles, dev, p.shift)
    rgbi = torch.from_numpy(sc.rgbi).to(dev)
    tiles_dev, tiles_off, flags = tables.plan(rgbi).run(rgbi)
    model = predictor.TorchvisionPredictor(device=dev, backbone="resnet18", batch_tiles=2, seed=0)
    det = model.forward(sc.stem, sc.tiles, tiles_dev, tiles_off, flags)
    assert det.boxes_net.is_cuda and det.probs.shape[1:] == (28, 28)
    n = int(det.scores.shape[0])
    assert n > 0, "a random-init ROI head with score threshold 0.3 yields detections"
    assert float(det.probs.min()) >= 0.0 and float(det.probs.max()) <= 1.0          # sigmoid probabilities
    assert bool((det.inst_tile[1:] >= det.inst_tile[:-1]).all())                     # tile-major
    tile_dims = torch.from_numpy(det.tile_dims).to(dev)
    table = pipeline.predict_stage(det.boxes_net, det.scores, det.probs, det.inst_tile, tile_dims, tables.tile_tf,
                                   tables.tile_boxes, p)
    # the same outputs as host fixtures through the oracle
    host = synth.Detections(det.boxes_net.cpu().numpy(), det.scores.cpu().numpy(), det.probs.cpu().numpy(),
                            det.inst_tile.cpu().numpy(), det.tile_dims, det.tile_ids, sc.tiles)
    rings, conf = port.predict_stage(host, sc.tiles, paste="torch")
    got_off = table.ring_off.cpu().numpy()
    got = table.verts.cpu().numpy()
    assert len(rings) == len(got_off) - 1
    for k, r in enumerate(rings):
        np.testing.assert_array_equal(got[got_off[k]:got_off[k + 1]], np.array(r).reshape(-1, 2))
    np.testing.assert_array_equal(table.conf.cpu().numpy(), np.array(conf, dtype=np.float64).reshape(-1))
