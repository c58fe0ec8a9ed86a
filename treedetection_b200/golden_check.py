"""Comparison of a device crown layer with the committed full-size golden of BASELINE config 2
(tests/golden/config2.npz, produced by the CPU oracle: tests/golden/make_golden_config2.py).
Used by tests/test_gpu_configs.py and by bench.py after its timed region (the bench's own crowns are
the golden's scene).  Reads the fixture only -- never the oracle."""
from __future__ import annotations

import hashlib
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "config2.npz")


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def golden_matches_workload(size_px, seed, density):
    if not os.path.exists(GOLDEN):
        return False
    with np.load(GOLDEN) as g:
        return [int(v) for v in g["workload"]] == [int(size_px), int(seed), int(density)]


def check_layer(host: dict, variant: str, n_candidates=None, atol_height=0.0):
    """host: dict of numpy arrays (api.features_to_host); variant: "split" (0.2 m nDSM) or "combined" (1 m).
    Raises AssertionError on the first difference; returns a short summary string."""
    with np.load(GOLDEN) as g:
        if n_candidates is not None:
            assert int(n_candidates) == int(g["n_candidates"][0]), \
                f"{n_candidates} candidate crowns, golden {int(g['n_candidates'][0])}"
        k = variant
        np.testing.assert_array_equal(host["poly_id"], g[f"{k}_poly_id"], err_msg="poly_id")
        np.testing.assert_array_equal(host["area"], g[f"{k}_area"], err_msg="Area")
        np.testing.assert_array_equal(host["tree_height"], g[f"{k}_height"], err_msg="TreeHeight")
        np.testing.assert_array_equal(host["centroid"], g[f"{k}_centroid"], err_msg="Centroid")
        np.testing.assert_array_equal(host["is_contained"].astype(bool), g[f"{k}_is_contained"], err_msg="is_contained")
        np.testing.assert_array_equal(host["num_contained"], g[f"{k}_num_contained"], err_msg="num_contained")
        np.testing.assert_array_equal(np.diff(host["ring_off"]).astype(np.int32), g[f"{k}_ring_len"], err_msg="ring lengths")
        assert _sha(host["verts"].astype(np.float64)) == str(g[f"{k}_verts_sha256"]), "vertex digest"
        return (f"{len(host['poly_id'])} crowns equal the CPU oracle's golden ({variant}): ids, areas, heights, "
                f"centroids, containment columns, ring lengths, vertex SHA-256")


def check_table(verts, ring_off, conf):
    """the stitched table (geojson_predictions/<image>.gpkg) against the golden"""
    with np.load(GOLDEN) as g:
        np.testing.assert_array_equal(np.diff(ring_off).astype(np.int32), g["table_ring_len"], err_msg="table ring lengths")
        np.testing.assert_array_equal(conf, g["table_conf"], err_msg="Confidence_score")
        assert _sha(np.asarray(verts, dtype=np.float64)) == str(g["table_verts_sha256"]), "table vertex digest"
