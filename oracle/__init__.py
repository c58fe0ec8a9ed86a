"""CPU oracle for the post-model crown pipeline -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import it, and only as the checker / the CPU
baseline -- never as the thing shipped.  ``treedetection_b200`` never imports
this package.

Layers
------
``oracle.refshim``   (container only) loads the *real* reference modules from
                     ``/root/reference/TreeDetection`` over a NumPy-backed
                     ``cupy`` shim and minimal shapely/rasterio/affine
                     stand-ins, so the reference's own functions execute on CPU.
                     Used to generate ``tests/golden`` and to pin the port.
``oracle.port``      NumPy / torch-CPU / cv2 restatement of every §8 row
                     (SURVEY.md), each function citing the reference file:line
                     it follows.  This is what travels to the GPU box.
``oracle.geom``      GEOS-semantics geometry used by both (Polygon, simplify,
                     area, bounds) -- shapely/GEOS are absent in this image, so
                     this restatement *defines* those results (parity unpinned
                     at the third-party boundary).
``oracle.synth``     seeded synthetic mosaics / nDSM / detection fixtures
                     (SURVEY.md §8d).

Parity status: the reference ships no tests, golden vectors or fixtures for
this path, so the port is pinned against outputs of the
reference's own functions run here through ``oracle.refshim``
(``tests/golden/make_golden.py`` is the generating script).  Stages whose
arithmetic lives in absent third-party code (detectron2 paste, GEOS simplify,
GDAL decimation, rasterio window rounding) are restated from their published
algorithms: "parity unpinned" for those, stated again in DESIGN.md.
"""
