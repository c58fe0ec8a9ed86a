"""P10 -- two-model fusion against a forest outline (config 3).

Reference: ``fuse_predictions`` (TreeDetection/helpers.py:703-834), tile flags
``only_forest`` / ``only_urban`` (TreeDetection/preprocessing.py:67-96) and the per-model
tile exclusion (TreeDetection/prediction.py:79-93).

STATUS (round 1): not built yet.  The per-model tile exclusion is implemented
(``predictor.FixturePredictor(exclude_vars=...)``); the polygon predicates against the
union of the forest polygons (``intersects`` / ``within`` / ``contains``, GEOS in the
reference) are the remaining piece and are listed as open in DESIGN.md.  Calling into this
module fails loudly rather than producing an unverified result."""
from __future__ import annotations


class ForestIndex:
    @classmethod
    def from_file(cls, path):
        raise NotImplementedError("forrest_outline (two-model fusion, SURVEY P10) is not built yet")

    def flags(self, minx, miny, maxx, maxy, bounds):
        raise NotImplementedError("forrest_outline (two-model fusion, SURVEY P10) is not built yet")


def fuse_predictions(urban_fold, forrest_fold, forrest_outline, output_fold, logger=None):
    raise NotImplementedError("fuse_predictions (two-model fusion, SURVEY P10) is not built yet")
