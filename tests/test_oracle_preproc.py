"""Region-mask preprocessing oracle (Topography.get_highvel_boundary): pinned by the reference's own output and, for the
mode filter, by PIL (the reference's third-party dependency) where it is installed."""
import os

import numpy as np
import pytest

from cases import highvel_case_inputs
from oracle import preproc_oracle as P

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def test_highvel_boundary_matches_reference_golden():
    hb = highvel_case_inputs()
    got = P.highvel_boundary(hb["velx"], hb["vely"], hb["threshold"], hb["grounded"], hb["ocean"], hb["distance_max"], hb["xx"],
                             hb["yy"], smooth_mode=hb["smooth_mode"])
    gold = np.load(os.path.join(GOLD, "highvel_boundary.npz"))["mask_final"]
    assert got.dtype == gold.dtype and np.array_equal(got, gold)
    assert 0 < gold.sum() < gold.size


@pytest.mark.parametrize("size", [1, 3, 4, 5, 10, 15])
def test_mode_filter_restatement_matches_pil(size):
    Image = pytest.importorskip("PIL.Image")
    from PIL import ImageFilter
    g = np.random.default_rng(size)
    for shape, p in [((12, 15), 0.5), ((40, 33), 0.3), ((2, 3), 0.5), ((1, 1), 1.0), ((25, 60), 0.8)]:
        img = ((g.random(shape) < p) * 255).astype(np.uint8)
        want = np.array(Image.fromarray(img).filter(ImageFilter.ModeFilter(size=size)))
        assert np.array_equal(P.mode_filter_binary(img, size), want), (shape, size)
