"""Whole-grid SGS with bounds on the GPU (gstatsim_custom/interpolate.sgs): for the same seed the realisation equals the
reference's (golden, where no distance tie straddles a cut-off) and the oracle's (deterministic tie order) to 1e-9;
batches of seeds equal single runs; realisations honour the data and the bounds."""
import os
import warnings

import numpy as np
import pytest

from cases import SGS_GRID_CASES, sgs_grid_inputs
from oracle import sgs_oracle as S

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
TOL = 1e-9


def _gpu(case, gi, seed):
    from mcmc_gpu_b200.gstatsim_custom import interpolate
    return interpolate.sgs(gi["xx"], gi["yy"], gi["cond"], gi["vario"], radius=case["radius"], num_points=case["num_points"],
                           bounds=gi["bounds"], seed=seed, quiet=True)


@pytest.mark.parametrize("name", sorted(SGS_GRID_CASES))
def test_same_seed_reproduces_oracle_and_reference(name):
    case = SGS_GRID_CASES[name]
    gi = sgs_grid_inputs(case)
    got = _gpu(case, gi, case["seed"])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ora, _ = S.sgs_grid(gi["xx"], gi["yy"], gi["cond"], gi["vario"], case["radius"], case["num_points"],
                            np.random.default_rng(case["seed"]), bounds=gi["bounds"])          # TIE_ORDER = "stable"
    scale = np.abs(ora).max()
    assert np.abs(got - ora).max() <= TOL * scale
    data = np.isfinite(gi["cond"])
    assert np.abs(got[data] - gi["cond"][data]).max() <= 1e-7 * scale       # conditioning data survive the normal-score round trip
    if name == "free":                      # isotropic, no tie at a cut-off: also the unmodified reference's realisation
        gold = np.load(os.path.join(GOLD, f"sgs_grid_{name}.npz"))["sim"]
        assert np.abs(got - gold).max() <= TOL * scale
    if case["bounds"]:
        free = ~data & (gi["bounds"][1] > -9000)
        assert (got[free] <= gi["bounds"][1][free] + 1e-6).all()
        pinned = ~data & (gi["bounds"][1] <= -9000)
        assert pinned.any() and np.allclose(got[pinned], np.nanmin(gi["cond"]))               # lower == upper: the bound itself


def test_batch_of_seeds_equals_single_runs_and_differs_between_seeds():
    from mcmc_gpu_b200.gstatsim_custom import interpolate
    case = SGS_GRID_CASES["bounded_k48"]
    gi = sgs_grid_inputs(case)
    many = interpolate.sgs_many(gi["xx"], gi["yy"], gi["cond"], gi["vario"], [3, 4, 5], radius=case["radius"],
                                num_points=case["num_points"], bounds=gi["bounds"])
    assert many.shape == (3,) + gi["cond"].shape and np.isfinite(many).all()
    for k, seed in enumerate((3, 4, 5)):
        assert np.array_equal(many[k], _gpu(case, gi, seed))
    assert np.abs(many[0] - many[1]).max() > 1.0


def test_sim_mask_and_loud_errors():
    from mcmc_gpu_b200.gstatsim_custom import interpolate
    case = SGS_GRID_CASES["free"]
    gi = sgs_grid_inputs(case)
    mask = np.zeros(gi["cond"].shape, dtype=bool)
    mask[5:20, 4:25] = True
    out = interpolate.sgs(gi["xx"], gi["yy"], gi["cond"], gi["vario"], radius=case["radius"], num_points=16, sim_mask=mask, seed=1)
    data = np.isfinite(gi["cond"])
    assert np.isfinite(out[mask | data]).all() and np.isnan(out[~mask & ~data]).all()
    with pytest.raises(ValueError):
        interpolate.sgs(gi["xx"], gi["yy"], gi["cond"], {"vtype": "matern"}, seed=1)
    with pytest.raises(NotImplementedError):
        interpolate.sgs(gi["xx"], gi["yy"], gi["cond"], gi["vario"], ktype="sk", seed=1)


def test_search_radius_is_widened_like_the_reference_when_a_node_finds_no_data():
    """Data in one corner only and a 1.5 km radius: most nodes find nothing at first and search again with
    radius + 100 km (interpolate.py:149-155).  Same seed -> the oracle's realisation."""
    case = SGS_GRID_CASES["free"]
    gi = sgs_grid_inputs(case)
    corner = np.full(gi["cond"].shape, np.nan)
    corner[:6, :7] = gi["cond"][:6, :7]
    corner[2, 3], corner[4, 1], corner[0, 5] = 310.0, 295.0, 330.0
    from mcmc_gpu_b200.gstatsim_custom import interpolate
    got = interpolate.sgs(gi["xx"], gi["yy"], corner, gi["vario"], radius=1.5e3, num_points=16, seed=5)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        ora, _ = S.sgs_grid(gi["xx"], gi["yy"], corner, gi["vario"], 1.5e3, 16, np.random.default_rng(5))
    assert np.isfinite(got).all()
    assert np.abs(got - ora).max() <= TOL * np.abs(ora).max()
