"""Host-side logic of the API mirror that needs no GPU: block tables, tapers, weights, seeds, error behaviour."""
import numpy as np
import pytest

from cases import TRAJECTORY_CASES, build_case_grids
from gpu_helpers import bits_equal, quiet
from oracle import crf_oracle as O


def _rf(case):
    from mcmc_gpu_b200 import MCMC
    kw = case["rf_kw"]
    rf = quiet(MCMC.RandField, kw["range_min_x"], kw["range_max_x"], kw["range_min_y"], kw["range_max_y"], kw["scale_min"],
               kw["scale_max"], kw["nugget_max"], kw["model_name"], kw["isotropic"], smoothness=kw.get("smoothness"),
               rng_seed=case["rf_seed"])
    rf.set_block_sizes(*case["blocks"])
    return rf


@pytest.mark.parametrize("name", sorted(TRAJECTORY_CASES))
def test_pairs_tapers_and_weights_equal_the_oracle(name):
    case = TRAJECTORY_CASES[name]
    g = build_case_grids(case)
    rf = _rf(case)
    rf.set_weight_param(*case["logistic"], case["max_dist"], g["resolution"])
    pairs = O.block_size_pairs(*case["blocks"])
    assert np.array_equal(rf.pairs, pairs)
    ref = O.edge_taper_masks(pairs, case["logistic"], case["max_dist"], g["resolution"])   # KD-tree, like the reference
    for a, b in zip(rf.edge_masks, ref):
        assert bits_equal(a, b)
    w = rf.get_crf_weight(g["xx"], g["yy"], g["data_mask"])[0]
    assert bits_equal(w, O.crf_data_weight(g["xx"], g["yy"], g["data_mask"], tuple(case["logistic"]), case["max_dist"]))


def test_taper_with_awkward_resolution():
    from mcmc_gpu_b200 import MCMC
    rf = quiet(MCMC.RandField, 1.0, 2.0, 1.0, 2.0, 1.0, 2.0, 0.0, "Gaussian", True, rng_seed=0)
    rf.set_block_sizes(6, 22, 8, 30, steps=4)
    rf.set_weight_param(1.0, 0.3, 5.0, 0.1, 0.7, 0.1)          # res = 0.1: j*res differences are not exact
    ref = O.edge_taper_masks(rf.pairs, (1.0, 0.3, 5.0, 0.1), 0.7, 0.1)
    for a, b in zip(rf.edge_masks, ref):
        assert bits_equal(a, b)


def test_constructor_and_setter_errors_match_the_reference():
    from mcmc_gpu_b200 import MCMC
    with pytest.raises(Exception, match="valid model_name"):
        MCMC.RandField(1, 2, 1, 2, 1, 2, 0, "Cubic", True)
    with pytest.raises(Exception, match="smoothness"):
        MCMC.RandField(1, 2, 1, 2, 1, 2, 0, "Matern", True)
    with pytest.raises(ValueError):
        MCMC.RandField(1, 2, 1, 2, 1, 2, 0, "Gaussian", True, rng_seed="abc")
    rf = quiet(MCMC.RandField, 1, 2, 1, 2, 1, 2, 0, "Gaussian", True)
    with pytest.raises(Exception, match="set_block_sizes"):
        rf.set_weight_param(2, 0, 6, 1, 1.0, 1.0)
    g = build_case_grids(TRAJECTORY_CASES["ragged_rf"])
    with pytest.raises(Exception, match="shape"):
        MCMC.chain_crf(g["xx"], g["yy"], g["bed0"], g["surf"][:-1], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"],
                       g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    ch = quiet(MCMC.chain_crf, g["xx"], g["yy"], g["bed0"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"],
               g["cond_bed"], g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    with pytest.raises(ValueError):
        ch.set_update_region(True, np.ones((3, 3)))
    with pytest.raises(ValueError):
        ch.set_update_type("nope")
    with pytest.raises(ValueError):
        ch.set_random_generator(1.5)
    quiet(ch.set_update_region, True, g["highvel_mask"] * 2)
    ch.set_loss_type(5.0, True)
    with pytest.raises(ValueError, match="0/1"):
        ch._static_args()


def test_philox_keys_are_distinct_and_stable():
    from mcmc_gpu_b200.MCMC import philox_key
    keys = {philox_key(s, s) for s in range(4096)}
    assert len(keys) == 4096
    assert philox_key(5, 5) == philox_key(5) and philox_key(5, 6) != philox_key(5, 5)
