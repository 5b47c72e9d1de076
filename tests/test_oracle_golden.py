"""The numpy oracle must reproduce the UNMODIFIED reference bit-for-bit (fixtures from oracle/make_golden.py)."""
import os

import numpy as np
import pytest

from cases import FIELD_CASES, TRAJECTORY_CASES, build_case_grids, residual_case_inputs
from oracle import crf_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _eq(a, b):
    return np.array_equal(np.asarray(a), np.asarray(b), equal_nan=True)


def test_residual_and_loss_match_reference():
    ri = residual_case_inputs()
    gold = np.load(os.path.join(GOLD, "residual_loss.npz"))
    res = O.mass_conservation_residual(ri["bed"], ri["surf"], ri["velx"], ri["vely"], ri["dhdt"], ri["smb"],
                                       ri["resolution"])
    assert _eq(res, gold["residual"])
    assert np.isnan(res).sum() >= 3          # NaN inputs propagate to their stencil neighbours
    loss = O.masked_loss(res, ri["mask"], ri["sigma_mc"])
    assert _eq(np.array(loss, dtype=np.float64), gold["loss"])


@pytest.mark.parametrize("name", sorted(FIELD_CASES))
def test_spectral_field_matches_reference(name):
    fc = FIELD_CASES[name]
    gold = np.load(os.path.join(GOLD, "spectral_fields.npz"))
    fp = O.FieldParams(**fc["rf_kw"], resolution=fc["res"])
    rng = np.random.default_rng(fc["seed"])
    for k, shape in enumerate(fc["shapes"]):
        fld, _ = O.spectral_field(fp, rng, tuple(shape))
        assert _eq(fld, gold[f"{name}__{k}"]), (name, shape)


@pytest.mark.parametrize("name", sorted(TRAJECTORY_CASES))
def test_trajectory_matches_reference(name):
    case = TRAJECTORY_CASES[name]
    gold = np.load(os.path.join(GOLD, f"traj_{name}.npz"))
    g = build_case_grids(case)
    cs, fp = O.setup_from_grids(g, sigma_mc=case["sigma_mc"], logistic=case["logistic"], max_dist=case["max_dist"],
                                blocks=case["blocks"], rf_kw=case["rf_kw"], update_in_region=case["update_in_region"],
                                block_type=case["block_type"])
    assert _eq(fp.pairs, gold["pairs"])
    assert _eq(fp.edge_masks[0], gold["edge_mask0"]) and _eq(fp.edge_masks[-1], gold["edge_mask_last"])
    assert _eq(cs.crf_weight, gold["crf_weight"])
    out = O.run_chain(cs, fp, g["bed0"], case["n_iter"], np.random.default_rng(case["chain_seed"]),
                      np.random.default_rng(case["rf_seed"]))
    for key in ("bed", "loss", "loss_mc", "loss_data", "steps", "resampled_times", "blocks"):
        assert _eq(out[key], gold[key]), key
    assert 0.05 < out["steps"].mean() < 0.95


def test_tutorial_anchor_value():
    """the survey's independent run of the reference printed this final loss."""
    gold = np.load(os.path.join(GOLD, "traj_tutorial200.npz"))
    assert gold["loss"][-1] == 348.2912467039959
