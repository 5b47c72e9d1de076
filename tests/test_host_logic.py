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
    # the logistic on a given distance map is host logic; the distance map itself is a GPU kernel (tests/test_gpu_residual.py)
    from scipy.spatial import cKDTree
    sel = g["data_mask"] == 1
    dist = cKDTree(np.column_stack([g["xx"][sel], g["yy"][sel]])).query(
        np.column_stack([g["xx"].ravel(), g["yy"].ravel()]))[0].reshape(g["xx"].shape)
    w = rf.get_crf_weight_from_dist(g["xx"], g["yy"], dist)[0]
    assert bits_equal(w, O.crf_data_weight(g["xx"], g["yy"], g["data_mask"], tuple(case["logistic"]), case["max_dist"]))


def test_crf_weight_has_no_cpu_fallback():
    """north_star: no CPU fallback - without a CUDA device the weight setup raises instead of quietly using scipy."""
    import torch
    from mcmc_gpu_b200._lib import GmcError
    if torch.cuda.is_available():
        pytest.skip("needs a host without a CUDA device")
    case = TRAJECTORY_CASES[sorted(TRAJECTORY_CASES)[0]]
    g = build_case_grids(case)
    rf = _rf(case)
    rf.set_weight_param(*case["logistic"], case["max_dist"], g["resolution"])
    with pytest.raises(GmcError):
        rf.get_crf_weight(g["xx"], g["yy"], g["data_mask"])


def test_context_cache_key_follows_the_randfield_contents():
    """a setter call on RF (or a re-created RF) must change the key the chain caches its device context under."""
    case = TRAJECTORY_CASES[sorted(TRAJECTORY_CASES)[0]]
    g = build_case_grids(case)
    rf = _rf(case)
    rf.set_weight_param(*case["logistic"], case["max_dist"], g["resolution"])
    k0 = rf._config_key()
    assert _rf_same(case, g)._config_key() == k0                   # same contents, different object: same key
    rf.set_block_sizes(case["blocks"][0] + 2, *case["blocks"][1:])
    rf.set_weight_param(*case["logistic"], case["max_dist"], g["resolution"])
    assert rf._config_key() != k0
    rf2 = _rf_same(case, g)
    L = list(case["logistic"])
    rf2.set_weight_param(L[0], L[1], L[2] + 1.0, L[3], case["max_dist"], g["resolution"])
    assert rf2._config_key() != k0


def _rf_same(case, g):
    rf = _rf(case)
    rf.set_weight_param(*case["logistic"], case["max_dist"], g["resolution"])
    return rf


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


def test_sgs_octant_stencil_reproduces_the_reference_search():
    """The host-built, pre-sorted octant offset lists select exactly the neighbours of neighbors.py:4-64 (stable ties)."""
    from mcmc_gpu_b200 import sgs_tables as T
    from oracle import sgs_oracle as S
    off, cnt, hw = T.octant_stencil(500.0, 500.0, 4e3)
    H = W = 41
    xx, yy = np.meshgrid(np.arange(W) * 500.0, np.arange(H) * 500.0)
    rng = np.random.default_rng(0)
    grid = rng.standard_normal((H, W))
    cond = rng.random((H, W)) < 0.7
    grid[~cond] = np.nan
    for (i, j) in [(20, 20), (0, 5), (40, 40), (3, 38), (17, 0)]:
        cond[i, j] = False
        grid[i, j] = np.nan
        nb = S.octant_neighbors(i, j, xx, yy, grid, cond, 4e3, 24, hw)
        mine = []
        for o in range(8):
            k = 0
            for t in range(cnt[o]):
                ci, cj = i + off[o, t, 0], j + off[o, t, 1]
                if 0 <= ci < H and 0 <= cj < W and cond[ci, cj]:
                    mine.append((ci, cj))
                    k += 1
                    if k == 3:
                        break
        assert np.array_equal(np.array(mine), nb[:, 3:5].astype(int)), (i, j)


def test_sgs_covariance_lut_matches_the_kriging_matrices():
    from scipy.spatial.distance import pdist, squareform
    from mcmc_gpu_b200 import sgs_tables as T
    from oracle import sgs_oracle as S
    for vario in (dict(azimuth=30.0, nugget=0.0, major_range=4000.0, minor_range=2500.0, sill=900.0, vtype="Exponential"),
                  dict(azimuth=0, nugget=0.1, major_range=3000.0, minor_range=3000.0, sill=1.0, vtype="Matern", s=1.2259),
                  dict(azimuth=75.0, nugget=0.0, major_range=2000.0, minor_range=900.0, sill=2.0, vtype="Spherical"),
                  dict(azimuth=0, nugget=0.0, major_range=1500.0, minor_range=1500.0, sill=1.0, vtype="Gaussian")):
        hw = 6
        lut = T.covariance_lut(500.0, 250.0, hw, vario)
        rng = np.random.default_rng(1)
        pts = rng.integers(-hw, hw + 1, size=(12, 2))                        # (di, dj) offsets of neighbours
        xy = np.stack([pts[:, 1] * 500.0, pts[:, 0] * 250.0], axis=1)
        R = S.rotation_matrix(vario["azimuth"], vario["major_range"], vario["minor_range"])
        Sigma = S.covariance(vario["vtype"], squareform(pdist(xy @ R)), vario["sill"], vario["nugget"], vario.get("s"))
        mine = np.array([[lut[a[0] - b[0] + 2 * hw, a[1] - b[1] + 2 * hw] for b in pts] for a in pts])
        assert np.allclose(mine, Sigma, rtol=1e-12, atol=1e-13), vario["vtype"]


def test_whole_grid_sgs_argument_checks_and_no_cpu_fallback():
    """gstatsim_custom.interpolate.sgs validates like the reference (interpolate.py:262-330) and never computes on the CPU."""
    import torch
    from mcmc_gpu_b200._lib import GmcError
    from mcmc_gpu_b200.gstatsim_custom import interpolate
    xx, yy = np.meshgrid(np.arange(6) * 100.0, np.arange(5) * 100.0)
    grid = np.full(xx.shape, np.nan)
    grid[1, 2], grid[3, 4] = 1.0, 2.0
    vario = dict(azimuth=0.0, nugget=0.0, major_range=300.0, minor_range=300.0, sill=1.0, vtype="exponential")
    with pytest.raises(ValueError, match="Variogram missing"):
        interpolate.sgs(xx, yy, grid, {"vtype": "exponential"})
    with pytest.raises(ValueError, match="same shape"):
        interpolate.sgs(xx, yy[:-1], grid, vario)
    with pytest.raises(ValueError, match="Matern covariance requires"):
        interpolate.sgs(xx, yy, grid, dict(vario, vtype="matern"))
    with pytest.raises(ValueError, match="sim_mask"):
        interpolate.sgs(xx, yy, grid, vario, sim_mask=np.ones((2, 2), dtype=bool))
    with pytest.raises(NotImplementedError):
        interpolate.sgs(xx, yy, grid, vario, ktype="sk")
    with pytest.raises(NotImplementedError):
        interpolate.sgs(xx, yy, grid, vario, num_points=64)
    with pytest.raises(NotImplementedError):
        interpolate.sgs(xx, yy, grid, dict(vario, sill=np.ones(xx.shape)))
    if not torch.cuda.is_available():
        with pytest.raises(GmcError, match="no CPU fallback"):
            interpolate.sgs(xx, yy, grid, vario, radius=500.0, num_points=8, seed=1)


def test_whole_grid_sgs_search_levels_are_prefixes_of_the_widest_lists():
    from mcmc_gpu_b200.gstatsim_custom.interpolate import search_levels
    from mcmc_gpu_b200.sgs_tables import octant_stencil
    off, cnt, hw, radii = search_levels(500.0, 500.0, 120, 150, 3e3)
    assert radii == [3e3, 103e3] and cnt.shape == (2, 8) and (cnt[1] >= cnt[0]).all()
    off0, cnt0, hw0 = octant_stencil(500.0, 500.0, 3e3)
    assert np.array_equal(cnt[0], cnt0)                                  # level 0 is exactly the search for `radius`
    for o in range(8):
        assert np.array_equal(off[o, :cnt0[o]], off0[o, :cnt0[o]])       # ... and its offsets are a prefix of the wide lists
    # a radius that already covers the grid needs no second level
    assert len(search_levels(500.0, 500.0, 20, 20, 50e3)[3]) == 1
