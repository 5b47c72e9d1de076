"""A5 (randomization-method proposal) on the CPU: the closed-form radial inversion and the anisotropy mapping are
checked against the published correlation functions of the models — E[cos(k . d)] over the sampled wave vectors is the
covariance of the generated field — and the summation against a literal double loop.  gstools itself is absent
(parity unpinned, see oracle/randmeth_oracle.py)."""
import numpy as np
import pytest

from oracle import randmeth_oracle as R
from philox_ref import randmeth_modes


@pytest.mark.parametrize("model,nu", [("Gaussian", 1.0), ("Exponential", 1.0), ("Matern", 0.9), ("Matern", 2.5)])
def test_sampled_wave_vectors_reproduce_the_model_correlation(model, nu):
    g = np.random.default_rng(3)
    n = 4_000_000
    range_x, range_y, angle = 30e3, 12e3, 35.0
    kx, ky = R.grid_wave_vectors(model, g.random(n), g.random(n), range_x, range_y, angle, nu)
    l1, l2 = R.model_lengths(model, range_x, range_y)
    th = np.deg2rad(angle)
    for d in [(2e3, 0.0), (0.0, 5e3), (7e3, -4e3), (-1.5e4, 9e3)]:
        # isometrized lag: rotate into the main-axis frame, stretch the minor axis by 1 / anis = l1 / l2
        a = np.cos(th) * d[0] + np.sin(th) * d[1]
        b = (-np.sin(th) * d[0] + np.cos(th) * d[1]) * (l1 / l2)
        want = float(R.correlation(model, np.hypot(a, b), l1, nu))
        got = float(np.cos(kx * d[0] + ky * d[1]).mean())
        assert abs(got - want) < 4.0 / np.sqrt(n) * 1.5 + 1e-3, (model, d, got, want)


def test_ppf_inverts_the_numerically_integrated_radial_density():
    for model, nu, dens in [("Gaussian", 1.0, lambda r, l: np.exp(-(r * l / np.sqrt(np.pi)) ** 2)),
                            ("Exponential", 1.0, lambda r, l: (1 + (r * l) ** 2) ** -1.5),
                            ("Matern", 1.3, lambda r, l: (1 + (r * l) ** 2 / 1.3) ** -2.3)]:
        l = 4e3
        r = np.linspace(0, 60.0 / l, 2_000_001)
        pdf = r * dens(r, l)
        cdf = np.concatenate([[0.0], np.cumsum(0.5 * (pdf[1:] + pdf[:-1]) * np.diff(r))])
        total = {"Gaussian": np.pi / (2 * l * l), "Exponential": 1 / (l * l), "Matern": 1.3 / (2 * 1.3 * l * l)}[model]
        for u in (0.01, 0.3, 0.5, 0.9, 0.97):
            rq = float(R.radial_ppf(model, u, l, nu))
            assert abs(np.interp(rq, r, cdf) / total - u) < 2e-4, (model, u)


def test_summation_matches_a_literal_double_loop_and_scales():
    g = np.random.default_rng(0)
    n, ny, nx, res = 37, 6, 9, 500.0
    kx, ky = g.normal(size=n) * 1e-4, g.normal(size=n) * 1e-4
    z1, z2, zn = g.normal(size=n), g.normal(size=n), g.normal(size=(ny, nx))
    got = R.randmeth_field(kx, ky, z1, z2, (ny, nx), res, 40.0, 0.25, zn)
    want = np.zeros((ny, nx))
    for y in range(ny):
        for x in range(nx):
            ph = kx * (x * res) + ky * (y * res)
            want[y, x] = 40.0 * (np.sqrt(1.0 / n) * np.sum(z1 * np.cos(ph) + z2 * np.sin(ph)) + 0.5 * zn[y, x])
    assert np.allclose(got, want, rtol=1e-12, atol=1e-12)


def test_emulated_device_modes_have_unit_variance_fields():
    """Device-RNG emulation end to end: pointwise variance of the unit-scale field over many keys is 1 (var = 1)."""
    vals = []
    for key in range(300):
        kx, ky, z1, z2 = randmeth_modes(key * 7919 + 1, 3, 200, "Matern", 20e3, 20e3, 0.0, 0.9)
        vals.append(R.randmeth_field(kx, ky, z1, z2, (2, 2), 500.0, 1.0)[1, 1])
    assert abs(np.var(vals) - 1.0) < 0.25 and abs(np.mean(vals)) < 0.2
