"""CPU restatement of the reference's *other* proposal generator — TEST INFRASTRUCTURE ONLY.

Reference call site: `RandField.get_random_field` (/root/reference/gstatsMCMC/MCMC.py:625-687), selected by
`set_generation_method(False)` (MCMC.py:514-522) and consumed by `get_rfblock` (MCMC.py:764-768):

    model = gstools.Gaussian|Exponential|Matern(dim=2, var=1, len_scale=[range1, range2]/(sqrt(3)|3|2),
                                                angles=angle*pi/180, nugget=nug[, nu=smoothness])
    field = gstools.SRF(model).structured([X, Y]).T * scale

PARITY UNPINNED.  The arithmetic lives in gstools 1.7.0 / gstools-cython 1.1.0 (gstatsMCMC.yml:14-15), which is neither
vendored under /root/reference nor installed here, and the reference passes no seed to SRF, so even the reference is
not reproducible run to run.  What follows restates the published algorithm (randomization method, Hesse et al. 2014,
as implemented by gstools.field.generator.RandMeth) and the published model definitions:

  RandMeth.__call__ :  field(p) = sqrt(var / N) * sum_m [ z1_m cos(k_m . p') + z2_m sin(k_m . p') ]  (+ sqrt(nugget) n(p))
                       N = mode_no = 1000 (SRF default), z1, z2 ~ N(0,1), p' = isometrized position
  RandMeth.reset_seed: k_m = r_m * (cos phi_m, sin phi_m), phi uniform on the circle (sample_sphere, dim 2), r_m drawn
                       from the model's radial spectral density (inversion of the cdf where gstools has a ppf,
                       otherwise an MCMC sampler of the same density)
  CovModel.isometrize: p' = diag(1, 1/anis) R(angle)^T p,   anis = len_scale[1] / len_scale[0], main length len_scale[0]
  correlation, h = rescale * d / len:   Gaussian exp(-h^2), rescale sqrt(pi)/2;   Exponential exp(-h), rescale 1;
                       Matern 2^(1-nu)/Gamma(nu) (sqrt(nu) h)^nu K_nu(sqrt(nu) h), rescale 1

In two dimensions the radial cdf of all three spectra inverts in closed form (`radial_ppf`), which is what the device
uses for every model; tests/test_oracle_randmeth.py checks the inversion against the correlation functions above
(E[cos(k . d)] over sampled wave vectors must equal rho(d)), i.e. against the published definitions, not against gstools.
Only tests/, __graft_entry__.smoke() and bench.py's cpu legs may import this module.
"""
from __future__ import annotations

import numpy as np
from scipy import special

MODELS = ("Gaussian", "Exponential", "Matern")


def model_lengths(model, range_x, range_y):
    """len_scale = [range1, range2] / sqrt(3) | 3 | 2                                        MCMC.py:657-676"""
    dv = {"Gaussian": np.sqrt(3.0), "Exponential": 3.0, "Matern": 2.0}[model]
    return range_x / dv, range_y / dv


def correlation(model, d, length, nu=1.0):
    """rho(d) of the isotropic model with main length `length` (gstools definitions, rescale factors included)."""
    d = np.asarray(d, dtype=np.float64)
    if model == "Gaussian":
        return np.exp(-(np.pi / 4.0) * (d / length) ** 2)
    if model == "Exponential":
        return np.exp(-d / length)
    h = np.sqrt(nu) * d / length
    with np.errstate(invalid="ignore", divide="ignore"):
        r = 2.0 ** (1.0 - nu) / special.gamma(nu) * h ** nu * special.kv(nu, h)
    return np.where(h == 0.0, 1.0, r)


def radial_ppf(model, u, length, nu=1.0):
    """Inverse of the radial spectral cdf in two dimensions (u in (0,1) -> wave number r).

    The 2-D spectral densities are  Gaussian ~ exp(-(r L/2)^2) with L = 2 length/sqrt(pi);  Exponential
    ~ (1 + (r l)^2)^(-3/2);  Matern ~ (1 + (r l)^2/nu)^(-(nu+1)).  Their radial cdfs int_0^r 2 pi k S(k) dk are
    1 - exp(-(r L/2)^2),  1 - (1 + (r l)^2)^(-1/2),  1 - (1 + (r l)^2/nu)^(-nu)."""
    u = np.asarray(u, dtype=np.float64)
    if model == "Gaussian":
        return np.sqrt(-np.log1p(-u)) * 1.7724538509055159 / length
    if model == "Exponential":
        return np.sqrt(u * (2.0 - u)) / (1.0 - u) / length
    return np.sqrt(nu * np.expm1(-np.log1p(-u) / nu)) / length


def grid_wave_vectors(model, u_rad, u_ang, range_x, range_y, angle_deg, nu=1.0):
    """Wave vectors in the grid frame: k = R(theta) diag(1, l1/l2) k', so that k . p = k' . isometrize(p)."""
    l1, l2 = model_lengths(model, range_x, range_y)
    r = radial_ppf(model, u_rad, l1, nu)
    c, s = np.cos(2.0 * np.pi * np.asarray(u_ang)), np.sin(2.0 * np.pi * np.asarray(u_ang))
    k0, k1 = r * c, r * s * (l1 / l2)
    th = angle_deg * np.pi / 180.0                                                         # MCMC.py:660
    return np.cos(th) * k0 - np.sin(th) * k1, np.sin(th) * k0 + np.cos(th) * k1


def randmeth_field(kx, ky, z1, z2, shape, res, scale, nug=0.0, z_nug=None):
    """The summation of RandMeth.__call__ on the block grid X = arange(nx) res, Y = arange(ny) res, times `scale`
    (MCMC.py:681: `srf.structured([X, Y]).T * scale`, nugget noise included in the scaled field) -> [ny, nx]."""
    ny, nx = shape
    X = np.arange(nx, dtype=np.float64) * res
    Y = np.arange(ny, dtype=np.float64) * res
    kx, ky, z1, z2 = (np.asarray(a, dtype=np.float64) for a in (kx, ky, z1, z2))
    out = np.zeros((ny, nx))
    step = 50                                          # modes per slab (bounds the temporary to step*ny*nx doubles)
    for m0 in range(0, len(kx), step):
        sl = slice(m0, m0 + step)
        phase = ky[sl, None, None] * Y[None, :, None] + kx[sl, None, None] * X[None, None, :]
        out += (z1[sl, None, None] * np.cos(phase) + z2[sl, None, None] * np.sin(phase)).sum(0)
    out *= np.sqrt(1.0 / len(kx))
    if nug > 0.0:
        out = out + np.sqrt(nug) * np.asarray(z_nug, dtype=np.float64).reshape(ny, nx)
    return out * scale
