"""Mirror of the part of the reference's gstatsMCMC/gstatsim_custom package that feeds the hot path: the whole-grid
Sequential Gaussian Simulation that produces the initial bed of every large-scale chain."""
from . import interpolate  # noqa: F401
