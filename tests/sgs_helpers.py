"""Builders shared by the SGS oracle/GPU tests."""
from __future__ import annotations

import numpy as np

from cases import build_sgs_inputs
from oracle import sgs_oracle as S


def sgs_vario_dict(v):
    """The dict chain_sgs.run hands to sgs() (MCMC.py:1682-1702), from set_variogram's arguments (MCMC.py:1515-1535)."""
    if v["isotropic"]:
        az, major, minor = 0, v["range"], v["range"]
    else:
        az, major, minor = v["azimuth"], v["range"][0], v["range"][1]
    d = dict(azimuth=az, nugget=v["nugget"], major_range=major, minor_range=minor, sill=v["sill"], vtype=v["vtype"])
    if v["vtype"] == "Matern":
        d["s"] = v["smoothness"]
    return d


def oracle_sgs_setup(case):
    g = build_sgs_inputs(case)
    nst = S.NormalScore(g["quantiles"], g["references"]) if case["transform"] else None
    su = S.SgsSetup(xx=g["xx"], yy=g["yy"], surf=g["surf"], velx=g["velx"], vely=g["vely"], dhdt=g["dhdt"], smb=g["smb"],
                    cond_bed=g["cond_bed"], grounded_ice_mask=g["grounded_ice_mask"], region_mask=g["highvel_mask"],
                    mc_region_mask=g["highvel_mask"], resolution=g["resolution"], sigma_mc=case["sigma_mc"], trend=g["trend"],
                    nst=nst, vario=sgs_vario_dict(case["vario"]), num_points=case["neighbors"], radius=case["radius"],
                    block=tuple(case["blocks"]))
    return g, su


def product_sgs_chain(case, g=None):
    """mcmc_gpu_b200's chain_sgs configured like the reference tutorial (T4_SmallScaleChain.ipynb cells 20-38)."""
    from gpu_helpers import quiet
    from mcmc_gpu_b200 import MCMC
    g = build_sgs_inputs(case) if g is None else g
    ch = quiet(MCMC.chain_sgs, g["xx"], g["yy"], g["bed_init"], g["surf"], g["velx"], g["vely"], g["dhdt"], g["smb"], g["cond_bed"],
               g["data_mask"], g["grounded_ice_mask"], g["resolution"])
    quiet(ch.set_update_region, True, g["highvel_mask"])
    ch.set_loss_type(sigma_mc=case["sigma_mc"], massConvInRegion=True)
    ch.set_block_sizes(*case["blocks"])
    ch.set_normal_transformation(g.get("nst"), do_transform=case["transform"])
    ch.set_trend(g["trend"], detrend_map=case["detrend"])
    v = case["vario"]
    quiet(ch.set_variogram, v["vtype"], v["range"], v["sill"], v["nugget"], isotropic=v["isotropic"],
          vario_smoothness=v["smoothness"], vario_azimuth=v["azimuth"])
    quiet(ch.set_sgs_param, case["neighbors"], case["radius"])
    ch.set_random_generator(case["seed"])
    return ch, g
