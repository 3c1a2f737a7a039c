"""Hyper-parameter containers and nugget / rescaling scalar formulas (host side, O(n d) work).

Field-for-field mirrors of the reference dataclasses so that user scripts keep working:
HparaOptzVal (base/GpHpara.py:12-19), LkdInfo (optz/CalcLkd.py:14-26), HparaOptzInfo (optz/GpHparaOptz.py:18-31).
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass
class HparaOptzVal:
    beta: np.ndarray = None       # coefficients of the mean function
    theta: np.ndarray = None      # kernel length-scale hyper-parameters
    kernel: float = None          # extra kernel hyper-parameter (none for the Gaussian kernel)
    varK: float = None            # process variance
    var_fval: float = None        # noise variance on the function values
    var_fgrad: float = None       # noise variance on the gradients


@dataclass
class LkdInfo:
    hp_beta: np.ndarray = None
    hp_beta_grad: np.ndarray = None
    hp_varK: float = None
    hp_varK_grad: np.ndarray = None
    ln_det_Kmat: float = None
    ln_det_Kmat_grad: np.ndarray = None
    ln_lkd: float = None
    ln_lkd_grad: np.ndarray = None
    cond: float = None
    cond_grad: np.ndarray = None
    data_vec: np.ndarray = None


@dataclass(frozen=True)
class HparaOptzInfo:
    n_hp: int = None
    has_theta: bool = False
    idx_theta: np.ndarray = None
    has_kernel: bool = False
    idx_kernel: np.ndarray = None
    has_varK: bool = False
    idx_varK: np.ndarray = None
    has_var_fval: bool = False
    idx_var_fval: np.ndarray = None
    has_var_fgrad: bool = False
    idx_var_fgrad: np.ndarray = None
    bvec_log_optz: np.ndarray = None


def vreq_rescale_origin(n_eval: int, dim: int) -> float:
    """Required minimum pairwise distance of the rescaling method (base/GpWellCond.py:26-40)."""
    if n_eval == 1:
        return 1.0
    dist_star = 2.0 * np.sqrt(dim)
    root = np.sqrt(4.0 + 2.0 * np.exp(2.0) * np.log((n_eval - 1) * (1.0 + dist_star) / 2.0))
    return float(min((2.0 + root) / np.exp(1.0), dist_star))


def nugget_Kbase(n_eval: int, cond_max: float) -> float:
    """eta = n / (cond_max - 1)  (base/GpWellCond.py:109-114)."""
    return n_eval / (cond_max - 1.0)


def nugget_Kfull_vreq(n_eval: int, dim: int, cond_max: float, vmin: float | None = None) -> float:
    """Nugget of the rescaling method (base/GpWellCond.py:78-99)."""
    if vmin is None:
        vmin = vreq_rescale_origin(n_eval, dim)
    if n_eval == 1:
        return n_eval / (cond_max - 1.0)
    assert vmin >= np.sqrt(2) - 1e-12, f"this method requires vmin = {vmin} >= sqrt(2)"
    v_frac = 2.0 * np.sqrt(dim) / vmin
    return (1.0 + (n_eval - 1) * v_frac * np.exp(1.0 / v_frac - 1.0)) / (cond_max - 1.0)


def nugget_precon_sqexp(n_eval: int, dim: int, cond_max: float) -> float:
    """Gershgorin-bound nugget of the preconditioned Gaussian-kernel matrix (base/GpWellCond.py:126-138)."""
    root = np.sqrt(1.0 + 4.0 * dim)
    ub = 0.5 * (n_eval - 1) * (1.0 + root) * np.exp(-(1.0 + 2.0 * dim - root) / (4.0 * dim))
    return (1.0 + ub) / (cond_max - 1.0)


def nugget_precon_matern52(n_eval: int, dim: int, cond_max: float) -> float:
    """Gershgorin-bound nugget of the preconditioned Matern-5/2 matrix (base/GpWellCond.py:130-134)."""
    r3 = np.sqrt(3.0 * dim)
    al = (r3 - 1.0 + np.sqrt(15.0 * dim + 2.0 * r3 + 1.0)) / (2.0 * (3.0 * dim + r3))
    ub = (n_eval - 1) * (1.0 + (dim + r3) * al + dim * (1.0 + r3) * al ** 2) * np.exp(-r3 * al)
    return (1.0 + ub) / (cond_max - 1.0)
