"""BASELINE configs[0] through the mirror API: the computation of gpgradpy/plt/plt_cond.py (2-D Rosenbrock, n = 20 points
in [0.9, 1.1]^2, log10 condition number and log-marginal-likelihood over a grid of length scales, one nugget for every
conditioning mode) without the plotting.  A user of the reference only changes the import line.

    python examples/config1_cond_lkd_grid.py [n_gamma]
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gpgradpy_b200.gp import GaussianProcess        # reference: from gpgradpy.src import GaussianProcess  # noqa: E402

n_eval, dim, varK, cond_max = 20, 2, 0.1, 1e10
n_gamma = int(sys.argv[1]) if len(sys.argv) > 1 else 12
range_gamma = np.array([1e-4, 1e8])


def calc_obj(xy, a=10):
    return np.sum(a * (xy[:, 1:] - xy[:, :-1] ** 2) ** 2 + (1 - xy[:, :-1]) ** 2, axis=1)


def calc_grad(xy, a=10):
    g = np.zeros_like(xy)
    g[:, :-1] += -2 * (1 - xy[:, :-1]) - 4 * a * xy[:, :-1] * (xy[:, 1:] - xy[:, :-1] ** 2)
    g[:, 1:] += 2 * a * (xy[:, 1:] - xy[:, :-1] ** 2)
    return g


x_eval = np.random.default_rng(2).uniform(0.9, 1.1, (n_eval, dim))
obj_eval, grad_eval = calc_obj(x_eval), calc_grad(x_eval)
nugget = GaussianProcess(dim, use_grad=True, kernel_type="SqExp", wellcond_mtd="precon").calc_nugget(n_eval)[1]
theta_vec = np.logspace(*np.log10(GaussianProcess.gamma2theta(range_gamma)), n_gamma)

for wellcond_mtd in ("base", "rescale_origin", "precon"):
    GP = GaussianProcess(dim, True, "SqExp", wellcond_mtd)
    GP.cond_eta_is_const = True
    GP.set_data(x_eval, obj_eval, np.zeros(n_eval), grad_eval, np.zeros(grad_eval.shape))
    GP._etaK = GP._eta_Kgrad = nugget
    GP.cond_max = cond_max
    lkd = np.full((n_gamma, n_gamma), np.nan)
    cond = np.full((n_gamma, n_gamma), np.nan)
    for i in range(n_gamma):
        for j in range(n_gamma):
            hp = GP.make_hp_class(varK=varK, theta=np.array([theta_vec[i], theta_vec[j]]), kernel=GP.hp_kernel_default)
            info = GP.calc_lkd_all(hp, calc_lkd=True, calc_cond=True, calc_grad=False)[0]
            lkd[i, j] = np.nan if info.ln_lkd is None else info.ln_lkd
            cond[i, j] = info.cond
    ok = np.isfinite(lkd)
    i, j = np.unravel_index(np.nanargmax(lkd), lkd.shape)
    print(f"{wellcond_mtd:15s}: Cholesky ok at {ok.sum()}/{ok.size} grid points; log10 cond in "
          f"[{np.log10(np.nanmin(cond)):.1f}, {np.log10(np.nanmax(cond)):.1f}]; fraction with cond <= 1e10: "
          f"{np.mean(cond <= cond_max):.2f}; max LML {np.nanmax(lkd):.3f} at gamma = "
          f"({np.sqrt(2 * theta_vec[i]):.2e}, {np.sqrt(2 * theta_vec[j]):.2e})")
