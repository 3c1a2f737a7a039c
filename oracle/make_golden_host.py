"""Generate tests/golden/host_logic_table.npz by running the REFERENCE: the host-side bookkeeping of GaussianProcess
(flags set by set_data, nuggets, hyper-parameter index maps, scaled data, noise vectors, vector <-> dataclass maps, initial
hyper-parameters, LHS start points and box bounds from the history) for every conditioning mode x noise model x
use_grad.  TEST INFRASTRUCTURE ONLY; runs in the build container (needs /root/reference).

    python oracle/make_golden_host.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from oracle import gegp_oracle as O  # noqa: E402

MODES = ("precon", "base", "rescale_origin", "rescale_eta_vary", "dflt_vmin", "dflt_vmax")
NOISES = ("zero", "known", "unknown")
ATTRS = ("wellcond_mtd", "b_use_cond_cstr", "b_use_data_scl", "b_has_noisy_data", "b_optz_var_fval", "b_optz_var_fgrad",
         "b_fval_zero", "b_fgrad_zero", "n_eval", "n_grad", "n_data", "_etaK", "_eta_Kbase", "_eta_Kgrad",
         "cond_eta_is_const", "known_eps_fval", "known_eps_fgrad")
HPINFO = ("n_hp", "has_theta", "has_kernel", "has_varK", "has_var_fval", "has_var_fgrad", "idx_theta", "idx_varK",
          "idx_var_fval", "idx_var_fgrad", "bvec_log_optz")


def collect(GPclass, x, f, g):
    """Same walk for the reference (here) and for the mirror class (tests/test_host_logic.py)."""
    out = {}

    def put(key, val):
        if val is None:
            out[key] = np.array("None")
        elif isinstance(val, str):
            out[key] = np.array(val)
        else:
            out[key] = np.asarray(val, dtype=float)

    n, d = x.shape
    for mode in MODES:
        for noise in NOISES:
            for use_grad in (True, False):
                tag = f"{mode}|{noise}|{int(use_grad)}|"
                G = GPclass(d, use_grad, "SqExp", mode)
                sf, sg = {"zero": (np.zeros(n), np.zeros((n, d))), "known": (0.01 * np.ones(n), 0.05 * np.ones((n, d))),
                          "unknown": (None, None)}[noise]
                G.init_optz_surr(4)
                if use_grad:
                    G.set_data(x, f, sf, g, sg)
                else:
                    G.set_data(x, f, sf)
                for a in ATTRS:
                    put(tag + a, getattr(G, a, None))
                hi = G.hp_info_optz_lkd
                for a in HPINFO:
                    put(tag + "hpinfo." + a, getattr(hi, a, None))
                put(tag + "x_scl", G.get_scl_x_w_dist()[0])
                for i, nm in enumerate(("f_scl", "stdf_scl", "g_scl", "stdg_scl")):
                    put(tag + nm, G.get_scl_eval_data()[i])
                vf = 0.3 if noise == "unknown" else None
                vg = 0.4 if (noise == "unknown" and use_grad) else None
                hp = G.make_hp_class(theta=np.array([0.1, 0.2, 0.3]), varK=2.0, var_fval=vf, var_fgrad=vg)
                if not (noise == "unknown" and not use_grad):      # the reference crashes there (kernel/Kernel.py:336)
                    put(tag + "noise_vec", G.calc_noise_vec(hp))
                v = np.linspace(-1.5, 0.5, hi.n_hp)
                dc = G.hp_vec2dataclass(hi, v.copy())
                for nm in ("theta", "varK", "var_fval", "var_fgrad"):
                    put(tag + "vec2dc." + nm, getattr(dc, nm))
                ih = G.get_init_hp_vals()
                for nm in ("theta", "varK", "var_fval", "var_fgrad", "beta"):
                    put(tag + "init_hp." + nm, getattr(ih, nm))
                G.hp_theta_all[0] = [1e-2, 2e-2, 5e-3]
                G.hp_varK_all[0] = 3.0
                G.hp_var_fval_all[0] = 1e-3
                G.hp_var_fgrad_all[0] = 2e-3
                x0, b = G.get_hp_x0_lhs_median(1, hi, 5)
                put(tag + "lhs_x0", x0)
                put(tag + "box_lb", b.lb)
                put(tag + "box_ub", b.ub)
    return out


if __name__ == "__main__":
    from gpgradpy.src.GaussianProcess import GaussianProcess  # the reference
    x, f, g = O.synthetic_problem(14, 3, 0)
    table = collect(GaussianProcess, x, f, g)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "host_logic_table.npz"), **table)
    print("entries", len(table))
