"""tests/golden/direct_*.npz: the reference's likelihood in its DIRECT (non-adjoint) form, lkd_use_adj_mtd = False
(optz/CalcLkd.py:64-85, 135-147, 238-241), which additionally returns hp_beta_grad, hp_varK_grad and ln_det_Kmat_grad.
TEST INFRASTRUCTURE ONLY; build container only (needs /root/reference).

    python oracle/make_golden_direct.py
"""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(HERE, "ref_shim"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
from gpgradpy.src.GaussianProcess import GaussianProcess  # noqa: E402  (the reference)
from oracle import gegp_oracle as O  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def case(name, kernel, hp_kernel, n, d, mode, std=None, varK=None, seed=0):
    x, f, g = O.synthetic_problem(n, d, seed)
    th = O.bench_theta(d) * 3.0
    GP = GaussianProcess(d, True, kernel, mode)
    sf, sg = (0.0, 0.0) if std is None else std
    GP.set_data(x, f, sf * np.ones(n), g, sg * np.ones(g.shape))
    hp = GP.make_hp_class(theta=th, kernel=hp_kernel, varK=varK)
    adj, ok1 = GP.calc_lkd_all(hp, calc_grad=True, lkd_use_adj_mtd=True)
    dr, ok2 = GP.calc_lkd_all(hp, calc_grad=True, lkd_use_adj_mtd=False)
    assert ok1 and ok2
    out = dict(kernel=kernel, hp_kernel=np.nan if hp_kernel is None else float(hp_kernel), x=x, fval=f, grad=g, theta=th,
               mode=mode, eta=GP._etaK, std_f=sf, std_g=sg, varK=np.nan if varK is None else varK,
               ln_lkd=dr.ln_lkd, ln_lkd_grad=dr.ln_lkd_grad, ln_lkd_grad_adjoint=adj.ln_lkd_grad,
               hp_beta_grad=dr.hp_beta_grad, ln_det_Kmat_grad=dr.ln_det_Kmat_grad,
               hp_varK_grad=np.zeros(0) if dr.hp_varK_grad is None else dr.hp_varK_grad)
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **out)
    kern = (kernel, hp_kernel)
    if varK is None:
        o = O.lkd_direct(*GP.get_scl_x_w_dist()[:1], *[a for a in (GP.get_scl_eval_data()[0], GP.get_scl_eval_data()[2])], th,
                         "precon" if mode == "precon" else "base", GP._etaK, kernel=kern)
    else:
        o = O.lkd_direct(x, f, g, th, mode, GP._etaK, kernel=kern, varK=varK, noise_vec=GP.calc_noise_vec(hp))
    rel = lambda a, b: float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / np.max(np.abs(b)))  # noqa: E731
    print(name, "oracle vs reference: lkd_grad", rel(o[0], dr.ln_lkd_grad), "beta_grad", rel(o[1], dr.hp_beta_grad),
          "logdet_grad", rel(o[3], dr.ln_det_Kmat_grad), "| direct vs adjoint", rel(dr.ln_lkd_grad, adj.ln_lkd_grad), flush=True)


if __name__ == "__main__":
    case("direct_sqexp_d3_n20_precon", "SqExp", None, 20, 3, "precon")
    case("direct_sqexp_d3_n20_base", "SqExp", None, 20, 3, "base")
    case("direct_sqexp_d2_n16_rescale_origin", "SqExp", None, 16, 2, "rescale_origin")
    case("direct_ratqu_d3_n20_precon", "RatQu", 1.7, 20, 3, "precon")
    case("direct_ma5f2_d3_n20_precon", "Ma5f2", None, 20, 3, "precon")
    case("direct_sqexp_d3_n20_noisy_precon", "SqExp", None, 20, 3, "precon", std=(1e-2, 5e-2), varK=350.0)
    case("direct_ratqu_d3_n20_noisy_base", "RatQu", 1.7, 20, 3, "base", std=(1e-2, 5e-2), varK=350.0)
