"""GPU: 2-norm condition number, its hyper-parameter gradient and the constrained fits of the base / rescale modes
(optz/GpHparaCon.py:161-235, optz/OptzLkd.py:116-333) against the reference outputs in tests/golden/cond_*.npz and
fit_*.npz (oracle/make_golden_cond.py) and against the CPU oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _gp(g, mode):
    from gpgradpy_b200.gp import GaussianProcess
    x = g["x"]
    n, d = x.shape
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, g["fval"], float(g.get("std_f", 0.0)) * np.ones(n), g["grad"],
                float(g.get("std_g", 0.0)) * np.ones(g["grad"].shape))
    return GP


@pytest.mark.parametrize("N", [5, 64, 300, 1100])
def test_extreme_eig_vs_eigh(N):
    """Restarted Lanczos (gegp_symv + gegp_lanczos_step + gegp_lincomb) against numpy.linalg.eigh."""
    import torch
    from gpgradpy_b200 import backend as bk
    rng = np.random.default_rng(N)
    Q = np.linalg.qr(rng.standard_normal((N, N)))[0]
    w = np.sort(10.0 ** rng.uniform(-4, 2, N))
    M = (Q * w) @ Q.T
    M = 0.5 * (M + M.T)
    ld = bk.ld_of(N)
    Md = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    Md[:, :N] = bk.to_dev(M)
    lam, v, resid, cycles = bk.extreme_eig(Md, N)
    wr, Vr = np.linalg.eigh(M)
    assert abs(lam - wr[-1]) < 1e-11 * wr[-1]
    vh = v.cpu().numpy()
    assert abs(np.linalg.norm(vh) - 1.0) < 1e-12
    assert 1.0 - abs(vh @ Vr[:, -1]) < 1e-9
    y = torch.empty(ld, dtype=torch.float64, device="cuda")
    x = bk.to_dev(np.pad(rng.standard_normal(N), (0, ld - N)))
    bk.symv(Md, x, y, N)
    ref = M @ x.cpu().numpy()[:N]
    assert np.max(np.abs(y.cpu().numpy()[:N] - ref)) < 1e-12 * np.max(np.abs(ref))


@pytest.mark.parametrize("n,d,noisy", [(14, 2, False), (40, 5, False), (150, 3, False), (16, 3, True)])
def test_quad_grad_vs_oracle(n, d, noisy):
    """v^T (dKcov/dhp) v with the derivative matrices generated on the fly against the materialised stack."""
    from gpgradpy_b200 import backend as bk, _lib as L
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(n, d, 3)
    th = O.bench_theta(d) * 3.0
    N = n * (d + 1)
    v = np.random.default_rng(5).standard_normal(N)
    v /= np.linalg.norm(v)
    eta, varK = 1e-6, 7.5
    if noisy:
        K = O.kern_grad(x, x, th)
        D = O.kcov_grad_hp_noisy(x, th, K, "base", eta, varK, True, True)
    else:
        D = O.kerngrad_hp(x, th, "base", eta)
    ref = np.array([v @ D[i] @ v for i in range(D.shape[0])])
    out = bk.quad_grad(x, th, v, eta=eta, noisy=noisy, varK=varK).cpu().numpy()
    got = out[L.OUT_GRAD:L.OUT_GRAD + d]
    assert np.max(np.abs(got - ref[:d])) < 1e-11 * np.max(np.abs(ref[:d]))
    if noisy:
        extra = np.array([out[L.OUT_DVARK], out[L.OUT_DVARF], out[L.OUT_DVARG]])
        assert np.max(np.abs(extra - ref[d:])) < 1e-11 * np.max(np.abs(ref[d:]))


COND = [("cond_d2_n12_base", 1e-8), ("cond_d3_n20_base", 1e-9), ("cond_d2_n16_rescale_origin", 1e-9),
        ("cond_d3_n14_precon", 1e-7), ("cond_d2_n14_noisy_base", 1e-7), ("cond_d4_n40_base_illcond", 1e-4)]


@pytest.mark.parametrize("name,tol", COND)
def test_cond_and_grad_through_the_api(golden_dir, name, tol):
    """GP.calc_lkd_all(calc_cond=True, calc_grad=True/False) -> LkdInfo.cond / cond_grad as the reference returns them
    (tolerance ~ eps * kappa: lambda_min is only defined to an absolute eps * lambda_max in either implementation)."""
    g = _load(golden_dir, name)
    mode = str(g["mode"])
    GP = _gp(g, mode)
    varK = None if np.isnan(g["varK"]) else float(g["varK"])
    hp = GP.make_hp_class(theta=g["theta"], varK=varK)
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
    assert ok
    assert abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert abs(info.cond - g["cond"]) < tol * g["cond"]
    if "cond_grad" in g:
        assert np.max(np.abs(info.cond_grad - g["cond_grad"])) < 20 * tol * np.max(np.abs(g["cond_grad"]))
    else:
        assert info.cond_grad is None
    info0, ok0 = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=False)
    assert ok0 and abs(info0.cond - g["cond_nograd"]) < tol * g["cond_nograd"]
    # calc_all_K_w_chofac(calc_cond=True): element [4] of the 7-tuple (kernel/Kernel.py:240,280)
    c = GP.calc_all_K_w_chofac(None, hp, calc_chofac=False, calc_cond=True, varK=varK if varK else 1)[4]
    assert abs(c - g["cond_nograd"]) < tol * g["cond_nograd"]


def test_cond_medium_vs_oracle():
    """N = 1100 (several leaves, restarts): kappa and gradient against the NumPy oracle."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    n, d = 100, 10
    x, f, g = O.synthetic_problem(n, d, 2)
    GP = GaussianProcess(d, True, "SqExp", "base")
    GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
    th = O.bench_theta(d) * 4.0
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=th), calc_cond=True, calc_grad=True)
    c, cg = O.cond_wo_noise(x, th, "base", GP._etaK)
    assert ok
    tol = max(1e-9, 1e-15 * c)
    assert abs(info.cond - c) < tol * c
    assert np.max(np.abs(info.cond_grad - cg)) < 100 * tol * np.max(np.abs(cg))


def test_failed_cholesky_reports_cond():
    """Numerically singular matrix (tiny theta, base mode, no usable nugget): b_chofac_good False and a huge but finite
    condition number as the substitute objective (optz/OptzLkd.py:75-77)."""
    from gpgradpy_b200.gp import GaussianProcess
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(30, 3, 0)
    GP = GaussianProcess(3, True, "SqExp", "base")
    GP.cond_eta_set_mtd = "dflt_eta"
    GP.cond_eta_dflt = 0.0
    GP.set_data(x, f, np.zeros(30), g, np.zeros((30, 3)))
    info, ok = GP.calc_lkd_all(GP.make_hp_class(theta=1e-7 * np.ones(3)), calc_cond=True, calc_grad=True)
    if not ok:
        assert np.isfinite(info.cond) and info.cond > 1e12
        assert info.cond_grad is not None and info.cond_grad.shape == (3,)
        GP._last_hp_vec = None
        assert GP.return_optz_val(np.log10(1e-7 * np.ones(3))) == info.cond   # minimised objective -LML := +cond


@pytest.mark.parametrize("name", ["fit_d2_n20_base", "fit_d2_n20_rescale_origin"])
def test_constrained_fit_vs_reference(golden_dir, name):
    """set_hpara('optz') in the modes that carry the kappa <= cond_max constraint (SLSQP NonlinearConstraint with the
    device condition number and gradient; rescale modes add the outer re-scaling loop).  Started from the start point the
    reference's own candidate scan picked; the optimum must be feasible and as good as the reference's."""
    g = _load(golden_dir, name)
    mode = str(g["mode"])
    from gpgradpy_b200.gp import GaussianProcess
    from scipy.optimize import Bounds
    x, f, gr = g["x"], g["fval"], g["grad"]
    n, d = x.shape
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.init_optz_surr(3)
    GP.set_data(x[:1], f[:1], np.zeros(1), gr[:1], np.zeros((1, d)))
    GP.set_hpara("optz", 0)
    GP.set_data(x, f, np.zeros(n), gr, np.zeros((n, d)))
    bound = Bounds(g["lb"], g["ub"], keep_feasible=True)
    if "rescale" in mode:
        best, cond_val, info = GP.optz_hp_max_lkd_mtd_rescale(1, g["hp_x0"], bound)
    else:
        best, cond_val, info = GP.optz_hp_max_lkd(g["hp_x0"], bound)
    hp = GP.optz_closed_form_hp(GP.hp_vec2dataclass(GP.hp_info_optz_lkd, best))
    res, ok = GP.calc_lkd_all(hp, calc_cond=True)
    assert ok
    assert res.cond < 1.01 * GP.cond_max
    assert res.ln_lkd > float(g["ln_lkd"]) - 1e-3 * max(1.0, abs(float(g["ln_lkd"])))
    assert info["hp_optz_con_good"] == 1.0
    if "rescale" in mode:
        assert np.allclose(GP.DataScl.xvec_scale, g["xvec_scale"], rtol=5e-2)
    # the whole public path as well
    GP.set_hpara("optz", 1)
    assert GP.hp_vals.varK > 0 and np.isfinite(GP.Kcov_cond_all[1])


def test_variable_nugget_mode(golden_dir):
    """wellcond_mtd='rescale_eta_vary': Gershgorin nugget from the device row sums (gegp_row_abs_sum), then LML,
    gradient, condition number and posterior against the reference."""
    from gpgradpy_b200.gp import GaussianProcess
    g = _load(golden_dir, "etavary_d3_n18")
    n, d = g["x"].shape
    GP = GaussianProcess(d, True, "SqExp", "rescale_eta_vary")
    GP.set_data(g["x"], g["fval"], np.zeros(n), g["grad"], np.zeros((n, d)))
    hp = GP.make_hp_class(theta=g["theta"])
    tup = GP.calc_all_K_w_chofac(None, hp, varK=1)
    assert abs(tup[5] - g["eta"]) < 1e-12 * g["eta"] and tup[6] == int(g["idx"])
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
    assert ok and abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert np.max(np.abs(info.ln_lkd_grad - g["ln_lkd_grad"])) < 1e-8 * np.max(np.abs(g["ln_lkd_grad"]))
    assert abs(info.cond - g["cond"]) < 1e-8 * g["cond"]
    assert np.max(np.abs(info.cond_grad - g["cond_grad"])) < 1e-7 * np.max(np.abs(g["cond_grad"]))
    GP.set_hpara("set", 1, GP.make_hp_class(theta=g["theta"], varK=float(g["hp_varK"]), beta=g["hp_beta"]))
    mu, sig = GP.eval_model(g["x_test"])[:2]
    assert np.max(np.abs(mu - g["mu"])) < 1e-8 * np.max(np.abs(g["mu"]))
    assert np.max(np.abs(sig - g["sig"])) < 1e-6 * np.max(np.abs(g["sig"]))


@pytest.mark.parametrize("name", ["condfro_d2_n12_base", "condfro_d2_n14_noisy_base"])
def test_frobenius_cond_through_the_api(golden_dir, name):
    """GP.cond_norm = 'fro' (optz/GpHparaCon.py:237-261): |K|_F |K^-1|_F from two device reductions and its gradient from
    sum((frac K - K^-3 / frac) .* dKcov/dhp) with dKcov/dhp generated on the fly (gegp_weighted_grad)."""
    g = _load(golden_dir, name)
    from gpgradpy_b200.gp import GaussianProcess
    x = g["x"]
    n, d = x.shape
    GP = GaussianProcess(d, True, "SqExp", "base")
    GP.cond_norm = "fro"
    GP.set_data(x, g["fval"], float(g["std_f"]) * np.ones(n), g["grad"], float(g["std_g"]) * np.ones((n, d)))
    varK = None if np.isnan(g["varK"]) else float(g["varK"])
    hp = GP.make_hp_class(theta=g["theta"], varK=varK)
    info, ok = GP.calc_lkd_all(hp, calc_cond=True, calc_grad=True)
    assert ok and abs(info.ln_lkd - g["ln_lkd"]) < 1e-8 * abs(g["ln_lkd"])
    assert abs(info.cond - g["cond"]) < 1e-7 * g["cond"]
    assert np.max(np.abs(info.cond_grad - g["cond_grad"])) < 1e-6 * np.max(np.abs(g["cond_grad"]))
    c = GP.calc_all_K_w_chofac(None, hp, calc_chofac=False, calc_cond=True, varK=varK if varK else 1)[4]
    assert abs(c - g["cond_nograd"]) < 1e-7 * g["cond_nograd"]


@pytest.mark.parametrize("mode", ["base", "precon", "rescale_origin"])
def test_config1_plt_cond_grid(golden_dir, mode):
    """BASELINE configs[0]: the loop of gpgradpy/plt/plt_cond.py:154-207 (2-D Rosenbrock, n = 20, grid over the length
    scales, one nugget for every mode, LML and condition number per grid point through
    calc_lkd_all(calc_lkd=True, calc_cond=True, calc_grad=False)) against the values the reference produces."""
    from gpgradpy_b200.gp import GaussianProcess
    g = _load(golden_dir, "c1_plt_cond_grid")
    x, f, gr, tv = g["x"], g["fval"], g["grad"], g["theta_vec"]
    GP = GaussianProcess(2, True, "SqExp", mode)
    GP.cond_eta_is_const = True
    GP.set_data(x, f, np.zeros(20), gr, np.zeros(gr.shape))
    GP._etaK = float(g["nugget"])
    GP._eta_Kgrad = float(g["nugget"])
    GP.cond_max = 1e10
    assert GP.Rtensor_init is not None or mode == "rescale_origin"
    n_cmp = 0
    for i in range(tv.size):
        for j in range(tv.size):
            hp = GP.make_hp_class(varK=float(g["varK"]), theta=np.array([tv[i], tv[j]]), kernel=GP.hp_kernel_default)
            info, ok = GP.calc_lkd_all(hp, calc_lkd=True, calc_cond=True, calc_grad=False)
            ref_l, ref_c = g[mode + "_lkd"][i, j], g[mode + "_cond"][i, j]
            if ref_c > 1e13:
                # numerically singular: whether cond passes cond_max_abs = 1e16 (kernel/Kernel.py:282-283) or the
                # Cholesky survives is decided by rounding noise in either implementation
                assert (not ok) or np.isfinite(info.ln_lkd)
                assert info.cond > 1e11
                continue
            assert ok == bool(np.isfinite(ref_l)), (i, j, ok, ref_l, ref_c)
            if not ok:
                assert info.ln_lkd is None and info.cond > 1e12          # "numerically singular" either way
                continue
            # LML: relative 1e-8 where cond(K) allows it (two LAPACK runs differ by ~eps * cond)
            tol = max(1e-8, 1e-15 * ref_c)
            assert abs(info.ln_lkd - ref_l) < tol * max(1.0, abs(ref_l)), (i, j, info.ln_lkd, ref_l, ref_c)
            if ref_c < 1e12:
                assert abs(info.cond - ref_c) < max(1e-8, 1e-15 * ref_c) * ref_c, (i, j, info.cond, ref_c)
                n_cmp += 1
    assert n_cmp >= 10


def test_noisy_fit_vs_reference(golden_dir):
    """Fit with known observation noise (varK becomes a numerically optimised hyper-parameter, optz/CalcLkd.py:185-251):
    from the start point the reference's scan picked, the optimum must be as good as the reference's."""
    g = _load(golden_dir, "fit_d2_n20_noisy_precon")
    from gpgradpy_b200.gp import GaussianProcess
    from scipy.optimize import Bounds
    x, f, gr = g["x"], g["fval"], g["grad"]
    n, d = x.shape
    sf, sg = float(g["std_f"]), float(g["std_g"])
    GP = GaussianProcess(d, True, "SqExp", "precon")
    GP.init_optz_surr(3)
    GP.set_data(x[:1], f[:1], sf * np.ones(1), gr[:1], sg * np.ones((1, d)))
    GP.set_hpara("optz", 0)
    GP.set_data(x, f, sf * np.ones(n), gr, sg * np.ones((n, d)))
    assert GP.b_has_noisy_data and GP.hp_info_optz_lkd.has_varK
    best, _, info = GP.optz_hp_max_lkd(g["hp_x0"], Bounds(g["lb"], g["ub"], keep_feasible=True))
    hp = GP.optz_closed_form_hp(GP.hp_vec2dataclass(GP.hp_info_optz_lkd, best))
    res, ok = GP.calc_lkd_all(hp, calc_grad=True)
    assert ok
    assert res.ln_lkd > float(g["ln_lkd"]) - 1e-4 * abs(float(g["ln_lkd"]))
    glog = res.ln_lkd_grad * np.hstack((hp.theta, [hp.varK])) * np.log(10)
    assert np.max(np.abs(glog)) < 1e-2 * abs(res.ln_lkd)
    assert np.allclose(hp.theta, g["theta"], rtol=5e-2) and abs(hp.varK - float(g["varK"])) < 5e-2 * float(g["varK"])
    GP.set_hpara("optz", 1)                       # the whole public path, incl. the candidate scan
    assert GP.hp_vals.varK > 0
