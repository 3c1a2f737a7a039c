"""CPU: host-side mirror of the reference API (no device work): nuggets, hp packing, bounds, rescaling, sharding."""
import os

import numpy as np
import pytest

from gpgradpy_b200 import hpara as H
from gpgradpy_b200.gp import GaussianProcess
from gpgradpy_b200.parallel import shard_bounds
from oracle import gegp_oracle as O


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name + ".npz"), allow_pickle=False)
    return {k: z[k] for k in z.files}


def _gp(x, f, g, mode, **kw):
    n, d = x.shape
    GP = GaussianProcess(d, True, "SqExp", mode)
    GP.set_data(x, f, np.zeros(n), g, np.zeros(g.shape), **kw)
    return GP


@pytest.mark.parametrize("name,mode", [("c1_d2_n20_precon", "precon"), ("c1_d2_n20_base", "base"),
                                       ("d3_n25_rescale_origin", "rescale_origin"), ("c2_d10_n500_precon", "precon")])
def test_nugget_matches_reference(golden_dir, name, mode):
    g = _load(golden_dir, name)
    GP = _gp(g["x"], g["fval"], g["grad"], mode)
    assert abs(GP._etaK - float(g["eta"])) <= 1e-15 * float(g["eta"])
    n, d = g["x"].shape
    assert abs(GP._etaK - O.nugget(n, d, mode)[1]) <= 1e-15 * GP._etaK


def test_rescale_origin_matches_reference(golden_dir):
    g = _load(golden_dir, "d3_n25_rescale_origin")
    GP = _gp(g["x"], g["fval"], g["grad"], "rescale_origin")
    xs = GP.get_scl_x_w_dist()[0]
    fs, _, gs, _ = GP.get_scl_eval_data()
    assert np.array_equal(xs, g["x_scl"]) and np.allclose(fs, g["fval_scl"], rtol=1e-15, atol=0)
    assert np.allclose(gs, g["grad_scl"], rtol=1e-15, atol=0)
    # round trip of the prediction scaling
    mu, sig = GP.data_scl_2_init(fs, np.abs(fs))[:2]
    assert np.allclose(mu, g["fval"], rtol=1e-13, atol=1e-12)


def test_flags_and_data_vector():
    x, f, g = O.synthetic_problem(7, 3, 0)
    GP = _gp(x, f, g, "precon")
    assert GP.n_data == 7 * 4 and GP.b_has_noisy_data is False and GP.b_use_cond_cstr is False
    assert np.array_equal(GP.make_data_vec(f, g), O.make_data_vec(f, g))
    hi = GP.hp_info_optz_lkd
    assert hi.n_hp == 3 and hi.has_theta and not hi.has_varK and hi.bvec_log_optz.all()
    hp = GP.hp_vec2dataclass(hi, np.array([-1.0, 0.0, 0.5]))
    assert np.allclose(hp.theta, [0.1, 1.0, 10 ** 0.5])
    # noisy data with unknown noise variance adds varK / var_fval / var_fgrad to the optimised vector
    GP2 = GaussianProcess(3, True, "SqExp", "precon")
    GP2.set_data(x, f, None, g, None)
    hi2 = GP2.hp_info_optz_lkd
    assert GP2.b_has_noisy_data and hi2.n_hp == 6 and hi2.idx_varK == 3 and hi2.idx_var_fval == 4 and hi2.idx_var_fgrad == 5
    nv = GP2.calc_noise_vec(GP2.make_hp_class(var_fval=0.25, var_fgrad=2.0))
    assert nv.shape == (28,) and np.all(nv[:7] == 0.25) and np.all(nv[7:] == 2.0)
    # known noise: gradient std is flattened Fortran-order (kernel/Kernel.py:353)
    GP3 = GaussianProcess(3, True, "SqExp", "base")
    sg = np.arange(21, dtype=float).reshape(7, 3) + 1
    GP3.set_data(x, f, 0.1 * np.ones(7), g, sg)
    nv3 = GP3.calc_noise_vec(GP3.make_hp_class())
    assert np.allclose(nv3[7:], (sg ** 2).reshape(-1, order="F")) and np.allclose(nv3[:7], 0.01)


def test_mode_table():
    for mode, scl, cstr in [("precon", False, False), ("base", False, True), ("rescale_origin", True, True),
                            ("dflt_vmin", True, True)]:
        GP = GaussianProcess(2, True, "SqExp", mode)
        assert GP.b_use_data_scl == scl and GP.b_use_cond_cstr == cstr
    assert GaussianProcess(2, False, "SqExp", "precon").wellcond_mtd == "base"   # GaussianProcess.py:202-203
    with pytest.raises(AssertionError):
        GaussianProcess(2, True, "SqExp", "req_vmin")                              # stale name of the reference tests
    with pytest.raises(Exception):
        GaussianProcess(2, True, "Matern")                                         # kernel/Kernel.py:105-107
    for kt, has_hp in (("SqExp", False), ("Ma5f2", False), ("RatQu", True)):       # kernel/Kernel.py:109
        GP = GaussianProcess(2, True, kt)
        assert GP.kernel_has_hp == has_hp and (GP.hp_kernel_default == 2) == has_hp


def test_lhs_bounds_follow_history():
    x, f, g = O.synthetic_problem(9, 2, 0)
    GP = _gp(x, f, g, "precon")
    GP.init_optz_surr(4)
    GP.hp_theta_all[0, :] = [1e-2, 1e-2]
    GP.hp_theta_all[1, :] = [1.0, 4.0]
    x0, bounds = GP.get_hp_x0_lhs_median(2, GP.hp_info_optz_lkd, 40)
    med = np.log10(np.median(GP.hp_theta_all[:2], axis=0))
    assert x0.shape == (40, 2)
    assert np.allclose(bounds.lb, med - 5) and np.allclose(bounds.ub, med + 5)
    assert np.all(x0 >= med - 3 - 1e-12) and np.all(x0 <= med + 3 + 1e-12)
    # one sample per stratum in every dimension (Latin hypercube)
    strata = np.floor((x0 - (med - 3)) / 6.0 * 40).astype(int)
    assert all(len(set(strata[:, j])) == 40 for j in range(2))


def test_shard_bounds_partition():
    for B in (1, 7, 40, 1024):
        for size in (1, 2, 3, 8):
            if size > B:
                continue
            parts = [shard_bounds(B, r, size) for r in range(size)]
            assert parts[0][0] == 0 and parts[-1][1] == B
            assert all(parts[i][1] == parts[i + 1][0] for i in range(size - 1))
            lens = [hi - lo for lo, hi in parts]
            assert max(lens) - min(lens) <= 1


def test_chain_rule_cache_and_failure_contract(monkeypatch):
    """calc_store_likelihood: log10 chain factor (optz/OptzLkd.py:65-70), one evaluation per point, -cond on failure."""
    x, f, g = O.synthetic_problem(9, 2, 0)
    GP = _gp(x, f, g, "precon")
    calls = []

    def fake(hp_vals, calc_lkd=True, calc_cond=False, calc_grad=False, lkd_use_adj_mtd=None):
        calls.append(hp_vals.theta.copy())
        return H.LkdInfo(ln_lkd=-3.0, ln_lkd_grad=np.array([2.0, -1.0]), cond=7.0), True

    monkeypatch.setattr(GP, "calc_lkd_all", fake)
    v = np.array([-1.0, 0.5])
    assert GP.return_optz_val(v) == 3.0
    gr = GP.return_optz_grad(v)
    assert len(calls) == 1                                   # second callback served from the cache
    assert np.allclose(gr, -np.array([2.0, -1.0]) * 10 ** v * np.log(10))
    GP.return_optz_val(v + 1e-3)
    assert len(calls) == 2
    monkeypatch.setattr(GP, "calc_lkd_all", lambda *a, **k: (H.LkdInfo(cond=123.0), False))
    GP._last_hp_vec = None
    assert GP.return_optz_val(v) == 123.0                    # objective = -(-cond)


@pytest.mark.parametrize("kernel_type", ["SqExp", "Ma5f2", "RatQu"])
@pytest.mark.parametrize("b_return_vec", [True, False])
def test_reference_unit_test_precon_grad(b_return_vec, kernel_type):
    """gpgradpy/unit_test/test_precon_grad.py:22-95 (the one reference test module that passes as shipped; it loops over
    the three kernels): analytic dP/dtheta of the preconditioner against a forward finite difference, vector and
    matrix form."""
    from gpgradpy_b200.gp import GaussianProcess
    eps, dim, n_eval = 1e-6, 2, 1
    theta = np.linspace(2.5, 3, dim)
    GP = GaussianProcess(dim, True, kernel_type, "precon")
    c = 5.0 / 3.0 if kernel_type == "Ma5f2" else 2.0
    np.testing.assert_allclose(GP.theta2gamma(theta), np.sqrt(c * theta))
    np.testing.assert_allclose(GP.gamma2theta(GP.theta2gamma(theta)), theta)
    pvec, pvec_inv, grad = GP.calc_Kern_precon(n_eval, n_eval, theta, calc_grad=True, b_return_vec=b_return_vec)
    n_data = n_eval * (dim + 1)
    fd = np.zeros((n_data, dim)) if b_return_vec else np.zeros((dim, n_data, n_data))
    for i in range(dim):
        tp = theta.copy()
        tp[i] += eps
        pe = GP.calc_Kern_precon(n_eval, n_eval, tp, calc_grad=False, b_return_vec=b_return_vec)[0]
        if b_return_vec:
            fd[:, i] = (pe - pvec) / eps
        else:
            fd[i] = (pe - pvec) / eps
    np.testing.assert_allclose(grad, fd, rtol=1e-4, atol=1e-8)
    np.testing.assert_allclose(pvec @ pvec_inv if not b_return_vec else pvec * pvec_inv,
                               np.eye(n_data) if not b_return_vec else np.ones(n_data))


def test_host_bookkeeping_table_vs_reference(golden_dir):
    """Every conditioning mode x noise model x use_grad: the flags set_data derives, nuggets, hyper-parameter index
    maps, scaled data, noise vectors, vector <-> dataclass maps, initial hyper-parameters, LHS start points and box
    bounds must equal what the reference produces (tests/golden/host_logic_table.npz, oracle/make_golden_host.py walks
    both classes with the same function).  The LHS sampler is scipy's here and in the reference run (its `smt` import
    is shimmed), so start points are comparable."""
    import os
    from oracle.make_golden_host import collect
    from gpgradpy_b200.gp import GaussianProcess
    z = np.load(os.path.join(golden_dir, "host_logic_table.npz"))
    x, f, g = O.synthetic_problem(14, 3, 0)
    mine = collect(GaussianProcess, x, f, g)
    assert set(mine) == set(z.files)
    bad = []
    for k in z.files:
        a, b = z[k], mine[k]
        if a.dtype.kind in "US" or b.dtype.kind in "US":
            ok = str(a) == str(b)
        else:
            ok = a.shape == b.shape and np.allclose(a, b, rtol=1e-13, atol=0, equal_nan=True)
        if not ok:
            bad.append(k)
    assert not bad, bad[:10]


def test_device_matrix_is_built_on_first_use_only():
    """The matrices of calc_all_K_w_chofac's 7-tuple are produced when (if) somebody reads them -- kernel/Kernel.py:140-307
    returns them all, but at N = 21000 each is 3.5 GB and the fit only ever asks for the condition number."""
    import torch
    from gpgradpy_b200.gp import DeviceMatrix
    calls = []

    def make():
        calls.append(1)
        return torch.arange(6, dtype=torch.float64).reshape(2, 3)

    m = DeviceMatrix(make, shape=(2, 3))
    assert m.shape == (2, 3) and calls == []            # the shape is known without building
    assert np.asarray(m)[1, 2] == 5.0 and calls == [1]
    assert m[0, 1] == 1.0 and m.tensor.shape == (2, 3) and calls == [1]   # built once, cached
    eager = DeviceMatrix(torch.ones(2, 2, dtype=torch.float64))
    assert eager.shape == (2, 2) and np.asarray(eager).sum() == 4.0
