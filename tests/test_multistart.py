"""CPU: the lock-step multi-start driver (gpgradpy_b200/multistart.py, batch point B of optz/OptzLkd.py:249-270).
Every SLSQP instance must follow exactly the trajectory it follows alone, while the objective requests of all live
instances arrive as batches; plus the world-size-2 (gloo) sharding of start rows through the GaussianProcess method
with the device evaluator replaced by the CPU oracle."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from scipy.optimize import Bounds, minimize

from gpgradpy_b200 import multistart


def _rosen(x):
    return float(np.sum(10.0 * (x[1:] - x[:-1] ** 2) ** 2 + (1 - x[:-1]) ** 2))


def _rosen_grad(x):
    g = np.zeros_like(x)
    g[:-1] += -40.0 * x[:-1] * (x[1:] - x[:-1] ** 2) - 2 * (1 - x[:-1])
    g[1:] += 20.0 * (x[1:] - x[:-1] ** 2)
    return g


def test_lockstep_equals_sequential():
    rng = np.random.default_rng(0)
    x0 = rng.uniform(-1.5, 1.5, (6, 4))
    bounds = Bounds(-2 * np.ones(4), 2 * np.ones(4), keep_feasible=True)
    opt = {"ftol": 1e-12, "eps": 1e-12, "maxiter": 250, "disp": False}
    sizes = []

    def batch(X):
        sizes.append(X.shape[0])
        return np.array([_rosen(x) for x in X]), np.array([_rosen_grad(x) for x in X])

    res, ev = multistart.minimize_lockstep(batch, x0, bounds, opt)
    for i in range(6):
        ref = minimize(_rosen, x0[i], method="SLSQP", jac=_rosen_grad, bounds=bounds, options=opt)
        assert np.array_equal(res[i].x, ref.x) and res[i].fun == ref.fun and res[i].nit == ref.nit
    assert sizes[0] == 6 and max(sizes) == 6 and min(sizes) >= 1          # batched while several starts are alive
    assert ev.n_evals == sum(sizes) and ev.n_batches == len(sizes)
    assert ev.n_batches < ev.n_evals                                       # fewer device calls than evaluations


def test_lockstep_propagates_errors():
    def bad(X):
        raise ValueError("boom")
    with pytest.raises(RuntimeError):
        multistart.minimize_lockstep(bad, np.zeros((3, 2)), None, {"maxiter": 5})


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _make_gp(monkeypatch_obj=None):
    """GaussianProcess whose device evaluator is replaced by the CPU oracle (host logic under test only)."""
    from gpgradpy_b200.gp import GaussianProcess
    from gpgradpy_b200 import _lib as L
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(10, 2, 0)
    GP = GaussianProcess(2, True, "SqExp", "precon")
    GP.set_data(x, f, np.zeros(10), g, np.zeros((10, 2)))
    eta = GP._etaK

    def fake_eval_rows(theta_rows, *, want_grad, **kw):
        th = np.atleast_2d(np.asarray(theta_rows, dtype=float))
        out = np.zeros((th.shape[0], L.out_len(2)))
        for r in range(th.shape[0]):
            o = O.lkd_wo_noise(x, f, g, th[r], "precon", eta, calc_grad=want_grad)
            out[r, L.OUT_LML], out[r, L.OUT_SIGMA2], out[r, L.OUT_BETA] = o.ln_lkd, o.hp_varK, o.hp_beta[0]
            if want_grad:
                out[r, L.OUT_GRAD:] = o.ln_lkd_grad
        return torch.as_tensor(out)

    GP._eval_rows = fake_eval_rows
    GP._final_cond = lambda best: np.nan     # the closing condition number is a device computation (GPU tests cover it)
    return GP


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    GP = _make_gp()
    x0 = np.log10(10.0 ** np.random.default_rng(1).uniform(-2.5, 0.0, (5, 2)))
    bound = Bounds(-5 * np.ones(2), 1 * np.ones(2), keep_feasible=True)
    best, cond, info = GP.optz_hp_max_lkd(x0, bound)
    q.put((rank, best, info, GP._lockstep_stats))
    dist.barrier()
    dist.destroy_process_group()


def test_multistart_sharded_world2():
    """5 start rows over 2 ranks (3 + 2): both ranks end with the same best point, equal to the one-rank lock-step run
    and to the sequential loop."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=240) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
    GP = _make_gp()
    x0 = np.log10(10.0 ** np.random.default_rng(1).uniform(-2.5, 0.0, (5, 2)))
    bound = Bounds(-5 * np.ones(2), 1 * np.ones(2), keep_feasible=True)
    best1, _, info1 = GP.optz_hp_max_lkd(x0, bound)                 # one rank, lock step
    GP.lockstep_multistart = False
    best_seq, _, info_seq = GP.optz_hp_max_lkd(x0, bound)           # the reference's sequential loop
    assert np.array_equal(best1, best_seq) and info1["hp_optz_iter_mean"] == info_seq["hp_optz_iter_mean"]
    for rank, best, info, stats in res:
        assert np.array_equal(best, best1)
        assert info["hp_optz_iter_mean"] == info1["hp_optz_iter_mean"]
        assert stats["batch_sizes"][0] == (3 if rank == 0 else 2)


def test_trust_constr_option_reaches_the_same_optimum():
    """optz_mtd='trust-constr' (optz/OptzLkd.py:216-221) through the same callbacks: same optimum as SLSQP."""
    x0 = np.array([[-1.0, -0.5]])
    bound = Bounds(-5 * np.ones(2), 1 * np.ones(2), keep_feasible=True)
    GP = _make_gp()
    b1 = GP.optz_hp_max_lkd(x0, bound)[0]
    GP.optz_mtd = "trust-constr"
    b2 = GP.optz_hp_max_lkd(x0, bound)[0]
    f1, f2 = GP.return_optz_val(b1), GP.return_optz_val(b2)
    assert abs(f1 - f2) < 1e-6 * max(1.0, abs(f1))
    GP.optz_mtd = "nope"
    with pytest.raises(Exception):
        GP.optz_hp_max_lkd(x0, bound)
