"""GPU parity of the dense fp64 engine (DMMA GEMM, trapezoid Cholesky, solves) through the C ABI."""
import numpy as np
import pytest
import scipy.linalg as sla

pytestmark = pytest.mark.gpu


def _spd(N, seed, cond_shift=None):
    rng = np.random.default_rng(seed)
    G = rng.standard_normal((N, N + 8))
    K = G @ G.T / (N + 8)
    K += (cond_shift if cond_shift is not None else 0.5) * np.eye(N)
    return K


@pytest.mark.parametrize("N,extra", [(7, 0), (60, 2), (128, 1), (129, 3), (185, 2), (300, 0), (640, 5), (1000, 2),
                                     (2500, 130)])
def test_potrf_trapezoid(N, extra):
    import torch
    from gpgradpy_b200 import backend as bk
    K = _spd(N, N)
    rng = np.random.default_rng(N + 1)
    R = rng.standard_normal((extra, N))
    ld = bk.ld_of(N)
    A = torch.full((N + extra, ld), float("nan"), dtype=torch.float64, device="cuda")
    A[:N, :N] = torch.as_tensor(np.tril(K)).cuda() + torch.triu(torch.full((N, N), float("nan"), dtype=torch.float64, device="cuda"), 1)
    if extra:
        A[N:, :N] = torch.as_tensor(R).cuda()
    info = bk.potrf(A, N, extra)
    torch.cuda.synchronize()
    assert int(info.item()) == 0
    Lg = np.tril(A[:N, :N].cpu().numpy())
    Lr = np.linalg.cholesky(K)
    err = np.abs(Lg - Lr).max() / np.abs(Lr).max()
    res = np.linalg.norm(Lg @ Lg.T - K) / np.linalg.norm(K)
    print(f"N={N} L err {err:.2e} residual {res:.2e}")
    assert err < 1e-11 and res < 1e-13
    if extra:
        Zr = sla.solve_triangular(Lr, R.T, lower=True).T
        Zg = A[N:, :N].cpu().numpy()
        e2 = np.abs(Zg - Zr).max() / np.abs(Zr).max()
        print(f"   solved rows err {e2:.2e}")
        assert e2 < 1e-10


def test_potrf_not_pd_reports_info():
    import torch
    from gpgradpy_b200 import backend as bk
    N = 200
    K = _spd(N, 3)
    K[150, 150] = -1.0
    ld = bk.ld_of(N)
    A = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
    A[:, :N] = torch.as_tensor(np.tril(K)).cuda()
    info = bk.potrf(A, N, 0)
    assert int(info.item()) == 151


@pytest.mark.parametrize("N,r", [(50, 3), (128, 200), (333, 17), (1100, 260)])
def test_trsm_rows(N, r):
    import torch
    from gpgradpy_b200 import backend as bk
    K = _spd(N, 10 + N)
    Lr = np.linalg.cholesky(K)
    ld = bk.ld_of(N)
    Lt = torch.full((N, ld), float("nan"), dtype=torch.float64, device="cuda")
    Lt[:, :N] = torch.as_tensor(np.tril(Lr)).cuda() + torch.triu(torch.full((N, N), float("nan"), dtype=torch.float64, device="cuda"), 1)
    rng = np.random.default_rng(5)
    Bm = rng.standard_normal((r, N))
    Bt = torch.zeros((r, ld), dtype=torch.float64, device="cuda")
    Bt[:, :N] = torch.as_tensor(Bm).cuda()
    bk.trsm_rows(Lt, N, Bt)
    Zr = sla.solve_triangular(Lr, Bm.T, lower=True).T
    e = np.abs(Bt[:, :N].cpu().numpy() - Zr).max() / np.abs(Zr).max()
    print(f"N={N} r={r} err {e:.2e}")
    assert e < 1e-10
