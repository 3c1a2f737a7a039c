"""Thin PyTorch-side plumbing over the C ABI: device buffers, streams and workspace management.

Every function here hands raw device pointers to libgegp.so; no arithmetic of the hot path is done in
Python/PyTorch, and there is no CPU fallback (a missing CUDA device or library raises).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib as L

F64 = torch.float64


def device() -> torch.device:
    if not torch.cuda.is_available():
        raise RuntimeError("gpgradpy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


def to_dev(a, dtype=F64):
    """Host array / tensor -> contiguous device tensor (no copy if already there)."""
    if a is None:
        return None
    if isinstance(a, torch.Tensor):
        return a.to(device=device(), dtype=dtype).contiguous()
    return torch.as_tensor(np.ascontiguousarray(a), dtype=dtype).to(device())


def pinned_like(buf, a):
    """Copy the host array `a` into a page-locked fp64 staging tensor (re-using `buf` when the shape matches)."""
    a = np.ascontiguousarray(a, dtype=np.float64)
    if buf is None or tuple(buf.shape) != a.shape:
        buf = torch.empty(a.shape, dtype=F64, pin_memory=True)
    buf.copy_(torch.from_numpy(a))
    return buf


def _p(t) -> int:
    return 0 if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _check(rc: int, what: str):
    if rc != 0:
        raise RuntimeError(f"{what} failed with code {rc} (negative: bad argument index / -1000-cudaError)")


def dmma_peak_tflops(reps: int = 3) -> float:
    """gegp_dmma_peak: issue peak of DMMA.8x8x4 on the current device in TFLOP/s (the fp64 roofline denominator)."""
    lib = L.load()
    sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
    scratch = torch.empty(512 * sms, dtype=F64, device=device())
    out = C.c_double(0.0)
    torch.cuda.synchronize()
    rc = lib.gegp_dmma_peak(_p(scratch), scratch.numel(), int(reps), C.byref(out), _stream())
    _check(rc, "gegp_dmma_peak")
    return float(out.value)


def _kern(kernel):
    """kernel = None | name | id | (name-or-id, hp)  ->  (GEGP_KERNEL_* id, hp float)."""
    if kernel is None:
        return L.KERNEL_SQEXP, 0.0
    hp = 0.0
    if isinstance(kernel, (tuple, list)):
        kernel, hp = kernel[0], (0.0 if kernel[1] is None else float(np.ravel(kernel[1])[0]))
    kid = L.KERNEL_IDS[kernel] if isinstance(kernel, str) else int(kernel)
    return kid, hp


def ld_of(N: int) -> int:
    return int(L.load().gegp_ld(N))


def slot_from_mask(mask, n):
    """bvec_use_grad (bool[n]) -> (slot int32[n] device tensor or None, n_g)."""
    if mask is None:
        return None, n
    m = np.asarray(mask, dtype=bool)
    assert m.size == n
    slot = np.where(m, np.cumsum(m) - 1, -1).astype(np.int32)
    ng = int(m.sum())
    if ng == n:
        return None, n
    return torch.as_tensor(slot).to(device()), ng


_ws_cache: dict = {}


def workspace(nbytes: int) -> torch.Tensor:
    """Grow-only byte workspace per device (torch caching allocator owns the memory)."""
    key = torch.cuda.current_device()
    buf = _ws_cache.get(key)
    if buf is None or buf.numel() < nbytes:
        _ws_cache.pop(key, None)
        buf = None
        buf = torch.empty(int(nbytes), dtype=torch.uint8, device=device())
        _ws_cache[key] = buf
    return buf


def free_workspace():
    _graph_cache.clear()
    _ws_cache.clear()


def build_cov(X, theta, *, n_g=None, slot=None, noise=None, mode=L.MODE_BASE, eta=0.0, varK=1.0, uplo=0, out=None,
              kernel=None):
    """gegp_build_cov -> (K [N, N] strided view over an [N, ld] buffer, p [2N] or None)."""
    lib = L.load()
    X = to_dev(X)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    N = n + n_g * d
    theta = to_dev(theta)
    noise = to_dev(noise)
    ld = ld_of(N)
    if out is None:
        out = torch.empty((N, ld), dtype=F64, device=device())
    p = torch.empty(2 * N, dtype=F64, device=device()) if mode == L.MODE_PRECON else None
    kid, khp = _kern(kernel)
    rc = lib.gegp_build_cov(n, n_g, d, _p(X), _p(slot), _p(theta), kid, khp, _p(noise), mode, float(eta), float(varK),
                            _p(out), out.stride(0), _p(p), int(uplo), _stream())
    _check(rc, "gegp_build_cov")
    return out[:, :N], p


def cross_cov(X, Xs, theta, *, n_g=None, slot=None, pinv=None, kernel=None):
    lib = L.load()
    X, Xs, theta, pinv = to_dev(X), to_dev(Xs), to_dev(theta), to_dev(pinv)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    N = n + n_g * d
    nx = Xs.shape[0]
    ld = ld_of(N)
    out = torch.empty((nx, ld), dtype=F64, device=device())
    kid, khp = _kern(kernel)
    rc = lib.gegp_cross_cov(n, n_g, d, _p(X), _p(slot), _p(Xs), nx, _p(theta), kid, khp, _p(pinv), _p(out), ld, _stream())
    _check(rc, "gegp_cross_cov")
    return out[:, :N]


def dinv_buffer(N: int) -> torch.Tensor:
    """Device buffer for the inverse-transposed diagonal blocks that belong to a factor of order N."""
    return torch.empty(int(L.load().gegp_dinv_doubles(N)), dtype=F64, device=device())


def potrf(A: torch.Tensor, N: int, n_extra: int = 0, dinv: torch.Tensor | None = None):
    """In-place trapezoid Cholesky of A[(N+n_extra), ld]; returns (device info int tensor, dinv blocks)."""
    lib = L.load()
    assert A.is_cuda and A.dtype == F64 and A.stride(1) == 1 and A.shape[0] >= N + n_extra
    info = torch.zeros(1, dtype=torch.int32, device=A.device)
    if dinv is None:
        dinv = dinv_buffer(N)
    rc = lib.gegp_potrf(N, n_extra, _p(A), A.stride(0), _p(dinv), _p(info), _stream())
    _check(rc, "gegp_potrf")
    return info, dinv


def trsm_rows(Lfac: torch.Tensor, dinv: torch.Tensor, N: int, B: torch.Tensor):
    lib = L.load()
    rc = lib.gegp_trsm_rows(N, _p(Lfac), Lfac.stride(0), _p(dinv), _p(B), B.stride(0), B.shape[0], _stream())
    _check(rc, "gegp_trsm_rows")
    return B


def potri(Lfac: torch.Tensor, dinv: torch.Tensor, N: int, U: torch.Tensor | None = None,
          Kinv: torch.Tensor | None = None):
    """Explicit inverse from the factor -> (U = L^-T in its upper triangle, Kinv full symmetric), [N, ld] buffers."""
    lib = L.load()
    ld = ld_of(N)
    if U is None:
        U = torch.empty((N, ld), dtype=F64, device=device())
    if Kinv is None:
        Kinv = torch.empty((N, ld), dtype=F64, device=device())
    rc = lib.gegp_potri(N, _p(Lfac), Lfac.stride(0), _p(dinv), _p(U), U.stride(0), _p(Kinv), Kinv.stride(0), _stream())
    _check(rc, "gegp_potri")
    return U, Kinv


def dgemm(A: torch.Tensor, B: torch.Tensor, C: torch.Tensor, *, transb: bool, alpha=1.0, beta=0.0):
    """C = alpha * A @ (B.T if transb else B) + beta * C on the DMMA engine (row-major 2-D device tensors)."""
    lib = L.load()
    M, K = A.shape
    N = B.shape[0] if transb else B.shape[1]
    assert (B.shape[1] if transb else B.shape[0]) == K and tuple(C.shape) == (M, N)
    rc = lib.gegp_dgemm(int(transb), M, N, K, float(alpha), _p(A), A.stride(0), _p(B), B.stride(0), float(beta), _p(C),
                        C.stride(0), _stream())
    _check(rc, "gegp_dgemm")
    return C


def _lml_workspace(lib, op, n, n_g, d, B, max_ws_bytes):
    """Workspace for B candidates (or the largest chunk of them that fits); cudaMemGetInfo only when it must grow."""
    need_all = int(lib.gegp_workspace_bytes(op, n, n_g, d, B)) + 4 * B + 256
    cached = _ws_cache.get(torch.cuda.current_device())
    if cached is not None and cached.numel() >= need_all and (max_ws_bytes is None or need_all <= max_ws_bytes):
        return cached
    per1 = int(lib.gegp_workspace_bytes(op, n, n_g, d, 1))
    if max_ws_bytes is None:
        free, _total = torch.cuda.mem_get_info()
        max_ws_bytes = int(0.6 * (free + (cached.numel() if cached is not None else 0)))
    if per1 + 4096 + 4 * B > max_ws_bytes:
        raise MemoryError(f"one candidate of this size needs a {per1 / 2**30:.1f} GiB workspace (N = {n + n_g * d}, "
                          f"{'with' if op == L.OP_LML_GRAD else 'without'} gradient) but only {max_ws_bytes / 2**30:.1f} GiB "
                          "may be used on this device")
    chunk = max(1, min(B, (max_ws_bytes - 4096 - 4 * B) // per1))
    return workspace(int(lib.gegp_workspace_bytes(op, n, n_g, d, chunk)) + 4 * B + 256)


def lml_eval(X, y, theta_batch, *, n_g=None, slot=None, mode=L.MODE_PRECON, eta=0.0, noise=None, varK_batch=None,
             pnlt_grad=0.0, want_grad=True, want_alpha=False, max_ws_bytes=None, out=None, kernel=None, kernel_hp_batch=None):
    """gegp_lml_eval for B candidate rows -> (out [B, 9+d] device tensor, alpha [B, N] or None)."""
    lib = L.load()
    X, y = to_dev(X), to_dev(y)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    N = n + n_g * d
    th = to_dev(theta_batch).reshape(-1, d)
    B = th.shape[0]
    noisy = noise is not None
    noise = to_dev(noise)
    vk = to_dev(varK_batch).reshape(-1) if noisy else None
    if noisy:
        assert vk.numel() == B
    if out is None:
        out = torch.empty((B, L.out_len(d)), dtype=F64, device=device())
    alpha = torch.empty((B, N), dtype=F64, device=device()) if want_alpha else None
    op = L.OP_LML_GRAD if want_grad else L.OP_LML
    ws = _lml_workspace(lib, op, n, n_g, d, B, max_ws_bytes)
    kid, khp = _kern(kernel)
    khp_dev = None
    if kid == L.KERNEL_RATQUAD:      # one alpha per candidate row (a scalar is broadcast)
        khp_dev = to_dev(kernel_hp_batch).reshape(-1) if kernel_hp_batch is not None else torch.full(
            (B,), khp, dtype=F64, device=device())
        assert khp_dev.numel() == B
    rc = lib.gegp_lml_eval(B, _p(th), _p(vk), kid, _p(khp_dev), n, n_g, d, _p(X), _p(slot), _p(y), _p(noise), int(mode),
                           float(eta), int(noisy), float(pnlt_grad), int(bool(want_grad)), _p(out), _p(alpha), _p(ws),
                           ws.numel(), _stream())
    _check(rc, "gegp_lml_eval")
    return out, alpha


class LmlGraph:
    """One captured CUDA graph of gegp_lml_eval for a fixed problem (data, shapes, mode, flags).

    The ~300 kernel launches of one evaluation are recorded once and replayed with new hyper-parameters, which
    are written into static device buffers before the replay.  Used for the optimiser's inner loop, where the
    same-shaped evaluation is repeated hundreds of times (optz/OptzLkd.py:249-270).
    """

    def __init__(self, X, y, B, *, n_g, slot, mode, eta, noise, noisy, pnlt_grad, want_grad, kernel=None):
        n, d = X.shape
        self.key_tensors = (X, y, slot, noise)          # keep the captured buffers alive
        self.theta = torch.empty((B, d), dtype=F64, device=device())
        self.varK = torch.ones(B, dtype=F64, device=device()) if noisy else None
        kid = _kern(kernel)[0]
        self.khp = torch.full((B,), 2.0, dtype=F64, device=device()) if kid == L.KERNEL_RATQUAD else None
        self.out = torch.empty((B, L.out_len(d)), dtype=F64, device=device())
        kw = dict(n_g=n_g, slot=slot, mode=mode, eta=eta, noise=noise if noisy else None, varK_batch=self.varK,
                  pnlt_grad=pnlt_grad, want_grad=want_grad, out=self.out, kernel=kid, kernel_hp_batch=self.khp)
        self.theta.fill_(1.0)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):                   # warm-up: function attributes, tensor maps, workspace growth
            lml_eval(X, y, self.theta, **kw)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self.graph = torch.cuda.CUDAGraph()
        L.profile_begin(False)
        with torch.cuda.graph(self.graph):
            lml_eval(X, y, self.theta, **kw)
        self.n_launches = L.profile_end()["launches"]          # kernels recorded in the graph
        self.ws = _ws_cache.get(torch.cuda.current_device())   # the captured workspace must stay allocated

    def __call__(self, theta_rows, varK_rows=None, khp_rows=None):
        src = theta_rows if isinstance(theta_rows, torch.Tensor) else torch.as_tensor(
            np.ascontiguousarray(theta_rows, dtype=np.float64))
        self.theta.copy_(src.reshape(self.theta.shape), non_blocking=True)
        if self.varK is not None:
            self.varK.copy_(torch.as_tensor(np.ascontiguousarray(varK_rows, dtype=np.float64)).reshape(-1))
        if self.khp is not None:
            self.khp.copy_(torch.as_tensor(np.ascontiguousarray(khp_rows, dtype=np.float64)).reshape(-1))
        self.graph.replay()
        replay_stats["replays"] += 1
        replay_stats["kernel_launches"] += self.n_launches
        return self.out


_graph_cache: dict = {}
replay_stats = {"replays": 0, "kernel_launches": 0}   # kernels launched through graph replays (bench.py reports them)


def lml_eval_graphed(X, y, theta_batch, *, n_g=None, slot=None, mode=L.MODE_PRECON, eta=0.0, noise=None,
                     varK_batch=None, pnlt_grad=0.0, want_grad=True, kernel=None, kernel_hp_batch=None):
    """Same result as lml_eval(...)[0], through a cached CUDA graph.  X, y, slot, noise must be device tensors that
    stay alive and unchanged in place between calls (the graph holds their addresses)."""
    n, d = X.shape
    n_g = n if n_g is None else n_g
    B = int(np.prod(theta_batch.shape)) // d
    noisy = noise is not None
    key = (torch.cuda.current_device(), X.data_ptr(), y.data_ptr(), _p(slot), _p(noise), n, n_g, d, B, int(mode),
           float(eta), float(pnlt_grad), bool(want_grad), noisy, _kern(kernel)[0])
    g = _graph_cache.get(key)
    ws_now = _ws_cache.get(torch.cuda.current_device())
    if g is not None and g.ws is not ws_now:   # the workspace was re-allocated since the capture
        g = None
    if g is None:
        if len(_graph_cache) >= 8:
            _graph_cache.clear()
        g = LmlGraph(X, y, B, n_g=n_g, slot=slot, mode=mode, eta=eta, noise=noise, noisy=noisy, pnlt_grad=pnlt_grad,
                     want_grad=want_grad, kernel=kernel)
        _graph_cache[key] = g
    khp = kernel_hp_batch
    if g.khp is not None and khp is None:
        khp = np.full(B, _kern(kernel)[1])
    return g(theta_batch, varK_batch, khp)


def free_graphs():
    _graph_cache.clear()


class PredictState:
    """Factor + solved residual row kept on the device between setup_eval_model and eval_model."""

    def __init__(self, A, dinv, p, info, alpha, n, n_g, d, N, X, slot, theta, mode, beta, kernel=(0, 0.0)):
        self.A, self.dinv, self.p, self.info, self.alpha = A, dinv, p, info, alpha
        self.n, self.n_g, self.d, self.N = n, n_g, d, N
        self.X, self.slot, self.theta, self.mode, self.beta = X, slot, theta, mode, beta
        self.kid, self.khp = kernel


def predict_setup(X, y, theta, beta, *, n_g=None, slot=None, noise=None, mode=L.MODE_PRECON, eta=0.0,
                  want_alpha=True, kernel=None) -> PredictState:
    lib = L.load()
    X, y, theta, noise = to_dev(X), to_dev(y), to_dev(theta), to_dev(noise)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    N = n + n_g * d
    ld = ld_of(N)
    A = torch.empty((N + 1, ld), dtype=F64, device=device())
    dinv = dinv_buffer(N)
    p = torch.empty(2 * N, dtype=F64, device=device())
    info = torch.zeros(1, dtype=torch.int32, device=device())
    alpha = torch.empty(N, dtype=F64, device=device()) if want_alpha else None
    kid, khp = _kern(kernel)
    rc = lib.gegp_predict_setup(n, n_g, d, _p(X), _p(slot), _p(theta), kid, khp, _p(noise), int(mode), float(eta), _p(y),
                                float(beta), _p(A), ld, _p(dinv), _p(p), _p(alpha), _p(info), _stream())
    _check(rc, "gegp_predict_setup")
    return PredictState(A, dinv, p, info, alpha, n, n_g, d, N, X, slot, theta, mode, float(beta), (kid, khp))


def predict(st: PredictState, Xs, varK: float, *, chunk_bytes: int = 1 << 30):
    """gegp_predict -> (mu, sig, sig2, n_negative) device tensors."""
    lib = L.load()
    Xs = to_dev(Xs)
    nx = Xs.shape[0]
    mu = torch.empty(nx, dtype=F64, device=device())
    sig = torch.empty(nx, dtype=F64, device=device())
    sig2 = torch.empty(nx, dtype=F64, device=device())
    nneg = torch.zeros(1, dtype=torch.int32, device=device())
    row_bytes = ld_of(st.N) * 8
    cx = max(1, min(nx, chunk_bytes // row_bytes))
    ws = workspace(cx * row_bytes)
    rc = lib.gegp_predict(st.n, st.n_g, st.d, _p(st.X), _p(st.slot), _p(st.theta), st.kid, st.khp, _p(st.A), st.A.stride(0), _p(st.dinv),
                          _p(st.p), int(st.mode), st.beta, float(varK), _p(Xs), nx, _p(mu), _p(sig), _p(sig2), _p(nneg), _p(ws),
                          cx * row_bytes, _stream())
    _check(rc, "gegp_predict")
    return mu, sig, sig2, nneg


def predict_grad(st: PredictState, Xs, varK: float, *, chunk_bytes: int = 1 << 30):
    """gegp_predict_grad -> (mu, sig, sig2, dmudx [nx, d], dsigdx [nx, d], n_negative) device tensors."""
    lib = L.load()
    Xs = to_dev(Xs)
    nx, d = Xs.shape[0], st.d
    dev = device()
    mu, sig, sig2 = (torch.empty(nx, dtype=F64, device=dev) for _ in range(3))
    dmu, dsg = (torch.empty((nx, d), dtype=F64, device=dev) for _ in range(2))
    nneg = torch.zeros(1, dtype=torch.int32, device=dev)
    per_x = (d + 1) * ld_of(st.N) * 8
    cx = max(1, min(nx, chunk_bytes // per_x))
    ws = workspace(cx * per_x)
    rc = lib.gegp_predict_grad(st.n, st.n_g, st.d, _p(st.X), _p(st.slot), _p(st.theta), st.kid, st.khp, _p(st.A), st.A.stride(0),
                               _p(st.dinv), _p(st.p), int(st.mode), st.beta, float(varK), _p(Xs), nx, _p(mu), _p(sig),
                               _p(sig2), _p(dmu), _p(dsg), _p(nneg), _p(ws), cx * per_x, _stream())
    _check(rc, "gegp_predict_grad")
    return mu, sig, sig2, dmu, dsg, nneg


def predict_hess(st: PredictState, xs, varK: float):
    """gegp_predict_hess at ONE test point -> (mu, sig, sig2, dmudx [d], dsigdx [d], hess3 [3, d, d], n_negative)."""
    lib = L.load()
    xs = to_dev(xs).reshape(-1)
    d = st.d
    assert xs.numel() == d and st.alpha is not None
    dev = device()
    mu, sig, sig2 = (torch.empty(1, dtype=F64, device=dev) for _ in range(3))
    dmu, dsg = (torch.empty(d, dtype=F64, device=dev) for _ in range(2))
    h3 = torch.empty((3, d, d), dtype=F64, device=dev)
    nneg = torch.zeros(1, dtype=torch.int32, device=dev)
    nbytes = (d + 2) * ld_of(st.N) * 8
    ws = workspace(nbytes)
    rc = lib.gegp_predict_hess(st.n, st.n_g, st.d, _p(st.X), _p(st.slot), _p(st.theta), st.kid, st.khp, _p(st.A), st.A.stride(0),
                               _p(st.dinv), _p(st.p), _p(st.alpha), int(st.mode), st.beta, float(varK), _p(xs), _p(mu),
                               _p(sig), _p(sig2), _p(dmu), _p(dsg), _p(h3), _p(nneg), _p(ws), nbytes, _stream())
    _check(rc, "gegp_predict_hess")
    return mu, sig, sig2, dmu, dsg, h3, nneg


# ----------------------------------------------------------------------------------------------------------------
# 2-norm condition number and its hyper-parameter gradient (kernel/Kernel.py:240,280; optz/GpHparaCon.py:161-235)
# ----------------------------------------------------------------------------------------------------------------
def symv(M: torch.Tensor, x: torch.Tensor, y: torch.Tensor, N: int):
    """y[:N] = M[:N, :N] @ x[:N] (gegp_symv; M is an [N, ld] row-major device buffer)."""
    rc = L.load().gegp_symv(N, _p(M), M.stride(0), _p(x), _p(y), _stream())
    _check(rc, "gegp_symv")
    return y


def row_abs_sum(M: torch.Tensor, N: int) -> torch.Tensor:
    """Gershgorin row sums sum_c |M[r, c]| of an [N, ld] device buffer (gegp_row_abs_sum)."""
    out = torch.empty(N, dtype=F64, device=M.device)
    rc = L.load().gegp_row_abs_sum(N, _p(M), M.stride(0), _p(out), _stream())
    _check(rc, "gegp_row_abs_sum")
    return out


def fro_norm(M: torch.Tensor, N: int) -> float:
    """Frobenius norm of M[:N, :N] (gegp_row_sq_sum, the N row sums added on the host)."""
    out = torch.empty(N, dtype=F64, device=M.device)
    rc = L.load().gegp_row_sq_sum(N, _p(M), M.stride(0), _p(out), _stream())
    _check(rc, "gegp_row_sq_sum")
    return float(np.sqrt(np.sum(out.cpu().numpy())))


def weighted_grad(X, theta, W: torch.Tensor, *, n_g=None, slot=None, eta=0.0, noisy=False, varK=1.0, kernel=None):
    """gegp_weighted_grad: sum(W .* dKcov/dhp) for every hyper-parameter (base mode) -> device row (GEGP_OUT_* layout)."""
    lib = L.load()
    X, theta = to_dev(X), to_dev(theta)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    N = n + n_g * d
    out = torch.zeros(L.out_len(d), dtype=F64, device=device())
    vk = to_dev(np.array([float(varK)]))
    nbytes = int(lib.gegp_quad_grad_work_bytes(n, n_g, d)) + 8 * (N + 2)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device())
    kid, khp = _kern(kernel)
    rc = lib.gegp_weighted_grad(n, n_g, d, _p(X), _p(slot), _p(theta), kid, khp, _p(W), W.stride(0), L.MODE_BASE, float(eta),
                                int(bool(noisy)), _p(vk), _p(out), _p(ws), nbytes, _stream())
    _check(rc, "gegp_weighted_grad")
    return out


def cond_fro(Kfull: torch.Tensor, Kinv: torch.Tensor, N: int, want_weight: bool):
    """Frobenius condition number |K|_F |K^-1|_F and, if asked, the weight matrix of its gradient
    d cond / dK = frac K - K^-3 / frac, frac = |K^-1|_F / |K|_F  (optz/GpHparaCon.py:237-261); the two matrix
    products run on the DMMA GEMM engine (K^-1 is symmetric, so K^-1 K^-1 = K^-1 (K^-1)^T)."""
    nk, ni = fro_norm(Kfull, N), fro_norm(Kinv, N)
    cond = nk * ni
    if not want_weight:
        return cond, None
    ld = ld_of(N)
    K2 = torch.empty((N, ld), dtype=F64, device=Kinv.device)
    K3 = torch.empty((N, ld), dtype=F64, device=Kinv.device)
    dgemm(Kinv[:, :N], Kinv[:, :N], K2[:, :N], transb=True)
    frac = ni / nk
    K3[:, :N].copy_(Kfull[:, :N])                                            # W = frac K - (K^-2 K^-1) / frac, in one GEMM
    dgemm(K2[:, :N], Kinv[:, :N], K3[:, :N], transb=True, alpha=-1.0 / frac, beta=frac)
    return cond, K3


def extreme_eig(M: torch.Tensor, N: int, *, k: int = 40, tol: float = 1e-12, max_cycles: int = 12, v0=None,
                chunk: int = 8):
    """Largest eigenpair of the symmetric device matrix M[:N, :N]: restarted Lanczos with full re-orthogonalisation.

    The matrix-vector products and the orthogonalisation run on the device (gegp_symv, gegp_lanczos_step,
    gegp_lincomb); every `chunk` steps the small tridiagonal eigenproblem is solved on the host and the iteration
    stops as soon as the residual estimate |beta_j s_j| <= tol |lambda| (or an invariant subspace is hit); after k steps
    it restarts from the Ritz vector.  `v0` (device vector) warm-starts the iteration -- e.g. the eigenvector of the
    previous optimiser iterate; it is mixed with the fixed pseudo-random start vector so that no eigen-direction is
    ever excluded.  Returns (lambda, v [N] device tensor with |v| = 1, relative residual estimate, matvecs)."""
    from scipy.linalg import eigh_tridiagonal
    lib = L.load()
    k = int(max(2, min(k, N, 200)))
    ld = ld_of(N)
    dev = device()
    V = torch.zeros((k + 1, ld), dtype=F64, device=dev)
    w = torch.zeros(ld, dtype=F64, device=dev)
    r = torch.zeros(ld, dtype=F64, device=dev)
    ab = torch.zeros((2, k), dtype=F64, device=dev)
    coef = torch.zeros(k, dtype=F64, device=dev)
    one = torch.ones(1, dtype=F64, device=dev)
    rnd = to_dev(np.random.default_rng(12345).standard_normal(N))   # fixed: deterministic results
    if v0 is None:
        V[0, :N] = rnd
    else:
        V[0, :N] = v0[:N] + (1e-3 / np.sqrt(N)) * rnd
    rc = lib.gegp_lincomb(N, 1, _p(V), ld, _p(one), _p(r), _stream())      # normalise
    _check(rc, "gegp_lincomb")
    V[0].copy_(r)
    lam, resid, matvecs = float("nan"), float("inf"), 0
    for _cycle in range(max_cycles):
        j, done, m = 0, False, k
        while j < k and not done:
            jn = min(k, j + chunk)
            for jj in range(j, jn):
                rc = lib.gegp_symv(N, _p(M), M.stride(0), _p(V[jj]), _p(w), _stream())
                _check(rc, "gegp_symv")
                rc = lib.gegp_lanczos_step(N, jj, _p(V), ld, _p(w), _p(ab[0]), _p(ab[1]), _stream())
                _check(rc, "gegp_lanczos_step")
            matvecs += jn - j
            j = jn
            h = ab.cpu().numpy()
            a, b = h[0], h[1]
            scale = max(float(np.max(np.abs(a[:j]))), 1e-300)
            m = j
            for q in range(j):            # breakdown: an invariant subspace was found after q + 1 steps
                if not (b[q] > 1e-14 * scale):
                    m = q + 1
                    break
            evals, evecs = eigh_tridiagonal(a[:m], b[:m - 1]) if m > 1 else (a[:1], np.ones((1, 1)))
            lam, s_vec = float(evals[-1]), evecs[:, -1]
            resid = abs(float(b[m - 1]) * float(s_vec[-1])) / max(abs(lam), 1e-300)
            done = resid <= tol or m < j
        coef[:m] = to_dev(np.ascontiguousarray(s_vec))
        rc = lib.gegp_lincomb(N, m, _p(V), ld, _p(coef), _p(r), _stream())
        _check(rc, "gegp_lincomb")
        V[0].copy_(r)
        if done:
            break
    return lam, r[:N].clone(), resid, matvecs


def cond2(Kfull: torch.Tensor, Kinv: torch.Tensor, N: int, *, tol: float = 1e-12, warm=None):
    """kappa_2 = lambda_max(K) * lambda_max(K^-1) with both extreme eigenvectors.  `warm`: a previous result of this
    function for a nearby matrix of the same order (its eigenvectors warm-start the two iterations).
    -> dict(cond, lam_max, lam_min, v_max, v_min, resid_max, resid_min, cycles = matvecs of the two iterations)."""
    w1 = w2 = None
    if warm is not None and warm.get("v_max") is not None and warm["v_max"].numel() == N:
        w1, w2 = warm["v_max"], warm["v_min"]
    lmax, vmax, r1, c1 = extreme_eig(Kfull, N, tol=tol, v0=w1)
    imax, vmin, r2, c2 = extreme_eig(Kinv, N, tol=tol, v0=w2)
    lmin = 1.0 / imax
    return dict(cond=lmax * imax, lam_max=lmax, lam_min=lmin, v_max=vmax, v_min=vmin, resid_max=r1, resid_min=r2,
                cycles=(c1, c2))


def quad_grad(X, theta, v, *, n_g=None, slot=None, eta=0.0, noisy=False, varK=1.0, kernel=None):
    """gegp_quad_grad: v^T (dKcov/dhp) v for every hyper-parameter (base mode) -> device row laid out as GEGP_OUT_*."""
    lib = L.load()
    X, theta, v = to_dev(X), to_dev(theta), to_dev(v)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    out = torch.zeros(L.out_len(d), dtype=F64, device=device())
    vk = to_dev(np.array([float(varK)]))
    nbytes = int(lib.gegp_quad_grad_work_bytes(n, n_g, d))
    ws = torch.empty(nbytes, dtype=torch.uint8, device=device())
    kid, khp = _kern(kernel)
    rc = lib.gegp_quad_grad(n, n_g, d, _p(X), _p(slot), _p(theta), kid, khp, _p(v), L.MODE_BASE, float(eta), int(bool(noisy)),
                            _p(vk), _p(out), _p(ws), nbytes, _stream())
    _check(rc, "gegp_quad_grad")
    return out


def lml_direct_terms(X, theta, *, n_g=None, slot=None, mode=L.MODE_PRECON, eta=0.0, noisy=False, varK=1.0, kernel=None):
    """gegp_lml_direct_terms right after a B = 1 lml_eval(want_grad=True) on this device -> dict of per-hyper-parameter
    rows (GEGP_OUT_* layout, NumPy): aDa = a^T D a, hDa = h^T D a, trKinvD = tr(K^-1 D), and the scalars HKH, Ha."""
    lib = L.load()
    X, theta = to_dev(X), to_dev(theta)
    n, d = X.shape
    n_g = n if n_g is None else n_g
    N = n + n_g * d
    ws = _ws_cache.get(torch.cuda.current_device())
    assert ws is not None, "no likelihood evaluation has run on this device yet"
    ol = L.out_len(d)
    out = torch.zeros(4 * ol + 2, dtype=F64, device=device())
    scratch = torch.empty(4 * ld_of(N), dtype=F64, device=device())
    vk = to_dev(np.array([float(varK)]))
    kid, khp = _kern(kernel)
    rc = lib.gegp_lml_direct_terms(n, n_g, d, _p(X), _p(slot), _p(theta), kid, khp, int(mode), float(eta), int(bool(noisy)),
                                   _p(vk), _p(ws), ws.numel(), _p(scratch), _p(out), _stream())
    _check(rc, "gegp_lml_direct_terms")
    o = out.cpu().numpy()
    rows = o[:4 * ol].reshape(4, ol)
    return dict(aDa=rows[0], hDa=0.25 * (rows[1] - rows[2]), trKinvD=rows[3], HKH=float(o[4 * ol]), Ha=float(o[4 * ol + 1]))


def lml_views(n: int, n_g: int, d: int):
    """Views of candidate 0's arrays inside the shared workspace, valid right after a B = 1 lml_eval(want_grad=True)
    (eager or graph replay) on this device: dict(A [N+2, ld] factor, U [N, ld], Kinv [N, ld], ld)."""
    lib = L.load()
    o = (C.c_int64 * 8)()
    rc = lib.gegp_lml_layout(n, n_g, d, 1, 1, o)
    _check(rc, "gegp_lml_layout")
    header, ld, per, offA, offP, _offD, offU, offK = (int(x) for x in o)
    ws = _ws_cache.get(torch.cuda.current_device())
    assert ws is not None and ws.numel() >= header + 8 * per, "no likelihood evaluation has run on this device yet"
    dbl = ws[header:header + 8 * per].view(F64)
    N = n + n_g * d
    return dict(A=dbl[offA:offA + (N + 2) * ld].view(N + 2, ld), U=dbl[offU:offU + N * ld].view(N, ld),
                Kinv=dbl[offK:offK + N * ld].view(N, ld), p=dbl[offP:offP + N], pinv=dbl[offP + ld:offP + ld + N], ld=ld)


def cond_fro_of_matrix(K: torch.Tensor, N: int) -> float:
    """Frobenius condition number of an explicit SPD device matrix (factor a copy, explicit inverse, two norms)."""
    ld = ld_of(N)
    A = torch.empty((N, ld), dtype=F64, device=K.device)
    A[:, :N] = K[:, :N]
    info, dinv = potrf(A, N, 0)
    if int(info.item()) != 0:
        return 1e18
    _, Kinv = potri(A, dinv, N)
    return cond_fro(K, Kinv, N, False)[0]


def cond2_of_matrix(K: torch.Tensor, N: int, *, tol: float = 1e-12):
    """Condition number of an explicit symmetric positive-definite device matrix K[:N, :N] ([N, ld] buffer):
    factor a copy, explicit inverse, then cond2().  A failed factorisation is retried once on K + delta I and the
    shift is taken out of lambda_min again (K is then numerically singular: kappa ~ 1 / eps)."""
    ld = ld_of(N)
    shift = 0.0
    for attempt in range(2):
        A = torch.empty((N, ld), dtype=F64, device=K.device)
        A[:, :N] = K[:, :N]
        if shift:
            A[:, :N].diagonal().add_(shift)
        info, dinv = potrf(A, N, 0)
        if int(info.item()) == 0:
            break
        if attempt == 1:   # still not positive definite: report "numerically singular"
            return dict(cond=1e18, lam_max=float("nan"), lam_min=0.0, v_max=None, v_min=None, shift=shift)
        shift = 1e-13 * float(torch.sum(torch.abs(K[:, :N]), dim=1).max().item())
    U, Kinv = potri(A, dinv, N)
    res = cond2(K if K.stride(0) % 2 == 0 and K.data_ptr() % 16 == 0 else K.contiguous(), Kinv, N, tol=tol)
    if shift:
        lmin = max(res["lam_min"] - shift, 2.3e-16 * res["lam_max"])
        res["lam_min"], res["cond"] = lmin, res["lam_max"] / lmin
    res["shift"] = shift
    return res
