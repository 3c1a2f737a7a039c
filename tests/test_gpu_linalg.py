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
    info, dinv = bk.potrf(A, N, extra)
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
    info, _ = bk.potrf(A, N, 0)
    assert int(info.item()) == 151


def _factor_on_gpu(K, extra_rows=0):
    import torch
    from gpgradpy_b200 import backend as bk
    N = K.shape[0]
    ld = bk.ld_of(N)
    A = torch.full((N + extra_rows, ld), float("nan"), dtype=torch.float64, device="cuda")
    A[:N, :N] = torch.as_tensor(np.tril(K)).cuda() + torch.triu(
        torch.full((N, N), float("nan"), dtype=torch.float64, device="cuda"), 1)
    info, dinv = bk.potrf(A, N, 0)
    assert int(info.item()) == 0
    return A, dinv


@pytest.mark.parametrize("N,r", [(50, 3), (128, 200), (333, 17), (1100, 260)])
def test_trsm_rows(N, r):
    import torch
    from gpgradpy_b200 import backend as bk
    K = _spd(N, 10 + N)
    Lr = np.linalg.cholesky(K)
    Lt, dinv = _factor_on_gpu(K)
    ld = bk.ld_of(N)
    rng = np.random.default_rng(5)
    Bm = rng.standard_normal((r, N))
    Bt = torch.zeros((r, ld), dtype=torch.float64, device="cuda")
    Bt[:, :N] = torch.as_tensor(Bm).cuda()
    bk.trsm_rows(Lt, dinv, N, Bt)
    Zr = sla.solve_triangular(Lr, Bm.T, lower=True).T
    e = np.abs(Bt[:, :N].cpu().numpy() - Zr).max() / np.abs(Zr).max()
    print(f"N={N} r={r} err {e:.2e}")
    assert e < 1e-10


@pytest.mark.parametrize("N", [5, 31, 32, 33, 64, 65, 96, 97, 127, 128])
def test_leaf_factor_and_inverse_blocks(N):
    """One-CTA leaf: L and the stored inverse-transposed block U = L^-T for every 32-block count and ragged size."""
    K = _spd(N, 77 + N)
    A, dinv = _factor_on_gpu(K)
    Lr = np.linalg.cholesky(K)
    Lg = np.tril(A[:N, :N].cpu().numpy())
    assert np.abs(Lg - Lr).max() / np.abs(Lr).max() < 1e-12
    U = dinv.cpu().numpy().reshape(128, 128)
    Ur = np.linalg.inv(Lr).T
    assert np.abs(U[:N, :N] - Ur).max() / np.abs(Ur).max() < 1e-11
    assert np.all(np.tril(U, -1) == 0.0)


@pytest.mark.parametrize("N", [60, 128, 200, 385, 1000, 2177])
def test_potri_explicit_inverse(N):
    import torch
    from gpgradpy_b200 import backend as bk
    K = _spd(N, 31 + N)
    A, dinv = _factor_on_gpu(K)
    ld = bk.ld_of(N)
    # poison the outputs: nothing outside what potri writes may be read
    U = torch.full((N, ld), float("nan"), dtype=torch.float64, device="cuda")
    Kinv = torch.full((N, ld), float("nan"), dtype=torch.float64, device="cuda")
    bk.potri(A, dinv, N, U, Kinv)
    Kr = np.linalg.inv(K)
    Kg = Kinv[:, :N].cpu().numpy()
    e = np.abs(Kg - Kr).max() / np.abs(Kr).max()
    Ug = np.triu(U[:, :N].cpu().numpy())
    Ur = np.linalg.inv(np.linalg.cholesky(K)).T
    eu = np.abs(Ug - Ur).max() / np.abs(Ur).max()
    print(f"N={N} Kinv err {e:.2e} U err {eu:.2e}")
    assert e < 1e-10 and eu < 1e-10
    assert np.array_equal(Kg, Kg.T)


@pytest.mark.parametrize("M,N,K,transb", [(1, 1, 1, True), (7, 5, 3, False), (130, 70, 33, True), (128, 128, 128, False),
                                          (300, 257, 129, True), (1000, 900, 515, False), (2048, 2048, 512, True)])
def test_dgemm(M, N, K, transb):
    import torch
    from gpgradpy_b200 import backend as bk
    rng = np.random.default_rng(M + N + K)
    pad = lambda c: (c + 1) // 2 * 2  # even leading dimensions  # noqa: E731
    A = torch.zeros((M, pad(K)), dtype=torch.float64, device="cuda")
    A[:, :K] = torch.as_tensor(rng.standard_normal((M, K))).cuda()
    bs = (N, K) if transb else (K, N)
    B = torch.zeros((bs[0], pad(bs[1])), dtype=torch.float64, device="cuda")
    B[:, :bs[1]] = torch.as_tensor(rng.standard_normal(bs)).cuda()
    C0 = rng.standard_normal((M, N))
    C = torch.zeros((M, pad(N)), dtype=torch.float64, device="cuda")
    C[:, :N] = torch.as_tensor(C0).cuda()
    bk.dgemm(A[:, :K], B[:, :bs[1]], C[:, :N], transb=transb, alpha=-1.5, beta=0.5)
    Bn = B[:, :bs[1]].cpu().numpy()
    ref = -1.5 * A[:, :K].cpu().numpy() @ (Bn.T if transb else Bn) + 0.5 * C0
    e = np.abs(C[:, :N].cpu().numpy() - ref).max() / np.abs(ref).max()
    assert e < 1e-13


@pytest.fixture
def force_tma():
    """Route every GEMM with M, N >= 128 through the TMA kernel (normally only >= 1000 tiles)."""
    from gpgradpy_b200 import _lib
    lib = _lib.load()
    old = lib.gegp_set_option(_lib.OPT_TMA_MIN_TILES, 1)
    assert old >= 1
    yield
    lib.gegp_set_option(_lib.OPT_TMA_MIN_TILES, old)


@pytest.mark.parametrize("M,N,K,transb", [(128, 128, 16, True), (130, 257, 33, True), (300, 257, 129, True),
                                          (1000, 900, 515, True), (4096, 4096, 256, True)])
def test_dgemm_tma_kernel(force_tma, M, N, K, transb):
    test_dgemm(M, N, K, transb)


@pytest.mark.parametrize("N,extra", [(300, 0), (640, 5), (1000, 2), (2500, 130)])
def test_potrf_trapezoid_tma(force_tma, N, extra):
    test_potrf_trapezoid(N, extra)


@pytest.mark.parametrize("N", [385, 1000, 2177])
def test_potri_tma(force_tma, N):
    test_potri_explicit_inverse(N)


def test_tma_and_cpasync_kernels_agree():
    """Same product on both engines (different k order inside a k-tile): equal to rounding."""
    import torch
    from gpgradpy_b200 import backend as bk, _lib
    lib = _lib.load()
    A = torch.randn((700, 520), dtype=torch.float64, device="cuda")
    B = torch.randn((400, 520), dtype=torch.float64, device="cuda")
    C1 = torch.zeros((700, 400), dtype=torch.float64, device="cuda")
    C2 = torch.zeros((700, 400), dtype=torch.float64, device="cuda")
    bk.dgemm(A, B, C1, transb=True)
    old = lib.gegp_set_option(_lib.OPT_TMA_MIN_TILES, 1)
    try:
        bk.dgemm(A, B, C2, transb=True)
    finally:
        lib.gegp_set_option(_lib.OPT_TMA_MIN_TILES, old)
    ref = A @ B.T
    assert (C1 - ref).abs().max().item() < 1e-12 and (C2 - ref).abs().max().item() < 1e-12


def test_tma_gemm_beside_foreign_kernels_is_reproducible(force_tma):
    """The TMA kernel must give the same bits whether or not other kernels run beside it.  With two CTAs of it per SM --
    sharing the SM with whatever else was resident -- products came out wrong beside cuBLAS DGEMMs on another stream (a
    few tiles, up to NaN in a factorisation; see csrc/gemm_tma.cu); it now takes an SM to itself."""
    import torch
    from gpgradpy_b200 import backend as bk
    g = torch.Generator(device="cuda").manual_seed(3)
    A = torch.randn((6144, 1024), dtype=torch.float64, device="cuda", generator=g)
    B = torch.randn((4096, 1024), dtype=torch.float64, device="cuda", generator=g)
    C0 = torch.randn((6144, 4096), dtype=torch.float64, device="cuda", generator=g)
    quiet = C0.clone()
    bk.dgemm(A, B, quiet, transb=True, alpha=-1.0, beta=1.0)
    torch.cuda.synchronize()
    ref = C0 - A @ B.T
    assert (quiet - ref).abs().max().item() < 1e-10
    Fa = torch.randn((4096, 4096), dtype=torch.float64, device="cuda", generator=g)
    Fb = torch.randn_like(Fa)
    Fc = torch.empty_like(Fa)
    side = torch.cuda.Stream()
    expect = C0.clone()
    for _step in range(4):
        bk.dgemm(A, B, expect, transb=True, alpha=-1.0, beta=1.0)
    torch.cuda.synchronize()
    bad = 0
    for _rep in range(12):
        C = C0.clone()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _load in range(6):
                torch.matmul(Fa, Fb, out=Fc)          # foreign load on another stream
        for _step in range(4):                        # the product under test, beside it: C0 - 4 A B^T in four steps
            bk.dgemm(A, B, C, transb=True, alpha=-1.0, beta=1.0)
        torch.cuda.synchronize()
        bad += int(not torch.equal(C, expect))
    assert bad == 0, f"{bad} of 12 repetitions beside foreign kernels differ from the quiet result"


@pytest.mark.parametrize("M,N,K", [(128, 128, 1408), (100, 130, 78), (300, 128, 128)])
def test_small_and_regular_tiles_agree_bit_for_bit(M, N, K):
    """32 x 32-tile and 64 x 64-tile cp.async kernels accumulate every output element in the same k order:
    identical results (which is what allows the tile choice to depend on how much work is in flight)."""
    import torch
    from gpgradpy_b200 import backend as bk, _lib
    lib = _lib.load()
    A = torch.randn((M, K), dtype=torch.float64, device="cuda")
    B = torch.randn((N, K), dtype=torch.float64, device="cuda")
    C0 = torch.randn((M, N), dtype=torch.float64, device="cuda")
    C1, C2 = C0.clone(), C0.clone()
    old = lib.gegp_set_option(_lib.OPT_SMALL_TILE_MAX, 1000)
    try:
        bk.dgemm(A, B, C1, transb=True, alpha=-1.0, beta=1.0)
        lib.gegp_set_option(_lib.OPT_SMALL_TILE_MAX, 0)
        bk.dgemm(A, B, C2, transb=True, alpha=-1.0, beta=1.0)
    finally:
        lib.gegp_set_option(_lib.OPT_SMALL_TILE_MAX, old)
    assert torch.equal(C1, C2)
    assert (C1 - (C0 - A @ B.T)).abs().max().item() < 1e-11


def test_potrf_reentrant_two_host_threads():
    """include/gegp.h promises re-entrancy: two host threads factoring different matrices on their own streams of the
    same device at the same time must not share look-ahead state.  Each concurrent result is compared bit for bit with
    the same factorisation run alone (reference call sites: kernel/Kernel.py:251,291, one cho_factor per thread)."""
    import threading
    import torch
    from gpgradpy_b200 import backend as bk
    Ns = (1500, 1100)
    Ks = [_spd(N, 40 + i) for i, N in enumerate(Ns)]

    def stage(i):
        N = Ns[i]
        A = torch.zeros((N + 1, bk.ld_of(N)), dtype=torch.float64, device="cuda")
        A[:N, :N] = torch.as_tensor(np.tril(Ks[i])).cuda()
        A[N, :N] = 1.0
        return A

    serial = []
    for i in range(2):
        A = stage(i)
        info, dinv = bk.potrf(A, Ns[i], 1)
        torch.cuda.synchronize()
        assert int(info.item()) == 0
        serial.append((A.clone(), dinv.clone()))
    for rep in range(3):
        bufs = [stage(i) for i in range(2)]
        streams = [torch.cuda.Stream() for _ in range(2)]
        torch.cuda.synchronize()
        out, errs = [None, None], []
        gate = threading.Barrier(2)

        def work(i):
            try:
                torch.cuda.set_device(0)
                with torch.cuda.stream(streams[i]):
                    gate.wait()
                    for _ in range(4):                      # several back-to-back calls per thread widen the overlap window
                        bufs[i].copy_(stage(i))
                        out[i] = bk.potrf(bufs[i], Ns[i], 1)
                streams[i].synchronize()
            except Exception as exc:                        # noqa: BLE001
                errs.append(exc)

        th = [threading.Thread(target=work, args=(i,)) for i in range(2)]
        for t in th:
            t.start()
        for t in th:
            t.join()
        assert not errs, errs
        torch.cuda.synchronize()
        for i in range(2):
            N = Ns[i]
            assert int(out[i][0].item()) == 0
            assert torch.equal(torch.tril(bufs[i][:N, :N]), torch.tril(serial[i][0][:N, :N]))
            assert torch.equal(bufs[i][N, :N], serial[i][0][N, :N])
            assert torch.equal(out[i][1], serial[i][1])


def test_chain_cluster_size_does_not_change_results():
    """The factorisation's chain step runs on a cluster of 1, 2 or 4 CTAs depending on the batch count; every size does
    the same arithmetic per strip / per element, so the factor must be bit-identical (a candidate evaluated alone, inside
    a batch or on another rank gives the same bits)."""
    import torch
    from gpgradpy_b200 import backend as bk, _lib as L
    N = 1000
    K = _spd(N, 5)
    lib = L.load()
    res = []
    for cs in (1, 2, 4, 0):
        old = lib.gegp_set_option(L.OPT_CHAIN_CLUSTER, cs)
        try:
            A = torch.zeros((N + 2, bk.ld_of(N)), dtype=torch.float64, device="cuda")
            A[:N, :N] = torch.as_tensor(np.tril(K)).cuda()
            A[N:, :N] = torch.as_tensor(np.random.default_rng(1).standard_normal((2, N))).cuda()
            info, dinv = bk.potrf(A, N, 2)
            torch.cuda.synchronize()
            assert int(info.item()) == 0
            res.append((torch.tril(A[:N, :N]).clone(), A[N:, :N].clone(), dinv.clone()))
        finally:
            lib.gegp_set_option(L.OPT_CHAIN_CLUSTER, old)
    for r in res[1:]:
        assert all(torch.equal(a, b) for a, b in zip(r, res[0]))
    Lr = np.linalg.cholesky(K)
    assert np.abs(res[0][0].cpu().numpy() - Lr).max() / np.abs(Lr).max() < 1e-11
