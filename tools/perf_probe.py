"""Quick device-side timing probe of the main phases (CUDA events). Not the bench; used while tuning."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O

def ev(fn, reps=3):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

cfgs = [(500, 10), (1000, 20)] if len(sys.argv) < 2 else [tuple(map(int, a.split(','))) for a in sys.argv[1:]]
for n, d in cfgs:
    N = n * (d + 1)
    x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
    y = O.make_data_vec(f, g)
    X = bk.to_dev(x); Y = bk.to_dev(y); TH = bk.to_dev(th[None, :])
    ld = bk.ld_of(N)
    buf = torch.empty((N + 2, ld), dtype=torch.float64, device="cuda")
    dinv = bk.dinv_buffer(N)
    ms = ev(lambda: bk.build_cov(X, TH[0], mode=L.MODE_PRECON, eta=eta, out=buf[:N]))
    print(f"n={n} d={d} N={N}: build full {ms:.3f} ms  {8*N*N/ms*1e-6:.0f} GB/s")
    ms = ev(lambda: bk.build_cov(X, TH[0], mode=L.MODE_PRECON, eta=eta, out=buf[:N], uplo=1))
    print(f"   build lower {ms:.3f} ms  {4*N*(N+1)/ms*1e-6:.0f} GB/s (lower bytes)")
    def fac():
        bk.build_cov(X, TH[0], mode=L.MODE_PRECON, eta=eta, out=buf[:N], uplo=1)
        bk.potrf(buf, N, 0, dinv)
    ms_b = ev(lambda: bk.build_cov(X, TH[0], mode=L.MODE_PRECON, eta=eta, out=buf[:N], uplo=1))
    ms = ev(fac) - ms_b
    print(f"   potrf {ms:.3f} ms  {N**3/3/ms*1e-9:.2f} TFLOP/s ({N**3/3/ms*1e-9/37.13*100:.1f}% of DMMA peak)")
    ms0 = ev(lambda: bk.lml_eval(X, Y, TH, mode=L.MODE_PRECON, eta=eta, want_grad=False))
    ms1 = ev(lambda: bk.lml_eval(X, Y, TH, mode=L.MODE_PRECON, eta=eta, want_grad=True))
    print(f"   lml only {ms0:.3f} ms ; lml+grad {ms1:.3f} ms -> {1e3/ms1:.2f} evals/s, {N**3/ms1*1e-9:.2f} TFLOP/s overall")
    # explicit inverse (trtri + lauum on the DMMA GEMM engine)
    U = torch.empty((N, ld), dtype=torch.float64, device="cuda"); Ki = torch.empty((N, ld), dtype=torch.float64, device="cuda")
    fac()
    ms = ev(lambda: bk.potri(buf, dinv, N, U, Ki))
    print(f"   potri {ms:.3f} ms  {2*N**3/3/ms*1e-9:.2f} TFLOP/s ({2*N**3/3/ms*1e-9/37.13*100:.1f}% of DMMA peak)")
    del U, Ki, buf

# raw DMMA GEMM engine
for (M, Nn, K, tb) in [(4096, 4096, 4096, True), (8192, 8192, 8192, True), (8192, 8192, 8192, False), (16384, 16384, 128, True),
                       (16384, 128, 128, False), (2048, 2048, 2048, True)]:
    A = torch.randn((M, K), dtype=torch.float64, device="cuda")
    Bm = torch.randn((Nn, K) if tb else (K, Nn), dtype=torch.float64, device="cuda")
    Cm = torch.zeros((M, Nn), dtype=torch.float64, device="cuda")
    ms = ev(lambda: bk.dgemm(A, Bm, Cm, transb=tb), reps=5)
    print(f"dgemm M={M} N={Nn} K={K} transb={tb}: {ms:.3f} ms  {2*M*Nn*K/ms*1e-9:.2f} TFLOP/s")
    del A, Bm, Cm
