"""BASELINE configs 3 (posterior at 10k test points) and 5 (N=51000 build + Cholesky, three conditioning modes)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O

def ev(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e9
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    return best

which = sys.argv[1] if len(sys.argv) > 1 else "c3"
if which == "c3":
    n, d, nx = 1000, 20, 10000
    N = n * (d + 1)
    x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
    y = O.make_data_vec(f, g)
    xs = np.random.default_rng(1).uniform(-2, 2, (nx, d))
    out, _ = bk.lml_eval(x, y, th[None, :], mode=L.MODE_PRECON, eta=eta, want_grad=False)
    o = out.cpu().numpy()[0]
    beta, s2 = float(o[L.OUT_BETA]), float(o[L.OUT_SIGMA2])
    t0 = time.perf_counter(); st = bk.predict_setup(x, y, th, beta, mode=L.MODE_PRECON, eta=eta); torch.cuda.synchronize()
    print(f"c3 predict_setup (build+factor+alpha): {(time.perf_counter()-t0)*1e3:.1f} ms  info={int(st.info.item())}")
    XS = bk.to_dev(xs)
    ms = ev(lambda: bk.predict(st, XS, s2, chunk_bytes=4 << 30))
    print(f"c3 predict nx={nx}: {ms:.1f} ms  solve {N*N*nx/ms*1e-9:.2f} TFLOP/s ({N*N*nx/ms*1e-9/37.13*100:.1f}% of DMMA peak), {nx/ms*1e3:.0f} points/s")
    mu, sig, sig2, nneg = bk.predict(st, XS[:64], s2)
    mu_r, sig_r, _, _ = O.eval_model(x[:], f, g, th, s2, beta, xs[:64], "precon", eta) if N <= 6000 else (None, None, None, None)
    print("   n_negative_var", int(nneg.item()), "mu[:3]", mu[:3].cpu().numpy(), "sig[:3]", sig[:3].cpu().numpy())
else:
    n, d = 1000, 50
    N = n * (d + 1)
    x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d)
    X = bk.to_dev(x); TH = bk.to_dev(th)
    ld = bk.ld_of(N)
    buf = torch.empty((N, ld), dtype=torch.float64, device="cuda")
    dinv = bk.dinv_buffer(N)
    for mode_name in ("base", "rescale_origin", "precon"):
        if mode_name == "rescale_origin":   # same GP in rescaled coordinates: x_s = (x - x_last) c, theta_s = theta / c^2
            xr, _, _, c = O.rescale_origin(x, f, g, O.vreq_rescale_origin(n, d))[:4]
            X = bk.to_dev(xr); TH = bk.to_dev(th / c ** 2)
        else:
            X = bk.to_dev(x); TH = bk.to_dev(th)
        eta = O.nugget(n, d, mode_name)[1]
        mode = L.MODE_PRECON if mode_name == "precon" else L.MODE_BASE
        ms_b = ev(lambda: bk.build_cov(X, TH, mode=mode, eta=eta, out=buf, uplo=0), reps=2)
        ms_l = ev(lambda: bk.build_cov(X, TH, mode=mode, eta=eta, out=buf, uplo=1), reps=2)
        def fac():
            bk.build_cov(X, TH, mode=mode, eta=eta, out=buf, uplo=1)
            return bk.potrf(buf, N, 0, dinv)[0]
        info = fac(); torch.cuda.synchronize()
        ms_f = ev(fac, reps=1) - ms_l
        print(f"c5 {mode_name:15s} N={N}: build full {ms_b:.2f} ms {8*N*N/ms_b*1e-6:.0f} GB/s | lower {ms_l:.2f} ms {4*N*(N+1)/ms_l*1e-6:.0f} GB/s | "
              f"potrf {ms_f:.1f} ms {N**3/3/ms_f*1e-9:.2f} TFLOP/s ({N**3/3/ms_f*1e-9/37.13*100:.1f}% of DMMA peak) info={int(info.item())}", flush=True)
