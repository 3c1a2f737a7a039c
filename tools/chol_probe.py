"""Cholesky-only timing probe: this library (chain cluster sizes 1 / 2 / 4 / auto, look-ahead on / off) beside cuSOLVER
potrf (torch.linalg.cholesky_ex) on the same matrix, plus an optional per-launch timeline (GEGP_TIMELINE=<file>)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O

def ev(fn, reps=5):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

cfgs = [(500, 10), (1000, 20)] if len(sys.argv) < 2 else [tuple(map(int, a.split(','))) for a in sys.argv[1:]]
lib = L.load()
for n, d in cfgs:
    N = n * (d + 1)
    x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
    X = bk.to_dev(x); TH = bk.to_dev(th)
    ld = bk.ld_of(N)
    buf = torch.empty((N + 2, ld), dtype=torch.float64, device="cuda")
    dinv = bk.dinv_buffer(N)
    build = lambda: bk.build_cov(X, TH, mode=L.MODE_PRECON, eta=eta, out=buf[:N], uplo=1)
    ms_b = ev(build)[0]
    def fac():
        build(); bk.potrf(buf, N, 0, dinv)
    # the product runs the factorisation inside a captured CUDA graph (backend.LmlGraph): no host launch overhead
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    def raw():
        rc = lib.gegp_potrf(N, 0, buf.data_ptr(), buf.stride(0), dinv.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    def graphed_ms():
        side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            build(); raw()
        torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            build(); raw()
        gb = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gb):
            build()
        return ev(gr.replay)[0] - ev(gb.replay)[0]
    for win in ([None] if os.environ.get("GEGP_LA_WINDOW") else [None]):
        mg = graphed_ms()
        print(f"N={N} graph replay: potrf {mg:.3f} ms  {N**3/3/mg*1e-9:.2f} TFLOP/s ({N**3/3/mg*1e-9/37.13*100:.1f}% of DMMA peak), info {int(info.item())}", flush=True)
    for cs in (0, 1, 2, 4):
        lib.gegp_set_option(L.OPT_CHAIN_CLUSTER, cs)
        mn, md = ev(fac)
        print(f"N={N} chain cluster {cs}: potrf {mn - ms_b:.3f} ms (median {md - ms_b:.3f})  {N**3/3/(mn-ms_b)*1e-9:.2f} TFLOP/s", flush=True)
    lib.gegp_set_option(L.OPT_CHAIN_CLUSTER, 0)
    lib.gegp_set_option(L.OPT_LOOKAHEAD, 0)
    mn, md = ev(fac, 3)
    print(f"N={N} single stream: potrf {mn - ms_b:.3f} ms", flush=True)
    lib.gegp_set_option(L.OPT_LOOKAHEAD, 1)
    Kf = bk.build_cov(X, TH, mode=L.MODE_PRECON, eta=eta)[0].contiguous()
    mn, md = ev(lambda: torch.linalg.cholesky_ex(Kf), 3)
    print(f"N={N} cuSOLVER potrf (torch.linalg.cholesky_ex): {mn:.3f} ms  {N**3/3/mn*1e-9:.2f} TFLOP/s", flush=True)
    Lc = torch.linalg.cholesky_ex(Kf)[0]
    fac(); torch.cuda.synchronize()
    err = float((torch.tril(buf[:N, :N]) - Lc).abs().max() / Lc.abs().max())
    print(f"N={N} max |L - L_cusolver| / max|L| = {err:.2e}", flush=True)
    del Kf, Lc, buf
