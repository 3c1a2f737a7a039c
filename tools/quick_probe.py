"""Graph-replayed timings of the factorisation, the LML and the LML+gradient evaluation (device-resident) for a list of
n,d pairs: the three numbers an experiment on the schedule has to move.  Environment switches (GEGP_*) apply."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O

def ev(fn, reps=7):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    return min(ts), float(np.median(ts))

def graphed(fn):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        fn()
    return gr

cfgs = [(500, 10)] if len(sys.argv) < 2 else [tuple(map(int, a.split(','))) for a in sys.argv[1:]]
lib = L.load()
tag = " ".join(f"{k}={v}" for k, v in os.environ.items() if k.startswith("GEGP_"))
for n, d in cfgs:
    N = n * (d + 1)
    x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
    y = O.make_data_vec(f, g)
    X = bk.to_dev(x); Y = bk.to_dev(y); TH = bk.to_dev(th); THB = bk.to_dev(th[None, :])
    ld = bk.ld_of(N)
    buf = torch.empty((N + 2, ld), dtype=torch.float64, device="cuda")
    dinv = bk.dinv_buffer(N)
    info = torch.zeros(1, dtype=torch.int32, device="cuda")
    build = lambda: bk.build_cov(X, TH, mode=L.MODE_PRECON, eta=eta, out=buf[:N], uplo=1)
    def fac():
        build()
        rc = lib.gegp_potrf(N, 0, buf.data_ptr(), buf.stride(0), dinv.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream)
        assert rc == 0
    ms_b = ev(graphed(build).replay)[0]
    ms_f = ev(graphed(fac).replay)[0] - ms_b
    res = {}
    for grad in (False, True):
        gr = graphed(lambda: res.__setitem__(grad, bk.lml_eval(X, Y, THB, mode=L.MODE_PRECON, eta=eta, want_grad=grad)[0]))
        res[("ms", grad)] = ev(gr.replay)[0]
    o = res[True][0].cpu().numpy()
    print(f"[{tag}] N={N}: potrf {ms_f:.3f} ms ({N**3/3/ms_f*1e-9:.2f} TF/s)  lml {res[('ms', False)]:.3f} ms  lml+grad {res[('ms', True)]:.3f} ms "
          f"({N**3/res[('ms', True)]*1e-9:.2f} TF/s)  info {int(info.item())}  lml {o[L.OUT_LML]:.12e} |grad| {np.linalg.norm(o[L.OUT_GRAD:L.OUT_GRAD+d]):.9e}", flush=True)
