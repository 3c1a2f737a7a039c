"""Wall-clock trace of the factorisation's critical chain (globaltimer stamps written by the chain kernels themselves;
needs tools/libgegp_dbg.so, tools/build_dbg.sh).  Prints per-leaf period, kernel durations and the gaps between them."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgegp_dbg.so")
from gpgradpy_b200 import backend as bk
from oracle import gegp_oracle as O
lib = _lib.load()
n, d = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (500, 10)
N = n * (d + 1)
x, f, g = O.synthetic_problem(n, d, 0); th = O.bench_theta(d); eta = O.nugget(n, d, "precon")[1]
X = bk.to_dev(x); TH = bk.to_dev(th)
ld = bk.ld_of(N)
buf = torch.empty((N + 2, ld), dtype=torch.float64, device="cuda")
dinv = bk.dinv_buffer(N)
out = (ctypes.c_ulonglong * (3 * 1024))()
graphed = os.environ.get("GRAPH", "1") == "1"
info = torch.zeros(1, dtype=torch.int32, device="cuda")
def run():
    rc = lib.gegp_potrf(N, 0, buf.data_ptr(), buf.stride(0), dinv.data_ptr(), info.data_ptr(), torch.cuda.current_stream().cuda_stream)
    assert rc == 0
bk.build_cov(X, TH, mode=_lib.MODE_PRECON, eta=eta, out=buf[:N], uplo=1)
run(); torch.cuda.synchronize()
if graphed:
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        run()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        run()
for it in range(4):
    bk.build_cov(X, TH, mode=_lib.MODE_PRECON, eta=eta, out=buf[:N], uplo=1)
    info.zero_()
    torch.cuda.synchronize()
    lib.gegp_debug_chain_ts(out, 1024)
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    if graphed: gr.replay()
    else: run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    cnt = lib.gegp_debug_chain_ts(out, 1024)
print("graph replay" if graphed else "eager launches", "info", int(info.item()))
r = np.array(list(out)[:3 * cnt], dtype=np.float64).reshape(cnt, 3)
r = r[np.argsort(r[:, 1])]
t0 = r[0, 1]
print(f"N={N}: potrf {ms:.3f} ms, {cnt} chain kernels; chain spans {(r[-1, 2] - t0) * 1e-3:.1f} us")
fac = r[r[:, 0] == 0]; prep = r[r[:, 0] == 1]
print(f"factor: mean {np.mean(fac[:, 2] - fac[:, 1]) * 1e-3:.1f} us (min {np.min(fac[:, 2] - fac[:, 1]) * 1e-3:.1f}, max {np.max(fac[:, 2] - fac[:, 1]) * 1e-3:.1f})")
print(f"chain step: mean {np.mean(prep[:, 2] - prep[:, 1]) * 1e-3:.1f} us (min {np.min(prep[:, 2] - prep[:, 1]) * 1e-3:.1f}, max {np.max(prep[:, 2] - prep[:, 1]) * 1e-3:.1f})")
gaps = r[1:, 1] - r[:-1, 2]
g_fp = gaps[(r[:-1, 0] == 0)]; g_pf = gaps[(r[:-1, 0] == 1)]
print(f"gap factor->step: median {np.median(g_fp) * 1e-3:.1f} us, mean {np.mean(g_fp) * 1e-3:.1f}; gap step->factor: median {np.median(g_pf) * 1e-3:.1f}, mean {np.mean(g_pf) * 1e-3:.1f}")
per = np.diff(fac[:, 1]) * 1e-3
print("per-leaf period (us):", " ".join(f"{p:.0f}" for p in per))
print("gaps before each chain step (us):", " ".join(f"{p * 1e-3:.0f}" for p in g_fp))
print("gaps before each factor (us):", " ".join(f"{p * 1e-3:.0f}" for p in g_pf))
