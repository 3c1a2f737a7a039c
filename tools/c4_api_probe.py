"""Where the time of GaussianProcess.calc_lkd_batch goes at config 4 (1024 candidates): wall clock of the call against
the device time of the batched evaluation inside it."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import GaussianProcess, backend as bk, _lib as L
from oracle import gegp_oracle as O
n, d, B = 200, 5, int(sys.argv[1]) if len(sys.argv) > 1 else 1024
x, f, g = O.synthetic_problem(n, d, 0)
G = GaussianProcess(d, True, "SqExp", "precon")
G.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
rows = np.random.default_rng(0).uniform(-5.0, 1.0, (B, d))
orig = bk.lml_eval
stat = {}
def timed(*a, **k):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); r = orig(*a, **k); e1.record()
    t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    stat["host_enqueue_ms"] = (t1 - t0) * 1e3; stat["wall_ms"] = (t2 - t0) * 1e3; stat["device_ms"] = e0.elapsed_time(e1)
    return r
for grad in (False, True):
    G.calc_lkd_batch(rows, calc_grad=grad); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        tab = G.calc_lkd_batch(rows, calc_grad=grad)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    bk.lml_eval = timed
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tab = G.calc_lkd_batch(rows, calc_grad=grad)
    tot = (time.perf_counter() - t0) * 1e3
    bk.lml_eval = orig
    print(f"B={B} grad={grad}: calc_lkd_batch {best:.2f} ms wall (best of 3); instrumented call {tot:.2f} ms: lml_eval host enqueue "
          f"{stat['host_enqueue_ms']:.2f} ms, device {stat['device_ms']:.2f} ms, wall {stat['wall_ms']:.2f} ms; eta {G._etaK:.3e}", flush=True)
