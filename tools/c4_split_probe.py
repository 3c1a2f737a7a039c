"""Experiment: a shard of B candidates (config 4, N = 1200) evaluated as S concurrent sub-batches on S caller streams
(each with its own workspace; the C ABI is re-entrant) instead of one batched call.  LML only."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import backend as bk, _lib as L
from oracle import gegp_oracle as O
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
grad = (sys.argv[2] != "0") if len(sys.argv) > 2 else False
n, d = 200, 5
lib = L.load()
x, f, g = O.synthetic_problem(n, d, 0); y = O.make_data_vec(f, g); eta = O.nugget(n, d, "precon")[1]
cand = 10.0 ** np.random.default_rng(0).uniform(-5, 1, (B, d))
X, Y, C = bk.to_dev(x), bk.to_dev(y), bk.to_dev(cand)
op = L.OP_LML_GRAD if grad else L.OP_LML
ref = None
for S in (1, 2, 4, 8):
    if B % S: continue
    b = B // S
    streams = [torch.cuda.Stream() for _ in range(S)]
    wss = [torch.empty(int(lib.gegp_workspace_bytes(op, n, n, d, b)) + 4 * b + 256, dtype=torch.uint8, device="cuda") for _ in range(S)]
    out = torch.empty((B, L.out_len(d)), dtype=torch.float64, device="cuda")
    def run():
        cur = torch.cuda.current_stream()
        for i, st in enumerate(streams):
            st.wait_stream(cur)
            th = C[i * b:(i + 1) * b]; o = out[i * b:(i + 1) * b]
            rc = lib.gegp_lml_eval(b, th.data_ptr(), 0, 0, 0, n, n, d, X.data_ptr(), 0, Y.data_ptr(), 0, int(L.MODE_PRECON), float(eta), 0, 0.0,
                                   int(grad), o.data_ptr(), 0, wss[i].data_ptr(), wss[i].numel(), st.cuda_stream)
            assert rc == 0, rc
        for st in streams:
            cur.wait_stream(st)
    for _ in range(2): run()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(5):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1))
    t = out.cpu().numpy()
    if ref is None: ref = t
    print(f"B={B} grad={grad} as {S} concurrent sub-batch(es) of {b}: {best:.3f} ms  identical to one batch: {np.array_equal(t, ref)}", flush=True)
