"""Phase clocks of potf2_inv_kernel and chain_prep_kernel (needs tools/libgegp_dbg.so built with -DGEGP_LEAF_CLOCKS:
tools/build_dbg.sh).  N = 256: one leaf factor, one chain step, one more leaf factor; the clocks of the LAST launch of
each kernel are read back."""
import os, sys, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gpgradpy_b200 import _lib
_lib.LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "libgegp_dbg.so")
from gpgradpy_b200 import backend as bk
lib = _lib.load()
N = 256
rng = np.random.default_rng(0)
G = rng.standard_normal((N, N + 8)); K = G @ G.T / (N + 8) + 0.5 * np.eye(N)
ld = bk.ld_of(N)
Kt = torch.as_tensor(np.tril(K)).cuda()
A = torch.zeros((N, ld), dtype=torch.float64, device="cuda")
dinv = bk.dinv_buffer(N)
for cs in (4, 2):
    lib.gegp_set_option(_lib.OPT_CHAIN_CLUSTER, cs)
    for it in range(3):
        A[:, :N] = Kt
        torch.cuda.synchronize()
        bk.potrf(A, N, 0, dinv)
        torch.cuda.synchronize()
        out = (ctypes.c_longlong * 32)()
        lib.gegp_debug_leaf_clocks(out)
        c = list(out)
        if it < 2:
            continue
        names = ["load", "panel0", "upd0", "panel1", "upd1", "panel2", "upd2", "panel3", "upd3(none)", "tail"]
        print("potf2 total cycles", c[10] - c[0])
        for i, nm in enumerate(names):
            print(f"   {nm:12s} {c[i + 1] - c[i]:8d}")
        p = c[16:24]
        pn = ["loads+stage_factor", "solve+store", "cluster_sync1", "X exchange", "cluster_sync2", "syrk", "tail"]
        print(f"chain_prep<{cs}> total cycles (rank 0, thread 0)", p[7] - p[0])
        for i, nm in enumerate(pn):
            print(f"   {nm:20s} {p[i + 1] - p[i]:8d}")
Lg = np.tril(A[:, :N].cpu().numpy()); Lr = np.linalg.cholesky(K)
print("max err", np.abs(Lg - Lr).max())
