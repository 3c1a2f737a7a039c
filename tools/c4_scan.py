"""BASELINE config 4: d=5, n=200 (N=1200), B candidate thetas, batched LML(+gradient), sharded over the ranks.

    python tools/c4_scan.py [B] [grad]                                   # one GPU
    torchrun --nproc-per-node G --master-addr 127.0.0.1 ... tools/c4_scan.py [B] [grad]

Every rank evaluates its contiguous slice of the candidate table with the batched kernels; one all_gather of the
[B/G, 9+d] result rows (NCCL).  Prints candidates/s (device time, max over ranks) and checks that the gathered table
is identical to a one-rank evaluation of a sample of rows (bit for bit) and has the same argmax."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gpgradpy_b200 import backend as bk, _lib as L, parallel
from oracle import gegp_oracle as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
grad = (sys.argv[2] != "0") if len(sys.argv) > 2 else True
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n, d = 200, 5
x, f, g = O.synthetic_problem(n, d, 0)
y = O.make_data_vec(f, g)
eta = O.nugget(n, d, "precon")[1]
cand = 10.0 ** np.random.default_rng(0).uniform(-5, 1, (B, d))      # log10 theta ~ U[-5, 1]^d
X, Y, C = bk.to_dev(x), bk.to_dev(y), bk.to_dev(cand)
rows = lambda c: bk.lml_eval(X, Y, c, mode=L.MODE_PRECON, eta=eta, want_grad=grad)[0]
def step():
    return parallel.sharded_eval(rows, C)
for _ in range(2):
    table = step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
e0.record()
for _ in range(reps):
    table = step()
e1.record(); torch.cuda.synchronize()
ms = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
t = table.cpu().numpy()
ok = t[:, L.OUT_INFO] == 0
sample = np.arange(0, B, max(1, B // 16))
one = np.vstack([rows(C[i:i + 1]).cpu().numpy() for i in sample])
same = bool(np.array_equal(one, t[sample]))
if rank == 0:
    lml = np.where(ok, t[:, L.OUT_LML], -np.inf)
    print(f"c4 d={d} n={n} N={n*(d+1)} B={B} grad={grad} world={world}: {ms.item():.2f} ms per scan -> "
          f"{B/ms.item()*1e3:.0f} candidates/s ({B*(n*(d+1))**3/ms.item()*1e-9:.2f} TFLOP/s); chol ok {int(ok.sum())}/{B}; "
          f"argmax {int(np.argmax(lml))} lml {lml.max():.9e}; bit-identical to single evaluation: {same}", flush=True)
if world > 1:
    dist.destroy_process_group()
