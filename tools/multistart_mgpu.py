"""Multi-start ('lhs', 5 starts) fit with the start rows sharded over the ranks (torchrun); every rank must end with the
same optimum as a one-rank run.  torchrun --nproc-per-node G tools/multistart_mgpu.py [n] [d]"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from gpgradpy_b200.gp import GaussianProcess
from oracle import gegp_oracle as O
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
d = int(sys.argv[2]) if len(sys.argv) > 2 else 5
x, f, g = O.synthetic_problem(n, d, 0)
GP = GaussianProcess(d, True, "SqExp", "precon")
GP.lkd_optz_start_mtd = "lhs"
GP.init_optz_surr(3)
GP.set_data(x[:1], f[:1], np.zeros(1), g[:1], np.zeros((1, d)))
GP.set_hpara("optz", 0)
GP.set_data(x, f, np.zeros(n), g, np.zeros((n, d)))
torch.cuda.synchronize(); t0 = time.perf_counter()
GP.set_hpara("optz", 1)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
th = torch.as_tensor(GP.hp_vals.theta).cuda()
if world > 1:
    allth = [torch.empty_like(th) for _ in range(world)]
    dist.all_gather(allth, th)
    same = all(torch.equal(allth[0], t) for t in allth)
else:
    same = True
st = GP._lockstep_stats
if rank == 0:
    np.set_printoptions(precision=14, linewidth=200)
    print(GP._multistart_table[:, [0, d, d + 2]], flush=True)
print(f"rank {rank}/{world} on cuda:{torch.cuda.current_device()}: fit {dt:.2f} s, first batch {st['batch_sizes'][0]} starts, "
      f"theta[0] {GP.hp_vals.theta[0]:.12e}, identical on all ranks: {same}", flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
