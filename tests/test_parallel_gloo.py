"""CPU, world_size 2, gloo: candidate rows are sharded across ranks, evaluated independently and gathered
with one collective; every rank ends with the full table and the same argmax as the sequential loop.
The per-row evaluator is the CPU oracle here (the device evaluator needs a GPU); the sharding, padding and
gather logic under test is exactly what GaussianProcess.calc_lkd_batch runs over NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from gpgradpy_b200 import parallel
    from oracle import gegp_oracle as O
    x, f, g = O.synthetic_problem(12, 2, 0)
    eta = O.nugget(12, 2, "precon")[1]
    rng = np.random.default_rng(0)
    cand = torch.as_tensor(10.0 ** rng.uniform(-3, 1, (B, 2)))
    seen = []

    def eval_rows(rows):
        out = []
        for th in rows.numpy():
            seen.append(th.copy())
            o = O.lkd_wo_noise_lean(x, f, g, th, "precon", eta, calc_grad=True)
            out.append(np.hstack(([o.ln_lkd, o.hp_varK], o.ln_lkd_grad)))
        return torch.as_tensor(np.array(out))

    table = parallel.sharded_eval(eval_rows, cand, width=4)
    lo, hi = parallel.shard_bounds(B, rank, world)
    q.put((rank, table.numpy(), len(seen), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [7, 8, 1])   # B = 1: more ranks than rows (the surplus rank still enters the collective)
def test_sharded_candidate_scan_world2(B):
    from oracle import gegp_oracle as O
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort(key=lambda t: t[0])
    x, f, g = O.synthetic_problem(12, 2, 0)
    eta = O.nugget(12, 2, "precon")[1]
    cand = 10.0 ** np.random.default_rng(0).uniform(-3, 1, (B, 2))
    seq = np.array([np.hstack(([o.ln_lkd, o.hp_varK], o.ln_lkd_grad)) for o in
                    (O.lkd_wo_noise_lean(x, f, g, th, "precon", eta) for th in cand)])
    assert res[0][2] + res[1][2] == B                       # every row evaluated exactly once
    assert abs(res[0][2] - res[1][2]) <= 1
    assert all(t[1].shape[0] == B for t in res)
    for _, table, _, _ in res:
        assert table.shape == seq.shape
        assert np.array_equal(table, seq)                   # same arithmetic, same order: bit-identical
        assert int(np.argmax(table[:, 0])) == int(np.argmax(seq[:, 0]))
