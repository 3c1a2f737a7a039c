"""Lock-step multi-start optimisation: batch point B of the reference (optz/OptzLkd.py:249-270).

The reference runs one SciPy SLSQP per start row, one after the other; every iteration of every start needs one
LML + gradient evaluation.  The starts are independent, so here each SLSQP instance runs in its own host thread and
the threads meet at a barrier whenever they need the objective: the requests of all live starts are evaluated as ONE
batched device call (`gegp_lml_eval` with B = number of live starts), which fills the GPU where a single N ~ 5000
evaluation is bound by the latency of its leaf chain.  A batched candidate is evaluated bit-identically to a lone
one (tested), so every start follows exactly the trajectory it would follow in the sequential loop and the selected
optimum is the same.  Start rows can additionally be sharded over the ranks of a process group; the per-start
results are exchanged with one all_gather (parallel.gather_rows).
"""
from __future__ import annotations

import threading

import numpy as np
from scipy.optimize import minimize


class LockstepEvaluator:
    """Barrier-style batcher: `evaluate(i, x)` blocks until every live worker has asked, then one of them runs
    `batch_fn(list of (i, x))` -> dict i -> result for all requests at once."""

    def __init__(self, batch_fn, n_workers: int):
        self._fn = batch_fn
        self._cv = threading.Condition()
        self._active = n_workers
        self._pending: dict = {}
        self._results: dict = {}
        self._error = None
        self.n_batches = 0
        self.n_evals = 0
        self.batch_sizes: list = []

    def _flush_locked(self):
        req = sorted(self._pending.items())
        self._pending = {}
        try:
            out = self._fn(req)
        except BaseException as exc:   # noqa: BLE001 -- hand the failure to every waiting worker
            self._error = exc
            out = {i: None for i, _ in req}
        self.n_batches += 1
        self.n_evals += len(req)
        self.batch_sizes.append(len(req))
        self._results.update(out)
        self._cv.notify_all()

    def evaluate(self, idx: int, x):
        with self._cv:
            self._pending[idx] = np.array(x, dtype=float, copy=True)
            if len(self._pending) >= self._active:
                self._flush_locked()
            while idx not in self._results:
                self._cv.wait()
            res = self._results.pop(idx)
            if self._error is not None:
                raise RuntimeError("batched evaluation failed") from self._error
            return res

    def retire(self, idx: int):
        """Worker `idx` is done: the barrier no longer waits for it."""
        with self._cv:
            self._active -= 1
            if self._pending and len(self._pending) >= self._active:
                self._flush_locked()


def minimize_lockstep(batch_val_and_grad, x0_rows, bounds, options, constraints=(), thread_init=None):
    """SLSQP from every row of x0_rows, objective/gradient requests batched across the rows.

    batch_val_and_grad(X [b, n_hp]) -> (vals [b], grads [b, n_hp]) of the function to MINIMISE.
    thread_init: called first in every worker thread (the CUDA device is per-thread state: a worker that ends up
    running the batched device call must select the caller's device).
    Returns (list of scipy OptimizeResult in row order, LockstepEvaluator with its statistics)."""
    x0_rows = np.atleast_2d(np.asarray(x0_rows, dtype=float))
    n = x0_rows.shape[0]

    def batch_fn(req):
        X = np.vstack([x for _, x in req])
        vals, grads = batch_val_and_grad(X)
        return {i: (float(vals[k]), np.array(grads[k], dtype=float)) for k, (i, _) in enumerate(req)}

    ev = LockstepEvaluator(batch_fn, n)
    results = [None] * n
    errors = [None] * n

    def worker(i):
        last = {"x": None, "val": None, "grad": None}

        def fun(x):
            if last["x"] is None or not np.array_equal(x, last["x"]):
                v, g = ev.evaluate(i, x)
                last.update(x=np.array(x, copy=True), val=v, grad=g)
            return last["val"]

        def jac(x):
            fun(x)
            return last["grad"]

        try:
            if thread_init is not None:
                thread_init()
            results[i] = minimize(fun, x0_rows[i], method="SLSQP", jac=jac, bounds=bounds, constraints=constraints,
                                  options=options)
        except BaseException as exc:   # noqa: BLE001
            errors[i] = exc
        finally:
            ev.retire(i)

    threads = [threading.Thread(target=worker, args=(i,), daemon=True) for i in range(n)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    for e in errors:
        if e is not None:
            raise e
    return results, ev
