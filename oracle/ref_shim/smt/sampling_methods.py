"""`smt.sampling_methods.LHS` look-alike backed by scipy.stats.qmc (test infrastructure)."""
import numpy as np
from scipy.stats import qmc


class LHS:
    def __init__(self, xlimits, random_state=None, criterion=None):
        self.xlimits = np.atleast_2d(np.asarray(xlimits, dtype=float))
        self.random_state = random_state

    def __call__(self, n):
        d = self.xlimits.shape[0]
        unit = qmc.LatinHypercube(d=d, seed=self.random_state).random(n)
        lo, hi = self.xlimits[:, 0], self.xlimits[:, 1]
        return lo[None, :] + unit * (hi - lo)[None, :]
