"""Affine rescaling of x, f and grad f used by the `rescale_*` / `dflt_v*` conditioning modes.

Host-side O(n^2 d) work, restated from the reference's base/Rescaling.py (x: :72-125, objective: :134-214):
x_s = (x - x[idx_best]) * c with an isotropic c that sets the min (or max) pairwise distance, and
f_s = (f - f[idx_best]) * s with s = 100 / (max f - min f); gradients scale with s / c.
Constraint rescaling (lincon / nonlincon) is out of scope (it serves an external constrained optimiser).
"""
from __future__ import annotations

import numpy as np
from scipy.spatial.distance import pdist


class Rescaling:
    tol_min_range_obj = 1e-20
    tol_min_dist_x = 1e-14
    rangeobj_max_dflt = 100.0

    def __init__(self, x_init, idx_xbest=None, x_scl_method=None, dist_set=None):
        assert x_init.ndim == 2
        assert x_scl_method in ("set_vmin", "set_vmax", None)
        self.x_init = np.asarray(x_init, dtype=float)
        self.n_eval, self.dim = self.x_init.shape
        self.idx_xbest = self.n_eval - 1 if idx_xbest is None else idx_xbest
        self.x_scl_method = x_scl_method
        self.dist_set = 1.0 if dist_set is None else dist_set
        self._obj_set = False
        self.set_xscale_data()

    # ---- x ----
    def set_xscale_data(self, xvec_scale_in=None):
        self.x_shift = self.x_init[self.idx_xbest, :].copy()
        base = np.ones(self.dim) if xvec_scale_in is None else np.asarray(xvec_scale_in, dtype=float)
        assert np.all(base > 0)
        x_v1 = (self.x_init - self.x_shift[None, :]) * base[None, :]
        if self.n_eval == 1 or self.x_scl_method is None:
            coeff = 1.0
        else:
            dist = pdist(x_v1)
            if self.x_scl_method == "set_vmin":
                coeff = self.dist_set / max(self.tol_min_dist_x, float(dist.min()))
            else:
                coeff = self.dist_set / float(dist.max())
        self.xvec_scale = base * coeff
        self.x_scl = self.x_init_2_scl(self.x_init)
        if self._obj_set:
            self._scale_obj()

    def x_init_2_scl(self, x):
        x = np.asarray(x, dtype=float)
        return (x - self.x_shift) * self.xvec_scale if x.ndim == 1 else (x - self.x_shift[None, :]) * self.xvec_scale[None, :]

    def x_scl_2_init(self, x):
        x = np.asarray(x, dtype=float)
        return x / self.xvec_scale + self.x_shift if x.ndim == 1 else x / self.xvec_scale[None, :] + self.x_shift[None, :]

    def get_scl_x_w_dist(self):
        return self.x_scl, None

    # ---- objective ----
    def set_obj_data(self, obj, std_obj, grad, std_grad):
        self.obj_init, self.std_obj_init, self.grad_init, self.std_grad_init = obj, std_obj, grad, std_grad
        self.obj_shift = float(obj[self.idx_xbest])
        if obj.size == 1:
            self.obj_scale = 1.0
        else:
            self.obj_scale = self.rangeobj_max_dflt / max(self.tol_min_range_obj, float(np.max(obj) - np.min(obj)))
        self._obj_set = True
        self._scale_obj()

    def _scale_obj(self):
        self.obj_scl, self.std_obj_scl, self.grad_scl, self.std_grad_scl = self.obj_init_2_scl(
            self.obj_init, self.std_obj_init, self.grad_init, self.std_grad_init)[:4]

    def obj_init_2_scl(self, mu=None, sig=None, dmudx=None, dsigdx=None, d2mudx2=None, d2sigdx2=None):
        g = self.obj_scale / self.xvec_scale[None, :]
        h = self.obj_scale / self.xvec_scale[None, :] ** 2
        return (None if mu is None else (mu - self.obj_shift) * self.obj_scale,
                None if sig is None else sig * self.obj_scale,
                None if dmudx is None else dmudx * g, None if dsigdx is None else dsigdx * g,
                None if d2mudx2 is None else d2mudx2 * h, None if d2sigdx2 is None else d2sigdx2 * h)

    def obj_scl_2_init(self, mu=None, sig=None, dmudx=None, dsigdx=None, d2mudx2=None, d2sigdx2=None):
        g = self.xvec_scale[None, :] / self.obj_scale
        h = self.xvec_scale[None, :] ** 2 / self.obj_scale
        return (None if mu is None else mu / self.obj_scale + self.obj_shift,
                None if sig is None else sig / self.obj_scale,
                None if dmudx is None else dmudx * g, None if dsigdx is None else dsigdx * g,
                None if d2mudx2 is None else d2mudx2 * h, None if d2sigdx2 is None else d2sigdx2 * h)
