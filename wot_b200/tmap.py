"""Transport maps kept implicit.

A finished solve determines its coupling through O((I + J) d) numbers: the local-PCA coordinates, the cost's median,
the dual potentials and the final epsilon (`tmap_ij = exp((f_i + g_j - C_ij)/eps) * out_scale`).
`ImplicitTransportMap` keeps exactly those and pushes populations forward / pulls them back
(reference: TransportMapModel.push_forward / pull_back, wot/tmap/transport_map_model.py:235-365, whose inner
products are `p @ tmap.X` at :290 and `tmap.X @ p.T` at :356) with one pass of the online kernel per population,
so trajectories over a 20k x 20k day-pair need neither the 3.2 GB dense coupling nor its trip over PCIe.
"""
from __future__ import annotations

import numpy as np

from . import _lib


class ImplicitTransportMap:
    def __init__(self, x0, x1, f, g, median, eps_final, out_scale, scale=None, obs=None, var=None, t0=None, t1=None):
        self.x0 = np.ascontiguousarray(x0, dtype=np.float64)
        self.x1 = np.ascontiguousarray(x1, dtype=np.float64)
        self.scale = None if scale is None else np.ascontiguousarray(scale, dtype=np.float64)
        self.f = np.ascontiguousarray(f, dtype=np.float64)
        self.g = np.ascontiguousarray(g, dtype=np.float64)
        self.median, self.eps_final, self.out_scale = float(median), float(eps_final), float(out_scale)
        self.obs, self.var, self.t0, self.t1 = obs, var, t0, t1

    @property
    def shape(self):
        return (self.x0.shape[0], self.x1.shape[0])

    def _apply(self, p, forward, normalize):
        p = np.asarray(p, dtype=np.float64)
        single = p.ndim == 1
        p = np.ascontiguousarray(np.atleast_2d(p))
        n_in = self.shape[0] if forward else self.shape[1]
        if p.shape[1] != n_in:
            raise ValueError("population has %d entries, the map has %d cells on that side" % (p.shape[1], n_in))
        if np.any(p < 0):
            raise ValueError("populations must be non-negative measures")
        n_out = self.shape[1] if forward else self.shape[0]
        out = np.empty((p.shape[0], n_out))
        ctx = _lib.context()
        _lib.check(ctx.lib.wotb_coupling_apply_host(
            ctx.handle, _lib.ptr(self.x0), self.shape[0], _lib.ptr(self.x1), self.shape[1], self.x0.shape[1],
            _lib.ptr(self.scale), self.median, _lib.ptr(self.f), _lib.ptr(self.g), self.eps_final, self.out_scale,
            1 if forward else 0, _lib.ptr(p), p.shape[0], _lib.ptr(out)))
        if normalize:
            out = (out.T / out.sum(axis=1)).T               # transport_map_model.py:291-292, :357-358
        return out[0] if single else out

    def push_forward(self, p, normalize=False):
        """p [n_pop, I] (or [I]) over the cells at t0 -> p @ tmap, [n_pop, J]."""
        return self._apply(p, True, normalize)

    def pull_back(self, p, normalize=False):
        """p [n_pop, J] (or [J]) over the cells at t1 -> (tmap @ p.T).T, [n_pop, I]."""
        return self._apply(p, False, normalize)

    def row_sums(self):
        return self.pull_back(np.ones(self.shape[1]))

    def col_sums(self):
        return self.push_forward(np.ones(self.shape[0]))
