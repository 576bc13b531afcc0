"""Transport maps kept implicit, and the trajectory machinery on top of them.

A finished solve determines its coupling through O((I + J) d) numbers: the local-PCA coordinates, the cost's median,
the dual potentials and the final epsilon (`tmap_ij = exp((f_i + g_j - C_ij)/eps) * out_scale`).
`ImplicitTransportMap` keeps exactly those and pushes populations forward / pulls them back (reference:
TransportMapModel.push_forward / pull_back, wot/tmap/transport_map_model.py:235-365, whose inner products are
`p @ tmap.X` at :290 and `tmap.X @ p.T` at :356) on the GPU in float64, all populations in one sweep over the
coupling (one exponential per entry, then one FMA per population), so trajectories over a 20k x 20k day-pair need
neither the 3.2 GB dense coupling nor its trip over PCIe.

`ImplicitTransportMapModel` chains such maps over consecutive day-pairs with the reference's semantics:
push_forward / pull_back with `to_time` and per-step normalisation (:235-365), trajectories (:105-143), fates
(:40-69), transition_table (:71-103), and glue (wot/tmap/util.py:74-94) as composition.  The same model can be built
from a directory of written transport-map files (`ImplicitTransportMapModel.from_directory`, `StoredTransportMap`): the
consumer side of the `.h5ad` output layout (:652-732).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import pandas as pd

from . import _lib


class Population:
    """A measure over the cells of one time point (wot/population.py): `p` [n cells of that day]."""

    def __init__(self, time, p, name=None):
        self.time, self.p, self.name = time, np.asarray(p, dtype=np.float64), name

    def normalized(self):
        return Population(self.time, self.p / self.p.sum(), self.name)


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

    def _args(self):
        return (_lib.ptr(self.x0), self.shape[0], _lib.ptr(self.x1), self.shape[1], self.x0.shape[1], _lib.ptr(self.scale),
                self.median, _lib.ptr(self.f), _lib.ptr(self.g), self.eps_final, self.out_scale)

    def _apply(self, p, forward, normalize):
        p = np.asarray(p, dtype=np.float64)
        single = p.ndim == 1
        p = np.ascontiguousarray(np.atleast_2d(p))
        n_in = self.shape[0] if forward else self.shape[1]
        if p.shape[1] != n_in:
            raise ValueError("population has %d entries, the map has %d cells on that side" % (p.shape[1], n_in))
        n_out = self.shape[1] if forward else self.shape[0]
        out = np.empty((p.shape[0], n_out))
        ctx = _lib.context()
        _lib.check(ctx.lib.wotb_coupling_apply_host(ctx.handle, *self._args(), 1 if forward else 0, _lib.ptr(p),
                                                    p.shape[0], _lib.ptr(out)))
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

    def sample_pairs(self, interp_frac, uniforms):
        """The index pairs interpolate_with_ot draws (wot/ot/util.py:140-146) for the given uniform samples:
        p = tmap / colsum^(1 - interp_frac), flattened row-major and normalised; pair k is the first flattened index
        whose cumulative probability exceeds uniforms[k] (np.random.choice = searchsorted on the cumulative sum).
        Returns (rows [n], cols [n]) int64."""
        u = np.ascontiguousarray(uniforms, dtype=np.float64)
        w = np.power(self.col_sums(), -(1.0 - float(interp_frac)))
        w[~np.isfinite(w)] = 0.0                            # empty columns carry no mass
        mass = self.pull_back(w)                            # row masses of p, unnormalised
        cum = np.cumsum(mass)
        total = cum[-1]
        want = u * total
        rows = np.minimum(np.searchsorted(cum, want, side="right"), len(cum) - 1).astype(np.int64)
        targets = np.ascontiguousarray(want - np.where(rows > 0, cum[rows - 1], 0.0))
        cols = np.empty(len(u), dtype=np.int64)
        ctx = _lib.context()
        _lib.check(ctx.lib.wotb_coupling_sample_host(ctx.handle, *self._args(), _lib.ptr(w), _lib.ptr(rows), _lib.ptr(targets),
                                                     len(u), _lib.ptr(cols)))
        return rows, cols

    def to_dense(self, device=None):
        """The I x J coupling as float64 ndarray (what compute_transport_map would have returned)."""
        import torch
        ctx = _lib.context(device)
        dev = torch.device("cuda", ctx.device)
        xs0 = self.x0 if self.scale is None else self.x0 * self.scale
        xs1 = self.x1 if self.scale is None else self.x1 * self.scale
        t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (xs0, xs1, self.f, self.g)]
        out = torch.empty(self.shape, dtype=torch.float64, device=dev)
        P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
        torch.cuda.synchronize()
        _lib.check(ctx.lib.wotb_coupling_online_dev(ctx.handle, P(t[0]), self.shape[0], P(t[1]), self.shape[1],
                                                    self.x0.shape[1], self.median, P(t[2]), P(t[3]), self.eps_final,
                                                    self.out_scale, P(out), self.shape[1], _lib.F64, None))
        return out.cpu().numpy()


class StoredTransportMap:
    """A transport map that lives in a FILE written by compute_all_transport_maps ('{prefix}_{t0}_{t1}.h5ad' / .npz),
    with the interface of ImplicitTransportMap: the consumer side of the output layout
    (TransportMapModel.from_directory reads such files, wot/tmap/transport_map_model.py:652-732, and multiplies
    populations with them at :290 and :356).  The cell ids are read when the object is made (for .h5ad without touching
    /X); the I x J matrix is read and moved to the GPU on first use and a small number of matrices is kept there
    (`StoredTransportMap.resident`, least recently used first out).  Products are float64 library GEMMs on the device;
    there is no CPU path."""

    resident = 2          # matrices kept on the GPU at a time
    _cache = []           # [(map, device tensor)], most recently used last

    def __init__(self, path, t0=None, t1=None):
        from . import io as _io
        self.path, self.t0, self.t1 = str(path), t0, t1
        if self.path.lower().endswith(".h5ad") and not _io.HAVE_ANNDATA:
            from . import h5ad
            d = h5ad.read_h5ad(self.path, with_x=False)
            self.obs = pd.DataFrame({k: v for k, v in d["obs"].items()}, index=pd.Index(d["obs_index"].astype(str)))
            self.var = pd.DataFrame(index=pd.Index(d["var_index"].astype(str)))
        else:
            ds = _io.read_dataset(self.path)
            self.obs, self.var = ds.obs, ds.var
        self._shape = (len(self.obs), len(self.var))

    @property
    def shape(self):
        return self._shape

    def _matrix(self):
        import torch
        if not torch.cuda.is_available():
            raise _lib.WotB200Error("transport-map products need a CUDA device (there is no CPU fallback)")
        cache = StoredTransportMap._cache
        for k, (owner, t) in enumerate(cache):
            if owner is self:
                cache.append(cache.pop(k))
                return t
        from . import io as _io
        X = np.ascontiguousarray(np.asarray(_io.read_dataset(self.path).X), dtype=np.float64)
        if X.shape != self._shape:
            raise ValueError("%s: matrix is %s, ids say %s" % (self.path, X.shape, self._shape))
        while len(cache) >= max(1, StoredTransportMap.resident):
            cache.pop(0)
        t = torch.from_numpy(X).to(torch.device("cuda", _lib.context().device))
        cache.append((self, t))
        return t

    def _apply(self, p, forward, normalize):
        import torch
        p = np.asarray(p, dtype=np.float64)
        single = p.ndim == 1
        p = np.ascontiguousarray(np.atleast_2d(p))
        n_in = self._shape[0] if forward else self._shape[1]
        if p.shape[1] != n_in:
            raise ValueError("population has %d entries, the map has %d cells on that side" % (p.shape[1], n_in))
        X = self._matrix()
        pd_ = torch.from_numpy(p).to(X.device)
        out = (pd_ @ X if forward else (X @ pd_.T).T).cpu().numpy()          # :290 / :356
        if normalize:
            out = (out.T / out.sum(axis=1)).T
        return out[0] if single else out

    def push_forward(self, p, normalize=False):
        return self._apply(p, True, normalize)

    def pull_back(self, p, normalize=False):
        return self._apply(p, False, normalize)

    def row_sums(self):
        return self.pull_back(np.ones(self._shape[1]))

    def col_sums(self):
        return self.push_forward(np.ones(self._shape[0]))

    def to_dense(self, device=None):
        return self._matrix().cpu().numpy()


class GluedTransportMap:
    """tmap_0 @ tmap_1 (glue_transport_maps, wot/tmap/util.py:74-94) for implicit maps: the product is never formed,
    populations go through the factors one after the other.  The cells of the intermediate day must be in the
    same order in both maps (tmap_0.var.index == tmap_1.obs.index; the reference re-indexes, :90-91)."""

    def __init__(self, *maps):
        for a, b in zip(maps[:-1], maps[1:]):
            if a.shape[1] != b.shape[0]:
                raise ValueError("maps do not chain: %s then %s" % (a.shape, b.shape))
            if a.var is not None and b.obs is not None and list(a.var.index) != list(b.obs.index):
                raise ValueError("the cells of the intermediate day differ between the two maps")
        self.maps = list(maps)
        self.obs, self.var = maps[0].obs, maps[-1].var
        self.t0, self.t1 = maps[0].t0, maps[-1].t1

    @property
    def shape(self):
        return (self.maps[0].shape[0], self.maps[-1].shape[1])

    def push_forward(self, p, normalize=False):
        for m in self.maps:
            p = m.push_forward(p)
        return (p.T / p.sum(axis=-1)).T if normalize else p

    def pull_back(self, p, normalize=False):
        for m in reversed(self.maps):
            p = m.pull_back(p)
        return (p.T / p.sum(axis=-1)).T if normalize else p


def glue_transport_maps(tmap_0, tmap_1):
    """wot/tmap/util.py:74-94.  Implicit maps compose lazily; two dense AnnData maps are multiplied on the GPU
    (a plain float64 library GEMM through torch)."""
    if isinstance(tmap_0, (ImplicitTransportMap, GluedTransportMap)) and isinstance(tmap_1, (ImplicitTransportMap, GluedTransportMap)):
        parts = (tmap_0.maps if isinstance(tmap_0, GluedTransportMap) else [tmap_0]) + \
                (tmap_1.maps if isinstance(tmap_1, GluedTransportMap) else [tmap_1])
        return GluedTransportMap(*parts)
    import torch

    from ._anndata import AnnData
    idx = tmap_1.obs.index.get_indexer_for(tmap_0.var.index)
    dev = torch.device("cuda", _lib.context().device)
    a = torch.from_numpy(np.ascontiguousarray(tmap_0.X, dtype=np.float64)).to(dev)
    b = torch.from_numpy(np.ascontiguousarray(np.asarray(tmap_1.X, dtype=np.float64)[idx, :])).to(dev)
    return AnnData((a @ b).cpu().numpy(), tmap_0.obs.copy(), tmap_1.var.copy())


class ImplicitTransportMapModel:
    """The part of wot.tmap.TransportMapModel that consumes couplings, on implicit maps.

    tmaps: {(t0, t1): ImplicitTransportMap} for consecutive time points; meta: DataFrame indexed by cell id with a
    'day' column (what TransportMapModel.from_directory assembles, transport_map_model.py:700-732)."""

    def __init__(self, tmaps, meta=None, timepoints=None):
        self.tmaps = dict(tmaps)
        if timepoints is None:
            timepoints = sorted({t for pair in self.tmaps for t in pair})
        self.timepoints = list(timepoints)
        if meta is None:
            frames = []
            for k, t in enumerate(self.timepoints):
                if k + 1 < len(self.timepoints):
                    ids = self.tmaps[(t, self.timepoints[k + 1])].obs.index
                else:
                    ids = self.tmaps[(self.timepoints[k - 1], t)].var.index
                frames.append(pd.DataFrame(index=ids, data={"day": t}))
            meta = pd.concat(frames)
        self.meta = meta

    @classmethod
    def from_ot_model(cls, ot_model):
        """Solve every consecutive day-pair of an OTModel into implicit maps."""
        t = ot_model.timepoints
        return cls({(t[k], t[k + 1]): ot_model.compute_implicit_transport_map(t[k], t[k + 1]) for k in range(len(t) - 1)},
                   timepoints=t)

    @classmethod
    def from_directory(cls, tmap_out, with_covariates=False):
        """The model of a directory of transport-map files, as TransportMapModel.from_directory builds it
        (transport_map_model.py:652-732): files '{prefix}_{t0}_{t1}.h5ad' (also .npz / .txt here) next to `tmap_out`'s
        prefix, `meta` = every cell id with its day (rows of every map, columns of the last one), maps opened lazily
        (StoredTransportMap).  ValueError when no file matches, like the reference."""
        import os
        import re
        if with_covariates:
            raise ValueError("covariate-split maps ('_cv{a}_cv{b}') are not chained into a model here")
        tmap_dir, prefix = os.path.split(str(tmap_out))
        tmap_dir, prefix = tmap_dir or ".", prefix or "tmaps"
        day = r"([0-9]*\.?[0-9]+)"
        pattern = re.compile(re.escape(prefix) + "_" + day + "_" + day + r"\.(h5ad|npz|txt)$")
        found = {}
        for name in sorted(os.listdir(tmap_dir)):
            m = pattern.match(name)
            path = os.path.join(tmap_dir, name)
            if m is not None and os.path.isfile(path):
                found[(float(m.group(1)), float(m.group(2)))] = path
        if not found:
            raise ValueError("No transport maps found in " + tmap_dir + " with prefix " + prefix)
        keys = sorted(found)
        tmaps = {k: StoredTransportMap(found[k], t0=k[0], t1=k[1]) for k in keys}
        frames = [pd.DataFrame(index=tmaps[k].obs.index, data={"day": k[0]}) for k in keys]
        frames.append(pd.DataFrame(index=tmaps[keys[-1]].var.index, data={"day": keys[-1][1]}))
        return cls(tmaps, meta=pd.concat(frames), timepoints=sorted({t for k in keys for t in k}))

    # ---- populations ----------------------------------------------------------------------------------------
    def population_from_ids(self, *ids, at_time, names=None):
        """Indicator populations of cell-id lists at one time point (transport_map_model.py:404-441)."""
        day_ids = self.meta.index[self.meta["day"] == at_time]
        out = []
        for k, group in enumerate(ids):
            p = np.asarray(day_ids.isin(list(group)), dtype=np.float64)
            if p.sum() > 0:
                out.append(Population(at_time, p, None if names is None else names[k]))
        return out

    @staticmethod
    def _unique_time(populations):
        times = {p.time for p in populations}
        if len(times) > 1:
            raise ValueError("Several populations were given, but they are not from the same day")
        if not times:
            raise ValueError("No cells found at the given day")
        return next(iter(times))

    def can_push_forward(self, *populations):
        return self.timepoints.index(self._unique_time(populations)) < len(self.timepoints) - 1

    def can_pull_back(self, *populations):
        return self.timepoints.index(self._unique_time(populations)) > 0

    def push_forward(self, *populations, to_time=None, normalize=True, as_list=False):
        """transport_map_model.py:235-298."""
        i = self.timepoints.index(self._unique_time(populations))
        j = i + 1 if to_time is None else self.timepoints.index(to_time)
        if j >= len(self.timepoints):
            raise ValueError("No further timepoints. Unable to push forward")
        if i > j:
            raise ValueError("Destination timepoint is before source. Unable to push forward")
        p = np.vstack([pop.p for pop in populations])
        while i < j:
            p = np.atleast_2d(self.tmaps[(self.timepoints[i], self.timepoints[i + 1])].push_forward(p, normalize=normalize))
            i += 1
        result = [Population(self.timepoints[i], p[k], populations[k].name) for k in range(p.shape[0])]
        return result[0] if len(result) == 1 and not as_list else result

    def pull_back(self, *populations, to_time=None, normalize=True, as_list=False):
        """transport_map_model.py:300-365."""
        i = self.timepoints.index(self._unique_time(populations))
        j = i - 1 if to_time is None else self.timepoints.index(to_time)
        if i == 0:
            raise ValueError("No previous timepoints. Unable to pull back")
        if i < j:
            raise ValueError("Destination timepoint is after source. Unable to pull back")
        p = np.vstack([pop.p for pop in populations])
        while i > j:
            p = np.atleast_2d(self.tmaps[(self.timepoints[i - 1], self.timepoints[i])].pull_back(p, normalize=normalize))
            i -= 1
        result = [Population(self.timepoints[i], p[k], populations[k].name) for k in range(p.shape[0])]
        return result[0] if len(result) == 1 and not as_list else result

    # ---- consumers ------------------------------------------------------------------------------------------------
    def trajectories(self, populations):
        """Ancestor / descendant distributions of every population at every time point
        (transport_map_model.py:105-143).  Returns a DataFrame: rows all cells (day order), columns populations."""
        self._unique_time(populations)
        start = [p.normalized() for p in populations]
        blocks = [np.array([p.p for p in start]).T]
        cur = start
        while self.can_pull_back(*cur):
            cur = self.pull_back(*cur, as_list=True)
            blocks.insert(0, np.array([p.p for p in cur]).T)
        cur = start
        while self.can_push_forward(*cur):
            cur = self.push_forward(*cur, as_list=True)
            blocks.append(np.array([p.p for p in cur]).T)
        return pd.DataFrame(np.concatenate(blocks), index=self.meta.index, columns=[p.name for p in populations])

    def fates(self, populations):
        """Probability that a cell of an earlier (or the same) day ends up in each population
        (transport_map_model.py:40-69; the populations are completed by an 'Other' population of the remaining cells)."""
        start_day = self._unique_time(populations)
        pops = [Population(p.time, p.p.copy(), p.name) for p in populations]
        rest = 1.0 - np.clip(np.sum([p.p > 0 for p in pops], axis=0), 0, 1)
        if rest.sum() > 0:
            pops.append(Population(start_day, rest, "Other"))          # Population.copy(add_missing=True)
        blocks = [np.array([p.p for p in pops]).T]
        cur = pops
        while self.can_pull_back(*cur):
            cur = self.pull_back(*cur, as_list=True, normalize=False)
            blocks.insert(0, np.array([p.p for p in cur]).T)
        X = np.concatenate(blocks)
        X = X / X.sum(axis=1, keepdims=True)
        obs = self.meta[self.meta["day"] <= start_day]
        return pd.DataFrame(X, index=obs.index, columns=[p.name for p in pops])

    def transition_table(self, start_populations, end_populations):
        """transport_map_model.py:71-103 (without the 'Other' completion when the populations already cover the day)."""
        start_time = self._unique_time(start_populations)
        cur = list(end_populations)
        while self.can_pull_back(*cur) and self._unique_time(cur) > start_time:
            cur = self.pull_back(*cur, as_list=True, normalize=False)
        end_p = np.vstack([p.p for p in cur])
        start_p = np.vstack([p.p for p in start_populations])
        table = start_p @ end_p.T
        return pd.DataFrame(table / table.sum(), index=[p.name for p in start_populations],
                            columns=[p.name for p in end_populations])
