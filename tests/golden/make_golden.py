#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

Run from the repo root:  python tests/golden/make_golden.py
Needs /root/reference (read-only).  The GPU box has no reference, so only the vectors travel.

How the reference is driven
  * wot/ot/optimal_transport.py imports only logging + numpy, so it is loaded by file path.
  * The solver's locals (u, v, a, b, epsilon_i, current_iter) are harvested from the unmodified
    function with a frame trace: on the 'return' event for the final state, and on the line
    `_a = a * np.exp(u / epsilon_i)` (one hit per convergence check) for the per-stage batch counts.
  * The OTModel path (PCA -> cost -> solver -> growth columns) is run through the unmodified
    wot.ot.OTModel with a minimal AnnData stand-in and MagicMock for the plotting / IO packages
    that are not installed here (anndata, h5py, POT, matplotlib, statsmodels).
"""
import importlib.util
import os
import sys
import types
from unittest import mock

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from wot_b200 import synthetic  # noqa: E402

DEFAULTS = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, scaling_iter=3000,
                inner_iter_max=50, tolerance=1e-8, max_iter=1e7, batch_size=5, extra_iter=1000,
                growth_iters=1)


def load_ref_solver_module():
    spec = importlib.util.spec_from_file_location("ref_optimal_transport",
                                                  os.path.join(REF, "wot/ot/optimal_transport.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class Harvest:
    """Frame tracer for one reference solver call."""

    def __init__(self, func):
        self.code = func.__code__
        self.check_line = None
        import inspect
        src, first = inspect.getsourcelines(func)
        for k, line in enumerate(src):
            if line.strip().startswith("_a = a * np.exp(u / epsilon_i)"):
                self.check_line = first + k
        self.checks = []       # (stage e, current_iter) at every convergence check
        self.final = None

    def _local(self, frame, event, arg):
        if event == "line" and frame.f_lineno == self.check_line:
            self.checks.append((frame.f_locals["e"], frame.f_locals["current_iter"]))
        elif event == "return":
            loc = frame.f_locals
            self.final = {k: (np.array(loc[k]) if isinstance(loc[k], np.ndarray) else loc[k])
                          for k in ("u", "v", "a", "b", "epsilon_i") if k in loc}
            self.final["iters"] = loc.get("current_iter", loc.get("i"))
            self.final["gap"] = loc.get("duality_gap", np.nan)
        return self._local

    def _global(self, frame, event, arg):
        if event == "call" and frame.f_code is self.code:
            return self._local
        return None

    def __enter__(self):
        sys.settrace(self._global)
        return self

    def __exit__(self, *exc):
        sys.settrace(None)


def run_solver(mod, name, C, G, **over):
    params = dict(DEFAULTS, **over)
    func = getattr(mod, name)
    with Harvest(func) as h:
        tmap = func(C=C, G=G, **params)
    fin = h.final
    eps = fin["epsilon_i"]
    out = {
        "tmap": tmap,
        "f": fin["u"] + eps * np.log(fin["a"]),
        "g": fin["v"] + eps * np.log(fin["b"]),
        "eps_final": eps,
        "gap": fin["gap"],
    }
    if name == "optimal_transport_duality_gap":
        batches = np.zeros(6, dtype=np.int64)
        for e, _ in h.checks:
            batches[e] += 1
        out["batches"] = batches
        out["iters"] = fin["iters"]
    return out


def pair_cost(n0, n1, seed, d=30):
    """Cost the way the reference builds it (scipy cdist sqeuclidean / median) on synthetic coords."""
    from scipy.spatial.distance import cdist
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=d, seed=seed)
    C = cdist(x0, x1, metric="sqeuclidean")
    return C / np.median(C), growth


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def flat(prefix, d):
    return {prefix + "_" + k: np.asarray(v) for k, v in d.items()}


def solver_cases(mod):
    # 1. the reference's own golden case, tests/test_transport.py:20-32
    C3 = np.array([[0, 100, 100], [100, 0, 100], [100, 100, 0]], dtype=np.float64)
    out = {}
    out.update(flat("dg", run_solver(mod, "optimal_transport_duality_gap", C3, np.ones(3), epsilon=0.01)))
    out.update(flat("fx", run_solver(mod, "transport_stablev2", C3, np.ones(3), epsilon=0.01)))
    save("ref_3x3", C=C3, **out)

    # 2. default parameters on synthetic pairs (shape, seed); C is regenerated from the seed by the tests
    for tag, n0, n1, seed in (("small", 60, 75, 11), ("mid", 300, 340, 12)):
        C, G = pair_cost(n0, n1, seed)
        res = run_solver(mod, "optimal_transport_duality_gap", C, G)
        save("dg_" + tag, shape=np.array([n0, n1, seed]), C_checksum=np.array([C.sum(), C[0, 0], C[-1, -1]]),
             G=G, **flat("dg", res))

    # 3. parameter variations, 120 x 150
    C, G = pair_cost(120, 150, 13)
    variations = {
        "eps01": dict(epsilon=0.01),
        "lam10_100": dict(lambda1=10, lambda2=100),
        "loose": dict(epsilon=0.1, lambda1=0.1, lambda2=1),
        "batch7": dict(batch_size=7),
        "tau1_2": dict(tau=1.2),                     # forces tau absorptions (:137-141) + warm-check quirk
        "tau2_eps02": dict(tau=2.0, epsilon=0.02),
        "maxiter37": dict(max_iter=37),              # early return without /J (:143-145)
        "eps0_2": dict(epsilon0=2.0),                # final eps = epsilon0*epsilon quirk
        "tol1e-5": dict(tolerance=1e-5),
    }
    arrays = {"shape": np.array([120, 150, 13]), "G": G, "names": np.array(sorted(variations))}
    for tag in sorted(variations):
        arrays.update(flat(tag, run_solver(mod, "optimal_transport_duality_gap", C, G, **variations[tag])))
    save("dg_variations", **arrays)

    # 4. fixed-iteration solver
    C, G = pair_cost(100, 120, 14)
    arrays = {"shape": np.array([100, 120, 14]), "G": G}
    arrays.update(flat("default", run_solver(mod, "transport_stablev2", C, G)))
    arrays.update(flat("short", run_solver(mod, "transport_stablev2", C, G, scaling_iter=330, extra_iter=40,
                                           inner_iter_max=50)))
    arrays.update(flat("tau1_5", run_solver(mod, "transport_stablev2", C, G, scaling_iter=400, extra_iter=50,
                                            tau=1.5)))
    save("fixed_iters", **arrays)

    # 5. growth loop, growth_iters = 3 (optimal_transport.py:10-33)
    C, G = pair_cost(90, 110, 15)
    params = dict(DEFAULTS, growth_iters=3, C=C, G=G.copy())
    tmap, learned = mod.compute_transport_matrix(solver=mod.optimal_transport_duality_gap, **params)
    save("growth3", shape=np.array([90, 110, 15]), G=G, tmap=tmap, learned=np.array(learned))


class StubAnnData:
    """The few AnnData behaviours ot_model.py touches (X, obs, var, shape, row masks, copy)."""

    def __init__(self, X, obs=None, var=None):
        self.X = X
        self.obs = obs if obs is not None else pd.DataFrame(index=pd.RangeIndex(X.shape[0]).astype(str))
        self.var = var if var is not None else pd.DataFrame(index=pd.RangeIndex(X.shape[1]).astype(str))

    @property
    def shape(self):
        return self.X.shape

    def copy(self):
        return StubAnnData(self.X.copy(), self.obs.copy(), self.var.copy())

    def __getitem__(self, key):
        rows, cols = key if isinstance(key, tuple) else (key, slice(None))
        rows = np.asarray(rows)
        X = self.X[rows]
        obs = self.obs[rows] if rows.dtype == bool else self.obs.iloc[rows]
        if not (isinstance(cols, slice) and cols == slice(None)):
            X = X[:, cols]
        return StubAnnData(X, obs, self.var)


def import_ref_wot():
    stub = types.ModuleType("anndata")
    stub.AnnData = StubAnnData
    sys.modules["anndata"] = stub
    for name in ("h5py", "ot", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "matplotlib.colors",
                 "statsmodels", "statsmodels.stats", "statsmodels.stats.multitest", "loompy"):
        sys.modules.setdefault(name, mock.MagicMock())
    sys.path.insert(0, REF)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        import wot  # noqa: F401
        import wot.ot
    return wot


def model_cases():
    wot = import_ref_wot()
    # 6. default cost on given coordinates + singular values (ot_model.py:242-253)
    x0, x1, _ = synthetic.day_pair_coords(70, 90, d=30, seed=16)
    sv = np.linspace(40.0, 3.0, 30)
    C = wot.ot.OTModel.compute_default_cost_matrix(x0, x1, np.diag(sv))
    C_plain = wot.ot.OTModel.compute_default_cost_matrix(x0[:, :7], x1[:, :7])
    save("cost_default", shape=np.array([70, 90, 16]), sv=sv, C=C, C_plain7=C_plain)

    # 7. whole OTModel path: 3 days, PCA -> cost -> solver -> growth columns, growth_iters=2
    X, day, growth = synthetic.expression_matrix([150, 170, 160], n_genes=200, seed=17)
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = StubAnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    model = wot.ot.OTModel(adata, growth_iters=2)
    tm = model.compute_transport_map(0, 1)
    p0 = X[day == 0]
    p1 = X[day == 1]
    pca0, pca1, pca, mean = wot.ot.compute_pca(p0, p1, 30)
    save("otmodel_path", cells=np.array([150, 170, 160]), n_genes=np.array(200), seed=np.array(17),
         tmap=tm.X, g0=tm.obs["g0"].values, g1=tm.obs["g1"].values, g2=tm.obs["g2"].values,
         obs_index=np.array(tm.obs.index), var_index=np.array(tm.var.index),
         pca0=pca0, pca1=pca1, singular_values=pca.singular_values_)

    # 8. the reference's golden case through OTModel (custom cost bypasses PCA/median), test_transport.py:11-34
    rng = np.random.default_rng(18)
    adata = StubAnnData(rng.random((6, 1000)), pd.DataFrame({"day": [1, 1, 1, 2, 2, 2]}),
                        pd.DataFrame(index=np.arange(1000)))
    cost = np.array([[0, 100, 100], [100, 0, 100], [100, 100, 0]])
    tm = wot.ot.OTModel(adata, epsilon=0.01, lambda1=1, lambda2=50).compute_transport_map(1, 2, cost_matrix=cost)
    save("otmodel_3x3", tmap=tm.X, g0=tm.obs["g0"].values, g1=tm.obs["g1"].values)


def pca_cases():
    """Local PCA through the UNMODIFIED reference (wot.ot.compute_pca, wot/ot/util.py:240-255) at shapes where
    scikit-learn's svd_solver='auto' runs its randomized solver -- the path csrc/pca.cu restates on the GPU.  Both
    orientations of randomized_svd(transpose='auto'): more cells than genes, and fewer."""
    wot = import_ref_wot()
    out = {}
    for tag, cells, genes, k in (("tall", [600, 700], 300, 30), ("wide", [300, 280], 700, 10)):
        X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=3)
        p0, p1, pca, mean = wot.ot.compute_pca(X[day == 0], X[day == 1], k)
        assert pca._fit_svd_solver == "randomized"
        out.update({tag + "_cells": np.array(cells), tag + "_genes": np.array(genes), tag + "_k": np.array(k),
                    tag + "_pca0": p0, tag + "_pca1": p1, tag + "_sv": pca.singular_values_, tag + "_mean": mean})
    save("pca_randomized", seed=np.array(3), **out)


if __name__ == "__main__":
    np.seterr(all="ignore")
    if len(sys.argv) > 1 and sys.argv[1] == "pca":
        pca_cases()
    else:
        solver_cases(load_ref_solver_module())
        model_cases()
        pca_cases()
