"""Shared helpers for the parity tests: inputs regenerated from seeds and parity metrics."""
import numpy as np

from wot_b200 import synthetic

DEFAULTS = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, scaling_iter=3000,
                inner_iter_max=50, tolerance=1e-8, max_iter=1e7, batch_size=5, extra_iter=1000,
                growth_iters=1)

# north_star tolerance: max relative error <= 1e-4 on coupling entries (above a floor of
# 1e-12 * max entry; reference entries reach 1e-36), on row/column marginals, growth columns and
# dual potentials; final-stage batch count within +-1.
RTOL = 1e-4
ENTRY_FLOOR = 1e-12


def pair_cost(n0, n1, seed, d=30):
    """Median-normalised squared-Euclidean cost on synthetic coords, float64 (oracle arithmetic)."""
    from oracle import wot_oracle
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=d, seed=seed)
    return wot_oracle.compute_default_cost_matrix(x0, x1), growth


def max_rel_err(got, want, floor=ENTRY_FLOOR):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape
    mask = np.abs(want) >= floor * np.max(np.abs(want))
    return float(np.max(np.abs(got[mask] - want[mask]) / np.abs(want[mask])))


def coupling_report(got, want):
    return {
        "entries": max_rel_err(got, want),
        "rows": max_rel_err(got.sum(axis=1), want.sum(axis=1), floor=0.0),
        "cols": max_rel_err(got.sum(axis=0), want.sum(axis=0), floor=0.0),
    }


def assert_coupling_close(got, want, rtol=RTOL):
    rep = coupling_report(got, want)
    assert rep["entries"] <= rtol and rep["rows"] <= rtol and rep["cols"] <= rtol, rep
    return rep
