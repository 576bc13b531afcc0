"""Deterministic synthetic inputs for the transport-map hot path.

The reference ships no benchmark data for this path, so tests and ``bench.py`` draw day-pairs
from these generators.  Everything is seeded; the same (seed, shape) gives the same arrays on
every box.
"""
from __future__ import annotations

import numpy as np

ATLAS_N_PAIRS = 39          # BASELINE.json configs[1]
ATLAS_CELLS_LO, ATLAS_CELLS_HI = 5000, 20000


def day_pair_coords(n0, n1, d=30, seed=0, n_clusters=8):
    """Local-PCA-like coordinates for one day-pair.

    Returns (x0 [n0,d], x1 [n1,d], growth [n0]) float64.  Cluster centres are
    N(0,1)*linspace(3,.3,d); cells add N(0,1)*linspace(1,.2,d); the later day is shifted by
    0.5*N(0,1) per dimension.  With the default solver parameters this gives a normalised cost
    with min~0.02, median 1, max~4 and ~455-460 Sinkhorn iterations.
    """
    rng = np.random.default_rng(seed)
    spread = np.linspace(3.0, 0.3, d)
    noise = np.linspace(1.0, 0.2, d)
    centres = rng.standard_normal((n_clusters, d)) * spread
    lab0 = rng.integers(0, n_clusters, n0)
    lab1 = rng.integers(0, n_clusters, n1)
    x0 = centres[lab0] + rng.standard_normal((n0, d)) * noise
    shift = 0.5 * rng.standard_normal(d)
    x1 = centres[lab1] + shift + rng.standard_normal((n1, d)) * noise
    growth = np.exp(rng.normal(0.0, 0.3, n0))
    return x0, x1, growth


def atlas_day_sizes(seed=1, n_days=ATLAS_N_PAIRS + 1, lo=ATLAS_CELLS_LO, hi=ATLAS_CELLS_HI):
    """Cells per day for the reprogramming-atlas-shaped config: n_days draws of U[lo, hi]."""
    rng = np.random.default_rng(seed)
    return [int(v) for v in rng.integers(lo, hi + 1, n_days)]


def atlas_pairs(seed=1, scale=1.0):
    """[(n0, n1, pair_seed)] for the 39 consecutive day-pairs of the atlas-shaped config.

    ``scale`` < 1 shrinks every day by that factor (used for bounded CPU samples and tests).
    """
    sizes = atlas_day_sizes(seed)
    return [(max(2, int(sizes[k] * scale)), max(2, int(sizes[k + 1] * scale)), 1000 * seed + k)
            for k in range(len(sizes) - 1)]


def expression_matrix(cells_per_day, n_genes=1000, n_latent=50, seed=0, dtype=np.float64):
    """Non-negative expression-like matrix for the OTModel path (PCA + cost + solver).

    Returns (X [sum(cells), n_genes], day [sum(cells)], growth_rate [sum(cells)]).
    """
    rng = np.random.default_rng(seed)
    n = int(sum(cells_per_day))
    mix = rng.standard_normal((n_latent, n_genes)) / np.sqrt(n_latent)
    drift = rng.standard_normal(n_latent) * 0.3
    blocks, days = [], []
    for t, m in enumerate(cells_per_day):
        lat = rng.standard_normal((m, n_latent)) * np.linspace(2.0, 0.2, n_latent) + t * drift
        blocks.append(np.abs(lat @ mix + 0.1 * rng.standard_normal((m, n_genes))))
        days.append(np.full(m, float(t)))
    X = np.vstack(blocks).astype(dtype)
    growth = np.exp(rng.normal(0.0, 0.2, n))
    return X, np.concatenate(days), growth
