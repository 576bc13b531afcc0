#!/usr/bin/env python
"""BASELINE.json configs[1] through the PUBLIC model API, expression matrices in: OTModel(adata, growth_iters=3)
and one compute_transport_map per consecutive day-pair (local PCA on the GPU -> cost -> 3 solves -> float64 coupling
on the host), day-pairs kept in flight by wot_b200.pipeline exactly as OTModel.compute_all_transport_maps does
(the maps are dropped instead of written: 39 x 1.2 GB).  Usage: python tools/otmodel_atlas.py [n_pairs] [scale]"""
import json
import os
import sys
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import ot, synthetic  # noqa: E402
from wot_b200._anndata import AnnData  # noqa: E402
from wot_b200.pipeline import Pipeline  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
    sizes = [max(2, int(s * scale)) for s in synthetic.atlas_day_sizes(seed=1)[: n_pairs + 1]]
    X, day, growth = synthetic.expression_matrix(sizes, n_genes=1479, seed=1)
    day = day * 0.5
    obs = pd.DataFrame({"day": day, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    out = {"pairs": n_pairs, "cells": int(X.shape[0]), "genes": int(X.shape[1])}
    for streams in (1, 2):
        model = ot.OTModel(adata, growth_iters=3, streams=streams)
        t = model.timepoints
        pairs = [(t[k], t[k + 1]) for k in range(n_pairs)]
        # warm-up: workspaces, and the pinned output pool (cudaHostAlloc is ~0.3 ms per MiB, so the first maps of a
        # run pay for their page-locked blocks; afterwards the pool recycles them)
        big = max(range(n_pairs), key=lambda k: sizes[k] * sizes[k + 1])
        warm = [model.compute_transport_map(*pairs[big]) for _ in range(streams + 2)]
        del warm

        def one(ctx, pair):
            tm = model.compute_transport_map(*pair)
            return float(tm.obs["g3"].values.sum())

        if streams == 1 and os.environ.get("WOT_PROFILE"):
            import cProfile
            import pstats
            prof = cProfile.Profile()
            prof.enable()
            for p in pairs[:4]:
                one(None, p)
            prof.disable()
            pstats.Stats(prof, stream=sys.stderr).sort_stats("cumulative").print_stats(22)
        t0 = time.perf_counter()
        if streams > 1:
            costs = [sizes[k] * sizes[k + 1] for k in range(n_pairs)]
            with Pipeline(streams=streams + 1, compute_slots=streams) as pipe:      # first use: cold contexts
                pipe.map(one, [pairs[big]] * (streams + 1))
            t0 = time.perf_counter()
            with Pipeline(streams=streams + 1, compute_slots=streams) as pipe:      # contexts come back warm
                sums = pipe.map(one, pairs, costs=costs)
        else:
            sums = [one(None, p) for p in pairs]
        wall = time.perf_counter() - t0
        out["streams_%d" % streams] = {"wall_s": wall, "tmaps_per_s": n_pairs / wall, "mass": float(np.mean(sums))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
