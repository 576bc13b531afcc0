#!/usr/bin/env python
"""Batch counts and potentials of the GPU kernels against the float64 oracle over the sweep corners
(tests/test_gpu_parity.py::SWEEP_CORNERS) plus epsilon = 0.005: stored, online with the default fp16 hi/lo operands
('online_fast'), online with the precise 6-segment operands ('online_precise').
Usage: python tools/precision_probe.py [--shape 420x460]"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

DEFAULTS = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7,
                batch_size=5, growth_iters=1)
SETTINGS = [dict(epsilon=0.01, lambda1=0.1, lambda2=1), dict(epsilon=0.01, lambda1=50, lambda2=100),
            dict(epsilon=0.025, lambda1=10, lambda2=10), dict(epsilon=0.05, lambda1=0.1, lambda2=100),
            dict(epsilon=0.1, lambda1=50, lambda2=1), dict(epsilon=0.1, lambda1=1, lambda2=50),
            dict(epsilon=0.005, lambda1=1, lambda2=50), dict(epsilon=0.005, lambda1=10, lambda2=100),
            dict(epsilon=0.05, lambda1=1, lambda2=50), dict(epsilon=0.01, lambda1=1, lambda2=50)]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="420x460")
    ap.add_argument("--kernels", default="stored,online_fast,online_precise")
    args = ap.parse_args()
    from oracle import wot_oracle as orc
    from wot_b200 import ot, synthetic
    n0, n1 = (int(v) for v in args.shape.split("x"))
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=4)
    cost = orc.compute_default_cost_matrix(x0, x1)
    for st in SETTINGS:
        params = dict(DEFAULTS, **st)
        info = orc.SolveInfo()
        t0 = time.time()
        want = orc.optimal_transport_duality_gap(C=cost, G=growth, info=info, gap="marginal", **params)
        print("eps %-6g l1 %-4g l2 %-4g oracle: iters %6d batches %s (%.1f s)"
              % (st["epsilon"], st["lambda1"], st["lambda2"], info.iters, info.batches, time.time() - t0), flush=True)
        mask = want >= 1e-12 * want.max()
        for kernel in args.kernels.split(","):
            try:
                tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                                      G=growth.copy(), kernel=kernel, **params)
            except Exception as exc:  # noqa: BLE001
                print("    %-15s FAILED: %s" % (kernel, exc), flush=True)
                continue
            got = ot.last_solve_info()
            gi = got["infos"][0]
            err = float(np.max(np.abs(tmap[mask] - want[mask]) / want[mask]))
            eps = st["epsilon"]
            print("    %-15s batches %s  d(final) %+d  coupling %.2e  |df|/eps %.2e |dg|/eps %.2e  %.1f ms"
                  % (kernel, gi["batches"], gi["batches"][5] - info.batches[5], err,
                     np.max(np.abs(got["f"] - info.f)) / eps, np.max(np.abs(got["g"] - info.g)) / eps, gi["gpu_ms"]),
                  flush=True)


if __name__ == "__main__":
    main()
