#!/usr/bin/env python
"""Device time of whole online solves (coordinates resident in HBM, wotb_sinkhorn_online_dev, one stream):
iterations, ms, us per Sinkhorn iteration and the all-in fraction of the MUFU peak.
Usage: python tools/solve_time.py [12486x12405 ...] [--eps 0.05] [--reps 3]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import _lib, synthetic  # noqa: E402

DEFAULTS = dict(lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7, batch_size=5)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("shapes", nargs="*", default=["12486x12405"])
    ap.add_argument("--eps", type=float, default=0.05)
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    import torch
    ctx = _lib.context(0)
    lib, h = ctx.lib, ctx.handle
    peak = C.c_double()
    _lib.check(lib.wotb_bench_mufu_dev(h, C.byref(peak)))
    dev = "cuda:0"
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    for shape in args.shapes:
        n0, n1 = (int(v) for v in shape.split("x"))
        x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=6)
        X0, X1, G = (torch.from_numpy(a).to(dev) for a in (x0, x1, growth))
        med = C.c_double()
        _lib.check(lib.wotb_cost_median_dev(h, P(X0), n0, P(X1), n1, 30, None, C.byref(med)))
        prm = _lib.make_params(solver=_lib.SOLVER_DUALITY_GAP, kernel=_lib.KERNEL_ONLINE, epsilon=args.eps, **DEFAULTS)
        f = torch.empty(n0, dtype=torch.float64, device=dev)
        g = torch.empty(n1, dtype=torch.float64, device=dev)
        rows = torch.empty(n0, dtype=torch.float64, device=dev)
        best = None
        for _ in range(args.reps + 1):
            info = _lib.Info()
            _lib.check(lib.wotb_sinkhorn_online_dev(h, P(X0), n0, P(X1), n1, 30, med.value, P(G), C.byref(prm), P(f), P(g),
                                                    P(rows), C.byref(info)))
            i = info.as_dict()
            if best is None or i["gpu_ms"] < best["gpu_ms"]:
                best = i
        us = best["gpu_ms"] * 1e3 / best["iters"]
        frac = 2.0 * n0 * n1 * best["iters"] / (best["gpu_ms"] * 1e-3) / peak.value
        print("solve %6d x %6d eps=%g  iters %5d batches %s  %.2f ms  %.1f us/iteration  %.3f of MUFU peak all-in  "
              "checksum %.10e" % (n0, n1, args.eps, best["iters"], best["batches"], best["gpu_ms"], us, frac,
                                  float(rows.sum().item())), flush=True)


if __name__ == "__main__":
    main()
