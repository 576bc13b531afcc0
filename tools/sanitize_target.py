#!/usr/bin/env python
"""Small workload for compute-sanitizer (memcheck / racecheck / synccheck): one default solve per kernel family on a
pair small enough to finish under the tool's slowdown, plus the multi-population apply and the sampling kernel.
  compute-sanitizer --tool racecheck python tools/sanitize_target.py"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import ot, synthetic  # noqa: E402

PARAMS = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=60, batch_size=5,
              growth_iters=1)


def main():
    shape = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (600, 700)
    x0, x1, growth = synthetic.day_pair_coords(*shape, d=30, seed=3)
    kernels = () if "wide-only" in sys.argv else ("stored", "online_fast", "online_precise", "online_simt")
    for kernel in kernels:
        tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                              G=growth.copy(), kernel=kernel, **PARAMS)
        info = ot.last_solve_info()["infos"][0]
        print("%-15s iters %d status %d mass %.6e" % (kernel, info["iters"], info["status"], float(tmap.sum())), flush=True)
    # rows wider than one CTA: the cluster-fused stored kernel (k_fused_cl: DSMEM stores + remote mbarrier arrives)
    for wide in ((48, 24001), (40, 47000)):
        x0, x1, growth = synthetic.day_pair_coords(*wide, d=30, seed=4)
        tmap, _ = ot.compute_transport_matrix(ot.optimal_transport_duality_gap, coords=(x0, x1, None), C=None,
                                              G=growth.copy(), kernel="stored", **dict(PARAMS, max_iter=30))
        info = ot.last_solve_info()["infos"][0]
        print("stored %dx%d iters %d status %d mass %.6e" % (wide[0], wide[1], info["iters"], info["status"],
                                                             float(tmap.sum())), flush=True)


if __name__ == "__main__":
    main()
