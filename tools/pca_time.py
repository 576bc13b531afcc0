#!/usr/bin/env python
"""Time the local PCA on the GPU next to scikit-learn on the host, atlas-pair shapes."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import synthetic  # noqa: E402
from wot_b200.ot import util  # noqa: E402


def main():
    shapes = [([5000, 7000], 1479), ([12486, 12405], 1479), ([20000, 20000], 1479)]
    for cells, genes in shapes:
        X, day, _ = synthetic.expression_matrix(cells, n_genes=genes, seed=3)
        m1, m2 = X[day == 0], X[day == 1]
        util.compute_pca_gpu(m1[:600], m2[:600], 30)
        t = time.perf_counter()
        q1, q2, gp, _ = util.compute_pca_gpu(m1, m2, 30)
        t_gpu = time.perf_counter() - t
        t = time.perf_counter()
        p1, p2, sp, _ = util.compute_pca_sklearn(m1, m2, 30)
        t_cpu = time.perf_counter() - t
        err = np.max(np.abs(np.vstack([q1, q2]) - np.vstack([p1, p2])))
        print("cells %s genes %d: GPU %.1f ms wall (%.1f ms device incl. H2D of %.0f MB), sklearn %.0f ms on %d cores; "
              "max |component diff| %.1e, sv rel %.1e"
              % (cells, genes, 1e3 * t_gpu, gp.gpu_ms, X.nbytes / 1e6, 1e3 * t_cpu, os.cpu_count(), err,
                 np.max(np.abs(gp.singular_values_ - sp.singular_values_) / sp.singular_values_)), flush=True)


if __name__ == "__main__":
    main()
