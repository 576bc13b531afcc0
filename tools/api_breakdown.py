#!/usr/bin/env python
"""Where the wall time of OTModel.compute_all_transport_maps goes (6 atlas day-pairs, 1,479 genes): with .h5ad
files, with the writer replaced by a no-op, and the local PCA alone."""
import os
import shutil
import sys
import tempfile
import time

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import io as wio  # noqa: E402
from wot_b200 import ot, synthetic  # noqa: E402
from wot_b200._anndata import AnnData  # noqa: E402
from wot_b200.ot.util import compute_pca  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 6
    sizes = synthetic.atlas_day_sizes(seed=1)[: n_pairs + 1]
    X, day, growth = synthetic.expression_matrix(sizes, n_genes=1479, seed=1)
    obs = pd.DataFrame({"day": day * 0.5, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    model = ot.OTModel(adata, growth_iters=3)
    tmp = tempfile.mkdtemp(prefix="wotb_api_")
    try:
        for label, patch in (("h5ad files", None), ("writer = no-op", lambda ds, path, output_format="txt": None),
                             ("h5ad files again", None)):
            orig = wio.write_dataset
            if patch is not None:
                wio.write_dataset = patch
            try:
                t0 = time.perf_counter()
                model.compute_all_transport_maps(tmap_out=os.path.join(tmp, "tmaps"), output_file_format="h5ad")
                wall = time.perf_counter() - t0
            finally:
                wio.write_dataset = orig
            print("%-18s %.2f s  %.2f tmaps/s" % (label, wall, n_pairs / wall), flush=True)
        off = np.concatenate([[0], np.cumsum(sizes)])
        t0 = time.perf_counter()
        for k in range(n_pairs):
            _, _, pca, _ = compute_pca(X[off[k]:off[k + 1]], X[off[k + 1]:off[k + 2]], 30)
        wall = time.perf_counter() - t0
        print("local PCA alone    %.2f s  (%.0f ms per pair, %.0f ms of it on the device incl. upload)"
              % (wall, 1e3 * wall / n_pairs, pca.gpu_ms), flush=True)
        for threads in (1, 3, 6):
            import wot_b200.h5ad as h5
            big = np.random.default_rng(0).random((12000, 12500))
            t0 = time.perf_counter()
            with h5.AsyncWriter(depth=threads, threads=threads) as w:
                for k in range(6):
                    w.submit(lambda k=k: h5.write_h5ad(os.path.join(tmp, "w%d.h5ad" % k), big, ["a"] * 12000, [], ["b"] * 12500))
            wall = time.perf_counter() - t0
            print("6 x 1.2 GB files, %d writer thread(s): %.2f s = %.2f GB/s" % (threads, wall, 6 * big.nbytes / wall / 1e9), flush=True)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
