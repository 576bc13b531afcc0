#!/usr/bin/env python
"""cProfile of the public API path, serial (streams=1) so that everything runs on the profiled thread:
OTModel(adata, growth_iters=3).compute_all_transport_maps over the first atlas day-pairs with the file writer
replaced by a no-op.  Prints the top entries by cumulative time: where the host spends the time the GPU does not."""
import cProfile
import os
import pstats
import shutil
import sys
import tempfile
import time

import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import io as wio  # noqa: E402
from wot_b200 import ot, synthetic  # noqa: E402
from wot_b200._anndata import AnnData  # noqa: E402


def main():
    n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    streams = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    sizes = synthetic.atlas_day_sizes(seed=1)[: n_pairs + 1]
    X, day, growth = synthetic.expression_matrix(sizes, n_genes=1479, seed=1)
    obs = pd.DataFrame({"day": day * 0.5, "cell_growth_rate": growth}, index=["c%d" % i for i in range(len(day))])
    adata = AnnData(X, obs, pd.DataFrame(index=["g%d" % i for i in range(X.shape[1])]))
    model = ot.OTModel(adata, growth_iters=3, streams=streams)
    tmp = tempfile.mkdtemp(prefix="wotb_api_")
    wio.write_dataset = lambda ds, path, output_format="txt": None
    try:
        for rep in range(2):
            prof = cProfile.Profile()
            t0 = time.perf_counter()
            prof.enable()
            model.compute_all_transport_maps(tmap_out=os.path.join(tmp, "tmaps"), output_file_format="h5ad")
            prof.disable()
            wall = time.perf_counter() - t0
            print("pass %d: %d pairs, streams=%d, no-op writer: %.2f s = %.0f ms per pair" % (rep, n_pairs, streams, wall,
                                                                                       1e3 * wall / n_pairs), flush=True)
        pstats.Stats(prof).sort_stats("cumulative").print_stats(32)
    finally:
        shutil.rmtree(tmp, ignore_errors=True)


if __name__ == "__main__":
    main()
