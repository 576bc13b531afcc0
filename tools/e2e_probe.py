#!/usr/bin/env python
"""Where does the end-to-end (host buffers in, host coupling out) time go with 1 or 2 day-pairs in flight?
Usage: python tools/e2e_probe.py [steps]"""
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import _lib, _pinned, synthetic  # noqa: E402
from wot_b200.ot import optimal_transport as wot_ot  # noqa: E402
from wot_b200.pipeline import Pipeline  # noqa: E402

DEFAULTS = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7, batch_size=5)


def main():
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    pairs = synthetic.atlas_pairs(seed=1)[:steps]
    coords = {p: synthetic.day_pair_coords(p[0], p[1], d=30, seed=p[2]) for p in pairs}
    max_ij = max(p[0] * p[1] for p in pairs)
    pin = {}
    for p in pairs:
        bufs = []
        for arr in coords[p]:
            pa = _pinned.empty(arr.shape, np.float64)
            pa[...] = arr
            bufs.append(pa)
        pin[p] = bufs
    for streams in (1, 2):
        for want in (True, False):
            pipe = Pipeline(0, streams)
            outs = [_pinned.empty((max_ij,), np.float64) for _ in range(streams)]
            lat = [[] for _ in range(streams)]

            def body(k, lo):
                for s in range(lo + k, len(pairs), streams):
                    p = pairs[s]
                    t = time.perf_counter()
                    wot_ot.solve_coords(*pin[p], _lib.SOLVER_DUALITY_GAP, growth_iters=3, kernel="online",
                                        out=outs[k][: p[0] * p[1]].reshape(p[0], p[1]) if want else None,
                                        want_tmap=want, ctx=pipe.contexts[k], **DEFAULTS)
                    lat[k].append(time.perf_counter() - t)

            def block(lo):
                ts = [threading.Thread(target=body, args=(k, lo)) for k in range(streams)]
                [t.start() for t in ts]
                [t.join() for t in ts]
            big = max(range(len(pairs)), key=lambda s: pairs[s][0] * pairs[s][1])
            for k in range(streams):       # warm every context on the largest pair
                p = pairs[big]
                wot_ot.solve_coords(*pin[p], _lib.SOLVER_DUALITY_GAP, growth_iters=1, kernel="online",
                                    out=outs[k][: p[0] * p[1]].reshape(p[0], p[1]), ctx=pipe.contexts[k], **DEFAULTS)
            lat = [[] for _ in range(streams)]
            t0 = time.perf_counter()
            block(0)
            wall = time.perf_counter() - t0
            print("streams %d coupling-to-host %-5s: %.1f ms/step wall; per-worker mean latency %s ms"
                  % (streams, want, 1e3 * wall / len(pairs), ["%.1f" % (1e3 * np.mean(v)) for v in lat]), flush=True)
            pipe.close()
            del outs


if __name__ == "__main__":
    main()
