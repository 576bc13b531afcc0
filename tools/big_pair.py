#!/usr/bin/env python
"""BASELINE.json configs[2]: one large day-pair, stored-K vs online-K on one GPU.
Usage: python tools/big_pair.py 50000 50000 [max_iter]   (coordinates: synthetic seed 2, d=30)
Prints Sinkhorn iterations/s for both kernels and the agreement of their potentials / row sums."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from wot_b200 import _lib, synthetic  # noqa: E402


def main():
    n0, n1 = int(sys.argv[1]), int(sys.argv[2])
    max_iter = float(sys.argv[3]) if len(sys.argv) > 3 else 1e7
    kernels = sys.argv[4].split(",") if len(sys.argv) > 4 else ["stored", "online"]
    d = 30
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=d, seed=2)
    dev = torch.device("cuda:0")
    stream = torch.cuda.Stream()
    ctx = _lib.Context(0, stream.cuda_stream)
    lib, h = ctx.lib, ctx.handle
    X0, X1 = torch.from_numpy(x0).to(dev), torch.from_numpy(x1).to(dev)
    G = torch.from_numpy(growth).to(dev)
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    med = C.c_double()
    t0 = time.perf_counter()
    _lib.check(lib.wotb_cost_median_dev(h, P(X0), n0, P(X1), n1, d, None, C.byref(med)))
    t_med = time.perf_counter() - t0
    out = {"shape": [n0, n1], "median": med.value, "median_s": t_med}
    res = {}
    for kernel in kernels:
        prm = _lib.make_params(solver=_lib.SOLVER_DUALITY_GAP, max_iter=max_iter,
                               kernel=_lib.KERNEL_STORED if kernel == "stored" else _lib.KERNEL_ONLINE)
        f = torch.empty(n0, dtype=torch.float64, device=dev)
        g = torch.empty(n1, dtype=torch.float64, device=dev)
        rows = torch.empty(n0, dtype=torch.float64, device=dev)
        info = _lib.Info()
        if kernel == "stored":
            ld = (n1 + 31) // 32 * 32
            Cm = torch.empty(n0 * ld, dtype=torch.float32, device=dev)
            _lib.check(lib.wotb_cost_matrix_dev(h, P(X0), n0, P(X1), n1, d, None, med.value, P(Cm), ld, _lib.F32))
            _lib.check(lib.wotb_sinkhorn_stored_dev(h, P(Cm), ld, n0, n1, P(G), C.byref(prm), P(f), P(g), P(rows),
                                                    C.byref(info)))
            del Cm
        else:
            _lib.check(lib.wotb_sinkhorn_online_dev(h, P(X0), n0, P(X1), n1, d, med.value, P(G), C.byref(prm), P(f),
                                                    P(g), P(rows), C.byref(info)))
        i = info.as_dict()
        res[kernel] = (f.cpu().numpy(), g.cpu().numpy(), rows.cpu().numpy())
        out[kernel] = {"iters": i["iters"], "batches": i["batches"], "gpu_ms": i["gpu_ms"],
                       "iters_per_s": i["iters"] / (i["gpu_ms"] * 1e-3), "status": i["status"], "gap": i["gap"],
                       "workspace_gb": lib.wotb_workspace_bytes(h) / 1e9}
        lib.wotb_release_workspace(h)
        torch.cuda.empty_cache()
    if len(res) == 2:
        a, b = res["stored"], res["online"]
        out["online_vs_stored"] = {"max_abs_df_over_eps": float(np.max(np.abs(a[0] - b[0])) / 0.05),
                                   "max_abs_dg_over_eps": float(np.max(np.abs(a[1] - b[1])) / 0.05),
                                   "max_rel_rowsum": float(np.max(np.abs(a[2] - b[2]) / np.abs(a[2])))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
