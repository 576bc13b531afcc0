"""Launched by test_gpu_multi.py: a row-sharded online solve whose ranks are THREADS of this process, all on cuda:0,
exchanging over peer memory (plain device pointers): the complete protocol of wotb_online_attach_peers -- stores from the
passes' finishing code into every rank's buffer, flag barrier, rank-ordered sums -- on a single GPU.  Checked against
the float64 oracle and against the one-rank solve (identical iteration and batch counts)."""
import os
import sys
import threading

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from oracle import wot_oracle as orc
    from wot_b200 import _lib, parallel, synthetic
    from tests.helpers import DEFAULTS, max_rel_err

    n0, n1, world = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
    eps = float(sys.argv[4]) if len(sys.argv) > 4 else 0.05
    fixed = len(sys.argv) > 5 and sys.argv[5] == "stablev2"      # the fixed-iteration solver (optimal_transport.py:167-236)
    params = dict(DEFAULTS, epsilon=eps)
    if fixed:
        params.update(scaling_iter=330, extra_iter=40, inner_iter_max=50, solver=_lib.SOLVER_FIXED_ITERS)
    torch.cuda.set_device(0)
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=321)
    ctx = _lib.Context(0)
    import ctypes as C
    med = C.c_double()
    X0 = torch.from_numpy(x0).cuda()
    X1 = torch.from_numpy(x1).cuda()
    _lib.check(ctx.lib.wotb_cost_median_dev(ctx.handle, C.c_void_p(X0.data_ptr()), n0, C.c_void_p(X1.data_ptr()), n1,
                                            x0.shape[1], None, C.byref(med)))
    torch.cuda.synchronize()
    comms = parallel.ThreadComm.make(world)
    results, errors = [None] * world, [None] * world

    def run(r):
        try:
            results[r] = parallel.sharded_online_solve(x0, x1, growth, device=0, exchange="peer", comm=comms[r],
                                                       use_graph=False, median=med.value, **params)
        except BaseException as e:  # noqa: BLE001
            errors[r] = e
            try:
                comms[r]._sh["bar"].abort()
            except Exception:  # noqa: BLE001
                pass

    threads = [threading.Thread(target=run, args=(r,)) for r in range(world)]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    if any(e is not None for e in errors):
        print("FAIL", [repr(e) for e in errors])
        sys.exit(1)
    one = parallel.sharded_online_solve(x0, x1, growth, device=0, median=med.value, **params)
    info = orc.SolveInfo()
    cost = orc.compute_default_cost_matrix(x0, x1)
    if fixed:
        oparams = {k: v for k, v in params.items() if k != "solver"}
        want = orc.transport_stablev2(C=cost, G=growth, info=info, **oparams)
    else:
        want = orc.optimal_transport_duality_gap(C=cost, G=growth, info=info, gap="marginal", **params)
    ok = True
    covered = 0
    for r, res in enumerate(results):
        rows = parallel.local_coupling_rows(res)
        lo, hi = res["rows"]
        covered += hi - lo
        err = max_rel_err(rows, want[lo:hi]) if hi > lo else 0.0
        ferr = float(np.max(np.abs(res["f"].cpu().numpy() - info.f))) / eps
        rerr = float(np.max(np.abs(res["rowsum"].cpu().numpy() - want.sum(axis=1)) / want.sum(axis=1)))
        same = (res["info"]["batches"] == one["info"]["batches"] and res["info"]["iters"] == one["info"]["iters"])
        d1 = float(np.max(np.abs(res["f"].cpu().numpy() - one["f"].cpu().numpy()))) / eps
        good = (err <= 1e-4 and ferr <= 1e-4 and rerr <= 1e-4 and same and d1 <= 1e-5
                and (fixed or abs(res["info"]["batches"][5] - info.batches[5]) <= 1)
                and (not fixed or res["info"]["iters"] == info.iters))
        ok = ok and good
        print("rank %d/%d rows [%d,%d) coupling err %.2e f err %.2e rowsum err %.2e vs one rank %.1e batches %s vs %s %s"
              % (r, world, lo, hi, err, ferr, rerr, d1, res["info"]["batches"], info.batches, "OK" if good else "FAIL"),
              flush=True)
    # replicated state: every rank must hold the same bits
    for res in results[1:]:
        ok = ok and bool(torch.equal(res["f"], results[0]["f"])) and bool(torch.equal(res["g"], results[0]["g"]))
    ok = ok and covered == n0
    print("ALL OK" if ok else "FAIL")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
