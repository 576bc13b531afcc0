#!/usr/bin/env python
"""BASELINE.json configs[3]: ONE large day-pair, online kernel, rows sharded over the ranks (NCCL all-reduce of
the column partial sums per iteration).  Launch with torchrun (or plain python for one GPU):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29544 \
      tools/sharded_pair.py 100000 100000 [max_iter]

Rank 0 prints one JSON line: Sinkhorn iterations/s, the fraction of the aggregate MUFU.EX2 peak, batch counts,
and a float64 check of the unbalanced-Sinkhorn fixed point on a sample of this rank's rows."""
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from wot_b200 import _lib, parallel, synthetic

    n0, n1 = int(sys.argv[1]), int(sys.argv[2])
    max_iter = float(sys.argv[3]) if len(sys.argv) > 3 else 1e7
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if world > 1 else 0
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=3)
    prm = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=max_iter,
               batch_size=5)
    timers = {}
    res = parallel.sharded_online_solve(x0, x1, growth, timers=timers, **prm)       # includes the exact median
    res2 = parallel.sharded_online_solve(x0, x1, growth, median=res["median"], **prm)  # solver alone, warm
    info = res2["info"]
    ms = torch.tensor([info["gpu_ms"]], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms.item())
    # fixed point on a sample of this rank's rows, float64, independent of the library's kernels
    lo, hi = res2["rows"]
    f, g = res2["f"], res2["g"]
    eps, l1 = 0.05, 1.0
    rows = torch.arange(lo, hi, max(1, (hi - lo) // 512), device=f.device)
    X0, X1 = res2["coords"]
    dist2 = ((X0[rows] ** 2).sum(1)[:, None] + (X1 ** 2).sum(1)[None, :] - 2.0 * X0[rows] @ X1.T).clamp_(min=0)
    r = torch.exp((f[rows, None] + g[None, :] - dist2 / res["median"]) / eps).sum(1) / n1
    want = torch.from_numpy(growth).to(f.device)[rows] * torch.exp(-f[rows] / l1)
    fp_err = float(((r - want).abs() / want).max())
    rs_err = float(((res2["rowsum"][rows] - r).abs() / r).max())
    peak = C.c_double()
    _lib.check(res2["ctx"].lib.wotb_bench_mufu_dev(res2["ctx"].handle, C.byref(peak)))
    if rank == 0:
        its = info["iters"]
        print(json.dumps({
            "config": "single %dx%d pair, online-K, rows sharded over %d GPU(s), NCCL all-reduce of column sums" % (n0, n1, world),
            "n_gpus": world, "iters": its, "batches": info["batches"], "status": info["status"], "gap": info["gap"],
            "solve_ms": ms, "sinkhorn_iters_per_s": its / (ms * 1e-3),
            "mufu_frac_aggregate": 2.0 * n0 * n1 * its / (ms * 1e-3) / (world * peak.value),
            "first_call_ms_incl_median": res["info"]["gpu_ms"], "median": res["median"],
            "fixed_point_max_rel_err_sampled_rows": fp_err, "rowsum_vs_float64_max_rel_err": rs_err,
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
