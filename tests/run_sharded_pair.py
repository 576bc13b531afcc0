"""Launched by torchrun from test_gpu_multi.py: one row-sharded online solve over all ranks (NCCL), checked
against the float64 oracle on rank 0."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from oracle import wot_oracle as orc
    from wot_b200 import parallel, synthetic
    from tests.helpers import DEFAULTS, max_rel_err

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    n0, n1 = int(sys.argv[1]), int(sys.argv[2])
    eps = float(sys.argv[3]) if len(sys.argv) > 3 else 0.05      # below 0.02: precise operands
    exchange = sys.argv[4] if len(sys.argv) > 4 else "auto"      # nccl | peer | auto
    DEFAULTS = dict(DEFAULTS, epsilon=eps)
    x0, x1, growth = synthetic.day_pair_coords(n0, n1, d=30, seed=123)
    timers = {}
    res = parallel.sharded_online_solve(x0, x1, growth, exchange=exchange, timers=timers, **DEFAULTS)
    rows = parallel.local_coupling_rows(res)
    lo, hi = res["rows"]
    info = orc.SolveInfo()
    want = orc.optimal_transport_duality_gap(C=orc.compute_default_cost_matrix(x0, x1), G=growth, info=info,
                                             gap="marginal", **DEFAULTS)
    err = max_rel_err(rows, want[lo:hi]) if hi > lo else 0.0
    ferr = float(np.max(np.abs(res["f"].cpu().numpy() - info.f))) / eps
    rerr = float(np.max(np.abs(res["rowsum"].cpu().numpy() - want.sum(axis=1)) / want.sum(axis=1)))
    ok = err <= 1e-4 and ferr <= 1e-4 and rerr <= 1e-4 and abs(res["info"]["batches"][5] - info.batches[5]) <= 1
    if exchange in ("peer", "nccl") and int(os.environ.get("WORLD_SIZE", "1")) > 1:
        ok = ok and timers.get("exchange") == exchange
    print("rank %d rows [%d,%d) exchange %s coupling err %.2e f err %.2e rowsum err %.2e batches %s vs %s %s"
          % (rank, lo, hi, timers.get("exchange"), err, ferr, rerr, res["info"]["batches"], info.batches,
             "OK" if ok else "FAIL"), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
