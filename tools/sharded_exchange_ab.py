#!/usr/bin/env python
"""A/B of the two exchanges of the row-sharded solve (BASELINE.json configs[3]) inside ONE process group, alternating
so that clocks, temperature and the box are the same for both: nccl, peer, nccl, peer, ...

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29545 \
      tools/sharded_exchange_ab.py [n=100000] [reps=4] [bind=1]

Rank 0 prints one line per solve (device time, max over ranks) and a summary."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from wot_b200 import parallel, synthetic

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
    bind = (sys.argv[3] if len(sys.argv) > 3 else "1") == "1"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    bound = parallel.bind_host_to_gpu(local) if bind else 0
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    x0, x1, growth = synthetic.day_pair_coords(n, n, d=30, seed=3)
    prm = dict(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7, batch_size=5)
    first = parallel.sharded_online_solve(x0, x1, growth, exchange="nccl", **prm)      # median + warm-up
    med = first["median"]
    parallel.sharded_online_solve(x0, x1, growth, exchange="peer", median=med, **prm)  # mappings, warm-up
    times = {"nccl": [], "peer": []}
    caps = {"nccl": [], "peer": []}
    for rep in range(reps):
        for mode in ("nccl", "peer"):
            tm = {}
            res = parallel.sharded_online_solve(x0, x1, growth, exchange=mode, median=med, timers=tm, **prm)
            t = torch.tensor([res["info"]["gpu_ms"], res["info"]["graph_capture_ms"]], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            times[mode].append(float(t[0].item()))
            caps[mode].append(float(t[1].item()))
            if rank == 0:
                print("rep %d %-4s %.1f ms (of which graph capture on the host %.1f ms)  iters %d batches %s (%s)"
                      % (rep, mode, times[mode][-1], caps[mode][-1], res["info"]["iters"], res["info"]["batches"],
                         tm.get("exchange")), flush=True)
    if rank == 0:
        print(json.dumps({"shape": [n, n], "n_gpus": world, "host_cpus_bound_to_gpu": bound,
                          "nccl_ms": times["nccl"], "peer_ms": times["peer"],
                          "nccl_capture_ms": caps["nccl"], "peer_capture_ms": caps["peer"],
                          "nccl_median_ms": float(np.median(times["nccl"])), "peer_median_ms": float(np.median(times["peer"]))}),
              flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
