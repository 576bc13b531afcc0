#!/usr/bin/env python
"""BASELINE.json configs[4]: the 64-setting (epsilon, lambda1, lambda2) sweep on one 10k x 10k day-pair, settings
dealt to the GPUs through a dynamic queue (wot_b200.parallel.parameter_sweep).  Launch with torchrun for N > 1.

  python tools/sweep.py [cells] [kernel] [streams]      kernel: auto (default) | online | stored; streams per GPU (default 2)

Rank 0 prints one JSON line: wall seconds for the whole sweep, settings/s, total Sinkhorn iterations/s."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torch.distributed as dist
    from wot_b200 import parallel, synthetic

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
    kernel = sys.argv[2] if len(sys.argv) > 2 else "auto"
    streams = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("gloo")       # plumbing only: the queue counter and the final gather
    rank = dist.get_rank() if world > 1 else 0
    x0, x1, growth = synthetic.day_pair_coords(n, n, d=30, seed=4)
    grid = parallel.sweep_grid()
    common = dict(epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7, batch_size=5)
    parallel.parameter_sweep(x0, x1, growth, grid[21:22], kernel=kernel, queue_key="warm", **common)   # warm-up
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    res = parallel.parameter_sweep(x0, x1, growth, grid, kernel=kernel, streams=streams, **common)
    wall = time.perf_counter() - t0
    if rank == 0:
        iters = sum(r["iters"] for r in res)
        per_rank = [sum(1 for r in res if r["rank"] == k) for k in range(world)]
        print(json.dumps({
            "config": "validation-sweep-shaped: 64 (eps, lambda1, lambda2) settings x %dx%d pair, kernel=%s" % (n, n, kernel),
            "n_gpus": world, "streams_per_gpu": streams, "wall_s": wall, "settings_per_s": len(grid) / wall, "sinkhorn_iters": iters,
            "sinkhorn_iters_per_s": iters / wall, "settings_per_rank": per_rank,
            "iters_min_max": [min(r["iters"] for r in res), max(r["iters"] for r in res)],
            "not_converged": [r["setting"] for r in res if r["status"] != 0],
            "gpu_ms_sum": sum(r["gpu_ms"] for r in res),
        }), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
