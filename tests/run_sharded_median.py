"""Launched by torchrun from test_gpu_multi.py: the one-pass median split over row shards (NCCL all-reduce of the
counts, all-gather of the window keys) must equal the single-GPU exact median bit for bit on every rank."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    from wot_b200 import _lib, parallel, synthetic

    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    ok = True
    for n0, n1, ties in ((8000, 7601, False), (7999, 7601, False), (8300, 7700, True)):
        x0, x1, _ = synthetic.day_pair_coords(n0, n1, d=30, seed=n0)
        if ties:
            x0[:] = x0[np.random.default_rng(1).integers(0, 3, n0)]      # ties overflow the window: fallback path
        dev = torch.device("cuda", local)
        stream = torch.cuda.Stream(dev)
        ctx = _lib.Context(local, stream.cuda_stream)
        with torch.cuda.stream(stream):
            X0, X1 = torch.from_numpy(x0).to(dev), torch.from_numpy(x1).to(dev)
            got = parallel.sharded_median(ctx, X0, X1, rank, world)
            med = C.c_double()
            os.environ["WOTB_NO_MEDIAN_WINDOW"] = "1"          # the three-pass radix select as the independent answer
            _lib.check(ctx.lib.wotb_cost_median_dev(ctx.handle, C.c_void_p(X0.data_ptr()), n0, C.c_void_p(X1.data_ptr()), n1,
                                                    30, None, C.byref(med)))
            os.environ.pop("WOTB_NO_MEDIAN_WINDOW")
        same = got == med.value
        ok = ok and same
        print("rank %d median %d x %d ties=%s sharded %.17g full %.17g %s" % (rank, n0, n1, ties, got, med.value,
                                                                             "OK" if same else "FAIL"), flush=True)
    if dist.is_initialized():
        dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
