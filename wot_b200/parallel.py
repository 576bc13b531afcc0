"""Multi-GPU execution of the transport-map path: one process per GPU, torch.distributed for plumbing.

Two partitionings (BASELINE.json north_star):

1. Independent day-pairs (and parameter settings) share nothing: `shard_units` deals them to ranks,
   longest first (`parameter_sweep` uses a dynamic `WorkQueue` instead, iteration counts being unknown);
   `compute_all_transport_maps` runs a rank's pairs and gathers the learned-growth tables so the outputs (file names, `{prefix}_g.txt` row order, --no_overwrite skipping) equal the serial loop of
   the reference (ot_model.py:182-201).  No collective sits on the data path.

2. One very large pair is row-sharded: `sharded_online_solve` drives the stepping entry points of the
   library (wotb_online_*): each rank computes its slice of rows with the online kernel; per Sinkhorn iteration
   the ranks exchange their a-slices and partial column sums either inside the kernels over peer memory (NVLink
   stores from the passes' finishing code, default) or with one NCCL all-reduce of 2 I + J float64 values.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
import time

import numpy as np


# ------------------------------------------------------------------------------------------------
# 1. independent units
# ------------------------------------------------------------------------------------------------
def shard_units(costs, world_size):
    """Longest-processing-time-first assignment.  costs[k] ~ I*J*expected_iterations of unit k.
    Returns a list (per rank) of unit indices; each rank's list keeps the original order."""
    loads = [0.0] * world_size
    owner = [0] * len(costs)
    for k in sorted(range(len(costs)), key=lambda q: (-costs[q], q)):
        r = min(range(world_size), key=lambda q: (loads[q], q))
        owner[k] = r
        loads[r] += float(costs[k])
    return [[k for k in range(len(costs)) if owner[k] == r] for r in range(world_size)]


def bind_host_to_gpu(device):
    """Pin the calling thread (and every thread it starts afterwards) to the CPU cores next to GPU `device`
    (NVML's ideal-CPU set), so that page-locked buffers are first touched on that GPU's NUMA node and the threads that
    feed it run there.  With one process per GPU on a two-socket host the 1-3 GB couplings otherwise cross the socket
    interconnect on their way from the GPU into host memory.  Returns the number of CPUs in the set, or 0 when NVML
    is unavailable or declines (containers); WOTB_NO_NUMA_BIND=1 disables it."""
    if os.environ.get("WOTB_NO_NUMA_BIND", "") == "1":
        return 0
    try:
        import pynvml
        import torch
        pynvml.nvmlInit()
        props = torch.cuda.get_device_properties(int(device))
        bus = "%08x:%02x:%02x.0" % (getattr(props, "pci_domain_id", 0), props.pci_bus_id, props.pci_device_id)
        handle = pynvml.nvmlDeviceGetHandleByPciBusId(bus.encode())
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return len(os.sched_getaffinity(0))
    except Exception:
        return 0


def _rank_world(group=None):
    try:
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(group), dist.get_world_size(group)
    except Exception:
        pass
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def compute_all_transport_maps(model, tmap_out="tmaps", overwrite=True, output_file_format="h5ad",
                               cost_matrices=None, group=None):
    """Distributed form of OTModel.compute_all_transport_maps (ot_model.py:124-201, covariate-free case).

    Every rank calls it with the same model.  Day-pairs are dealt to ranks by cell-count product; each rank
    writes its own '{prefix}_{t0}_{t1}.{fmt}' files; rank 0 writes '{prefix}_g.txt' with the rows of all
    pairs in day-pair order, exactly as the serial loop concatenates them."""
    import pandas as pd
    import torch.distributed as dist

    from . import io as _io

    rank, world = _rank_world(group)
    tmap_dir, tmap_prefix = os.path.split(tmap_out) if tmap_out is not None else (None, None)
    tmap_prefix = tmap_prefix or "tmaps"
    tmap_dir = tmap_dir or "."
    os.makedirs(tmap_dir, exist_ok=True)
    day_pairs = model.day_pairs
    if day_pairs is None or len(day_pairs) == 0:
        t = model.timepoints
        day_pairs = [(t[k], t[k + 1]) for k in range(len(t) - 1)]
    else:
        day_pairs = list(day_pairs)
    if cost_matrices is None:
        cost_matrices = [None] * len(day_pairs)
    files = [_io.check_file_extension(os.path.join(tmap_dir, tmap_prefix + "_{}_{}".format(*p)), output_file_format)
             for p in day_pairs]
    # skip before dispatch; rank 0 decides and broadcasts, so that a rank arriving late cannot see files written by
    # faster ranks in THIS run and deal itself a different assignment
    todo = [k for k in range(len(day_pairs)) if overwrite or not os.path.exists(files[k])]
    if world > 1:
        box = [todo if rank == 0 else None]
        dist.broadcast_object_list(box, src=0, group=group)
        todo = box[0]
    days = model.matrix.obs[model.day_field]
    counts = days.value_counts()
    costs = [float(counts.get(day_pairs[k][0], 0)) * float(counts.get(day_pairs[k][1], 0)) for k in todo]
    mine = [todo[q] for q in shard_units(costs, world)[rank]]
    keep_growth = model.ot_config.get("growth_iters", 1) > 1
    frames = {}

    _io.check_output_format(output_file_format)
    writer = _io.TmapWriter(output_file_format)

    def one(k):
        tmap = model.compute_transport_map(*day_pairs[k], cost_matrix=cost_matrices[k])
        if tmap is None:
            return None
        writer.write(tmap, files[k])
        return tmap.obs if keep_growth else None

    from .ot import optimal_transport as _ot
    streams = int(getattr(model, "streams", 1))
    if model.solver in (_ot.optimal_transport_duality_gap, _ot.transport_stablev2) and streams > 1 and len(mine) > 1:
        # this rank's pairs, `streams` solves in flight on its GPU plus one coupling on its way to the host
        from .pipeline import Pipeline
        with Pipeline(device=int(os.environ.get("LOCAL_RANK", "0")), streams=streams + 1, compute_slots=streams) as pipe:
            cost_of = dict(zip(todo, costs))
            got = pipe.map(lambda ctx, k: one(k), mine, costs=[cost_of[k] for k in mine])
    else:
        got = [one(k) for k in mine]
    writer.close()
    for k, obs in zip(mine, got):
        if obs is not None:
            frames[k] = obs
    if keep_growth:
        gathered = [frames]
        if world > 1:
            gathered = [None] * world
            dist.all_gather_object(gathered, frames, group=group)
        if rank == 0:
            merged = {}
            for part in gathered:
                merged.update(part)
            if merged:
                pd.concat([merged[k] for k in sorted(merged)]).to_csv(
                    os.path.join(tmap_dir, tmap_prefix + "_g.txt"), sep="\t", index_label="id")
    if world > 1:
        dist.barrier(group=group)
    return [day_pairs[k] for k in mine]


def sweep_grid(epsilons=(0.01, 0.025, 0.05, 0.1), lambda1s=(0.1, 1, 10, 50), lambda2s=(1, 10, 50, 100)):
    """The 64-setting grid of BASELINE.json configs[4] (shape of optimal_transport_validation_parameter_sweep.wdl:
    113-146, one `wot optimal_transport_validation` task per (epsilon, lambda1, lambda2)); the reference ships no
    parameter file, so the values are this build's.  Order: epsilon slowest."""
    return [dict(epsilon=float(e), lambda1=float(l1), lambda2=float(l2))
            for e in epsilons for l1 in lambda1s for l2 in lambda2s]


class WorkQueue:
    """Dynamic queue over the ranks of a torch.distributed job: a shared counter in the rendezvous store
    (`store.add` is atomic), so a rank that finishes early takes the next unit.  Nothing travels on the data
    path.  `order` is the dispatch order (longest expected unit first).  Without an initialised process group
    it is a plain serial iterator."""

    def __init__(self, order, store=None, key="wot_b200/queue"):
        self.order, self.key = list(order), key
        self.store = store
        self._next = 0
        if store is None:
            try:
                import torch.distributed as dist
                if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
                    from torch.distributed.distributed_c10d import _get_default_store
                    self.store = _get_default_store()
            except Exception:
                self.store = None

    def __iter__(self):
        return self

    def __next__(self):
        k = self.store.add(self.key, 1) - 1 if self.store is not None else self._next
        self._next += 1
        if k >= len(self.order):
            raise StopIteration
        return self.order[k]


_sweep_calls = 0


def expected_sweep_cost(setting):
    """Relative cost guess used only to order the queue: iteration counts grow roughly like
    (lambda1 + lambda2) / epsilon at fixed shape (70 .. >30k iterations)."""
    return (setting.get("lambda1", 1.0) + setting.get("lambda2", 50.0)) / setting.get("epsilon", 0.05)


def parameter_sweep(x0, x1, G, settings, solve=None, group=None, store=None, queue_key="wot_b200/sweep",
                    growth_iters=1, kernel="online", streams=1, **common):
    """BASELINE.json configs[4]: many (epsilon, lambda1, lambda2) settings on ONE day-pair, settings dealt to
    the ranks through a dynamic queue (independent units; no data-path collective).

    `solve(x0, x1, G, **params) -> dict` runs one setting; the default runs the GPU growth loop from
    coordinates without materialising the coupling and reports iteration/batch counts, the duality gap,
    the final row sums and the potentials.  Returns, on every rank, the list of per-setting results in the
    order of `settings` (each tagged with the rank that ran it).  `streams` > 1 (GPU solves only): every rank keeps
    that many settings in flight on separate CUDA streams (wot_b200.pipeline), each worker drawing from the same
    queue."""
    rank, world = _rank_world(group)
    if solve is None:
        solve = _sweep_solve_gpu
    else:
        streams = 1
    order = sorted(range(len(settings)), key=lambda k: (-expected_sweep_cost(settings[k]), k))
    mine = {}
    global _sweep_calls
    _sweep_calls += 1                  # every rank makes the same sequence of calls: a fresh counter per sweep
    queue_key = "%s/%d" % (queue_key, _sweep_calls)
    if world > 1:
        import torch.distributed as dist
        dist.barrier(group=group)      # start drawing together
    queue = WorkQueue(order, store=store, key=queue_key)
    queue_lock = threading.Lock()

    def draw():
        with queue_lock:                      # one store round trip at a time per rank
            return next(queue, None)

    def work(ctx=None):
        while True:
            k = draw()
            if k is None:
                return
            params = dict(common)
            params.update(settings[k])
            res = solve(x0, x1, G, growth_iters=growth_iters, kernel=kernel, **params)
            res["rank"] = rank
            res["setting"] = dict(settings[k])
            mine[k] = res

    if streams > 1:
        from .pipeline import Pipeline
        with Pipeline(device=int(os.environ.get("LOCAL_RANK", "0")), streams=streams) as pipe:
            for fut in [pipe.submit(work) for _ in range(streams)]:
                fut.result()
    else:
        work()
    parts = [mine]
    if world > 1:
        import torch.distributed as dist
        parts = [None] * world
        dist.all_gather_object(parts, mine, group=group)
    merged = {}
    for part in parts:
        merged.update(part)
    return [merged[k] for k in range(len(settings))]


def _sweep_solve_gpu(x0, x1, G, growth_iters=1, kernel="online", **params):
    from . import _lib
    from .ot import optimal_transport as wot_ot
    device = int(os.environ.get("LOCAL_RANK", "0"))      # a pipeline worker's bound context takes precedence
    _, learned = wot_ot.solve_coords(x0, x1, G, _lib.SOLVER_DUALITY_GAP, growth_iters=growth_iters, kernel=kernel,
                                     want_tmap=False, device=device, **params)
    last = wot_ot.last_solve_info()
    infos = last["infos"]
    return {"iters": sum(i["iters"] for i in infos), "batches": infos[-1]["batches"], "gap": infos[-1]["gap"],
            "status": infos[-1]["status"], "gpu_ms": sum(i["gpu_ms"] for i in infos), "rowsum": learned[-1].copy(),
            "f": np.array(last["f"]), "g": np.array(last["g"])}


# ------------------------------------------------------------------------------------------------
# 2. one pair, rows sharded
# ------------------------------------------------------------------------------------------------
OP_BEGIN_A, OP_BEGIN_B, OP_ROW, OP_COL_PARTIAL, OP_COL_FINISH, OP_GAP_ROWS, OP_CHECK, OP_FINAL_ROWS = range(8)


def hi_rows(n, rank, world, block=256):
    """rows of the slice of 256-row blocks rank `rank` owns (the split of wotb_online_open)"""
    blocks = -(-n // block)
    lo, hi = blocks * rank // world, blocks * (rank + 1) // world
    return max(0, min(hi * block, n) - min(lo * block, n))


def sharded_median(ctx, X0, X1, rank, world, group=None):
    """Exact np.median of the I*J squared distances (ot_model.py:252) with the one pass over the distances split over
    the ranks' row shards: every rank draws the same sample (the coordinates are replicated), counts and gathers its
    rows' share of the window, `below` is all-reduced, the gathered keys are all-gathered, and every rank finishes the
    select on the union -> the same bits everywhere.  Falls back to the full select on every rank when the problem is
    small, there is one rank, or the window check fails."""
    import torch
    import torch.distributed as dist

    from . import _lib
    lib, h = ctx.lib, ctx.handle
    n_i, n_j, d = X0.shape[0], X1.shape[0], X0.shape[1]
    P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
    med = C.c_double()
    cap = C.c_int64(0)
    _lib.check(lib.wotb_cost_median_window_cap(n_i, n_j, C.byref(cap)))
    if world > 1 and cap.value > 0:
        dev = X0.device
        lo, hi = n_i * rank // world, n_i * (rank + 1) // world
        keys = torch.empty(cap.value, dtype=torch.int64, device=dev)
        cnt = torch.zeros(2, dtype=torch.int32, device=dev)
        below = torch.zeros(1, dtype=torch.int64, device=dev)
        _lib.check(lib.wotb_cost_median_window_rows_dev(h, P(X0), n_i, P(X1), n_j, d, None, lo, hi, P(keys), cap.value,
                                                        P(cnt), P(below)))
        counts = torch.zeros(world, dtype=torch.int64, device=dev)
        counts[rank] = cnt[0].to(torch.int64)
        dist.all_reduce(counts, group=group)
        dist.all_reduce(below, group=group)
        counts_h = [int(v) for v in counts.cpu()]
        if max(counts_h) <= cap.value and sum(counts_h) < 2 ** 31:
            width = max(1, max(counts_h))
            parts = [torch.empty(width, dtype=torch.int64, device=dev) for _ in range(world)]
            dist.all_gather(parts, keys[:width].contiguous(), group=group)
            union = torch.cat([p[:c] for p, c in zip(parts, counts_h)]) if sum(counts_h) else keys[:0]
            ok = C.c_int32(0)
            _lib.check(lib.wotb_cost_median_window_finish_dev(h, n_i, n_j, P(union) if union.numel() else P(keys), union.numel(),
                                                              C.c_uint64(int(below.item())), C.byref(med), C.byref(ok)))
            if ok.value:
                return med.value
    _lib.check(lib.wotb_cost_median_dev(h, P(X0), n_i, P(X1), n_j, d, None, C.byref(med)))
    return med.value


class DistComm:
    """Host-side plumbing of a row-sharded solve over torch.distributed: rank / world, a barrier, an object
    all-gather (IPC handles) and the NCCL all-reduce of the `exchange="nccl"` mode."""
    in_process = False

    def __init__(self, group=None):
        self.group = group
        self.rank, self.world = _rank_world(group)

    def barrier(self):
        import torch.distributed as dist
        if self.world > 1:
            dist.barrier(group=self.group)

    def all_gather_object(self, obj):
        import torch.distributed as dist
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def all_reduce(self, tensor):
        import torch.distributed as dist
        if self.world > 1:
            dist.all_reduce(tensor, op=dist.ReduceOp.SUM, group=self.group)


class ThreadComm:
    """The same plumbing for `world` ranks that are THREADS of one process (each with its own library context and
    stream, on one GPU or several): peer buffers are plain device pointers, there is no NCCL.  Used by the tests to run
    the complete peer-memory protocol on a single GPU; make one with `ThreadComm.make(world)` and hand comm[r] to
    rank r's thread."""
    in_process = True

    def __init__(self, rank, world, shared):
        self.rank, self.world, self._sh = rank, world, shared

    @classmethod
    def make(cls, world):
        shared = {"bar": threading.Barrier(world), "slots": [None] * world}
        return [cls(r, world, shared) for r in range(world)]

    def barrier(self):
        self._sh["bar"].wait(timeout=120)

    def all_gather_object(self, obj):
        self._sh["slots"][self.rank] = obj
        self.barrier()
        out = list(self._sh["slots"])
        self.barrier()
        return out

    def all_reduce(self, tensor):
        raise RuntimeError("in-process ranks exchange over peer memory only")


class _PeerBuffers:
    """One exchange buffer per rank, mapped into every rank (wotb_peer_*): cudaIpc handles between processes, plain
    pointers between threads of one process."""

    def __init__(self, ctx, solve, comm):
        from . import _lib
        self.ctx, self.comm = ctx, comm
        self.own = C.c_void_p()
        self.opened = []
        lib, h = ctx.lib, ctx.handle
        nbytes = C.c_int64()
        _lib.check(lib.wotb_online_peer_bytes(solve, comm.world, C.byref(nbytes)))
        handle = (C.c_ubyte * 64)()
        err = None
        try:
            _lib.check(lib.wotb_peer_alloc(h, nbytes.value, C.byref(self.own), handle))
        except Exception as e:  # noqa: BLE001 -- every rank must reach the gathers below
            err = repr(e)
        mine = self.own.value if comm.in_process else bytes(handle)
        infos = comm.all_gather_object((err, mine))
        ptrs = (C.c_void_p * comm.world)()
        if all(e is None for e, _ in infos):
            for w, (_, item) in enumerate(infos):
                if w == comm.rank or comm.in_process:
                    ptrs[w] = self.own.value if w == comm.rank else item
                    continue
                try:
                    p = C.c_void_p()
                    buf = (C.c_ubyte * 64).from_buffer_copy(item)
                    _lib.check(lib.wotb_peer_open(h, buf, C.byref(p)))
                    self.opened.append(p)
                    ptrs[w] = p.value
                except Exception as e:  # noqa: BLE001
                    err = repr(e)
                    break
            if err is None:
                try:
                    _lib.check(lib.wotb_online_attach_peers(solve, comm.world, ptrs))
                except Exception as e:  # noqa: BLE001
                    err = repr(e)
        errs = [e for e in comm.all_gather_object(err) if e is not None] + [e for e, _ in infos if e is not None]
        self.error = errs[0] if errs else None      # the same verdict on every rank (also the barrier after attach)
        if self.error is not None:
            self.close(barrier=False)

    def close(self, barrier=True):
        lib, h = self.ctx.lib, self.ctx.handle
        for p in self.opened:
            lib.wotb_peer_close(h, p)
        self.opened = []
        if barrier:
            self.comm.barrier()                      # nobody frees a buffer a peer still has mapped
        if self.own.value:
            lib.wotb_peer_free(h, self.own)
            self.own = C.c_void_p()


def sharded_online_solve(x0, x1, G, group=None, device=None, stream=None, timers=None, use_graph=True, exchange="auto",
                         comm=None, **params):
    """Solve one day-pair with its rows sharded over the ranks of `group` (online kernel, float64 state).

    x0 [I,d], x1 [J,d], G [I]: the full arrays on every rank (NumPy or CUDA tensors).  Returns a dict with
    f, g (replicated CUDA tensors), rowsum (replicated), rows=(lo, hi) of this rank, median, info.
    With world_size 1 no collective is issued (the stepping path is then a plain single-GPU solve).

    One batch of the device state machine -- operand repack if epsilon changed, 5 x (row half-step on the own rows,
    partial column sums, ONE all-reduce, column finish), convergence check -- is captured ONCE as a CUDA graph,
    NCCL all-reduces included, and replayed until the replicated state says done: between two looks at the state
    the host issues a single cudaGraphLaunch, so neither Python nor launch latency sits between the passes and the
    collectives (`use_graph=False` keeps the step-by-step launches).  Every kernel of the sequence is a no-op on
    the device when its work is not due, so all ranks replay the same graph the same number of times.
    `timers` (dict, optional) receives allreduce_us_per_iter (CUDA events around 20 all-reduces of the per-iteration
    payload on the solve's stream), allreduce_bytes and the launch mode.

    exchange: "nccl" = the all-reduce described above; "peer" = no collective at all: the pass kernels store their
    results into every rank's exchange buffer over NVLink while they run, one warp exchanges flags, one kernel adds
    the partial column sums in rank order (include/wot_b200.h, wotb_online_attach_peers); "auto" = peer when the buffers
    can be mapped on every rank (cudaIpc), else nccl.  WOTB_SHARD_EXCHANGE overrides "auto".
    comm: DistComm(group) by default; ThreadComm for ranks that are threads of one process.
    """
    import torch

    from . import _lib

    comm = comm or DistComm(group)
    rank, world = comm.rank, comm.world
    if exchange == "auto":
        exchange = os.environ.get("WOTB_SHARD_EXCHANGE", "auto")
    if exchange not in ("auto", "peer", "nccl"):
        raise ValueError("exchange must be 'auto', 'peer' or 'nccl'")
    if world == 1 or world > 8:
        exchange = "nccl"
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(device)
    dev = torch.device("cuda", device)
    # the library's kernels and the NCCL all-reduces must be ordered on ONE stream: a dedicated torch stream
    # that is made current while the solve runs (the legacy default stream has no usable handle)
    stream = stream or torch.cuda.Stream(dev)
    ctx = _lib.Context(device, stream.cuda_stream)
    lib, h = ctx.lib, ctx.handle

    def to_dev(a):
        if isinstance(a, torch.Tensor):
            return a.to(dev, torch.float64).contiguous()
        return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)

    with torch.cuda.stream(stream):
        X0, X1, Gd = to_dev(x0), to_dev(x1), to_dev(G)
        n_i, n_j, d = X0.shape[0], X1.shape[0], X0.shape[1]
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        median = params.pop("median", None)
        median_ms = 0.0
        if median is None:
            if comm.in_process:
                raise ValueError("in-process ranks need the median passed in")
            mev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            mev[0].record(stream)
            median = sharded_median(ctx, X0, X1, rank, world, group)
            mev[1].record(stream)
            mev[1].synchronize()
            median_ms = mev[0].elapsed_time(mev[1])
        solver = params.pop("solver", _lib.SOLVER_DUALITY_GAP)
        prm = _lib.make_params(solver=solver, kernel=_lib.KERNEL_ONLINE, **params)
        f = torch.empty(n_i, dtype=torch.float64, device=dev)
        g = torch.empty(n_j, dtype=torch.float64, device=dev)
        exch = torch.zeros(2 * n_i + n_j, dtype=torch.float64, device=dev)
        solve = C.c_void_p()
        _lib.check(lib.wotb_online_open(h, P(X0), n_i, P(X1), n_j, d, median, P(Gd), C.byref(prm), rank, world, P(f),
                                        P(g), C.byref(solve)))

        peers = None
        if exchange in ("auto", "peer"):
            stream.synchronize()
            peers = _PeerBuffers(ctx, solve, comm)
            if peers.error is not None:
                if exchange == "peer":
                    lib.wotb_online_close(solve)
                    raise RuntimeError("peer-memory exchange unavailable: " + peers.error)
                peers = None
        exchange = "peer" if peers is not None else "nccl"

        def step(op):
            _lib.check(lib.wotb_online_step(solve, op, P(exch)))

        def reduce(n):
            if world > 1 and peers is None:
                comm.all_reduce(exch[:n])

        slots = 5 if solver == _lib.SOLVER_DUALITY_GAP else 10

        def batch():
            step(OP_BEGIN_A)
            if solver == _lib.SOLVER_DUALITY_GAP:
                reduce(n_i)
            step(OP_BEGIN_B)
            for _ in range(slots):
                step(OP_ROW)             # a for this rank's rows (+ their row sums, for the lazy gap check)
                step(OP_COL_PARTIAL)     # needs only this rank's a: no exchange in between
                reduce(2 * n_i + n_j)    # ONE all-reduce per iteration: gathered a | row sums | column sums
                step(OP_COL_FINISH)
            step(OP_CHECK)

        info, done = _lib.Info(), C.c_int32(0)

        def state():
            _lib.check(lib.wotb_online_state(solve, C.byref(info), C.byref(done)))
            return bool(done.value)

        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        graph = None
        try:
            if world > 1 and peers is None:
                reduce(n_i)                       # NCCL communicator and channels exist before timing / capture
                exch.zero_()
            ev[0].record(stream)
            batch()                               # first batch eagerly: sizes the lazily grown workspaces
            finished = state()
            capture_ms = 0.0
            if use_graph and not finished:
                # host-side work with the GPU idle (stream capture + graph instantiation; with NCCL inside the graph it
                # also registers the collective's buffers): part of gpu_ms, reported separately as graph_capture_ms
                t_cap = time.perf_counter()
                graph = torch.cuda.CUDAGraph()
                with torch.cuda.graph(graph, stream=stream, capture_error_mode="thread_local"):
                    batch()
                capture_ms = 1e3 * (time.perf_counter() - t_cap)
            if graph is not None and peers is not None and not finished:
                # peer exchange: nothing in the graph is a collective, and every kernel of a batch that is not due is
                # a no-op on the device -- so the ranks need not replay the same number of batches, and the host can
                # keep `depth` batches in flight and look at the done flag (mapped page-locked memory) instead of
                # synchronising on every batch.  (With NCCL inside the graph every rank must issue the same replays.)
                depth = 3
                ring = [torch.cuda.Event() for _ in range(depth)]
                flag = C.c_int32(0)
                k = 0
                while True:
                    graph.replay()
                    ring[k % depth].record(stream)
                    k += 1
                    if k >= depth:
                        ring[k % depth].synchronize()              # the oldest batch in flight
                        _lib.check(lib.wotb_online_done(solve, C.byref(flag)))
                        if flag.value:
                            break
                stream.synchronize()
                finished = state()
            while not finished:
                if graph is not None:
                    graph.replay()
                else:
                    batch()
                finished = state()
            step(OP_FINAL_ROWS)
            reduce(n_i)
            rowsum = exch[:n_i].clone()
            ev[1].record(stream)
            ev[1].synchronize()
            lo, hi = C.c_int64(), C.c_int64()
            _lib.check(lib.wotb_online_rows(solve, C.byref(lo), C.byref(hi)))
            if timers is not None:
                timers["exchange"] = exchange
                if peers is not None:
                    timers["mode"] = ("cuda graph per batch" if graph is not None else "stepwise launches") + \
                        ", peer-memory exchange (stores from the passes' finishing code, flag barrier, no collective)"
                    timers["peer_bytes_out_per_iter"] = (world - 1) * (2 * (hi_rows(n_i, rank, world)) + n_j) * 8
                else:
                    timers["mode"] = "cuda graph per batch (NCCL captured)" if graph is not None else "stepwise launches"
                    timers["allreduce_bytes"] = (2 * n_i + n_j) * 8
                if world > 1 and peers is None:
                    probe = torch.zeros(2 * n_i + n_j, dtype=torch.float64, device=dev)
                    te = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                    for _ in range(3):
                        comm.all_reduce(probe)
                    te[0].record(stream)
                    for _ in range(20):
                        comm.all_reduce(probe)
                    te[1].record(stream)
                    te[1].synchronize()
                    timers["allreduce_us_per_iter"] = te[0].elapsed_time(te[1]) * 1e3 / 20
        finally:
            del graph
            stream.synchronize()
            if peers is not None:
                peers.close()
            lib.wotb_online_close(solve)
        out_info = info.as_dict()
        out_info["gpu_ms"] = ev[0].elapsed_time(ev[1])
        out_info["graph_capture_ms"] = capture_ms
        out_info["median_ms"] = median_ms
    return {"f": f, "g": g, "rowsum": rowsum, "rows": (lo.value, hi.value), "median": median, "info": out_info,
            "ctx": ctx, "coords": (X0, X1), "stream": stream}


def local_coupling_rows(result, out_dtype=np.float64):
    """Materialise this rank's rows of the coupling of a sharded solve (host ndarray [hi-lo, J])."""
    import torch

    from . import _lib
    ctx = result["ctx"]
    X0, X1 = result["coords"]
    lo, hi = result["rows"]
    n_j, d = X1.shape[0], X1.shape[1]
    info = result["info"]
    tdt = torch.float64 if np.dtype(out_dtype) == np.float64 else torch.float32
    out = torch.empty((hi - lo, n_j), dtype=tdt, device=X0.device)
    if hi > lo:
        P = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        _lib.check(ctx.lib.wotb_coupling_online_dev(
            ctx.handle, C.c_void_p(X0.data_ptr() + lo * d * 8), hi - lo, P(X1), n_j, d, result["median"],
            C.c_void_p(result["f"].data_ptr() + lo * 8), P(result["g"]), info["eps_final"], info["out_scale"], P(out),
            n_j, _lib.F64 if tdt == torch.float64 else _lib.F32, None))
    return out.cpu().numpy()
