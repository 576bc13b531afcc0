"""Several transport maps in flight on one GPU.

Day-pairs (ot_model.py:182-199) and sweep settings are independent, and one solve leaves the GPU idle in
places: the tail and launch gap of every pass kernel (each is a one-wave grid that must drain before the next
half-iteration can start), the convergence checks, the median selection, the copy of the coupling to the host.
A second solve on its own CUDA stream fills those holes: both kernels are one CTA per SM, so the hardware hands
every SM that a draining kernel frees to the other stream's kernel.  Measured on B200 (atlas-shaped config):
see DESIGN.md section 7.

`Pipeline(device, streams)` owns `streams` library contexts (one CUDA stream and workspace set each) and one
worker thread per context; ctypes releases the GIL inside the library, so the workers really overlap.
Results come back in submission order.
"""
from __future__ import annotations

import queue
import threading
from concurrent.futures import Future

from . import _lib


# Library contexts are kept between pipelines: their grow-only workspaces are sized by the largest day-pair seen,
# and growing one (cudaFree + cudaMalloc) synchronises the whole device, i.e. stalls every stream of a pipeline.
_idle_contexts = {}      # device -> [Context]
_idle_lock = threading.Lock()

# wotb_set_compute_slots / wotb_set_pdl are process-wide library settings; several pipelines may be open at once
# (two OTModel.compute_all_transport_maps calls in different threads, a parameter sweep alongside), so they are
# reference counted: slots follow the most recent pipeline that asked for them and return to "no limit" when the
# last such pipeline closes; programmatic dependent launch stays off while ANY multi-stream pipeline is open.
_settings_lock = threading.Lock()
_open_slot_limits = []   # compute_slots of the open pipelines that set one, in opening order
_open_multi = 0          # open pipelines with more than one stream


def _settings_open(slots, multi):
    global _open_multi
    lib = _lib.load()
    with _settings_lock:
        if slots:
            _open_slot_limits.append(slots)
            lib.wotb_set_compute_slots(int(slots))
        if multi:
            _open_multi += 1
            lib.wotb_set_pdl(0)


def _settings_close(slots, multi):
    global _open_multi
    lib = _lib.load()
    with _settings_lock:
        if slots:
            _open_slot_limits.remove(slots)
            lib.wotb_set_compute_slots(int(_open_slot_limits[-1]) if _open_slot_limits else 0)
        if multi:
            _open_multi -= 1
            if _open_multi == 0:
                lib.wotb_set_pdl(1)


class Pipeline:
    def __init__(self, device=None, streams=2, make_stream=None, compute_slots=0):
        """`make_stream(k)` may return a raw cudaStream_t handle for context k (e.g. a torch stream's
        .cuda_stream, kept alive by the caller); by default every context creates its own stream.
        `compute_slots` n > 0: at most n host-buffer solves are in their solve phase at once (process-wide,
        wotb_set_compute_slots); with streams = n + 1 the extra context overlaps its coupling's trip over PCIe
        with the other contexts' solves."""
        if streams < 1:
            raise ValueError("streams must be >= 1")
        self._slots = max(0, int(compute_slots))
        # programmatic dependent launch helps a lone solve and hurts interleaved ones (csrc/online_tc.cuh)
        self._multi = streams > 1
        _settings_open(self._slots, self._multi)
        self._settings_held = True
        if device is None:
            device = _lib.context().device
        self.device = int(device)
        self._cached = make_stream is None        # contexts on caller-owned streams die with the pipeline
        self.contexts = []
        for k in range(streams):
            ctx = None
            if self._cached:
                with _idle_lock:
                    idle = _idle_contexts.get(self.device, [])
                    ctx = idle.pop() if idle else None
            self.contexts.append(ctx or _lib.Context(self.device, make_stream(k) if make_stream else None))
        self._jobs = queue.Queue()
        self._threads = [threading.Thread(target=self._work, args=(ctx,), daemon=True) for ctx in self.contexts]
        for t in self._threads:
            t.start()

    def _work(self, ctx):
        _lib._thread.ctx = ctx       # library calls made on this thread without an explicit ctx use this context
        while True:
            job = self._jobs.get()
            if job is None:
                return
            fut, fn, args, kwargs = job
            if not fut.set_running_or_notify_cancel():
                continue
            try:
                fut.set_result(fn(ctx, *args, **kwargs))
            except BaseException as exc:  # noqa: BLE001 - handed to the caller through the future
                fut.set_exception(exc)

    def submit(self, fn, *args, **kwargs):
        """Run fn(ctx, *args, **kwargs) on the next free context; returns a concurrent.futures.Future."""
        fut = Future()
        self._jobs.put((fut, fn, args, kwargs))
        return fut

    def map(self, fn, items, costs=None):
        """fn(ctx, item) for every item, at most `streams` at a time; results in the order of `items`.
        With `costs`, items are started largest first (the first jobs size every context's workspaces once, and
        the tail of the run is made of short jobs)."""
        order = range(len(items)) if costs is None else sorted(range(len(items)), key=lambda k: (-costs[k], k))
        futs = {k: self.submit(fn, items[k]) for k in order}
        return [futs[k].result() for k in range(len(items))]


    def close(self):
        for _ in self._threads:
            self._jobs.put(None)
        for t in self._threads:
            t.join()
        for ctx in self.contexts:
            if self._cached:
                with _idle_lock:
                    _idle_contexts.setdefault(self.device, []).append(ctx)
            else:
                ctx.close()
        self.contexts = []
        if self._settings_held:
            _settings_close(self._slots, self._multi)
            self._settings_held = False

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def release_idle_contexts():
    """Destroy the cached contexts (frees their device workspaces)."""
    with _idle_lock:
        for ctxs in _idle_contexts.values():
            for ctx in ctxs:
                ctx.close()
        _idle_contexts.clear()
