"""ctypes binding of include/wot_b200.h.  There is no CPU fallback: a missing library or a missing
GPU raises, it never silently computes elsewhere."""
from __future__ import annotations

import ctypes as C
import math
import os
import threading

import numpy as np

from ._build import LIB

N_STAGES = 6
OK, ERR_INVALID, ERR_CUDA, ERR_NOMEM, ERR_NAN_GAP = range(5)
SOLVER_DUALITY_GAP, SOLVER_FIXED_ITERS = 0, 1
KERNEL_STORED, KERNEL_ONLINE = 0, 1
STATUS_CONVERGED, STATUS_MAX_ITER, STATUS_NAN = 0, 1, 2
F32, F64 = 0, 1


class Params(C.Structure):
    _fields_ = [("epsilon", C.c_double), ("lambda1", C.c_double), ("lambda2", C.c_double),
                ("epsilon0", C.c_double), ("tau", C.c_double), ("tolerance", C.c_double),
                ("max_iter", C.c_double), ("batch_size", C.c_int32), ("scaling_iter", C.c_int32),
                ("extra_iter", C.c_int32), ("inner_iter_max", C.c_int32), ("solver", C.c_int32),
                ("kernel", C.c_int32), ("use_graph", C.c_int32), ("reserved", C.c_int32)]


class Info(C.Structure):
    _fields_ = [("iters", C.c_int64), ("batches", C.c_int32 * N_STAGES), ("tau_absorptions", C.c_int32),
                ("status", C.c_int32), ("gap", C.c_double), ("primal", C.c_double), ("dual", C.c_double),
                ("eps_final", C.c_double), ("out_scale", C.c_double), ("gpu_ms", C.c_double),
                ("launches", C.c_int64), ("matvec_launches", C.c_int64)]

    def as_dict(self):
        return {"iters": int(self.iters), "batches": [int(b) for b in self.batches],
                "tau_absorptions": int(self.tau_absorptions), "status": int(self.status), "gap": float(self.gap),
                "primal": float(self.primal), "dual": float(self.dual), "eps_final": float(self.eps_final),
                "out_scale": float(self.out_scale), "gpu_ms": float(self.gpu_ms), "launches": int(self.launches),
                "matvec_launches": int(self.matvec_launches)}


_P = C.c_void_p
_I64 = C.c_int64
_I32 = C.c_int32
_D = C.c_double

# name -> (restype, argtypes); every symbol include/wot_b200.h declares
SIGNATURES = {
    "wotb_version": (C.c_char_p, []),
    "wotb_last_error": (C.c_char_p, []),
    "wotb_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "wotb_destroy": (None, [_P]),
    "wotb_sync": (C.c_int, [_P]),
    "wotb_workspace_bytes": (C.c_size_t, [_P]),
    "wotb_release_workspace": (None, [_P]),
    "wotb_set_compute_slots": (None, [C.c_int32]),
    "wotb_set_pdl": (None, [C.c_int32]),
    "wotb_set_sm_limit": (C.c_int, [_P, C.c_int32]),
    "wotb_coupling_apply_host": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.c_int32, _P, C.c_double, _P, _P, C.c_double,
                                           C.c_double, C.c_int32, _P, C.c_int32, _P]),
    "wotb_coupling_sample_host": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.c_int32, _P, C.c_double, _P, _P, C.c_double,
                                            C.c_double, _P, _P, _P, C.c_int64, _P]),
    "wotb_pca_host": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int64, C.c_int64, C.c_int32, _P, C.c_int32, C.c_int32, _P, _P, _P,
                                _P, _P]),
    "wotb_cost_median_dev": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, C.POINTER(_D)]),
    "wotb_cost_median_window_cap": (C.c_int, [_I64, _I64, C.POINTER(_I64)]),
    "wotb_cost_median_window_rows_dev": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, _I64, _I64, _P, _I64, _P, _P]),
    "wotb_cost_median_window_finish_dev": (C.c_int, [_P, _I64, _I64, _P, _I64, C.c_uint64, C.POINTER(_D), C.POINTER(_I32)]),
    "wotb_cost_matrix_dev": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, _D, _P, _I64, _I32]),
    "wotb_cost_to_f32_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _I64]),
    "wotb_sinkhorn_stored_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, C.POINTER(Params), _P, _P, _P,
                                           C.POINTER(Info)]),
    "wotb_sinkhorn_online_dev": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _P, C.POINTER(Params), _P, _P, _P,
                                           C.POINTER(Info)]),
    "wotb_coupling_dev": (C.c_int, [_P, _P, _I64, _I64, _I64, _P, _P, _D, _D, _P, _I64, _I32, _P]),
    "wotb_coupling_online_dev": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _P, _P, _D, _D, _P, _I64, _I32, _P]),
    "wotb_transport_map_from_cost_host": (C.c_int, [_P, _P, _I64, _I64, _P, C.POINTER(Params), _I32, _P, _I32, _P,
                                                    _P, _P, C.POINTER(Info)]),
    "wotb_transport_map_from_coords_host": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, _P, C.POINTER(Params), _I32,
                                                      _P, _I32, _P, _P, _P, C.POINTER(_D), C.POINTER(Info)]),
    "wotb_default_cost_matrix_host": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _P, _P, C.POINTER(_D)]),
    "wotb_online_open": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _P, C.POINTER(Params), _I32, _I32, _P, _P,
                                   C.POINTER(_P)]),
    "wotb_online_step": (C.c_int, [_P, _I32, _P]),
    "wotb_online_state": (C.c_int, [_P, C.POINTER(Info), C.POINTER(_I32)]),
    "wotb_online_done": (C.c_int, [_P, C.POINTER(_I32)]),
    "wotb_online_rows": (C.c_int, [_P, C.POINTER(_I64), C.POINTER(_I64)]),
    "wotb_online_close": (None, [_P]),
    "wotb_peer_alloc": (C.c_int, [_P, _I64, C.POINTER(_P), _P]),
    "wotb_peer_open": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "wotb_peer_close": (C.c_int, [_P, _P]),
    "wotb_peer_free": (C.c_int, [_P, _P]),
    "wotb_online_peer_bytes": (C.c_int, [_P, _I32, C.POINTER(_I64)]),
    "wotb_online_attach_peers": (C.c_int, [_P, _I32, C.POINTER(_P)]),
    "wotb_bench_matvec_dev": (C.c_int, [_P, _I64, _I64, _I32, C.POINTER(_D), C.POINTER(_D), C.POINTER(_D)]),
    "wotb_online_rowsums_dev": (C.c_int, [_P, _P, _I64, _P, _I64, _I32, _D, _P, _P, _I32, _I32, _P, C.POINTER(_D)]),
    "wotb_bench_mufu_dev": (C.c_int, [_P, C.POINTER(_D)]),
    "wotb_pinned_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "wotb_pinned_free": (None, [_P]),
}

_lib = None


def load():
    """dlopen csrc/libwot_b200.so (no CUDA call happens here) and set the prototypes."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("WOT_B200_LIB", LIB)   # developer knob: an alternative build of the same library
    if not os.path.exists(path):
        raise ImportError("%s is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(wot_b200 has no CPU fallback)" % path)
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


class WotB200Error(RuntimeError):
    pass


def check(rc):
    if rc == OK:
        return
    msg = load().wotb_last_error().decode("utf-8", "replace")
    if rc == ERR_NAN_GAP:
        # same exception type and text as optimal_transport.py:162-163
        raise RuntimeError(msg)
    if rc == ERR_INVALID:
        raise ValueError(msg)
    if rc == ERR_NOMEM:
        raise MemoryError(msg)
    raise WotB200Error(msg)


def make_params(epsilon=0.05, lambda1=1, lambda2=50, epsilon0=1, tau=10000, tolerance=1e-8, max_iter=1e7,
                batch_size=5, scaling_iter=3000, extra_iter=1000, inner_iter_max=50, solver=SOLVER_DUALITY_GAP,
                kernel=KERNEL_STORED, use_graph=True, fuse=True, online_simt=False, online_precise=None, online_batch=False,
                **ignored):
    """Pack the ot_config keys the solvers read (ot_model.py:85-87).  Unknown keys are ignored, like the
    reference solvers' **ignored."""
    p = Params()
    p.epsilon, p.lambda1, p.lambda2, p.epsilon0 = float(epsilon), float(lambda1), float(lambda2), float(epsilon0)
    p.tau = math.nan if tau is None else float(tau)
    p.tolerance, p.max_iter = float(tolerance), float(max_iter)
    p.batch_size, p.scaling_iter = int(batch_size), int(scaling_iter)
    p.extra_iter, p.inner_iter_max = int(extra_iter), int(inner_iter_max)
    p.solver, p.kernel, p.use_graph = int(solver), int(kernel), int(bool(use_graph))
    # bit0: disable the fused (K-read-once) iteration kernel; bit1: online kernel on the SIMT FP32 pass, not tcgen05;
    # bit2 / bit3: force / forbid the precise 6-segment operands of the tcgen05 pass (default: by final epsilon)
    p.reserved = (0 if fuse else 1) | (2 if online_simt else 0)
    if online_precise is not None:
        p.reserved |= 4 if online_precise else 8
    if online_batch:             # bit4: one persistent cooperative launch per batch of iterations (off by default)
        p.reserved |= 16
    return p


class Context:
    """One wotb_ctx: a CUDA stream plus grow-only workspaces.  The process-wide default (context()) serves plain calls;
    wot_b200.pipeline creates one per worker thread.  A context is not thread-safe."""

    def __init__(self, device=0, stream=None):
        lib = load()
        handle = _P()
        check(lib.wotb_create(int(device), _P(stream) if stream else None, C.byref(handle)))
        self.lib, self.handle, self.device = lib, handle, int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.wotb_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_contexts = {}
_thread = threading.local()      # .ctx: the context a pipeline worker thread is bound to


def context(device=None):
    """The context library calls of this thread use: the one a wot_b200.pipeline worker is bound to, else the
    process-wide context of `device` (default: LOCAL_RANK, else 0)."""
    bound = getattr(_thread, "ctx", None)
    if bound is not None and bound.handle and (device is None or int(device) == bound.device):
        return bound
    if device is None:
        device = int(os.environ.get("WOT_B200_DEVICE", os.environ.get("LOCAL_RANK", "0")))
    ctx = _contexts.get(device)
    if ctx is None:
        ctx = _contexts[device] = Context(device)
    return ctx


def ptr(arr):
    return None if arr is None else _P(arr.ctypes.data)


class PinnedArray:
    """Page-locked host ndarrays from the library (cudaHostAlloc), so couplings travel at PCIe speed."""

    def __init__(self, nbytes):
        self.lib = load()
        p = _P()
        check(self.lib.wotb_pinned_alloc(C.c_size_t(int(nbytes)), C.byref(p)))
        self.ptr, self.nbytes = p, int(nbytes)
        self.buf = (C.c_char * self.nbytes).from_address(p.value)

    def view(self, shape, dtype):
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        assert n <= self.nbytes
        arr = np.frombuffer(self.buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr

    def __del__(self):
        try:
            if self.ptr:
                self.lib.wotb_pinned_free(self.ptr)
                self.ptr = None
        except Exception:
            pass
