"""GPU solver callables with the reference's signatures (reference: wot/ot/optimal_transport.py).

  compute_transport_matrix(solver, **params)            :10-33   growth loop
  optimal_transport_duality_gap(C, G, lambda1, ...)     :67-164  default solver
  transport_stablev2(C, lambda1, lambda2, epsilon, ...) :167-236 fixed-iteration solver

NumPy arrays in, NumPy float64 coupling out; everything in between runs in libwot_b200.so on the GPU
(there is no CPU path).  Additive, not in the reference: `last_solve_info()` returns iteration / batch
counts, the dual potentials f, g and the row sums of the last call, which the parity criteria need.
"""
from __future__ import annotations

import ctypes as C
import logging
import threading

import numpy as np

from .. import _lib, _pinned

logger = logging.getLogger("wot")

_tls = threading.local()


def last_solve_info():
    """dict(infos=[per growth iteration], f, g, learned_growth, median) of the most recent solve made by the
    calling thread (the workers of wot_b200.pipeline each see their own)."""
    if not hasattr(_tls, "last"):
        _tls.last = {}
    return _tls.last


def resolve_kernel(kernel, n_i, n_j, d, epsilon=0.05):
    """'auto' -> 'online' or 'stored'.  The online kernel (K recomputed tile by tile on tcgen05 + MUFU, nothing of
    size I x J in memory) is the faster one on B200 and is used whenever the coordinates fit its K budget (d <= 46).
    Below a final epsilon of 0.02 the library switches its operands to the precise 6-segment form (leading limb on a
    fixed-point grid, exact accumulation of the large cancelling terms: exponent error ~1e-6 whatever epsilon, about
    twice the tensor-core work), so that slowly converging settings reproduce the reference's batch counts
    (tests/test_gpu_parity.py::test_sweep_settings_vs_oracle)."""
    if kernel != "auto":
        return kernel
    return "online" if d <= 46 else "stored"


def _kernel_id(kernel):
    """(wotb_kernel, force_simt, precise).  'online' uses the tcgen05 pass kernel when d <= 46 (precise operands
    below final epsilon 0.02) and the SIMT FP32 kernel otherwise; 'online_simt' forces the latter; 'online_fast' /
    'online_precise' pin the operand form of the tcgen05 pass."""
    if kernel in (None, "stored", _lib.KERNEL_STORED):
        return _lib.KERNEL_STORED, False, None
    if kernel in ("online", _lib.KERNEL_ONLINE):
        return _lib.KERNEL_ONLINE, False, None
    if kernel == "online_simt":
        return _lib.KERNEL_ONLINE, True, None
    if kernel == "online_fast":
        return _lib.KERNEL_ONLINE, False, False
    if kernel == "online_precise":
        return _lib.KERNEL_ONLINE, False, True
    raise ValueError("kernel must be 'stored', 'online', 'online_fast', 'online_precise' or 'online_simt'")


def _out_array(shape, out, out_dtype, pinned):
    out_dtype = np.dtype(out_dtype)
    if out_dtype not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise ValueError("out_dtype must be float32 or float64")
    if out is not None:
        if out.shape != tuple(shape) or out.dtype != out_dtype or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous %s array of shape %s" % (out_dtype, tuple(shape)))
        return out
    return _pinned.empty(shape, out_dtype) if pinned else np.empty(shape, dtype=out_dtype)


def _record(infos, f, g, learned, median=None):
    last = last_solve_info()
    last.clear()
    last.update(infos=[i.as_dict() for i in infos], f=f, g=g, learned_growth=learned, median=median)
    for i in infos:
        if i.status == _lib.STATUS_MAX_ITER:
            logger.warning("Reached max_iter with duality gap still above threshold. Returning")


def solve_cost(C_mat, G, solver_id, growth_iters=1, out=None, out_dtype=np.float64, want_tmap=True, pinned=True,
               device=None, ctx=None, **params):
    """Growth loop on a caller-supplied cost matrix.  Returns (tmap or None, learned_growth [g+1, I])."""
    C_mat = np.ascontiguousarray(C_mat, dtype=np.float64)
    if C_mat.ndim != 2:
        raise ValueError("C must be a 2-D cost matrix")
    n_i, n_j = C_mat.shape
    G = np.ascontiguousarray(G, dtype=np.float64)
    if G.shape != (n_i,):
        raise ValueError("G must have one entry per row of C")
    growth_iters = int(growth_iters)
    prm = _lib.make_params(solver=solver_id, **params)
    ctx = ctx or _lib.context(device)
    tmap = _out_array((n_i, n_j), out, out_dtype, pinned) if want_tmap else None
    learned = np.empty((growth_iters + 1, n_i))
    f, g = np.empty(n_i), np.empty(n_j)
    infos = (_lib.Info * growth_iters)()
    dt = _lib.F32 if (tmap is not None and tmap.dtype == np.float32) else _lib.F64
    rc = ctx.lib.wotb_transport_map_from_cost_host(ctx.handle, _lib.ptr(C_mat), n_i, n_j, _lib.ptr(G), C.byref(prm),
                                                   growth_iters, _lib.ptr(tmap), dt, _lib.ptr(learned), _lib.ptr(f),
                                                   _lib.ptr(g), infos)
    _lib.check(rc)
    _record(infos, f, g, learned)
    return tmap, learned


def solve_coords(x0, x1, G, solver_id, scale=None, growth_iters=1, kernel="auto", out=None, out_dtype=np.float64,
                 want_tmap=True, pinned=True, device=None, ctx=None, **params):
    """Default cost (ot_model.py:242-253) + growth loop from local-PCA coordinates, all on the GPU."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    if x0.ndim != 2 or x1.ndim != 2 or x0.shape[1] != x1.shape[1]:
        raise ValueError("x0 and x1 must be 2-D with the same number of columns")
    n_i, n_j, d = x0.shape[0], x1.shape[0], x0.shape[1]
    G = np.ascontiguousarray(G, dtype=np.float64)
    if G.shape != (n_i,):
        raise ValueError("G must have one entry per row of x0")
    if scale is not None:
        scale = np.ascontiguousarray(scale, dtype=np.float64)
        if scale.shape != (d,):
            raise ValueError("scale must have one entry per coordinate")
    growth_iters = int(growth_iters)
    # final epsilon: epsilon0 * epsilon for the duality-gap schedule (:102-120), epsilon for fixed_iters (:184-185)
    eps_final = float(params.get("epsilon", 0.05))
    if solver_id == _lib.SOLVER_DUALITY_GAP:
        eps_final *= float(params.get("epsilon0", 1.0))
    kernel = resolve_kernel(kernel, n_i, n_j, d, eps_final)
    kernel_id, simt, precise = _kernel_id(kernel)
    prm = _lib.make_params(solver=solver_id, kernel=kernel_id, online_simt=simt, online_precise=precise, **params)
    ctx = ctx or _lib.context(device)
    tmap = _out_array((n_i, n_j), out, out_dtype, pinned) if want_tmap else None
    learned = np.empty((growth_iters + 1, n_i))
    f, g = np.empty(n_i), np.empty(n_j)
    infos = (_lib.Info * growth_iters)()
    median = C.c_double(0.0)
    dt = _lib.F32 if (tmap is not None and tmap.dtype == np.float32) else _lib.F64
    rc = ctx.lib.wotb_transport_map_from_coords_host(ctx.handle, _lib.ptr(x0), n_i, _lib.ptr(x1), n_j, d,
                                                     _lib.ptr(scale), _lib.ptr(G), C.byref(prm), growth_iters,
                                                     _lib.ptr(tmap), dt, _lib.ptr(learned), _lib.ptr(f), _lib.ptr(g),
                                                     C.byref(median), infos)
    _lib.check(rc)
    _record(infos, f, g, learned, median.value)
    return tmap, learned


def default_cost_matrix(x0, x1, scale=None, device=None):
    """float64 median-normalised squared-Euclidean cost, computed on the GPU."""
    x0 = np.ascontiguousarray(x0, dtype=np.float64)
    x1 = np.ascontiguousarray(x1, dtype=np.float64)
    n_i, n_j, d = x0.shape[0], x1.shape[0], x0.shape[1]
    if scale is not None:
        scale = np.ascontiguousarray(scale, dtype=np.float64)
    ctx = _lib.context(device)
    out = np.empty((n_i, n_j))
    median = C.c_double(0.0)
    _lib.check(ctx.lib.wotb_default_cost_matrix_host(ctx.handle, _lib.ptr(x0), n_i, _lib.ptr(x1), n_j, d,
                                                     _lib.ptr(scale), _lib.ptr(out), C.byref(median)))
    return out


def optimal_transport_duality_gap(C, G, lambda1, lambda2, epsilon, batch_size, tolerance, tau, epsilon0, max_iter,
                                  **ignored):
    """Unbalanced entropic OT with the guarantee that the duality gap is at most `tolerance`
    (optimal_transport.py:67-164).  Returns the I x J transport map as float64 ndarray."""
    tmap, _ = solve_cost(C, G, _lib.SOLVER_DUALITY_GAP, growth_iters=1, lambda1=lambda1, lambda2=lambda2,
                         epsilon=epsilon, batch_size=batch_size, tolerance=tolerance, tau=tau, epsilon0=epsilon0,
                         max_iter=max_iter, **_extras(ignored))
    return tmap


def transport_stablev2(C, lambda1, lambda2, epsilon, scaling_iter, G, tau, epsilon0, extra_iter, inner_iter_max,
                       **ignored):
    """Fixed-iteration stabilised scaling (optimal_transport.py:167-236)."""
    if tau is None:
        # the reference evaluates `max(...) > None` at :211, which raises in Python 3
        raise TypeError("'>' not supported between instances of 'float' and 'NoneType'")
    tmap, _ = solve_cost(C, G, _lib.SOLVER_FIXED_ITERS, growth_iters=1, lambda1=lambda1, lambda2=lambda2,
                         epsilon=epsilon, scaling_iter=scaling_iter, tau=tau, epsilon0=epsilon0,
                         extra_iter=extra_iter, inner_iter_max=inner_iter_max, **_extras(ignored))
    return tmap


_SOLVER_IDS = {optimal_transport_duality_gap: _lib.SOLVER_DUALITY_GAP, transport_stablev2: _lib.SOLVER_FIXED_ITERS}
_EXTRA_KEYS = ("out", "out_dtype", "pinned", "device", "ctx", "use_graph", "fuse", "online_batch", "want_tmap")


def _extras(ignored):
    return {k: ignored[k] for k in _EXTRA_KEYS if k in ignored}


def compute_transport_matrix(solver, **params):
    """Growth-iteration loop (optimal_transport.py:10-33).  Returns (tmap, [G_0 .. G_{growth_iters-1}]).

    With one of this module's solvers the whole loop (solves, row sums, the final coupling) is one library
    call and the coupling crosses PCIe once.  Any other callable is driven exactly like the reference does.
    Additive: pass coords=(x0, x1, scale_or_None) instead of C to build the default cost on the GPU.
    """
    growth_iters = int(params["growth_iters"])
    solver_id = _SOLVER_IDS.get(solver)
    if solver_id is None:
        learned, rows, tmap = [], params["G"], None
        for it in range(growth_iters):
            if it > 0:
                rows = tmap.sum(axis=1)
            params["G"] = rows
            learned.append(rows)
            tmap = solver(**params)
        return tmap, learned
    coords = params.pop("coords", None)
    keys = ("lambda1", "lambda2", "epsilon", "batch_size", "tolerance", "tau", "epsilon0", "max_iter",
            "scaling_iter", "extra_iter", "inner_iter_max") + _EXTRA_KEYS
    kw = {k: params[k] for k in keys if k in params}
    if coords is not None and params.get("C") is None:
        x0, x1, scale = coords
        tmap, learned = solve_coords(x0, x1, params["G"], solver_id, scale=scale, growth_iters=growth_iters,
                                     kernel=params.get("kernel", "auto"), **kw)
    else:
        tmap, learned = solve_cost(params["C"], params["G"], solver_id, growth_iters=growth_iters, **kw)
    params["G"] = learned[growth_iters - 1]
    return tmap, [learned[k] for k in range(growth_iters)]
