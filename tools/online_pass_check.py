#!/usr/bin/env python
"""Kernel-level check and timing of one online-kernel pass (wotb_online_rowsums_dev): the SIMT FP32
kernel and the tcgen05 kernel against float64 NumPy, on offsets and scales shaped like a final-stage
Sinkhorn state.  Usage: python tools/online_pass_check.py [--check 3000x3301] [--time 12486x12405 ...]"""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import _lib, synthetic  # noqa: E402


def make_inputs(n_out, n_in, d, seed, eps=0.05):
    """Coordinates of a synthetic day pair, offsets of the shape c1*u - c2*|x|^2 with potentials that keep
    the row sums O(1), as in a converged solve."""
    x0, x1, _ = synthetic.day_pair_coords(n_out, n_in, d=d, seed=seed)
    rng = np.random.default_rng(seed + 1)
    sub0 = x0[rng.choice(n_out, min(n_out, 512), replace=False)]
    sub1 = x1[rng.choice(n_in, min(n_in, 512), replace=False)]
    med = np.median(((sub0[:, None, :] - sub1[None, :, :]) ** 2).sum(-1))
    c1 = np.log2(np.e) / eps
    c2 = c1 / med
    scale = np.sqrt(2 * c2)
    off_out = -c2 * (x0 ** 2).sum(1) + c1 * 0.05 * rng.standard_normal(n_out)
    off_in = -c2 * (x1 ** 2).sum(1) + c1 * 0.05 * rng.standard_normal(n_in) - np.log2(n_in)
    return x0, x1, scale, off_out, off_in


def reference(x0, x1, scale, off_out, off_in, block=1024):
    out = np.empty(x0.shape[0])
    ys = (scale * scale) * x1.T
    for r in range(0, x0.shape[0], block):
        e = x0[r:r + block] @ ys + off_out[r:r + block, None] + off_in[None, :]
        out[r:r + block] = np.exp2(e).sum(1)
    return out


def run(ctx, torch, x0, x1, scale, off_out, off_in, impl, reps):
    dev = "cuda:%d" % ctx.device
    t = [torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (x0, x1, off_out, off_in)]
    sums = torch.empty(x0.shape[0], dtype=torch.float64, device=dev)
    ms = C.c_double()
    P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.wotb_online_rowsums_dev(ctx.handle, P(t[0]), x0.shape[0], P(t[1]), x1.shape[0], x0.shape[1],
                                               float(scale), P(t[2]), P(t[3]), impl, reps, P(sums), C.byref(ms)))
    return sums.cpu().numpy(), ms.value


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", nargs="*", default=["300x340", "3000x3301"])
    ap.add_argument("--time", nargs="*", default=["12486x12405"])
    ap.add_argument("--d", type=int, default=30)
    ap.add_argument("--impls", default="0,1")
    ap.add_argument("--eps", type=float, nargs="*", default=[0.05], help="epsilon of the --check states")
    args = ap.parse_args()
    import torch
    ctx = _lib.context(0)
    impls = [int(v) for v in args.impls.split(",")]
    names = {0: "simt", 1: "tcgen05 ew4", 2: "tcgen05 ew8", 3: "tcgen05 precise"}
    for dbg in range(1, 64):
        for base in (1, 2, 3):
            names[base + 16 * dbg] = names[base] + "".join(t for b, t in ((1, " no-mma"), (2, " no-exp"), (4, " ld-only"), (8, " prof"), (16, " mma-x2"), (32, " no-accwait")) if dbg & b)
    peak = C.c_double()
    _lib.check(ctx.lib.wotb_bench_mufu_dev(ctx.handle, C.byref(peak)))
    print("MUFU.EX2 peak (measured): %.3f T ex2/s" % (peak.value / 1e12), flush=True)
    for shape in args.check:
        n_out, n_in = (int(v) for v in shape.split("x"))
        for eps in args.eps:
            x0, x1, scale, po, pi = make_inputs(n_out, n_in, args.d, seed=5, eps=eps)
            want = reference(x0, x1, scale, po, pi)
            for impl in impls:
                got, ms = run(ctx, torch, x0, x1, scale, po, pi, impl, 1)
                ok = want > 1e-30
                err = np.abs(got[ok] - want[ok]) / want[ok]
                bias = float(np.mean((got[ok] - want[ok]) / want[ok]))
                print("check %6d x %6d d=%d eps=%g %-14s max rel err %.3e  mean %.3e  bias %+.3e  (sum range %.3g..%.3g)"
                      % (n_out, n_in, args.d, eps, names[impl], err.max(), err.mean(), bias, want[ok].min(), want.max()),
                      flush=True)
    for shape in args.time:
        n_out, n_in = (int(v) for v in shape.split("x"))
        x0, x1, scale, po, pi = make_inputs(n_out, n_in, args.d, seed=6)
        for impl in impls:
            _, ms = run(ctx, torch, x0, x1, scale, po, pi, impl, 20)
            ent = n_out * n_in
            print("time  %6d x %6d d=%d %-22s %.1f us/pass  %.2f T entries/s  %.3f of MUFU peak"
                  % (n_out, n_in, args.d, names[impl], ms * 1e3, ent / ms / 1e9, ent / ms * 1e3 / peak.value), flush=True)


if __name__ == "__main__":
    main()
