#!/usr/bin/env python
"""How does tcgen05.mma (kind::f16, fp32 accumulate in TMEM) round?  Probe through wotb_online_rowsums_dev.

One in-side row, many out-side rows: sums[i] = exp2(D_i) with D_i = <X_i, Y> + P_i + Q, so log2(sums[i]) - D_i is the
error of the accumulated exponent.  Cases:
  grid   coordinates are multiples of 2^-q (hi part exact, lo part zero), offsets multiples of 2^-2q chosen so that
         D_i is a small integer: every addend is a multiple of one quantum -> exact if the hardware keeps >= 24 bits
  float  coordinates with full fp16x2 (22-bit) content, same magnitudes: shows the rounding mode (bias) and size
Usage: python tools/tc_accum_probe.py
"""
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import _lib  # noqa: E402


def run(ctx, torch, x_out, x_in, off_out, off_in, impl=2):
    dev = "cuda:%d" % ctx.device
    t = [torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev) for a in (x_out, x_in, off_out, off_in)]
    sums = torch.empty(x_out.shape[0], dtype=torch.float64, device=dev)
    P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.wotb_online_rowsums_dev(ctx.handle, P(t[0]), x_out.shape[0], P(t[1]), x_in.shape[0], x_out.shape[1],
                                               1.0, P(t[2]), P(t[3]), impl, 0, P(sums), None))
    return sums.cpu().numpy()


def case(ctx, torch, rng, n, d, amp, q, grid, dims_active=None, impl=2, by_target=False):
    scale = 2.0 ** q
    x = rng.uniform(-amp, amp, size=(n, d))
    y = rng.uniform(-amp, amp, size=(1, d))
    if dims_active is not None:
        x[:, dims_active:] = 0
        y[:, dims_active:] = 0
    if grid:
        x = np.rint(x * scale) / scale
        y = np.rint(y * scale) / scale
    dot = (x * y).sum(1)                       # exact in float64 for grid inputs (multiples of 2^-2q, < 2^30)
    target = -rng.integers(1, 30, size=n).astype(np.float64) if by_target else -rng.integers(1, 12, size=n).astype(np.float64)
    if by_target and not grid:
        target = target + rng.uniform(-0.5, 0.5, size=n)
    q_off = -np.floor(0.5 * (y * y).sum())     # integer
    p_off = target - dot - q_off
    if grid:
        # make P a multiple of 2^-2q (it is, since dot is) and check representability by the two fp16 slots
        assert np.all(np.abs(p_off * 4.0 ** q - np.rint(p_off * 4.0 ** q)) < 1e-9)
    got = run(ctx, torch, x, y, p_off, np.array([q_off]), impl=impl)
    err = np.log2(got) - target
    part = np.abs(dot).max()
    print("%-5s q=%d amp=%5.1f d_active=%2s  |<X,Y>|max %7.1f |P|max %7.1f  err: mean %+.3e  rms %.3e  max %.3e  exact rows %d/%d"
          % ("grid" if grid else "float", q, amp, dims_active or d, part, np.abs(p_off).max(), err.mean(),
             np.sqrt((err ** 2).mean()), np.abs(err).max(), int((err == 0).sum()), n), flush=True)
    if by_target:
        for lo, hi in ((0, 2), (2, 4), (4, 8), (8, 16), (16, 32)):
            sel = (-target >= lo) & (-target < hi)
            if sel.any():
                print("        impl %d  |D| in [%2d,%2d): n %5d  mean %+.3e  rms %.3e" % (impl, lo, hi, int(sel.sum()), err[sel].mean(),
                                                                                 np.sqrt((err[sel] ** 2).mean())), flush=True)


def main():
    import torch
    ctx = _lib.context(0)
    rng = np.random.default_rng(0)
    n, d = 4096, 30
    for amp, q in ((3.0, 6), (6.0, 6), (12.0, 6), (12.0, 5), (24.0, 5), (24.0, 4), (3.0, 8), (6.0, 7)):
        case(ctx, torch, rng, n, d, amp, q, True)
    for amp in (3.0, 6.0, 12.0, 24.0):
        case(ctx, torch, rng, n, d, amp, 0, False)
    # error of the exponent by its magnitude, default vs precise operands, at magnitudes of eps = 0.05 / 0.01 / 0.005
    for amp in (3.0, 6.0, 9.0):
        for impl in (2, 3):
            case(ctx, torch, rng, 16384, d, amp, 0, False, impl=impl, by_target=True)
    # few active dimensions: products large relative to the sum
    for amp, q in ((12.0, 6), (24.0, 5)):
        case(ctx, torch, rng, n, d, amp, q, True, dims_active=4)
        case(ctx, torch, rng, n, d, amp, q, False, dims_active=4)


if __name__ == "__main__":
    main()
