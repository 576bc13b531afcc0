#!/usr/bin/env python
"""Do the two roofs add up?  (1) stored fused iteration (HBM-bound) and online pass (MUFU-bound) as a function of the
number of SMs their grids are sized for; (2) both at once on complementary SM sets, two streams.
Usage: python tools/sm_split_probe.py"""
import ctypes as C
import os
import sys
import threading
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import _lib, synthetic  # noqa: E402
from tools.online_pass_check import make_inputs  # noqa: E402

I, J, D = 12486, 12405, 30


def stored_ms(ctx, reps=40):
    a, b, c = C.c_double(), C.c_double(), C.c_double()
    _lib.check(ctx.lib.wotb_bench_matvec_dev(ctx.handle, I, J, reps, C.byref(a), C.byref(b), C.byref(c)))
    return c.value


def online_ms(ctx, torch, t, scale, reps=40):
    sums = torch.empty(I, dtype=torch.float64, device="cuda:0")
    ms = C.c_double()
    P = lambda v: C.c_void_p(v.data_ptr())  # noqa: E731
    _lib.check(ctx.lib.wotb_online_rowsums_dev(ctx.handle, P(t[0]), I, P(t[1]), J, D, float(scale), P(t[2]), P(t[3]), 2, reps,
                                               P(sums), C.byref(ms)))
    return ms.value


def main():
    import torch
    x0, x1, scale, po, pi = make_inputs(I, J, D, seed=6)
    t = [torch.from_numpy(np.ascontiguousarray(a)).to("cuda:0") for a in (x0, x1, po, pi)]
    ca, cb = _lib.Context(0), _lib.Context(0)
    ld = (J + 31) // 32 * 32
    for n in (148, 112, 100, 74, 56, 48, 40, 32):
        ca.lib.wotb_set_sm_limit(ca.handle, n)
        ms_s = stored_ms(ca)
        ms_o = online_ms(ca, torch, t, scale)
        print("SMs %3d  stored fused iteration %.1f us = %.0f GB/s   online pass %.1f us = %.2f T exp/s"
              % (n, ms_s * 1e3, I * ld * 4 / ms_s / 1e6, ms_o * 1e3, I * J / ms_o / 1e9), flush=True)
    for n_s in (32, 40, 48, 56):
        ca.lib.wotb_set_sm_limit(ca.handle, n_s)
        cb.lib.wotb_set_sm_limit(cb.handle, 148 - n_s)
        ca.lib.wotb_set_pdl(0)
        res = {}

        def run_s():
            res["s"] = stored_ms(ca, reps=200)

        def run_o():
            res["o"] = online_ms(cb, torch, t, scale, reps=300)
        th = [threading.Thread(target=run_s), threading.Thread(target=run_o)]
        t0 = time.perf_counter()
        for x in th:
            x.start()
        for x in th:
            x.join()
        print("together: stored on %d SMs %.1f us/iteration (%.0f GB/s), online on %d SMs %.1f us/pass (%.2f T exp/s)  [wall %.3f s]"
              % (n_s, res["s"] * 1e3, I * ld * 4 / res["s"] / 1e6, 148 - n_s, res["o"] * 1e3, I * J / res["o"] / 1e9,
                 time.perf_counter() - t0), flush=True)


if __name__ == "__main__":
    main()
