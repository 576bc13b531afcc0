#!/usr/bin/env python
"""Isolated timing of the stored-K iteration kernels (row, column, fused) for a list of shapes.
Usage: python tools/matvec_bench.py 12486x12405 20000x20000 ..."""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from wot_b200 import _lib  # noqa: E402


def main():
    shapes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(12486, 12405)]
    ctx = _lib.context(0)
    for I, J in shapes:
        r, c, f = C.c_double(), C.c_double(), C.c_double()
        _lib.check(ctx.lib.wotb_bench_matvec_dev(ctx.handle, I, J, 20, C.byref(r), C.byref(c), C.byref(f)))
        ld = (J + 31) // 32 * 32
        gb = I * ld * 4 / 1e9
        print("%6d x %6d  K %.3f GB | row %.1f us %.0f GB/s | col %.1f us %.0f GB/s | fused %.1f us %.0f GB/s"
              % (I, J, gb, r.value * 1e3, gb / r.value * 1e3, c.value * 1e3, gb / c.value * 1e3, f.value * 1e3,
                 gb / f.value * 1e3 if f.value > 0 else 0), flush=True)


if __name__ == "__main__":
    main()
