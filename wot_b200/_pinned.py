"""Pool of page-locked host blocks for coupling outputs.

cudaHostAlloc costs ~0.3 ms per MiB, comparable to solving the day-pair whose coupling it holds, so blocks
are recycled: when the last ndarray viewing a block is garbage-collected the block returns to the pool and
the next transport map of similar size reuses it (compute_all_transport_maps drops each map after writing it).
"""
from __future__ import annotations

import weakref

import numpy as np

from . import _lib

_free = []          # PinnedArray blocks nobody views
_MAX_IDLE = 3


class _PinnedNd(np.ndarray):
    """ndarray view that keeps its pinned block alive."""
    _block = None


def _give_back(block):
    if len(_free) < _MAX_IDLE:
        _free.append(block)


def empty(shape, dtype):
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    best = None
    for k, blk in enumerate(_free):
        if blk.nbytes >= nbytes and blk.nbytes <= 2 * nbytes + (1 << 20) and (best is None or blk.nbytes < _free[best].nbytes):
            best = k
    block = _free.pop(best) if best is not None else _lib.PinnedArray(max(nbytes, 1))
    arr = block.view(shape, dtype).view(_PinnedNd)
    arr._block = block
    weakref.finalize(arr, _give_back, block)
    return arr


def drain():
    _free.clear()
