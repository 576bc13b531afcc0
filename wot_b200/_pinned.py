"""Pool of page-locked host blocks for coupling outputs.

cudaHostAlloc costs ~0.3 ms per MiB (350 ms for the 1.2 GB coupling of a mean atlas pair, more than solving the
pair) and synchronises the device, so blocks are recycled: when the last ndarray viewing a block is
garbage-collected the block returns to the pool (compute_all_transport_maps drops each map after writing it).
Day-pairs differ in size by up to 16x, so a new block is never smaller than the largest one handed out so far:
after the first few maps every idle block fits every request and the pool stops allocating.
"""
from __future__ import annotations

import threading
import weakref

import numpy as np

from . import _lib

_free = []          # PinnedArray blocks nobody views
_lock = threading.Lock()
_MAX_IDLE = 4
_largest = 0        # bytes of the largest block allocated so far
_GRANULE = 64 << 20


class _PinnedNd(np.ndarray):
    """ndarray view that keeps its pinned block alive."""
    _block = None


def _give_back(block):
    with _lock:
        if len(_free) < _MAX_IDLE:
            _free.append(block)


def empty(shape, dtype):
    global _largest
    nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
    with _lock:
        best = None
        for k, blk in enumerate(_free):
            if blk.nbytes >= nbytes and (best is None or blk.nbytes < _free[best].nbytes):
                best = k
        block = _free.pop(best) if best is not None else None
        if block is None:
            want = max(nbytes, 1)
            if want >= _GRANULE:                      # large outputs: size for the largest request seen so far
                want = max(-(-want // _GRANULE) * _GRANULE, _largest)
                _largest = want
    if block is None:
        block = _lib.PinnedArray(want)
    arr = block.view(shape, dtype).view(_PinnedNd)
    arr._block = block
    weakref.finalize(arr, _give_back, block)
    return arr


def drain():
    global _largest
    with _lock:
        _free.clear()
        _largest = 0
