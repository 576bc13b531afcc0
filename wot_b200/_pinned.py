"""Pool of page-locked host blocks for coupling outputs.

cudaHostAlloc costs ~0.3 ms per MiB (350 ms for the 1.2 GB coupling of a mean atlas pair, more than solving the
pair) and synchronises the device, so blocks are recycled.  Every array handed out views a fresh `_Lease`
object; NumPy keeps the object an array was made from as the `.base` of that array and of EVERY view derived
from it (np.asarray(tmap), tmap.view(np.ndarray), slices, pd.DataFrame(tmap) ...), so the lease lives exactly
as long as somebody can still read the memory, and its finalizer returns the block to the pool
(compute_all_transport_maps drops each map after writing it).
Day-pairs differ in size by up to 16x, so a new block is never smaller than the largest one handed out so far:
after the first few maps every idle block fits every request and the pool stops allocating.
"""
from __future__ import annotations

import threading
import weakref

import numpy as np

from . import _lib

_free = []          # blocks nobody views
_lock = threading.Lock()
_MAX_IDLE = 4
_largest = 0        # bytes of the largest block allocated so far
_GRANULE = 64 << 20
_alloc = None       # block factory: nbytes -> object with .buf (ctypes char array) and .nbytes; tests replace it


class _Lease:
    """One loan of a pinned block.  Exposes the block's memory through the buffer protocol (PEP 688) and is
    the ultimate `.base` of every ndarray viewing it."""

    def __init__(self, block):
        self._block = block

    def __buffer__(self, flags):
        return memoryview(self._block.buf)


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
        block = (_alloc or _lib.PinnedArray)(want)
    lease = _Lease(block)
    weakref.finalize(lease, _give_back, block)
    count = int(np.prod(shape))
    return np.frombuffer(lease, dtype=dtype, count=count).reshape(shape)


def drain():
    global _largest
    with _lock:
        _free.clear()
        _largest = 0
