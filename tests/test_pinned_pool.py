"""Host-side lifetime logic of the page-locked buffers (harmonypy_b200._cabi: pinned_empty, _ResultPool), on CPU: the
CUDA allocator is replaced by libc's so that frees can be counted.  Without a GPU the real allocator refuses and both
fall back to plain NumPy arrays."""
import ctypes
import gc

import numpy as np

from harmonypy_b200 import _cabi


class _Libc:
    def __init__(self):
        self.lib = ctypes.CDLL(None)
        self.lib.malloc.restype = ctypes.c_void_p
        self.lib.malloc.argtypes = [ctypes.c_size_t]
        self.lib.free.argtypes = [ctypes.c_void_p]
        self.live, self.freed = set(), []

    def alloc(self, device, nbytes):
        p = self.lib.malloc(nbytes)
        self.live.add(p)
        return p

    def free(self, p):
        assert p in self.live, "double free or foreign pointer"
        self.live.discard(p)
        self.freed.append(p)
        self.lib.free(ctypes.c_void_p(p))


def test_without_gpu_the_allocator_refuses_and_callers_fall_back():
    import torch
    if torch.cuda.is_available():
        return
    assert _cabi._pinned_alloc(0, 1 << 20) is None
    a = _cabi.pinned_empty((1000, 7), np.float32)
    assert a.shape == (1000, 7) and a.dtype == np.float32 and a.flags.owndata
    b = _cabi._ResultPool().array((3000, 1000), np.float32)
    assert b.shape == (3000, 1000)


def test_pinned_empty_is_freed_with_its_last_view(monkeypatch):
    libc = _Libc()
    monkeypatch.setattr(_cabi, "_pinned_alloc", libc.alloc)
    monkeypatch.setattr(_cabi, "_pinned_free", libc.free)
    a = _cabi.pinned_empty((257, 13), np.float32, device=3)
    assert a.shape == (257, 13) and a.dtype == np.float32 and not a.flags.owndata and a.flags.c_contiguous
    a[...] = 1.5
    v = a[10:20].T                   # a view of a view keeps the memory
    del a
    gc.collect()
    assert not libc.freed and float(v.sum()) == 1.5 * 10 * 13
    del v
    gc.collect()
    assert len(libc.freed) == 1 and not libc.live


def test_pool_reuses_a_buffer_only_after_every_view_is_gone(monkeypatch):
    libc = _Libc()
    monkeypatch.setattr(_cabi, "_pinned_alloc", libc.alloc)
    monkeypatch.setattr(_cabi, "_pinned_free", libc.free)
    pool = _cabi._ResultPool()
    shape = (1 << 20, 4)             # 16 MB >= MIN_BYTES
    a = pool.array(shape, np.float32, device=1)
    pa = a.ctypes.data
    s = a[5:9]                       # a slice handed on by the caller
    del a
    b = pool.array(shape, np.float32, device=1)
    assert b.ctypes.data != pa       # the first buffer is still referenced through the slice
    del s
    c = pool.array(shape, np.float32, device=1)
    assert c.ctypes.data == pa       # free again: handed out
    pb = b.ctypes.data
    del b, c
    d = pool.array(shape, np.float32, device=1)
    assert d.ctypes.data in (pa, pb)
    assert len(libc.live) == 2 and not libc.freed          # pool buffers live as long as the process
    # beyond MAX_PER_SIZE live results the pool hands out plain arrays instead of growing
    keep = [pool.array(shape, np.float32) for _ in range(5)]
    assert len(libc.live) == pool.MAX_PER_SIZE
    assert sum(1 for k in keep if k.base is not None and not k.base.flags.owndata) <= pool.MAX_PER_SIZE
    small = pool.array((10, 10), np.float64)
    assert small.flags.owndata
