#!/usr/bin/env python
"""Host-only: bandwidth of b200lz4_gather_host (pageable arrays -> page-locked buffer) by thread count, next to a plain
numpy copy.  This is the staging cost in front of every batch call when the caller's arrays are not page-locked."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from streamly_lz4_b200 import _lib, datagen
import streamly_lz4_b200 as lz
lib = _lib.load()
total, block = 1 << 30, 640000
host = datagen.make("mixed", 2, total)
offs = np.arange(0, total, block, dtype=np.int64); lens = np.minimum(block, total - offs).astype(np.int32)
arrays = [host[o:o + l].copy() for o, l in zip(offs, lens)]
ptrs = np.array([a.ctypes.data for a in arrays], dtype=np.uint64)
ctx = lz.Context(0)
dst = ctx.pinned("g", total + 64)
pageable = np.empty(total, dtype=np.uint8); pageable[::4096] = 1
for name, d in (("pinned", dst), ("pageable", pageable)):
    for threads in (1, 2, 4, 8, 16, 32):
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            assert lib.b200lz4_gather_host(d.ctypes.data, ptrs.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(lens), threads) == 0
            dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        print(f"gather_host -> {name:8s} threads={threads:2d}: {total / best / 1e9:6.2f} GB/s", flush=True)
t0 = time.perf_counter(); dst[:total] = host; dt = time.perf_counter() - t0
print(f"numpy copy  -> pinned   threads= 1: {total / dt / 1e9:6.2f} GB/s")
print("cpus", os.cpu_count())
