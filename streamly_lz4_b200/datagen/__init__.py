"""Seeded synthetic corpora (ctypes binding of datagen.c).

Replaces the reference's downloaded Canterbury corpora (download-corpora.sh:6-49,
benchmark/Main.hs:70-103) and restates its QuickCheck generators
(test/Main.hs:33-55) as seeded byte generators.
"""
from __future__ import annotations

import ctypes
import os
from concurrent.futures import ThreadPoolExecutor

import numpy as np

TEXT, RANDOM, SPARSE01, RECORDS, MIXED, BITS01, BIASED01, ZERO = range(8)
KINDS = {"text": TEXT, "random": RANDOM, "sparse01": SPARSE01, "records": RECORDS,
         "mixed": MIXED, "bits01": BITS01, "biased01": BIASED01, "zero": ZERO}

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def _lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libb200gen.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make` (or __graft_entry__.build()) first")
        lib = ctypes.CDLL(path)
        lib.b200gen_fill.argtypes = [ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_void_p]
        lib.b200gen_fill.restype = None
        _LIB = lib
    return _LIB


def fill(out: np.ndarray, kind, seed: int, offset: int = 0, threads: int = 0) -> np.ndarray:
    """Fill the uint8 array `out` with bytes [offset, offset+len) of corpus (kind, seed)."""
    if isinstance(kind, str):
        kind = KINDS[kind]
    assert out.dtype == np.uint8 and out.flags["C_CONTIGUOUS"]
    lib = _lib()
    n = out.size
    base = out.ctypes.data
    threads = threads or min(os.cpu_count() or 1, 16)
    piece = 1 << 24
    if n <= piece or threads == 1:
        lib.b200gen_fill(kind, seed, offset, n, base)
        return out
    jobs = [(s, min(piece, n - s)) for s in range(0, n, piece)]
    with ThreadPoolExecutor(threads) as ex:
        list(ex.map(lambda j: lib.b200gen_fill(kind, seed, offset + j[0], j[1], base + j[0]), jobs))
    return out


def make(kind, seed: int, n: int, offset: int = 0) -> np.ndarray:
    return fill(np.empty(n, dtype=np.uint8), kind, seed, offset)
