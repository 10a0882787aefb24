#!/usr/bin/env python
"""Small end-to-end pass (also the workload to put under compute-sanitizer where that tool is open): every path of both kernels on small inputs,
checked against the oracle.  usage: python tools/small_e2e_check.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import streamly_lz4_b200 as lz
from streamly_lz4_b200 import datagen
from oracle.oracle import Oracle

ora = Oracle("auto")
ctx = lz.Context(0)
n_ok = 0
for kind in ("mixed", "text", "sparse01", "records", "random", "zero", "biased01"):
    data = datagen.make(kind, 9, 300000 + 777)
    for bs, accel, linked in ((65536, 1, True), (100000, 400, False), (4096, 5, True), (300777, 1, False), (13, 1, True)):
        arrays = [data[i:i + bs].tobytes() for i in range(0, min(data.size, 40 * bs), bs)]
        cfg = lz.BlockConfig(independent=not linked)
        got = list(lz.compress_chunks(cfg, accel, arrays, ctx=ctx))
        want = ora.compress_chunks(arrays, accel, linked=linked)
        assert got == want, (kind, bs, accel, linked)
        back = list(lz.decompress_chunks_raw(cfg, want, ctx=ctx))
        assert back == arrays, (kind, bs, accel, linked)
        n_ok += 1
# BlockMax configs (4-byte headers, capacity = block maximum) and malformed payloads
data = datagen.make("text", 3, 200000)
arrays = [data[i:i + 50000].tobytes() for i in range(0, data.size, 50000)]
cfg = lz.BlockConfig(block_size=lz.BlockSize.BlockMax64KB)
framed = list(lz.compress_chunks(cfg, 1, arrays, ctx=ctx))
assert b"".join(lz.decompress_chunks_raw(cfg, framed, ctx=ctx)) == data.tobytes()
rng = np.random.default_rng(1)
good = ora.compress_chunks([arrays[0]], 1, linked=False)[0]
for _ in range(30):
    b = bytearray(good[8:]); b[int(rng.integers(0, len(b)))] ^= 1 << int(rng.integers(0, 8))
    arr = len(b).to_bytes(4, "little") + len(arrays[0]).to_bytes(4, "little") + bytes(b)
    try:
        list(lz.decompress_chunks_raw(lz.BlockConfig(independent=True), [arr], ctx=ctx))
    except lz.LZ4Error:
        pass
ctx.close()
print("sanitize_small ok:", n_ok, "cases")
