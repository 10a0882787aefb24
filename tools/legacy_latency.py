#!/usr/bin/env python
"""Per-call latency of the legacy single-block symbols (LZ4_compress_fast_continue / LZ4_decompress_safe_continue as the
unmodified reference would call them): one 64 KiB block per call, linked stream."""
import ctypes, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from streamly_lz4_b200 import _lib, datagen
lib = _lib.load()
d = datagen.make("text", 5, 256 * 65536)
cs, ds = lib.LZ4_createStream(), lib.LZ4_createStreamDecode()
bound = lib.LZ4_compressBound(65536)
dst = ctypes.create_string_buffer(bound)
outs = [ctypes.create_string_buffer(65536) for _ in range(2)]
tc, td = [], []
for i in range(256):
    a = d[i * 65536:(i + 1) * 65536].tobytes()
    t = time.perf_counter(); n = lib.LZ4_compress_fast_continue(cs, a, dst, 65536, bound, 1); tc.append(time.perf_counter() - t)
    assert n > 0
    t = time.perf_counter(); m = lib.LZ4_decompress_safe_continue(ds, dst.raw[:n], outs[i & 1], n, 65536); td.append(time.perf_counter() - t)
    assert m == 65536 and outs[i & 1].raw == a
tc, td = np.array(tc[16:]) * 1e6, np.array(td[16:]) * 1e6
print(f"LZ4_compress_fast_continue  64 KiB text block: median {np.median(tc):.0f} us  (p10 {np.percentile(tc,10):.0f}, p90 {np.percentile(tc,90):.0f})  -> {65536/np.median(tc):.1f} MB/s")
print(f"LZ4_decompress_safe_continue 64 KiB text block: median {np.median(td):.0f} us  (p10 {np.percentile(td,10):.0f}, p90 {np.percentile(td,90):.0f})  -> {65536/np.median(td):.1f} MB/s")
lib.LZ4_freeStream(cs); lib.LZ4_freeStreamDecode(ds)
