#!/usr/bin/env python
"""BASELINE.json configs 3-5 at (per-GPU) full size on one B200: kernel-resident and end-to-end timings plus parity
samples against the oracle.  One JSON line per config on stdout.  (configs[1] is bench.py; configs[0] is the CPU case.)

  config 3  d+640000 of an 8 GiB pre-compressed stream, 1 GiB stripe (= one of 8 GPUs), through re-frame + decode
  config 4  128 linked streams x 64 MiB (= one of 8 GPUs), 64 KiB blocks, accel 1, previous block as dictionary
  config 5  re-frame of a stream with 4 KiB-4 MiB blocks at the reference's read sizes
"""
import argparse, ctypes, hashlib, json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import _lib, datagen
    from oracle.oracle import Oracle, resize_chunks
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", type=int, default=128)
    ap.add_argument("--mib-per-stream", type=int, default=64)
    ap.add_argument("--c3-mib", type=int, default=1024)
    args = ap.parse_args()
    ora = Oracle("auto")
    ctx = lz.Context(0)
    lib = _lib.load()
    threads = min(os.cpu_count() or 1, 64)

    # ---------------------------------------------------------------- config 3
    total, bs = args.c3_mib << 20, 640000
    data = datagen.make("mixed", 3, total)
    offs = np.arange(0, total, bs, dtype=np.int64); lens = np.minimum(bs, total - offs).astype(np.int32); n = len(lens)
    ptrs = (data.ctypes.data + offs).astype(np.uint64)
    caps = (lens.astype(np.int64) + lens // 255 + 24).astype(np.int32)
    arena, dptrs, doffs = ora.slots(caps)
    out_len = np.zeros(n, dtype=np.int32)
    t0 = time.perf_counter()
    assert ora.compress_ptrs(ptrs, lens, dptrs, caps, out_len, 1, 8, np.arange(n + 1, dtype=np.int32), 0, threads) == 0
    t_cpu_c = time.perf_counter() - t0
    stream = np.concatenate([arena[o:o + 8 + l] for o, l in zip(doffs[:-1], out_len)])
    pin = ctx.pinned("c3_in", stream.size); pin[:stream.size] = stream
    back = ctx.pinned("c3_out", total + 64)
    boff = np.zeros(n + 8, dtype=np.int64); blen = np.zeros(n + 8, dtype=np.int32)
    found, used, ended = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    best = None
    for _ in range(3):
        t0 = time.perf_counter()
        rc = lib.b200lz4_reframe(pin.ctypes.data, stream.size, 8, 0, boff.ctypes.data, blen.ctypes.data, len(boff),
                                 ctypes.byref(found), ctypes.byref(used), ctypes.byref(ended))
        t_reframe = time.perf_counter() - t0
        assert rc == 0 and found.value == n and used.value == stream.size
        rc, doff, dlen = ctx.decompress_batch(pin[:stream.size], boff[:n].copy(), blen[:n].copy(), 8, 0, back)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    tm = ctx.timing()
    ok = hashlib.sha256(back[:total].tobytes()).digest() == hashlib.sha256(data.tobytes()).digest()
    # CPU reference decode of the same stream, all threads
    dcaps = lens.copy(); darena, ddptrs, ddoffs = ora.slots(dcaps)
    fptrs = (arena.ctypes.data + doffs[:-1]).astype(np.uint64); flens = (out_len + 8).astype(np.int32)
    dout = np.zeros(n, dtype=np.int32)
    t0 = time.perf_counter()
    assert ora.decompress_ptrs(fptrs, flens, ddptrs, dcaps, dout, 8, np.arange(n + 1, dtype=np.int32), 0, threads) == 0
    t_cpu_d = time.perf_counter() - t0
    print(json.dumps({"config": 3, "what": f"d+640000: {args.c3_mib} MiB stripe of the pre-compressed stream ({n} blocks, accel-1 oracle output, ratio {total / stream.size:.2f}): host re-frame + b200lz4_decompress_batch (pinned host in/out)",
                      "e2e_gbps": total / best / 1e9, "reframe_ms": 1e3 * t_reframe, "h2d_ms": tm["h2d_ms"], "kernel_ms": tm["kernel_ms"], "d2h_ms": tm["d2h_ms"],
                      "identical": bool(ok), "cpu_reference_decode_gbps": total / t_cpu_d / 1e9, "cpu_reference_compress_gbps": total / t_cpu_c / 1e9, "cpu_threads": threads}), flush=True)
    del data, stream, arena, darena

    # ---------------------------------------------------------------- config 4
    ns, per, bs = args.streams, args.mib_per_stream << 20, 65536
    total = ns * per
    data = datagen.make("mixed", 4, total)
    offs = np.arange(0, total, bs, dtype=np.int64); lens = np.full(len(offs), bs, dtype=np.int32); n = len(lens)
    bps = per // bs
    sf = (np.arange(ns + 1, dtype=np.int64) * bps).astype(np.int32)
    dst = ctx.pinned("c4_dst", int((lens.astype(np.int64) + lens // 255 + 24).sum()))
    psrc = ctx.pinned("c4_src", total)                     # page-locked input, as the Haskell shim stages it
    psrc[:total] = data
    rc, doff, olen = ctx.compress_batch(psrc[:total], offs, lens, 1, 8, dst, stream_first=sf)      # warm-up (allocations)
    t0 = time.perf_counter()
    rc, doff, olen = ctx.compress_batch(psrc[:total], offs, lens, 1, 8, dst, stream_first=sf)
    t_c = time.perf_counter() - t0
    assert rc == 0
    tmc = ctx.timing()
    # parity: 4 whole streams against the oracle
    check = [0, ns // 3, (2 * ns) // 3, ns - 1]
    same = True
    for s in check:
        arrays = [data[o:o + bs] for o in offs[sf[s]:sf[s + 1]]]
        want = ora.compress_chunks([a.tobytes() for a in arrays], 1, linked=True)
        for k, w in enumerate(want):
            b = sf[s] + k
            same &= dst[doff[b]:doff[b + 1]].tobytes() == w
    back = ctx.pinned("c4_back", total + 64)
    comp = dst[:doff[-1]]
    rc, boff2, blen2 = ctx.decompress_batch(comp, doff[:-1].copy(), np.diff(doff).astype(np.int32), 8, 0, back, stream_first=sf)
    t0 = time.perf_counter()
    rc, boff2, blen2 = ctx.decompress_batch(comp, doff[:-1].copy(), np.diff(doff).astype(np.int32), 8, 0, back, stream_first=sf)
    t_d = time.perf_counter() - t0
    assert rc == 0
    tmd = ctx.timing()
    rt = hashlib.sha256(back[:total].tobytes()).digest() == hashlib.sha256(data.tobytes()).digest()
    print(json.dumps({"config": 4, "what": f"{ns} linked streams x {args.mib_per_stream} MiB, 64 KiB blocks, accel 1 (one GPU's share of 1024 streams on 8 GPUs)",
                      "compress_e2e_gbps": total / t_c / 1e9, "compress_kernel_ms": tmc["kernel_ms"], "compress_kernel_gbps": total / tmc["kernel_ms"] / 1e6,
                      "decompress_e2e_gbps": total / t_d / 1e9, "decompress_kernel_ms": tmd["kernel_ms"], "decompress_kernel_gbps": total / tmd["kernel_ms"] / 1e6,
                      "ratio": total / int(doff[-1]), "streams_byte_identical_to_oracle": {"checked": check, "identical": bool(same)}, "round_trip_identical": bool(rt)}), flush=True)
    del data

    # ---------------------------------------------------------------- config 5
    rng = np.random.default_rng(5)
    sizes = np.exp(rng.uniform(np.log(4096), np.log(4 << 20), 400)).astype(np.int64)
    total = int(sizes.sum())
    data = datagen.make("mixed", 5, total)
    offs = np.zeros(len(sizes), dtype=np.int64); offs[1:] = np.cumsum(sizes[:-1])
    lens = sizes.astype(np.int32)
    dst = ctx.pinned("c5_dst", int((sizes + sizes // 255 + 24).sum()))
    rc, doff, olen = ctx.compress_batch(data, offs, lens, 1, 8, dst)
    assert rc == 0
    blob = dst[:doff[-1]].tobytes()
    res = {}
    for bufsize in (512, 6553, 65536, 655360, 640000):
        chunks = [blob[i:i + bufsize] for i in range(0, len(blob), bufsize)]
        t0 = time.perf_counter()
        got = list(lz.resize_chunks(lz.BlockConfig(), lz.default_frame_config, chunks))
        dt = time.perf_counter() - t0
        ok = [len(g) for g in got] == [int(x) for x in np.diff(doff)] and b"".join(got) == blob
        if bufsize in (6553, 655360):
            ok &= got == resize_chunks(chunks)
        res[str(bufsize)] = {"seconds": dt, "blocks": len(got), "bit_exact": bool(ok)}
    print(json.dumps({"config": 5, "what": f"re-frame of {len(sizes)} blocks of 4 KiB-4 MiB plaintext ({len(blob) >> 20} MiB compressed) at the reference's read sizes (Python mirror over b200lz4_reframe)", "by_read_size": res}), flush=True)
    ctx.close()


if __name__ == "__main__":
    main()
