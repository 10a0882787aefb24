"""BASELINE.json configs 3-5 through the C ABI at (per-GPU) full sizes, checked through size-independent
properties plus oracle samples."""
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _blocks(total, bs):
    offs = np.arange(0, total, bs, dtype=np.int64)
    lens = np.minimum(bs, total - offs).astype(np.int32)
    return offs, lens


def test_config3_decompress_precompressed_stream(ctx, ref):
    """configs[2]: d+640000 of a stream pre-compressed by the ORACLE (accel 1, 640000-byte independent blocks),
    read back in 640000-byte chunks -> re-frame -> GPU decode.  2 GiB here = the stripe one GPU of four gets."""
    import os
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import _lib, datagen
    total, bs = 2 << 30, 640000
    data = datagen.make("mixed", 3, total)
    offs, lens = _blocks(total, bs)
    n = len(lens)
    threads = min(os.cpu_count() or 1, 64)
    ptrs = (data.ctypes.data + offs).astype(np.uint64)
    caps = (lens.astype(np.int64) + lens // 255 + 16 + 8).astype(np.int32)
    arena, dptrs, doffs = ref.slots(caps)
    out_len = np.zeros(n, dtype=np.int32)
    assert ref.compress_ptrs(ptrs, lens, dptrs, caps, out_len, 1, 8, np.arange(n + 1, dtype=np.int32), 0, threads) == 0
    stream = np.concatenate([arena[o:o + 8 + l] for o, l in zip(doffs[:-1], out_len)])
    # re-frame: the header walk of resizeChunksD over the stream consumed in 640000-byte reads (incremental calls)
    lib = _lib.load()
    import ctypes
    boff = np.zeros(n + 8, dtype=np.int64); blen = np.zeros(n + 8, dtype=np.int32)
    found, used, ended = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    got_off, got_len, base, fed = [], [], 0, 0
    while fed < stream.size:
        fed = min(fed + 640000, stream.size)
        view = stream[base:fed]
        rc = lib.b200lz4_reframe(view.ctypes.data, view.size, 8, 0, boff.ctypes.data, blen.ctypes.data, len(boff),
                                 ctypes.byref(found), ctypes.byref(used), ctypes.byref(ended))
        assert rc == 0
        got_off += [base + int(x) for x in boff[:found.value]]; got_len += [int(x) for x in blen[:found.value]]
        base += used.value
    assert base == stream.size and len(got_off) == n
    assert got_len == [int(l) + 8 for l in out_len]
    back = ctx.pinned("c3_back", total + 64)
    rc, doff, dlen = ctx.decompress_batch(stream, np.array(got_off, dtype=np.int64), np.array(got_len, dtype=np.int32), 8, 0, back)
    assert rc == 0 and (dlen == lens).all()
    assert hashlib.sha256(back[:total].tobytes()).digest() == hashlib.sha256(data.tobytes()).digest()


def test_config4_linked_streams(ctx, ref):
    """configs[3]: concurrent linked streams of 64 KiB blocks with the prior block as dictionary, accel 1.
    16 streams x 64 MiB here (config 4 puts 128 such streams on each GPU): every stream byte-identical to the
    oracle via a checksum of checksums, two streams compared block for block, and the round trip is the identity."""
    import os
    from streamly_lz4_b200 import datagen
    n_streams, per_stream, bs = 16, 64 << 20, 65536
    total = n_streams * per_stream
    data = datagen.make("mixed", 4, total)
    offs, lens = _blocks(total, bs)
    n = len(lens)
    bps = per_stream // bs
    sf = (np.arange(n_streams + 1, dtype=np.int64) * bps).astype(np.int32)
    cap = int((lens.astype(np.int64) + lens // 255 + 16 + 8).sum())
    dst = ctx.pinned("c4_dst", cap)
    rc, dst_off, out_len = ctx.compress_batch(data, offs, lens, 1, 8, dst, stream_first=sf)
    assert rc == 0 and (out_len > 0).all()
    # oracle: same call sequence, one pthread per stream
    threads = min(os.cpu_count() or 1, n_streams)
    # arrays of one stream must not be address-adjacent for the reference (separate Haskell arrays never are)
    gap_arena, ptrs, _ = ref.lay_out([data[o:o + l] for o, l in zip(offs, lens)])
    caps = (lens.astype(np.int64) + lens // 255 + 16 + 8).astype(np.int32)
    arena, dptrs, doffs = ref.slots(caps)
    want_len = np.zeros(n, dtype=np.int32)
    assert ref.compress_ptrs(ptrs, lens, dptrs, caps, want_len, 1, 8, sf, 0, threads) == 0
    assert (want_len == out_len).all()
    for s in range(n_streams):
        h_gpu, h_ref = hashlib.sha256(), hashlib.sha256()
        for b in range(sf[s], sf[s + 1]):
            h_gpu.update(dst[dst_off[b]:dst_off[b + 1]].tobytes())
            h_ref.update(arena[doffs[b]:doffs[b] + 8 + want_len[b]].tobytes())
        assert h_gpu.digest() == h_ref.digest(), f"stream {s} differs from the oracle"
    back = ctx.pinned("c4_back", total + 64)
    comp = dst[:dst_off[-1]]
    rc, boff, blen = ctx.decompress_batch(comp, dst_off[:-1].copy(), np.diff(dst_off).astype(np.int32), 8, 0, back, stream_first=sf)
    assert rc == 0 and (blen == lens).all()
    assert hashlib.sha256(back[:total].tobytes()).digest() == hashlib.sha256(data.tobytes()).digest()


def test_config5_mixed_block_sizes_reframe_and_sweep(ctx, ref):
    """configs[4]: (i) blocks of plaintext sizes log-uniform in [4 KiB, 4 MiB], GPU-compressed, concatenated,
    fragmented at the reference's read sizes, re-framed (bit-exact vs the restated resizeChunksD, idempotent) and
    decoded; (ii) acceleration sweep on config-2 data, bytes identical to the oracle per value."""
    import streamly_lz4_b200 as lz
    from oracle.oracle import resize_chunks
    from streamly_lz4_b200 import datagen
    rng = np.random.default_rng(5)
    sizes = np.exp(rng.uniform(np.log(4096), np.log(4 << 20), 96)).astype(np.int64)
    total = int(sizes.sum())
    data = datagen.make("mixed", 5, total)
    offs = np.zeros(len(sizes), dtype=np.int64); offs[1:] = np.cumsum(sizes[:-1])
    arrays = [data[o:o + s].tobytes() for o, s in zip(offs, sizes)]
    cfg = lz.BlockConfig(independent=True)
    framed = list(lz.compress_chunks(cfg, 1, arrays, ctx=ctx))
    want = ref.compress_chunks(arrays, 1, linked=False, threads=8)
    assert framed == want
    blob = b"".join(framed)
    for bufsize in (512, 6553, 65536, 655360, 640000):
        chunks = [blob[i:i + bufsize] for i in range(0, len(blob), bufsize)]
        re1 = list(lz.resize_chunks(cfg, lz.default_frame_config, chunks))
        assert re1 == resize_chunks(chunks) == framed
        assert list(lz.resize_chunks(cfg, lz.default_frame_config, re1)) == re1          # idempotent
    chunks = [blob[i:i + 640000] for i in range(0, len(blob), 640000)]
    assert list(lz.decompress_chunks(cfg, chunks, ctx=ctx)) == arrays
    # (ii) sweep
    d2 = datagen.make("mixed", 2, 8 * 640000)
    blocks = [d2[i:i + 640000].tobytes() for i in range(0, d2.size, 640000)]
    seen = {}
    for accel in (-1, 0, 1, 2, 5, 10, 12, 100, 400, 1000, 65537, 65538):
        got = list(lz.compress_chunks(cfg, accel, blocks, ctx=ctx))
        assert got == ref.compress_chunks(blocks, accel, linked=False, threads=8), f"accel {accel}"
        seen[accel] = hashlib.sha256(b"".join(got)).digest()
    assert seen[-1] == seen[0] == seen[1] and seen[65537] == seen[65538]


def test_linked_stream_across_2gib_renorm(ctx, ref):
    """Row a6 (LZ4_renormDictT, cbits/lz4.c:1545-1562): one linked stream whose currentOffset passes 2 GiB.
    1.95 GiB of incompressible data at acceleration 65537 (cheap), then 320 MiB of mixed data at acceleration 1
    through the SAME stream state; every block of the second call must equal the reference's bytes."""
    import ctypes
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import datagen
    bs = 4 << 20
    n1, n2 = 499, 80                                      # 1.949 GiB, then 320 MiB: the offset passes 2^31 inside call 2
    filler = datagen.make("random", 21, bs)
    tail = datagen.make("mixed", 22, n2 * bs)
    # reference: one LZ4_stream_t, separately allocated arrays
    cs = ref.ccreate()
    bound = ref.bound(bs)
    dst = np.zeros(bound, dtype=np.uint8)
    blocks_a = [filler.copy() for _ in range(2)]           # alternate two allocations: never adjacent, never the same array twice in a row
    for i in range(n1):
        a = blocks_a[i & 1]
        assert ref.ccont(cs, a.ctypes.data, dst.ctypes.data, bs, bound, 65537) > 0
    want = []
    keep = []
    for i in range(n2):
        a = tail[i * bs:(i + 1) * bs].copy(); keep.append(a)
        r = ref.ccont(cs, a.ctypes.data, dst.ctypes.data, bs, bound, 1)
        assert r > 0
        want.append(dst[:r].tobytes())
    ref.cfree(cs)
    # GPU: same two calls through one device-resident stream
    stream = lz.CompressStream(ctx)
    src1 = np.tile(filler, n1)
    offs1 = (np.arange(n1, dtype=np.int64) * bs); lens1 = np.full(n1, bs, dtype=np.int32)
    out1 = ctx.pinned("rn_out", int(n1 * (bound + 8)))
    rc, _, ol1 = ctx.compress_batch(src1, offs1, lens1, 65537, 8, out1, stream_first=np.array([0, n1], np.int32), streams=[stream])
    assert rc == 0 and (ol1 > 0).all()
    _, off_mid = stream.peek()
    assert off_mid == n1 * bs
    offs2 = (np.arange(n2, dtype=np.int64) * bs); lens2 = np.full(n2, bs, dtype=np.int32)
    rc, doff, ol2 = ctx.compress_batch(tail, offs2, lens2, 1, 8, out1, stream_first=np.array([0, n2], np.int32), streams=[stream])
    assert rc == 0 and (ol2 > 0).all()
    for i in range(n2):
        assert out1[doff[i] + 8:doff[i + 1]].tobytes() == want[i], f"block {i} after {n1} filler blocks differs"
    _, off_end = stream.peek()
    assert off_end < (1 << 31)                             # renormalised (the reference resets to 64 KiB + later blocks)
    stream.free()


def test_device_reframe_feeds_device_decode(ctx, ref):
    """b200lz4_reframe_dev: the header walk on a stream that is already in HBM must report the same blocks as the host
    walk, and its output arrays feed b200lz4_decompress_dev directly (no host round trip)."""
    import ctypes
    import torch
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import _lib, datagen
    lib = _lib.load()
    rng = np.random.default_rng(9)
    sizes = np.exp(rng.uniform(np.log(4096), np.log(1 << 20), 200)).astype(np.int64)
    total = int(sizes.sum())
    data = datagen.make("mixed", 9, total)
    offs = np.zeros(len(sizes), dtype=np.int64); offs[1:] = np.cumsum(sizes[:-1])
    arrays = [data[o:o + s].tobytes() for o, s in zip(offs, sizes)]
    framed = ref.compress_chunks(arrays, 1, linked=False, threads=8)
    for end_mark in (0, 1):
        blob = np.frombuffer(b"".join(framed) + (b"\0\0\0\0junk" if end_mark else b""), dtype=np.uint8)
        dev = torch.device("cuda", 0)
        d_blob = torch.from_numpy(blob.copy()).to(dev)
        nmax = len(framed) + 5
        d_off = torch.zeros(nmax, dtype=torch.int64, device=dev); d_len = torch.zeros(nmax, dtype=torch.int32, device=dev)
        d_res = torch.zeros(4, dtype=torch.int64, device=dev)
        sh = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        p = lambda t: ctypes.c_void_p(t.data_ptr())
        assert lib.b200lz4_reframe_dev(p(d_blob), blob.size, 8, end_mark, p(d_off), p(d_len), nmax, p(d_res), sh) == 0
        res = d_res.cpu().numpy()
        assert res[0] == len(framed) and res[2] == end_mark and res[3] == 0
        assert res[1] == sum(len(f) for f in framed) + (4 if end_mark else 0)
        assert d_len[:len(framed)].cpu().tolist() == [len(f) for f in framed]
        # decode straight from the device-side block table
        n = len(framed)
        d_dst_off = torch.from_numpy(offs).to(dev); d_cap = torch.from_numpy(sizes.astype(np.int32)).to(dev)
        d_out = torch.zeros(total + 64, dtype=torch.uint8, device=dev); d_olen = torch.zeros(n, dtype=torch.int32, device=dev)
        d_scratch = torch.zeros(lib.b200lz4_scratch_bytes(), dtype=torch.uint8, device=dev)
        assert lib.b200lz4_decompress_dev(p(d_blob), p(d_off), p(d_len), n, None, 0, None, p(d_out), p(d_dst_off), p(d_cap),
                                          p(d_olen), 8, 0, p(d_scratch), sh) == 0
        torch.cuda.synchronize()
        assert d_olen.cpu().tolist() == [int(s) for s in sizes]
        assert bytes(d_out[:total].cpu().numpy()) == data.tobytes()
