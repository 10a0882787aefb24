"""Randomised parity: (i) hand-built LZ4 blocks that exercise decoder paths the encoder never produces
(offsets 1..3, long overlapping matches, 255-runs in both length fields, literal runs around the 15/32 limits,
matches reaching into the previous block), decoded by the GPU and by the oracle; (ii) random arrays / sizes /
accelerations / linkage through the whole compress -> decompress path against the oracle."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHIFT = int(os.environ.get("B200LZ4_FUZZ_SHIFT", "0"))      # soak runs: B200LZ4_FUZZ_SHIFT=1000 python -m pytest tests/test_gpu_fuzz.py


def _emit_len(out, v):
    v -= 15
    while v >= 255:
        out.append(255); v -= 255
    out.append(v)


def build_block(rng, target, dict_avail):
    """A valid LZ4 block of about `target` output bytes; returns (payload bytes, plaintext length)."""
    out = bytearray()
    produced = 0
    plain_src = rng.integers(0, 256, size=target + 4096, dtype=np.uint8).tobytes()
    while True:
        style = rng.integers(0, 8)
        lit = int(rng.choice([0, 0, 1, 3, 14, 15, 16, 31, 32, 33, 60, 270, 600])) if style else int(rng.integers(0, 20))
        if produced + lit + 40 >= target:
            break
        reach = produced + lit + dict_avail
        if reach < 1:
            lit = max(lit, 1); reach = produced + lit + dict_avail
        off = int(rng.choice([1, 2, 3, 4, 7, 8, 15, 16, 31, 32, 33, 47, 48, 63, 64, 65, 255, 256, 1000, 4095, 8191, 8192, 20000, 65535]))
        if style == 1:
            off = int(rng.integers(1, 65536))
        off = max(1, min(off, reach, 65535))
        ml = int(rng.choice([4, 5, 18, 19, 20, 33, 36, 64, 65, 68, 69, 100, 273, 274, 529, 2000])) if style else int(rng.integers(4, 40))
        ml = min(ml, target - produced - lit - 13)
        if ml < 4:
            break
        tok = (min(lit, 15) << 4) | min(ml - 4, 15)
        out.append(tok)
        if lit >= 15:
            _emit_len(out, lit)
        out += plain_src[produced:produced + lit]
        out += bytes([off & 255, off >> 8])
        if ml - 4 >= 15:
            _emit_len(out, ml - 4)
        produced += lit + ml
    last = max(target - produced, 6)                       # the block ends with >= 5 literals (12 after the last match start)
    out.append(min(last, 15) << 4)
    if last >= 15:
        _emit_len(out, last)
    out += plain_src[produced:produced + last]
    return bytes(out), produced + last


@pytest.mark.parametrize("seed", range(20))
def test_handbuilt_blocks_decode_like_the_reference(ctx, ref, seed):
    import streamly_lz4_b200 as lz
    rng = np.random.default_rng(1000 + seed + SHIFT)
    for linked in (False, True):
        framed = []
        prev = 0
        for _ in range(10):
            target = int(rng.choice([64, 300, 5000, 70000, 300000]))
            payload, n = build_block(rng, target, prev if linked else 0)
            framed.append(len(payload).to_bytes(4, "little") + n.to_bytes(4, "little") + payload)
            prev = n
        want = ref.decompress_chunks_raw(framed, linked=linked)
        got = list(lz.decompress_chunks_raw(lz.BlockConfig(independent=not linked), framed, ctx=ctx))
        assert [len(g) for g in got] == [len(w) for w in want]
        for i, (g, w) in enumerate(zip(got, want)):
            if g != w:
                at = int(np.argmax(np.frombuffer(g, np.uint8) != np.frombuffer(w, np.uint8)))
                raise AssertionError(f"seed {seed} linked {linked} block {i}: first difference at byte {at} of {len(w)}")


@pytest.mark.parametrize("seed", range(8))
def test_random_round_trips_against_oracle(ctx, ref, seed):
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import datagen
    rng = np.random.default_rng(2000 + seed + SHIFT)
    kinds = ["text", "random", "sparse01", "records", "mixed", "bits01", "biased01", "zero"]
    for _ in range(12):
        kind = kinds[int(rng.integers(0, len(kinds)))]
        total = int(rng.integers(1, 400000))
        data = datagen.make(kind, int(rng.integers(0, 1 << 30)), total)
        cuts = np.sort(rng.integers(0, total + 1, size=int(rng.integers(0, 12))))
        arrays = [data[a:b].tobytes() for a, b in zip(np.r_[0, cuts], np.r_[cuts, total])]   # includes empty arrays
        accel = int(rng.choice([-3, 0, 1, 2, 7, 33, 400, 5000, 65537, 100000]))
        linked = bool(rng.integers(0, 2))
        cfg = lz.BlockConfig(independent=not linked)
        got = list(lz.compress_chunks(cfg, accel, arrays, ctx=ctx))
        want = ref.compress_chunks(arrays, accel, linked=linked)
        assert got == want, f"{kind} total={total} accel={accel} linked={linked} cuts={list(cuts)}"
        assert list(lz.decompress_chunks_raw(cfg, want, ctx=ctx)) == arrays


@pytest.mark.parametrize("seed", range(6))
def test_linked_streams_split_over_random_batches(ctx, ref, seed):
    """One linked stream pushed through the library in random batch sizes (device-resident stream state carried from
    call to call, including tiny and empty arrays) must equal the reference's single pass, both ways; the same for
    hand-built blocks whose matches reach into the previous output."""
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import datagen
    rng = np.random.default_rng(3000 + seed + SHIFT)
    kind = ["text", "mixed", "records", "biased01"][seed % 4]
    total = int(rng.integers(200000, 900000))
    data = datagen.make(kind, 40 + seed, total)
    cuts = np.sort(rng.integers(0, total + 1, size=int(rng.integers(5, 40))))
    arrays = [data[a:b].tobytes() for a, b in zip(np.r_[0, cuts], np.r_[cuts, total])]
    accel = int(rng.choice([1, 1, 3, 50]))
    want = ref.compress_chunks(arrays, accel, linked=True)
    for batch_arrays in (1, int(rng.integers(2, 6)), int(rng.integers(6, 20))):
        got = list(lz.compress_chunks(lz.BlockConfig(), accel, arrays, ctx=ctx, batch_arrays=batch_arrays))
        assert got == want, f"{kind} accel={accel} batch_arrays={batch_arrays}"
        assert list(lz.decompress_chunks_raw(lz.BlockConfig(), want, ctx=ctx, batch_arrays=batch_arrays)) == arrays
    framed, prev = [], 0
    for _ in range(12):
        payload, n = build_block(rng, int(rng.choice([40, 64, 300, 5000, 70000])), prev)
        framed.append(len(payload).to_bytes(4, "little") + n.to_bytes(4, "little") + payload)
        prev = n
    plain = ref.decompress_chunks_raw(framed, linked=True)
    for batch_arrays in (1, 2, 5):
        assert list(lz.decompress_chunks_raw(lz.BlockConfig(), framed, ctx=ctx, batch_arrays=batch_arrays)) == plain


def _has_zero_offset(payload: bytes) -> bool:
    """Walk the sequences of a (possibly corrupt) payload; True if one carries offset 0 (the one input class where this
    decoder deliberately rejects what the reference accepts: the reference then copies unwritten memory)."""
    ip, n = 0, len(payload)
    try:
        while ip < n:
            tok = payload[ip]; ip += 1
            lit = tok >> 4
            if lit == 15:
                while True:
                    s = payload[ip]; ip += 1; lit += s
                    if s != 255:
                        break
            ip += lit
            if ip + 2 > n:
                return False
            if payload[ip] == 0 and payload[ip + 1] == 0:
                return True
            ip += 2
            if tok & 15 == 15:
                while True:
                    s = payload[ip]; ip += 1
                    if s != 255:
                        break
    except IndexError:
        pass
    return False


@pytest.mark.parametrize("seed", range(6))
def test_corrupted_blocks_accept_reject_like_the_reference(ctx, ref, seed):
    """Bit flips, truncations and garbage tails on hand-built and encoder-made blocks: the GPU decoder must make the same
    accept / reject decision as the reference and, when both accept, produce the same bytes."""
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import datagen
    rng = np.random.default_rng(4000 + seed + SHIFT)
    cfg = lz.BlockConfig(independent=True)
    sources = []
    for target in (64, 400, 5000, 70000):
        sources.append(build_block(rng, target, 0))
    d = datagen.make(["text", "mixed", "sparse01"][seed % 3], 60 + seed, 60000).tobytes()
    sources.append((ref.compress_chunks([d], 1, linked=False)[0][8:], len(d)))
    arrays, notes = [], []
    for payload, n in sources:
        for _ in range(25):
            b = bytearray(payload)
            kind = int(rng.integers(0, 4))
            if kind == 0:
                at = int(rng.integers(0, len(b))); b[at] ^= 1 << int(rng.integers(0, 8))
            elif kind == 1:
                at = int(rng.integers(0, min(len(b), 64))); b[at] = int(rng.integers(0, 256))
            elif kind == 2:
                b = b[:int(rng.integers(1, len(b)))]
            else:
                b += bytes(rng.integers(0, 256, size=int(rng.integers(1, 9)), dtype=np.uint8))
            cap = n if rng.integers(0, 3) else int(rng.integers(max(n - 20, 0), n + 20))
            arrays.append(len(b).to_bytes(4, "little") + cap.to_bytes(4, "little") + bytes(b))
            notes.append((kind, len(payload), n, cap))
    for arr, note in zip(arrays, notes):
        try:
            want = ref.decompress_chunks_raw([arr], linked=False)
        except RuntimeError:
            want = None
        try:
            got = list(lz.decompress_chunks_raw(cfg, [arr], ctx=ctx))
        except lz.LZ4Error:
            got = None
        if got != want:
            if got is None and want is not None and _has_zero_offset(arr[8:]):
                continue
            raise AssertionError(f"seed {seed} case {note}: gpu {'rejects' if got is None else 'accepts'}, "
                                 f"reference {'rejects' if want is None else 'accepts'}")


def _echo_arrays(rng, n_arrays):
    """Arrays that repeat, shift and lightly edit their predecessors: long matches into the dictionary, matches that run
    from the dictionary into the current array (the two-segment count, cbits/lz4.c:1080-1089), tiny dictionaries."""
    from streamly_lz4_b200 import datagen
    kind = ["text", "records", "random", "mixed"][int(rng.integers(0, 4))]
    base = datagen.make(kind, int(rng.integers(0, 1 << 30)), int(rng.integers(2000, 90000))).tobytes()
    out = [base]
    for _ in range(n_arrays - 1):
        prev = out[-1]
        op = int(rng.integers(0, 7))
        if op == 0:
            a = prev
        elif op == 1 and len(prev) > 10:
            k = int(rng.integers(1, min(len(prev), 5000))); a = prev[k:] + prev[:k]
        elif op == 2 and len(prev) > 10:
            b = bytearray(prev)
            for _ in range(int(rng.integers(1, 20))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
            a = bytes(b)
        elif op == 3:
            a = prev[-int(rng.integers(1, 70000)):] + prev[:int(rng.integers(0, 70000))]
        elif op == 4:
            a = prev[:int(rng.integers(0, 14))]                          # tiny array: a tiny (or dropped) dictionary follows
        elif op == 5:
            a = base[:int(rng.integers(1, len(base)))] * int(rng.integers(1, 4))
        else:
            a = prev + prev[:int(rng.integers(0, 3000))]
        out.append(a[:200000])
    return out


@pytest.mark.parametrize("seed", range(6))
def test_echoing_linked_streams_wide_and_dense(ctx, ref, seed):
    import streamly_lz4_b200 as lz
    rng = np.random.default_rng(5000 + seed + SHIFT)
    accel = int(rng.choice([1, 1, 2, 9, 400]))
    # (i) a handful of streams: the wide kernel (one stream per SM, dictionary in the shared-memory ring)
    # (ii) more streams than SMs: the dense kernel (dictionary in global memory)
    for n_streams, per in ((3, 14), (200, 4)):
        streams = [_echo_arrays(rng, per) for _ in range(n_streams)]
        arrays = [a for s in streams for a in s]
        sf = np.zeros(n_streams + 1, dtype=np.int32); sf[1:] = np.cumsum([len(s) for s in streams])
        lens = np.array([len(a) for a in arrays], dtype=np.int32)
        strides = (lens.astype(np.int64) + 16 + 15) // 16 * 16
        offs = np.zeros(len(arrays), dtype=np.int64); offs[1:] = np.cumsum(strides[:-1])
        src = ctx.pinned("echo_src", int(strides.sum()) + 64)
        for a, o in zip(arrays, offs):
            src[o:o + len(a)] = np.frombuffer(a, dtype=np.uint8)
        dst = ctx.pinned("echo_dst", int((lens.astype(np.int64) + lens // 255 + 24).sum()))
        rc, doff, olen = ctx.compress_batch(src[:int(strides.sum())], offs, lens, accel, 8, dst, stream_first=sf)
        assert rc == 0
        want = ref.compress_chunks(arrays, accel, linked=True, stream_first=sf, threads=8)
        for i, w in enumerate(want):
            assert dst[doff[i]:doff[i + 1]].tobytes() == w, f"seed {seed} streams {n_streams} accel {accel} array {i} (len {lens[i]})"
        back = ctx.pinned("echo_back", int(lens.sum()) + 64)
        comp = dst[:doff[-1]]
        rc, boff, blen = ctx.decompress_batch(comp, doff[:-1].copy(), np.diff(doff).astype(np.int32), 8, 0, back, stream_first=sf)
        assert rc == 0 and (blen == lens).all()
        assert back[:int(lens.sum())].tobytes() == b"".join(arrays)


@pytest.mark.parametrize("seed", range(4))
def test_corrupted_block_inside_a_linked_stream(ctx, ref, seed):
    """A block that fails in the MIDDLE of a linked stream must not disturb the chain: LZ4_decompress_safe_continue returns
    before it touches the stream state (cbits/lz4.c:2353 `if (result <= 0) return result`), so the next block still sees the
    last GOOD output as its dictionary.  The batch call reports the failure per block and goes on; the same call sequence
    on the reference's own lz4.c (one LZ4_streamDecode_t, LZ4_decompress_safe_continue per block) is the expectation.  In
    the wide decoder this is the path that restarts the shared-memory ring from global memory."""
    import ctypes
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import datagen
    rng = np.random.default_rng(7000 + seed + SHIFT)
    bs = int(rng.choice([3000, 65536, 200000]))
    data = datagen.make(["text", "mixed", "records", "sparse01"][seed % 4], 80 + seed, 6 * bs)
    arrays = [data[i:i + bs].tobytes() for i in range(0, data.size, bs)]
    good = ref.compress_chunks(arrays, 1, linked=True)                       # blocks 0..5, each with its predecessor as dictionary
    skip = ref.compress_chunks([arrays[1], arrays[3]], 1, linked=True)[1]    # block 3 compressed against block 1
    half = good[2][8:8 + (len(good[2]) - 8) // 2]                             # a truncated payload: fails (input ends inside a sequence)
    junk = len(half).to_bytes(4, "little") + good[2][4:8] + half
    garbage = (24).to_bytes(4, "little") + (bs).to_bytes(4, "little") + bytes(rng.integers(0, 256, 24, dtype=np.uint8))
    stream = [good[0], good[1], bytes(junk), garbage, skip, good[4]]         # good[4] needs block 3's plaintext as dictionary
    # the reference, call by call
    framed_arena, ptrs, lens = ref.lay_out(stream)
    ds = ref.dcreate()
    outs = [np.zeros(bs + 64, dtype=np.uint8) for _ in stream]               # separate, non-adjacent output arrays
    want_len, want = [], []
    for i, f in enumerate(stream):
        cap = int.from_bytes(f[4:8], "little", signed=True)
        r = ref.dcont(ds, int(ptrs[i]) + 8, outs[i].ctypes.data, len(f) - 8, cap)
        want_len.append(r); want.append(outs[i][:max(r, 0)].tobytes())
    ref.dfree(ds)
    assert want_len[0] == bs and want_len[1] == bs and want_len[2] < 0 and want_len[3] < 0
    assert want_len[4] == bs and want[4] == arrays[3] and want[5] == arrays[4]     # the chain survived the two failures
    # here, in one batch and split over two calls (the state crosses the call after the failed blocks)
    for cut in (len(stream), 4):
        dstream = lz.DecompressStream(ctx)
        got_len, got = [], []
        for part in (stream[:cut], stream[cut:]):
            if not part:
                continue
            src, offs, plens = lz.api._gather(ctx, "t_src", part)
            dst = ctx.pinned("t_dst", (bs + 64) * len(part))
            rc, doff, olen = ctx.decompress_batch(src, offs, plens, 8, 0, dst, np.array([0, len(part)], dtype=np.int32), [dstream])
            assert rc in (0, -4)
            for k in range(len(part)):
                got_len.append(int(olen[k])); got.append(dst[doff[k]:doff[k] + max(int(olen[k]), 0)].tobytes())
        dstream.free()
        for i in range(len(stream)):
            assert (got_len[i] < 0) == (want_len[i] < 0), f"block {i}: {got_len[i]} vs reference {want_len[i]}"
            if want_len[i] >= 0:
                assert got_len[i] == want_len[i] and got[i] == want[i], f"block {i} differs (cut {cut})"
