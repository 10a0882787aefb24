"""LZ4 frame interoperability: XXH32 on the device, a complete frame writer whose output the
stock frame rules (restated in oracle/frame.py) decode, and a complete reader for frames a stock writer produces."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_xxh32_kernel_matches_the_restatement(ctx):
    from oracle import frame
    rng = np.random.default_rng(3)
    buf = ctx.pinned("x_buf", 12 << 20)
    buf[:12 << 20] = rng.integers(0, 256, 12 << 20, dtype=np.uint8)
    lens = [0, 1, 3, 4, 5, 15, 16, 17, 31, 32, 33, 63, 64, 511, 512, 513, 4096, 65536, 640000, 1 << 20]
    lens += [int(x) for x in rng.integers(0, 70000, 200)]
    offs, at = [], 0
    for n in lens:
        offs.append(at + int(rng.integers(0, 7)))           # every alignment
        at = offs[-1] + n
    assert at < buf.size
    for seed in (0, 1, 0x9E3779B1):
        got = ctx.xxh32(buf, np.array(offs), np.array(lens), seed)
        want = [frame.xxh32(buf[o:o + n].tobytes(), seed) for o, n in zip(offs, lens)]
        assert [int(x) for x in got] == want


@pytest.mark.parametrize("independent", [False, True])
@pytest.mark.parametrize("checks", [(False, False, False), (True, True, True)])
def test_written_frames_decode_with_the_stock_rules(ctx, ref, independent, checks):
    """write_frame -> oracle/frame.py (header checksum, block checksums, content size, content checksum all verified there)."""
    import streamly_lz4_b200 as lz
    from oracle import frame
    from streamly_lz4_b200 import datagen
    bc, cs, cc = checks
    data = datagen.make("mixed", 23, 9 * 65536 + 777).tobytes()
    arrays = [data[i:i + 65536] for i in range(0, len(data), 65536)]
    blob = b"".join(lz.write_frame(lz.BlockSize.BlockMax64KB, 1, arrays, independent=independent, block_checksum=bc,
                                   content_checksum=cc, content_size=len(data) if cs else None, ctx=ctx, batch_arrays=4))
    d, blocks, _ = frame.parse(blob)
    assert d["independent"] == independent and d["block_checksum"] == bc and d["content_checksum"] == cc
    want = [f[4:] for f in ref.compress_chunks(arrays, 1, block_size="BlockMax64KB", linked=not independent)]
    assert any(stored for stored, _ in blocks)                                  # the incompressible segment
    for (stored, b), w, a in zip(blocks, want, arrays):
        assert b == (a if stored else w) and stored == (len(w) >= len(a))
    assert frame.decode(ref, blob) == data


@pytest.mark.parametrize("independent", [False, True])
def test_stock_frames_decode_here(ctx, ref, independent):
    """oracle/frame.py writes what a stock writer would (incompressible blocks stored uncompressed, all checksums);
    read_frame decodes it from fragmented input and rejects corrupted checksums."""
    import streamly_lz4_b200 as lz
    from oracle import frame
    from streamly_lz4_b200 import datagen
    data = datagen.make("mixed", 29, 11 * 65536 + 99).tobytes()
    arrays = [data[i:i + 65536] for i in range(0, len(data), 65536)]
    blob = frame.encode(ref, arrays, 4, 1, independent, True, True, True)
    assert any(stored for stored, _ in frame.parse(blob)[1])
    for bufsize in (7, 4096, 1 << 20):
        chunks = [blob[i:i + bufsize] for i in range(0, len(blob), bufsize)]
        assert b"".join(lz.read_frame(chunks, ctx=ctx)) == data
    bad = bytearray(blob); bad[len(blob) // 2] ^= 0x10
    with pytest.raises(lz.LZ4Error):
        b"".join(lz.read_frame([bytes(bad)], ctx=ctx))
    bad = bytearray(blob); bad[-1] ^= 1                      # content checksum
    with pytest.raises(lz.LZ4Error, match="content checksum"):
        b"".join(lz.read_frame([bytes(bad)], ctx=ctx))
