"""CPU tests: pin the oracle.

1. The restatement (oracle/lz4_oracle.c) must reproduce the committed golden vectors, which
   are outputs of the reference's own cbits/lz4.c (tests/golden/make_golden.py).
2. Where the reference build is present (oracle/_ref/libreflz4.so), the restatement must equal
   it byte for byte on fresh seeded inputs, including the hash-table-dependent linked mode.
3. The Haskell-level logic restated in Python (resize, headers) must satisfy the properties
   of test/Main.hs:189-245.
"""
import hashlib
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = json.load(open(os.path.join(HERE, "golden", "golden.json")))


def arrays_of(case):
    from streamly_lz4_b200 import datagen
    import importlib.util
    spec = importlib.util.spec_from_file_location("make_golden", os.path.join(HERE, "golden", "make_golden.py"))
    if case["name"] == "tiny_sizes":
        mg = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mg)
        d = datagen.make(case["kind"], case["seed"], 4096)
        out, at = [], 0
        for n in mg.TINY_SIZES:
            out.append(d[at:at + n].tobytes()); at += n
        return out
    d = datagen.make(case["kind"], case["seed"], case["total"])
    return [d[i:i + case["block"]].tobytes() for i in range(0, case["total"], case["block"])]


@pytest.mark.parametrize("case", GOLDEN["cases"], ids=[c["name"] for c in GOLDEN["cases"]])
def test_port_matches_golden(port, case):
    arrays = arrays_of(case)
    assert hashlib.sha256(b"".join(arrays)).hexdigest() == case["input_sha256"], "data generator drifted"
    framed = port.compress_chunks(arrays, case["accel"], block_size=case["block_size"], linked=case["linked"])
    assert [len(f) for f in framed] == case["framed_lens"]
    blob = b"".join(framed)
    assert hashlib.sha256(blob).hexdigest() == case["framed_sha256"]
    if "framed_hex" in case:
        assert blob.hex() == case["framed_hex"]
    back = port.decompress_chunks_raw(framed, block_size=case["block_size"], linked=case["linked"])
    assert back == arrays


def test_reference_matches_golden_when_present(built):
    from oracle.oracle import Oracle, available
    if not available("reference"):
        pytest.skip("reference build not present on this box")
    ref = Oracle("reference")
    for case in GOLDEN["cases"]:
        arrays = arrays_of(case)
        framed = ref.compress_chunks(arrays, case["accel"], block_size=case["block_size"], linked=case["linked"])
        assert hashlib.sha256(b"".join(framed)).hexdigest() == case["framed_sha256"], case["name"]


@pytest.mark.parametrize("kind", ["text", "random", "sparse01", "records", "mixed", "bits01", "biased01", "zero"])
def test_port_equals_reference(built, port, kind):
    from oracle.oracle import Oracle, available
    from streamly_lz4_b200 import datagen
    if not available("reference"):
        pytest.skip("reference build not present on this box")
    ref = Oracle("reference")
    d = datagen.make(kind, 4242, 3 << 20)
    for bs, accel, linked in [(65536, 1, True), (65536, 1, False), (640000, 400, False), (100000, 5, True),
                              (4096, 1, True), (1000, 12, True), (333, 0, True), (3 << 20, 1, False), (50000, 65537, True)]:
        arrays = [d[i:i + bs].tobytes() for i in range(0, min(d.size, 200 * bs), bs)]
        a = ref.compress_chunks(arrays, accel, linked=linked)
        b = port.compress_chunks(arrays, accel, linked=linked)
        assert a == b, (kind, bs, accel, linked)
        assert port.decompress_chunks_raw(a, linked=linked) == arrays
        assert ref.decompress_chunks_raw(b, linked=linked) == arrays


def test_port_equals_reference_edge_sizes(built, port):
    from oracle.oracle import Oracle, available
    from streamly_lz4_b200 import datagen
    if not available("reference"):
        pytest.skip("reference build not present on this box")
    ref = Oracle("reference")
    rng = np.random.default_rng(1)
    d = datagen.make("text", 1, 400000)
    for trial in range(30):
        sizes = [int(x) for x in rng.choice([0, 1, 2, 3, 4, 5, 11, 12, 13, 14, 15, 16, 17, 100, 255, 270, 4096, 65535, 65536, 65547],
                                            size=12)]
        arrays, at = [], 0
        for n in sizes:
            at = (at + 997) % (d.size - n - 1)
            arrays.append(d[at:at + n].tobytes())
        accel = int(rng.integers(-1, 13))
        a = ref.compress_chunks(arrays, accel, linked=True)
        assert a == port.compress_chunks(arrays, accel, linked=True), (sizes, accel)
        assert port.decompress_chunks_raw(a, linked=True) == arrays


def test_acceleration_clamp(port):
    """cbits/lz4.c:1577-1578: -5,-1,0,1 coincide; 65537, 65538, 10**6 coincide."""
    from streamly_lz4_b200 import datagen
    d = datagen.make("mixed", 3, 1 << 20)
    arrays = [d[i:i + 100000].tobytes() for i in range(0, d.size, 100000)]
    low = [port.compress_chunks(arrays, a) for a in (-5, -1, 0, 1)]
    assert all(x == low[0] for x in low)
    high = [port.compress_chunks(arrays, a) for a in (65537, 65538, 10 ** 6)]
    assert all(x == high[0] for x in high)
    assert low[0] != high[0]


def test_decoder_accept_reject_agrees_with_reference(built, port):
    from oracle.oracle import Oracle, available
    from streamly_lz4_b200 import datagen
    if not available("reference"):
        pytest.skip("reference build not present on this box")
    ref = Oracle("reference")
    d = datagen.make("text", 13, 100000).tobytes()
    payload = ref.compress_chunks([d], 1, linked=False)[0][8:]
    rng = np.random.default_rng(7)
    cases = [payload[:k] for k in (1, 2, 10, len(payload) // 2, len(payload) - 1)]
    for _ in range(300):
        b = bytearray(payload)
        at = int(rng.integers(0, len(b)))
        b[at] ^= 1 << int(rng.integers(0, 8))
        cases.append(bytes(b))
    agree = 0
    for c in cases:
        arr = len(c).to_bytes(4, "little") + len(d).to_bytes(4, "little") + c
        res = []
        for o in (ref, port):
            try:
                res.append(o.decompress_chunks_raw([arr], linked=False))
            except RuntimeError:
                res.append(None)
        assert res[0] == res[1]
        agree += 1
    assert agree == len(cases)


# ---- resizeChunksD restatement (pure Python) --------------------------------------------------

def _framed(port, n_arrays=20, bs=5000):
    from streamly_lz4_b200 import datagen
    d = datagen.make("biased01", 2, n_arrays * bs)
    arrays = [d[i:i + bs].tobytes() for i in range(0, d.size, bs)]
    return arrays, port.compress_chunks(arrays, 1)


@pytest.mark.parametrize("bufsize", [1, 7, 512, 4000, 32 * 1024, 256 * 1024])
def test_resize_restatement_refragments(port, bufsize):
    from oracle.oracle import resize_chunks
    arrays, framed = _framed(port)
    blob = b"".join(framed)
    chunks = [blob[i:i + bufsize] for i in range(0, len(blob), bufsize)]
    assert resize_chunks(chunks) == framed
    assert resize_chunks(resize_chunks(chunks)) == framed            # idempotent, test/Main.hs:189-201


def test_resize_restatement_end_mark_and_errors(port):
    from oracle.oracle import resize_chunks
    arrays, framed = _framed(port, 5, 1000)
    blob = b"".join(framed)
    with_mark = blob + b"\0\0\0\0" + b"trailing garbage is ignored"
    assert resize_chunks([with_mark], has_end_mark=True) == framed   # test/Main.hs:234-243
    with pytest.raises(RuntimeError, match="No end mark"):
        resize_chunks([blob], has_end_mark=True)
    with pytest.raises(RuntimeError, match="Incomplete block"):
        resize_chunks([blob[:-3]])
    assert resize_chunks([]) == []


def test_xxh32_known_answers(built):
    """XXH32 restated in oracle/lz4_oracle.c against published known answers, and the frame-header checksum bytes of the
    headers the stock `lz4` tool writes (04 22 4D 18 60 40 82 / 64 40 A7 / 64 70 B9)."""
    from oracle import frame
    assert frame.xxh32(b"") == 0x02CC5D05
    assert frame.xxh32(b"a") == 0x550D7456
    assert frame.xxh32(b"abc") == 0x32D153FF
    assert frame.xxh32(b"Nobody inspects the spammish repetition") == 0xE2293B2F
    for desc, hc in ((b"\x60\x40", 0x82), (b"\x64\x40", 0xA7), (b"\x64\x70", 0xB9)):
        assert (frame.xxh32(desc) >> 8) & 0xFF == hc
    assert frame.header(4, independent=True, content_checksum=True).hex() == "04224d186440a7"


def test_frame_restatement_round_trips(built):
    """oracle/frame.py: encode -> parse -> decode over every flag combination, stored (incompressible) blocks included;
    corrupted checksums are rejected."""
    import numpy as np
    from oracle import frame
    from oracle.oracle import Oracle
    from streamly_lz4_b200 import datagen
    ora = Oracle("auto")
    data = datagen.make("mixed", 17, 5 * 65536 + 1234).tobytes()
    arrays = [data[i:i + 65536] for i in range(0, len(data), 65536)]
    for ind in (False, True):
        for bc in (False, True):
            for cs in (False, True):
                for cc in (False, True):
                    blob = frame.encode(ora, arrays, 4, 1, ind, bc, cs, cc)
                    d, blocks, _ = frame.parse(blob)
                    assert d["frame_bytes"] == len(blob) and len(blocks) == len(arrays)
                    assert any(stored for stored, _ in blocks)              # the random segment is stored uncompressed
                    assert frame.decode(ora, blob) == data
    blob = bytearray(frame.encode(ora, arrays, 4, 1, False, True, False, True))
    blob[40] ^= 1
    with pytest.raises(ValueError):
        frame.decode(ora, bytes(blob))
