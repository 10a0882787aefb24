"""The speculate / verify / repair prototype (tools/specparse_proto.c: a greedy LZ4 parse split across workers, CPU only,
DESIGN.md section 7) must produce the reference's bytes -- for large independent blocks cut into segments and for linked
streams with one worker per block.  Not product code: this pins the ALGORITHM the next compressor kernel would use."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_spec = importlib.util.spec_from_file_location("specparse_proto", os.path.join(ROOT, "tools", "specparse_proto.py"))
proto = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(proto)


@pytest.fixture(scope="module")
def lib(built):
    return proto.build()


def _oracle(built):
    from oracle.oracle import Oracle
    return Oracle("auto")


KINDS = ["text", "mixed", "records", "sparse01", "bits01", "biased01", "random", "zero"]


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("accel", [1, 12, 400])
def test_segmented_block_is_byte_exact(built, lib, kind, accel):
    from streamly_lz4_b200 import datagen
    ora = _oracle(built)
    n = 400000 + 12345
    block = datagen.make(kind, 21, n)
    want = ora.compress_chunks([block.tobytes()], accel, linked=False)[0][8:]
    for seg, warm in ((65536, 65536), (100000, 131072), (50000, 4096), (n + 7, 65536), (65536, 0)):
        got, st = proto.compress(lib, block, accel, seg, warm)
        assert got == want, f"{kind} accel {accel} seg {seg} warm {warm}: bytes differ from the oracle"
        assert st[10] > 0 or kind in ("random",) or accel > 1


@pytest.mark.parametrize("n", [0, 1, 12, 13, 14, 100, 4095, 65535, 65536, 65547, 131079])
def test_segmented_block_edge_sizes(built, lib, n):
    from streamly_lz4_b200 import datagen
    ora = _oracle(built)
    block = datagen.make("text", 5, max(n, 1))[:n]
    want = ora.compress_chunks([block.tobytes()], 1, linked=False)[0][8:]
    for seg in (16, 1000, 32768):
        got, _ = proto.compress(lib, np.ascontiguousarray(block), 1, seg, 65536)
        assert got == want, f"n {n} seg {seg}"


@pytest.mark.parametrize("kind", KINDS)
@pytest.mark.parametrize("accel", [1, 400])
def test_linked_stream_is_byte_exact(built, lib, kind, accel):
    from streamly_lz4_b200 import datagen
    ora = _oracle(built)
    bs, nb = 65536, 12
    stream = datagen.make(kind, 22, bs * nb)
    arrays = [stream[i * bs:(i + 1) * bs].tobytes() for i in range(nb)]
    want = [o[8:] for o in ora.compress_chunks(arrays, accel, linked=True)]
    for warm_blocks in (1, 2, 3, 0):
        got, st = proto.compress_linked(lib, stream, [bs] * nb, accel, warm_blocks)
        assert got == want, f"{kind} accel {accel} warm {warm_blocks}: bytes differ from the oracle"
        # the kernel plan's two claims: a worker's access log is implied by its sequences + catch-up lengths, and the block-level
        # check run without a log (first access per bucket = minimum over access order) finds the same divergence as the log scan
        assert st[17] == 0, "reconstructed access log differs from the logged one"
        assert st[18] == nb - 1 and st[19] == 0, "log-free block check disagrees with the log scan"


@pytest.mark.parametrize("seed", range(6))
def test_linked_stream_ragged_blocks(built, lib, seed):
    """array sizes the reference's tests use (empty arrays mid-stream, 1-3 byte arrays that drop the dictionary, arrays
    below 64 KiB: the dictSmall rule, cbits/lz4.c:1581-1587, :1627) on the generators of test/Main.hs:33-47"""
    from streamly_lz4_b200 import datagen
    ora = _oracle(built)
    rng = np.random.default_rng(100 + seed)
    sizes = [int(x) for x in rng.choice([0, 1, 2, 3, 4, 12, 13, 100, 1000, 10000, 40000, 65535, 65536, 65537, 100000], size=24)]
    if seed == 0:
        sizes = [10000] * 20
    total = sum(sizes)
    kind = ["bits01", "biased01", "text", "mixed", "records", "sparse01"][seed % 6]
    stream = datagen.make(kind, 23 + seed, max(total, 1))[:total]
    offs = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    arrays = [stream[offs[i]:offs[i + 1]].tobytes() for i in range(len(sizes))]
    for accel in (1, 5):
        want = [o[8:] for o in ora.compress_chunks(arrays, accel, linked=True)]
        for warm_blocks in (1, 3):
            got, st = proto.compress_linked(lib, np.ascontiguousarray(stream), sizes, accel, warm_blocks)
            assert got == want, f"seed {seed} kind {kind} accel {accel} warm {warm_blocks} sizes {sizes}"
            assert st[17] == 0 and st[19] == 0
