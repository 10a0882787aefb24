"""CPU tests of the boundary: the C-ABI library loads without a GPU, exports every symbol
include/b200lz4.h declares, its host-only entry points (bounds, re-frame) agree with the oracle,
and the codec entry points fail loudly (no CPU fallback) when there is no device."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "b200lz4.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b((?:b200lz4|LZ4)_[A-Za-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported(built):
    from streamly_lz4_b200 import _lib
    names = declared_symbols()
    assert len(names) >= 30
    out = subprocess.check_output(["nm", "-D", "--defined-only", _lib.LIB_PATH], text=True)
    exported = set(line.split()[-1] for line in out.splitlines() if line.strip())
    missing = [n for n in names if n not in exported]
    assert not missing, f"declared in include/b200lz4.h but not exported: {missing}"
    assert set(_lib.PROTOTYPES) == set(names), "ctypes prototypes and header disagree"
    _lib.load()


def test_library_is_sm100a_cuda_code(built):
    from streamly_lz4_b200 import _lib
    out = subprocess.run(["cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump not available")
    assert "sm_100a" in out.stdout


def test_compress_bound_matches_reference(built, ref):
    from streamly_lz4_b200 import _lib
    lib = _lib.load()
    for n in [0, 1, 12, 13, 254, 255, 256, 65536, 640000, 4 << 20, 0x7E000000, 0x7E000001, -1]:
        assert lib.b200lz4_compress_bound(n) == ref.bound(n) == lib.LZ4_compressBound(n)
    assert lib.b200lz4_compress_bound(65536) == 65809 and lib.b200lz4_compress_bound(640000) == 642525


def test_no_gpu_fails_loudly(built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import _lib
    with pytest.raises(lz.LZ4Error, match="no CPU path"):
        lz.Context(0)
    lib = _lib.load()
    assert lib.b200lz4_device_count() == 0
    assert not lib.LZ4_createStream()
    assert list(lz.compress_chunks(lz.default_block_config, 1, [])) == [] if False else True


def _reframe_all(lib, blob: bytes, header: int, has_end_mark: bool, max_blocks=1 << 16):
    off = np.zeros(max_blocks, dtype=np.int64)
    ln = np.zeros(max_blocks, dtype=np.int32)
    nf, used, ended = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()
    buf = ctypes.create_string_buffer(blob, len(blob))
    rc = lib.b200lz4_reframe(buf, len(blob), header, int(has_end_mark), off.ctypes.data, ln.ctypes.data, max_blocks,
                             ctypes.byref(nf), ctypes.byref(used), ctypes.byref(ended))
    return rc, [(int(off[k]), int(ln[k])) for k in range(nf.value)], used.value, ended.value


def test_reframe_matches_resize_restatement(built, port):
    from oracle.oracle import resize_chunks
    from streamly_lz4_b200 import _lib, datagen
    lib = _lib.load()
    rng = np.random.default_rng(3)
    d = datagen.make("mixed", 31, 6 << 20)
    # config 5(i): plaintext sizes log-uniform in [4 KiB, 4 MiB] (here capped to keep the CPU suite short)
    sizes, at, arrays = [], 0, []
    while at < d.size:
        n = int(np.exp(rng.uniform(np.log(4096), np.log(1 << 20))))
        arrays.append(d[at:at + n].tobytes()); at += n
    for cfg, header in (("BlockHasSize", 8), ("BlockMax1MB", 4)):
        framed = port.compress_chunks(arrays, 1, block_size=cfg)
        blob = b"".join(framed)
        rc, blocks, used, ended = _reframe_all(lib, blob, header, False)
        assert rc == 0 and used == len(blob) and not ended
        assert [blob[o:o + n] for o, n in blocks] == framed == resize_chunks([blob], cfg)
        # a truncated tail is reported as "not consumed", never as a block
        rc, blocks2, used2, _ = _reframe_all(lib, blob[:-5], header, False)
        assert rc == 0 and blocks2 == blocks[:-1] and used2 == blocks[-1][0]
        # end mark stops the walk
        rc, blocks3, used3, ended3 = _reframe_all(lib, blob + b"\0\0\0\0junk", header, True)
        assert rc == 0 and blocks3 == blocks and ended3 == 1 and used3 == len(blob) + 4


@pytest.mark.parametrize("bufsize", [1, 512, 6553, 65536, 655360, 640000])
def test_resize_chunks_mirror_fragmented(built, port, bufsize):
    """benchmark strategy r+bufsize (benchmark/Main.hs:189-207) at the bufsizes of config 5(i)."""
    import streamly_lz4_b200 as lz
    from oracle.oracle import resize_chunks as ora_resize
    from streamly_lz4_b200 import datagen
    d = datagen.make("mixed", 8, (1 << 20) if bufsize > 1 else 60000)
    arrays = [d[i:i + 70000].tobytes() for i in range(0, d.size, 70000)]
    framed = port.compress_chunks(arrays, 5)
    blob = b"".join(framed)
    chunks = [blob[i:i + bufsize] for i in range(0, len(blob), bufsize)]
    got = list(lz.resize_chunks(lz.default_block_config, lz.default_frame_config, chunks))
    assert got == framed == ora_resize(chunks)
    again = list(lz.resize_chunks(lz.default_block_config, lz.default_frame_config, got))     # idempotent
    assert again == got


def test_resize_chunks_mirror_errors(built, port):
    import streamly_lz4_b200 as lz
    from streamly_lz4_b200 import datagen
    d = datagen.make("text", 8, 30000)
    framed = port.compress_chunks([d[:10000].tobytes(), d[10000:].tobytes()], 1)
    blob = b"".join(framed)
    endcfg = lz.set_frame_end_mark(True)(lz.default_frame_config)
    assert list(lz.resize_chunks(lz.default_block_config, endcfg, [blob + b"\0\0\0\0", b"ignored"])) == framed
    with pytest.raises(lz.LZ4Error, match="No end mark found"):
        list(lz.resize_chunks(lz.default_block_config, endcfg, [blob]))
    with pytest.raises(lz.LZ4Error, match="Incomplete block"):
        list(lz.resize_chunks(lz.default_block_config, lz.default_frame_config, [blob[:-1]]))
    bad = (-5).to_bytes(4, "little", signed=True) + bytes(20)
    with pytest.raises(lz.LZ4Error):
        list(lz.resize_chunks(lz.default_block_config, lz.default_frame_config, [bad]))


def test_config_mirror():
    import streamly_lz4_b200 as lz
    cfg = lz.default_block_config
    assert cfg.meta_size == 8 and cfg.max_block_size == 0x7E000000 and not cfg.independent
    c2 = lz.set_block_max_size(lz.BlockSize.BlockMax256KB)(cfg)
    assert c2.meta_size == 4 and c2.max_block_size == 256 * 1024
    assert lz.set_block_independence(True)(cfg).independent
    assert lz.set_frame_end_mark(True)(lz.default_frame_config).has_end_mark


def test_frame_header_parser_mirrors_the_reference():
    """simpleFrameParserD (src/Streamly/Internal/LZ4.hs:590-651): accepted headers, every rejection and its message."""
    import streamly_lz4_b200 as lz
    ok = bytes([4, 34, 77, 24, 64, 64, 0])                  # benchmark/Main.hs:92-100
    cfg, fc = lz.simple_frame_parser(ok)
    assert cfg.block_size is lz.BlockSize.BlockMax64KB and cfg.meta_size == 4 and fc.has_end_mark and not cfg.independent
    for code, bs in ((4, "BlockMax64KB"), (5, "BlockMax256KB"), (6, "BlockMax1MB"), (7, "BlockMax4MB")):
        assert lz.simple_frame_parser(ok[:5] + bytes([code << 4, 0]))[0].block_size is lz.BlockSize[bs]
        assert lz.frame_header(lz.BlockSize[bs], checksum=False) == ok[:5] + bytes([code << 4, 0])
    bad = [(bytes([5, 34, 77, 24, 64, 64, 0]), "does not match 407708164"),
           (ok[:4] + bytes([0x80, 64, 0]), "Version is not 01"), (ok[:4] + bytes([0x00, 64, 0]), "Version is not 01"),
           (ok[:4] + bytes([0x60, 64, 0]), "Block independence is not yet supported"),
           (ok[:4] + bytes([0x50, 64, 0]), "Block checksum is not yet supported"),
           (ok[:4] + bytes([0x48, 64, 0]), "Content size is not yet supported"),
           (ok[:4] + bytes([0x44, 64, 0]), "Content checksum is not yet supported"),
           (ok[:4] + bytes([0x41, 64, 0]), "Dict is not yet supported"),
           (ok[:5] + bytes([0x30, 0]), "parseBD: Unknown block max size")]
    for hdr, msg in bad:
        with pytest.raises(lz.LZ4Error, match=msg):
            lz.simple_frame_parser(hdr)
    cfg, _ = lz.simple_frame_parser(ok[:4] + bytes([0x60, 64, 0]), allow_independent=True)
    assert cfg.independent


def test_gather_host_is_a_parallel_memcpy(built):
    """b200lz4_gather_host is plain host code (the staging step in front of every batch call): usable without a GPU."""
    import ctypes
    from streamly_lz4_b200 import _lib
    lib = _lib.load()
    rng = np.random.default_rng(1)
    arrays = [rng.integers(0, 256, int(k), dtype=np.uint8) for k in rng.integers(0, 400000, 57)]
    lens = np.array([a.size for a in arrays], dtype=np.int32)
    offs = np.zeros(len(arrays), dtype=np.int64)
    offs[1:] = np.cumsum((lens[:-1].astype(np.int64) + 31) // 16 * 16)
    dst = np.zeros(int(offs[-1] + lens[-1] + 64), dtype=np.uint8)
    ptrs = np.array([a.ctypes.data for a in arrays], dtype=np.uint64)
    for threads in (1, 0, 5):
        dst[:] = 0
        assert lib.b200lz4_gather_host(dst.ctypes.data, ptrs.ctypes.data, offs.ctypes.data, lens.ctypes.data, len(arrays), threads) == 0
        for a, o in zip(arrays, offs):
            assert np.array_equal(dst[o:o + a.size], a)


def test_complete_frame_descriptor(built):
    """The descriptor beyond the reference's stub: the headers the stock `lz4` tool writes (known answers), every flag,
    header-checksum verification, reserved bits."""
    import streamly_lz4_b200 as lz
    from oracle import frame as oframe
    B = lz.BlockSize
    assert lz.frame_header(B.BlockMax64KB, independent=True).hex() == "04224d18604082"
    assert lz.frame_header(B.BlockMax64KB, independent=True, content_checksum=True).hex() == "04224d186440a7"
    assert lz.frame_header(B.BlockMax4MB, independent=True, content_checksum=True).hex() == "04224d186470b9"
    for bs, code in ((B.BlockMax64KB, 4), (B.BlockMax256KB, 5), (B.BlockMax1MB, 6), (B.BlockMax4MB, 7)):
        for ind in (False, True):
            for bc in (False, True):
                for cc in (False, True):
                    for size in (None, 0, 123456789012):
                        h = lz.frame_header(bs, independent=ind, block_checksum=bc, content_checksum=cc, content_size=size)
                        assert h == oframe.header(code, ind, bc, size, cc)
                        cfg, fc, info = lz.parse_frame_header(h + bytes(8))
                        assert (cfg.block_size, cfg.independent, fc.has_end_mark) == (bs, ind, True)
                        assert (info.block_checksum, info.content_checksum, info.content_size, info.header_len) == (bc, cc, size, len(h))
                        bad = bytearray(h); bad[-1] ^= 1
                        with pytest.raises(lz.LZ4Error, match="header checksum"):
                            lz.parse_frame_header(bytes(bad) + bytes(8))
    with pytest.raises(lz.LZ4Error, match="reserved"):
        lz.parse_frame_header(bytes.fromhex("04224d18624000"))
    with pytest.raises(lz.LZ4Error, match="Dict"):
        lz.parse_frame_header(bytes.fromhex("04224d18614000") + bytes(8))
