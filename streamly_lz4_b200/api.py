"""Host-side mirror of the reference's public API over the C ABI.

Names, argument meaning and error behaviour follow `Streamly.LZ4`
(src/Streamly/LZ4.hs:40-122) and `Streamly.Internal.LZ4`
(src/Streamly/Internal/LZ4.hs:342-567):

    compress_chunks(cfg, speed, chunks)      compressChunks   LZ4.hs:94-100   / Internal/LZ4.hs:353-394
    decompress_chunks(cfg, chunks)           decompressChunks LZ4.hs:114-122  (= resize then raw decode)
    decompress_chunks_raw(cfg, chunks)       decompressChunksRawD  Internal/LZ4.hs:539-567
    resize_chunks(cfg, frame_cfg, chunks)    resizeChunksD    Internal/LZ4.hs:432-523
    BlockConfig / BlockSize / FrameConfig    Internal/LZ4/Config.hs:41-161

A "stream" is any Python iterable of bytes-like arrays.  Where the reference
calls the codec once per array, this mirror gathers arrays into a pinned host
batch and makes ONE call into libb200lz4.so per batch (the behaviour the
patched Haskell shim in INTEGRATION.md has); results are yielded in order, one
output array per input array, with the reference's block header layout.

All codec work happens in the CUDA library; there is no CPU path here.
"""
from __future__ import annotations

import ctypes
import enum
import threading
from dataclasses import dataclass, replace
from typing import Iterable, Iterator, List, Optional, Sequence

import numpy as np

from . import _lib

LZ4_MAX_INPUT_SIZE = 0x7E000000


class BlockSize(enum.Enum):                 # Config.hs:104-119
    BlockHasSize = 0
    BlockMax64KB = 64 * 1024
    BlockMax256KB = 256 * 1024
    BlockMax1MB = 1024 * 1024
    BlockMax4MB = 4 * 1024 * 1024


@dataclass(frozen=True)
class BlockConfig:                           # Config.hs:121-134
    block_size: BlockSize = BlockSize.BlockHasSize
    independent: bool = False                # setBlockIndependence (a stub in the reference, Config.hs:142-146)

    @property
    def meta_size(self) -> int:              # metaSize, Internal/LZ4.hs:177-181
        return 8 if self.block_size is BlockSize.BlockHasSize else 4

    @property
    def max_block_size(self) -> int:         # Internal/LZ4.hs:275-281
        return LZ4_MAX_INPUT_SIZE if self.block_size is BlockSize.BlockHasSize else self.block_size.value


@dataclass(frozen=True)
class FrameConfig:                           # Config.hs:41-48
    has_end_mark: bool = False


default_block_config = BlockConfig()
default_frame_config = FrameConfig()


def set_block_max_size(bs: BlockSize):       # setBlockMaxSize, Config.hs:136-140
    return lambda cfg: replace(cfg, block_size=bs)


def set_block_independence(flag: bool):      # setBlockIndependence, Config.hs:142-146 (implemented here)
    return lambda cfg: replace(cfg, independent=flag)


def set_frame_end_mark(flag: bool):          # setFrameEndMark, Config.hs:60-63
    return lambda cfg: replace(cfg, has_end_mark=flag)


class LZ4Error(RuntimeError):
    pass


class Context:
    """One b200lz4_ctx (device + CUDA stream + staging arenas) with pinned batch buffers."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        rc = self.lib.b200lz4_ctx_create(device, ctypes.byref(h))
        if rc != 0:
            raise LZ4Error(f"b200lz4_ctx_create({device}) failed ({rc}): {_lib.last_error()}")
        self.handle = h
        self.device = device
        self._pins = {}

    def close(self):
        if self.handle:
            for p, _ in self._pins.values():
                self.lib.b200lz4_host_free(p)
            self._pins = {}
            self.lib.b200lz4_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def pinned(self, name: str, nbytes: int, write_combined: bool = False) -> np.ndarray:
        """A reusable page-locked uint8 buffer of at least nbytes (write_combined: for input the host only writes)."""
        cur = self._pins.get(name)
        if cur is None or cur[1].size < nbytes:
            if cur is not None:
                self.lib.b200lz4_host_free(cur[0])
            cap = max(int(nbytes * 1.25) + 4096, 1 << 16)
            p = self.lib.b200lz4_host_alloc_wc(cap) if write_combined else self.lib.b200lz4_ctx_host_alloc(self.handle, cap)
            if not p:
                raise LZ4Error("b200lz4_host_alloc failed: " + _lib.last_error())
            arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint8)), shape=(cap,))
            self._pins[name] = cur = (p, arr)
        return cur[1]

    def timing(self):
        a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        self.lib.b200lz4_last_timing(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"h2d_ms": a.value, "kernel_ms": b.value, "d2h_ms": c.value}

    def launch_count(self) -> int:
        return int(self.lib.b200lz4_launch_count(self.handle))

    def last_error(self) -> str:
        return self.lib.b200lz4_ctx_last_error(self.handle).decode("utf-8", "replace")

    def copy_probe(self, h_src: np.ndarray, h2d_bytes: int, h_dst: np.ndarray, d2h_bytes: int, check: bool = False) -> int:
        """One plain H2D + D2H copy pair on the ctx's copy streams (measurement aid, no kernels)."""
        rc = self.lib.b200lz4_copy_probe(self.handle, h_src.ctypes.data, h2d_bytes, h_dst.ctypes.data, d2h_bytes)
        if check and rc != 0:
            raise LZ4Error("b200lz4_copy_probe: " + _lib.last_error())
        return rc

    # ---- raw batch calls on numpy buffers --------------------------------
    def compress_batch(self, src: np.ndarray, src_off: np.ndarray, src_len: np.ndarray,
                       accel: int, header: int, dst: np.ndarray,
                       stream_first: Optional[np.ndarray] = None, streams: Optional[Sequence] = None):
        n = len(src_len)
        dst_off = np.zeros(n + 1, dtype=np.int64)
        out_len = np.zeros(n, dtype=np.int32)
        sf = None if stream_first is None else np.ascontiguousarray(stream_first, dtype=np.int32)
        ns = 0 if sf is None else len(sf) - 1
        sh = None
        if streams is not None:
            sh = (ctypes.c_void_p * ns)(*[s.handle for s in streams])
        rc = self.lib.b200lz4_compress_batch(
            self.handle, src.ctypes.data, src.size, src_off.ctypes.data, src_len.ctypes.data, n,
            None if sf is None else sf.ctypes.data, ns, sh, accel, header,
            dst.ctypes.data, dst.size, dst_off.ctypes.data, out_len.ctypes.data)
        return rc, dst_off, out_len

    def decompress_batch(self, src: np.ndarray, src_off: np.ndarray, src_len: np.ndarray,
                         header: int, max_block: int, dst: np.ndarray,
                         stream_first: Optional[np.ndarray] = None, streams: Optional[Sequence] = None):
        n = len(src_len)
        dst_off = np.zeros(n + 1, dtype=np.int64)
        out_len = np.zeros(n, dtype=np.int32)
        sf = None if stream_first is None else np.ascontiguousarray(stream_first, dtype=np.int32)
        ns = 0 if sf is None else len(sf) - 1
        sh = None
        if streams is not None:
            sh = (ctypes.c_void_p * ns)(*[s.handle for s in streams])
        rc = self.lib.b200lz4_decompress_batch(
            self.handle, src.ctypes.data, src.size, src_off.ctypes.data, src_len.ctypes.data, n,
            None if sf is None else sf.ctypes.data, ns, sh, header, max_block,
            dst.ctypes.data, dst.size, dst_off.ctypes.data, out_len.ctypes.data)
        return rc, dst_off, out_len


class CompressStream:
    """Device-resident LZ4_stream_t (c_createStream / c_freeStream, Internal/LZ4.hs:105-110)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = ctypes.c_void_p()
        rc = ctx.lib.b200lz4_cstream_create(ctx.handle, ctypes.byref(h))
        if rc != 0:
            raise LZ4Error("b200lz4_cstream_create: " + _lib.last_error())
        self.handle = h

    def peek(self):
        table = np.zeros(4096, dtype=np.uint32)
        off = ctypes.c_uint32()
        rc = self.ctx.lib.b200lz4_cstream_peek(self.handle, table.ctypes.data, ctypes.addressof(off))
        if rc != 0:
            raise LZ4Error(_lib.last_error())
        return table, off.value

    def free(self):
        if self.handle:
            self.ctx.lib.b200lz4_cstream_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.free()
        except Exception:
            pass


class DecompressStream:
    """Device-resident LZ4_streamDecode_t (c_createStreamDecode, Internal/LZ4.hs:113-118)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        h = ctypes.c_void_p()
        rc = ctx.lib.b200lz4_dstream_create(ctx.handle, ctypes.byref(h))
        if rc != 0:
            raise LZ4Error("b200lz4_dstream_create: " + _lib.last_error())
        self.handle = h

    def free(self):
        if self.handle:
            self.ctx.lib.b200lz4_dstream_free(self.handle)
            self.handle = None

    def __del__(self):
        try:
            if self.ctx.handle:
                self.free()
        except Exception:
            pass


class MultiContext:
    """b200lz4_mctx: one batch striped over several GPUs of the box by one process (host thread + ctx per device)."""

    def __init__(self, devices: Optional[Sequence[int]] = None, n: int = 0):
        self.lib = _lib.load()
        h = ctypes.c_void_p()
        if devices is not None:
            arr = (ctypes.c_int * len(devices))(*devices)
            rc = self.lib.b200lz4_mctx_create(arr, len(devices), ctypes.byref(h))
        else:
            rc = self.lib.b200lz4_mctx_create(None, n, ctypes.byref(h))
        if rc != 0:
            raise LZ4Error(f"b200lz4_mctx_create failed ({rc}): {_lib.last_error()}")
        self.handle = h
        self.size = int(self.lib.b200lz4_mctx_size(h))

    def close(self):
        if self.handle:
            self.lib.b200lz4_mctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def last_error(self) -> str:
        return self.lib.b200lz4_mctx_last_error(self.handle).decode("utf-8", "replace")

    def timing(self, i: int):
        a, b, c = ctypes.c_float(), ctypes.c_float(), ctypes.c_float()
        self.lib.b200lz4_last_timing(self.lib.b200lz4_mctx_ctx(self.handle, i), ctypes.byref(a), ctypes.byref(b), ctypes.byref(c))
        return {"h2d_ms": a.value, "kernel_ms": b.value, "d2h_ms": c.value}

    def launch_count(self) -> int:
        return sum(int(self.lib.b200lz4_launch_count(self.lib.b200lz4_mctx_ctx(self.handle, i))) for i in range(self.size))

    def compress_batch(self, src, src_off, src_len, accel, header, dst, stream_first=None):
        n = len(src_len)
        dst_off = np.zeros(n + 1, dtype=np.int64)
        out_len = np.zeros(n, dtype=np.int32)
        sf = None if stream_first is None else np.ascontiguousarray(stream_first, dtype=np.int32)
        rc = self.lib.b200lz4_compress_batch_multi(
            self.handle, src.ctypes.data, src.size, src_off.ctypes.data, src_len.ctypes.data, n,
            None if sf is None else sf.ctypes.data, 0 if sf is None else len(sf) - 1, accel, header,
            dst.ctypes.data, dst.size, dst_off.ctypes.data, out_len.ctypes.data)
        return rc, dst_off, out_len

    def decompress_batch(self, src, src_off, src_len, header, max_block, dst, stream_first=None):
        n = len(src_len)
        dst_off = np.zeros(n + 1, dtype=np.int64)
        out_len = np.zeros(n, dtype=np.int32)
        sf = None if stream_first is None else np.ascontiguousarray(stream_first, dtype=np.int32)
        rc = self.lib.b200lz4_decompress_batch_multi(
            self.handle, src.ctypes.data, src.size, src_off.ctypes.data, src_len.ctypes.data, n,
            None if sf is None else sf.ctypes.data, 0 if sf is None else len(sf) - 1, header, max_block,
            dst.ctypes.data, dst.size, dst_off.ctypes.data, out_len.ctypes.data)
        return rc, dst_off, out_len


_default_ctx: Optional[Context] = None


def default_context() -> Context:
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = Context(0)
    return _default_ctx


def _batches(chunks: Iterable, batch_arrays: int, batch_bytes: int):
    group: List[bytes] = []
    size = 0
    for c in chunks:
        b = c if isinstance(c, (bytes, bytearray, memoryview)) else np.ascontiguousarray(c, dtype=np.uint8).reshape(-1)
        group.append(b)
        size += len(b)
        if len(group) >= batch_arrays or size >= batch_bytes:
            yield group
            group, size = [], 0
    if group:
        yield group


def _gather(ctx: Context, name: str, arrays: Sequence):
    """Lay the arrays of one batch out in a page-locked buffer (b200lz4_gather_host: parallel memcpy in the library)."""
    n = len(arrays)
    views = [a if isinstance(a, np.ndarray) else np.frombuffer(a, dtype=np.uint8) for a in arrays]
    lens = np.fromiter((v.size for v in views), dtype=np.int32, count=n)
    offs = np.zeros(n, dtype=np.int64)
    # 16-byte aligned starts; a gap keeps consecutive arrays non-adjacent like separate Haskell arrays
    strides = (lens.astype(np.int64) + 16 + 15) // 16 * 16
    if n > 1:
        offs[1:] = np.cumsum(strides[:-1])
    total = int(strides.sum())
    buf = ctx.pinned(name, total)
    ptrs = np.fromiter((v.ctypes.data for v in views), dtype=np.uint64, count=n)
    rc = ctx.lib.b200lz4_gather_host(buf.ctypes.data, ptrs.ctypes.data, offs.ctypes.data, lens.ctypes.data, n, 0)
    if rc != 0:
        raise LZ4Error("b200lz4_gather_host: " + _lib.last_error())
    return buf[:total], offs, lens


def _staged(ctx: Context, name: str, groups):
    """Yield (group, src, offs, lens) with the gather of the NEXT group running on a helper thread while the caller
    works on the current one (two alternating pinned buffers; ctypes releases the GIL inside the library)."""
    it = iter(groups)
    slot = 0
    box = {}

    def stage(g, k):
        try:
            box["out"] = (g,) + _gather(ctx, f"{name}{k}", g)
        except BaseException as e:      # re-raised on the consumer side
            box["err"] = e

    def start():
        nonlocal slot
        g = next(it, None)
        if g is None:
            return None
        t = threading.Thread(target=stage, args=(g, slot))
        slot ^= 1
        t.start()
        return t

    t = start()
    while t is not None:
        t.join()
        if "err" in box:
            raise box.pop("err")
        cur = box.pop("out")
        t = start()
        yield cur


def compress_chunks(cfg: BlockConfig, speed: int, chunks: Iterable, *, ctx: Optional[Context] = None,
                    batch_arrays: int = 4096, batch_bytes: int = 256 << 20, copy: bool = True) -> Iterator[bytes]:
    """compressChunks cfg speed (LZ4.hs:94-100): each input array becomes one framed LZ4 block.

    Linked mode (default, like the reference): one device stream state for the whole
    stream, arrays strictly in order.  cfg.independent: fresh state per array.
    copy=False yields numpy views into the pinned result buffer instead of bytes objects (what the Haskell shim does
    with array slices); a view stays valid until the batch after the next one has been produced.
    """
    ctx = ctx or default_context()
    speed = max(int(speed), 0)                                   # Internal/LZ4.hs:364
    header = cfg.meta_size
    stream = None if cfg.independent else CompressStream(ctx)    # CompressInit, Internal/LZ4.hs:367-376

    def checked(groups):
        for group in groups:
            for a in group:
                if len(a) >= 2 * 1024 * 1024 * 1024:             # Internal/LZ4.hs:384-385
                    raise LZ4Error("compressChunksD: Array element > 2 GB encountered")
                if len(a) > cfg.max_block_size:                  # Internal/LZ4.hs:237-241
                    raise LZ4Error(f"compressChunk: Source array length {len(a)} exceeds the maximum block size "
                                   f"of {cfg.max_block_size}")
            yield group
    try:
        k = 0
        for group, src, offs, lens in _staged(ctx, "c_src", checked(_batches(chunks, batch_arrays, batch_bytes))):
            cap = int((lens.astype(np.int64) + lens // 255 + 16 + header).sum())
            dst = ctx.pinned(f"c_dst{k}", cap)
            k ^= 1
            sf = None if stream is None else np.array([0, len(group)], dtype=np.int32)
            rc, dst_off, out_len = ctx.compress_batch(src, offs, lens, speed, header, dst, sf,
                                                      None if stream is None else [stream])
            if rc != 0:                                          # Internal/LZ4.hs:257-260
                raise LZ4Error(f"compressChunk: c_compressFastContinue failed ({rc}): {ctx.last_error()}")
            for i in range(len(group)):
                blk = dst[dst_off[i]:dst_off[i + 1]]
                yield blk.tobytes() if copy else blk
    finally:
        if stream is not None:
            stream.free()                                        # CompressDone, Internal/LZ4.hs:393-394


def _decode_batches(chunks: Iterable, batch_arrays: int, batch_bytes: int, cap_of, out_budget: int):
    """Group framed arrays so that neither the compressed bytes nor the WORST-CASE output of a batch exceed their
    budgets (BlockMax* headers carry no size: every block may expand to the configured maximum)."""
    group: List = []
    size = out = 0
    for c in chunks:
        b = c if isinstance(c, (bytes, bytearray, memoryview)) else np.ascontiguousarray(c, dtype=np.uint8).reshape(-1)
        cap = cap_of(b)
        if group and (out + cap > out_budget):
            yield group
            group, size, out = [], 0, 0
        group.append(b)
        size += len(b)
        out += cap
        if len(group) >= batch_arrays or size >= batch_bytes:
            yield group
            group, size, out = [], 0, 0
    if group:
        yield group


def decompress_chunks_raw(cfg: BlockConfig, chunks: Iterable, *, ctx: Optional[Context] = None,
                          batch_arrays: int = 4096, batch_bytes: int = 256 << 20, out_budget: int = 1 << 30,
                          copy: bool = True) -> Iterator[bytes]:
    """decompressChunksRawD (Internal/LZ4.hs:539-567): every input array is exactly one framed block.
    A batch is flushed as soon as its worst-case output would exceed out_budget bytes (1 GiB)."""
    ctx = ctx or default_context()
    header = cfg.meta_size
    max_block = 0 if cfg.block_size is BlockSize.BlockHasSize else cfg.block_size.value
    stream = None if cfg.independent else DecompressStream(ctx)

    def cap_of(a) -> int:
        if header == 8:
            return max(int.from_bytes(bytes(a[4:8]), "little", signed=True), 0) if len(a) >= 8 else 0
        return max_block + 16
    try:
        k = 0
        for group, src, offs, lens in _staged(ctx, "d_src", _decode_batches(chunks, batch_arrays, batch_bytes // 2, cap_of, out_budget)):
            cap = sum(cap_of(a) for a in group)
            dst = ctx.pinned(f"d_dst{k}", cap + 64)
            k ^= 1
            sf = None if stream is None else np.array([0, len(group)], dtype=np.int32)
            rc, dst_off, out_len = ctx.decompress_batch(src, offs, lens, header, max_block, dst, sf,
                                                        None if stream is None else [stream])
            if rc != 0:                                          # Internal/LZ4.hs:309-330
                raise LZ4Error(f"decompressChunk: c_decompressSafeContinue failed ({rc}): {ctx.last_error()}")
            for i in range(len(group)):
                blk = dst[dst_off[i]:dst_off[i] + out_len[i]]
                yield blk.tobytes() if copy else blk
    finally:
        if stream is not None:
            stream.free()


def resize_chunks(cfg: BlockConfig, frame_cfg: FrameConfig, chunks: Iterable) -> Iterator[bytes]:
    """resizeChunksD (Internal/LZ4.hs:432-523): re-frame an arbitrarily fragmented compressed
    stream into one array per [header][block].  The header walk runs in b200lz4_reframe."""
    lib = _lib.load()
    header = cfg.meta_size
    buf = bytearray()
    max_blocks = 1 << 16
    off = np.zeros(max_blocks, dtype=np.int64)
    ln = np.zeros(max_blocks, dtype=np.int32)
    nf, used, ended = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int()

    def drain():
        while True:
            raw = (ctypes.c_char * len(buf)).from_buffer(buf) if buf else None
            rc = lib.b200lz4_reframe(raw, len(buf), header, int(frame_cfg.has_end_mark),
                                     off.ctypes.data, ln.ctypes.data, max_blocks,
                                     ctypes.byref(nf), ctypes.byref(used), ctypes.byref(ended))
            del raw
            if rc != 0:
                raise LZ4Error("resizeChunksD: " + _lib.last_error())
            out = [bytes(buf[off[k]:off[k] + ln[k]]) for k in range(nf.value)]
            del buf[:used.value]
            yield from out
            if ended.value or nf.value < max_blocks:
                return

    for c in chunks:
        buf += bytes(c)
        yield from drain()
        if ended.value:
            return                                               # RFooter -> Stop, Internal/LZ4.hs:506-522
    if buf:
        raise LZ4Error("resizeChunksD: Incomplete block")        # RAccumulate + Stop, Internal/LZ4.hs:505
    if frame_cfg.has_end_mark:
        raise LZ4Error("resizeChunksD: No end mark found")       # RInit + Stop, Internal/LZ4.hs:493-496


def decompress_chunks(cfg: BlockConfig, chunks: Iterable, *, ctx: Optional[Context] = None, **kw) -> Iterator[bytes]:
    """decompressChunks (LZ4.hs:114-122) = decompressChunksRawD . resizeChunksD."""
    return decompress_chunks_raw(cfg, resize_chunks(cfg, default_frame_config, chunks), ctx=ctx, **kw)


# ---------------------------------------------------------------------------------------------------------------
# LZ4 frame header, as far as the reference goes (Internal/LZ4.hs:569-651)

FRAME_MAGIC = 407708164                       # 0x184D2204, little endian on the stream (Internal/LZ4.hs:610)
_BD_CODES = {4: BlockSize.BlockMax64KB, 5: BlockSize.BlockMax256KB, 6: BlockSize.BlockMax1MB, 7: BlockSize.BlockMax4MB}


def simple_frame_parser(header: bytes, *, allow_independent: bool = False):
    """simpleFrameParserD (Internal/LZ4.hs:590-651) over the 7 header bytes [magic LE32][FLG][BD][HC]:
    returns (BlockConfig, FrameConfig) with hasEndMark = True; same rejections, same messages.
    allow_independent: accept the block-independence flag (the reference dies on it, :631-632; this codec supports it)."""
    if len(header) < 7:
        raise LZ4Error("simpleFrameParserD: input ended inside the frame header")
    magic = int.from_bytes(header[0:4], "little")
    if magic != FRAME_MAGIC:                                                    # :604-620
        raise LZ4Error(f"The parsed magic {magic} does not match {FRAME_MAGIC}")
    flg = header[4]
    if not (not (flg >> 7) & 1 and (flg >> 6) & 1):                            # :624
        raise LZ4Error("Version is not 01")
    independent = bool((flg >> 5) & 1)
    if independent and not allow_independent:
        raise LZ4Error("Block independence is not yet supported")               # :631-632
    if (flg >> 4) & 1:
        raise LZ4Error("Block checksum is not yet supported")
    if (flg >> 3) & 1:
        raise LZ4Error("Content size is not yet supported")
    if (flg >> 2) & 1:
        raise LZ4Error("Content checksum is not yet supported")
    if flg & 1:
        raise LZ4Error("Dict is not yet supported")
    bs = _BD_CODES.get(header[5] >> 4)                                           # :643-650
    if bs is None:
        raise LZ4Error("parseBD: Unknown block max size")
    return BlockConfig(block_size=bs, independent=independent), FrameConfig(has_end_mark=True)   # header checksum: any byte, :602


def frame_header(block_size: BlockSize, *, independent: bool = False) -> bytes:
    """The 7 bytes benchmark/Main.hs:92-100 writes in front of a framed stream (FLG = version 01, BD = block maximum, HC = 0)."""
    code = {v: k for k, v in _BD_CODES.items()}[block_size]
    return FRAME_MAGIC.to_bytes(4, "little") + bytes([0x40 | (0x20 if independent else 0), code << 4, 0])


def compress_chunks_frame(cfg: BlockConfig, frame_cfg: FrameConfig, speed: int, chunks: Iterable, **kw) -> Iterator[bytes]:
    """compressChunksFrame of benchmark/Main.hs:105-118: compressChunksD, then the 4-byte end mark if the frame has one."""
    yield from compress_chunks(cfg, speed, chunks, **kw)
    if frame_cfg.has_end_mark:
        yield b"\x00\x00\x00\x00"


def decompress_chunks_with(parser, chunks: Iterable, *, ctx: Optional[Context] = None, **kw) -> Iterator[bytes]:
    """decompressChunksWithD (Internal/LZ4.hs:569-577): run `parser` over the first 7 bytes of the stream, then
    decompressChunksRawD cfg . resizeChunksD cfg frameCfg on what follows."""
    it = iter(chunks)
    head = b""
    rest = None
    for c in it:
        head += bytes(c)
        if len(head) >= 7:
            rest = head[7:]
            break
    cfg, frame_cfg = parser(head[:7])

    def tail():
        if rest:
            yield rest
        yield from it
    return decompress_chunks_raw(cfg, resize_chunks(cfg, frame_cfg, tail()), ctx=ctx, **kw)
