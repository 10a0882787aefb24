"""ctypes front-end of the CPU checker (TEST INFRASTRUCTURE ONLY).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs may import this module.  The product package
(streamly_lz4_b200) never does.

Two back-ends with the same call-sequence driver (oracle/ref_driver.c):
  kind="reference"  oracle/_ref/libreflz4.so : the reference's own cbits/lz4.c (LZ4 1.9.3)
  kind="port"       oracle/liboracle.so      : the restatement in oracle/lz4_oracle.c

Host-side logic of the reference that is *not* in C is restated here in plain
Python, each function citing the Haskell it follows:
  frame / unframe        src/Streamly/Internal/LZ4.hs:177-207 (header layout)
  resize_chunks          src/Streamly/Internal/LZ4.hs:413-523 (resizeChunksD)
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Iterable, List, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_PATHS = {"reference": os.path.join(_HERE, "_ref", "libreflz4.so"),
          "port": os.path.join(_HERE, "liboracle.so")}

BLOCK_MAX = {"BlockHasSize": 0x7E000000, "BlockMax64KB": 64 << 10, "BlockMax256KB": 256 << 10,
             "BlockMax1MB": 1 << 20, "BlockMax4MB": 4 << 20}   # LZ4.hs:275-281


def meta_size(block_size: str) -> int:
    """metaSize / dataOffset, LZ4.hs:177-204."""
    return 8 if block_size == "BlockHasSize" else 4


def build(force: bool = False) -> None:
    """Compile the checker (gcc).  `_ref` is rebuilt only where /root/reference exists."""
    if force or not os.path.exists(_PATHS["port"]) or (
            os.path.exists("/root/reference/cbits/lz4.c") and not os.path.exists(_PATHS["reference"])):
        subprocess.check_call(["make", "-s", "-C", _HERE])


def available(kind: str) -> bool:
    return os.path.exists(_PATHS[kind])


_u8p = ctypes.POINTER(ctypes.c_uint8)


class Oracle:
    def __init__(self, kind: str = "reference"):
        if kind == "auto":
            kind = "reference" if available("reference") else "port"
        self.kind = kind
        path = _PATHS[kind]
        if not os.path.exists(path):
            build()
        self.lib = lib = ctypes.CDLL(path)
        assert lib.drv_kind() == (1 if kind == "reference" else 0)
        vp, ip = ctypes.c_void_p, ctypes.c_int
        lib.drv_compress.argtypes = [vp, vp, ip, vp, ip, vp, vp, vp, ip, ip, ip, ip]
        lib.drv_compress.restype = ip
        lib.drv_decompress.argtypes = [vp, vp, ip, vp, ip, vp, vp, vp, ip, ip, ip]
        lib.drv_decompress.restype = ip
        if kind == "reference":
            n = {"ccreate": "LZ4_createStream", "cfree": "LZ4_freeStream", "dcreate": "LZ4_createStreamDecode",
                 "dfree": "LZ4_freeStreamDecode", "bound": "LZ4_compressBound",
                 "ccont": "LZ4_compress_fast_continue", "dcont": "LZ4_decompress_safe_continue"}
        else:
            n = {"ccreate": "ora_cstream_create", "cfree": "ora_cstream_free", "dcreate": "ora_dstream_create",
                 "dfree": "ora_dstream_free", "bound": "ora_compress_bound",
                 "ccont": "ora_compress_continue", "dcont": "ora_decompress_continue"}
        self.ccreate = getattr(lib, n["ccreate"]); self.ccreate.restype = vp; self.ccreate.argtypes = []
        self.cfree = getattr(lib, n["cfree"]); self.cfree.argtypes = [vp]
        self.dcreate = getattr(lib, n["dcreate"]); self.dcreate.restype = vp; self.dcreate.argtypes = []
        self.dfree = getattr(lib, n["dfree"]); self.dfree.argtypes = [vp]
        self.bound = getattr(lib, n["bound"]); self.bound.restype = ip; self.bound.argtypes = [ip]
        self.ccont = getattr(lib, n["ccont"]); self.ccont.restype = ip
        self.ccont.argtypes = [vp, vp, vp, ip, ip, ip]
        self.dcont = getattr(lib, n["dcont"]); self.dcont.restype = ip
        self.dcont.argtypes = [vp, vp, vp, ip, ip]

    # ---- arena helpers: arrays of one stream must not be address-adjacent ----
    @staticmethod
    def lay_out(arrays: Sequence, gap: int = 64):
        """Copy arrays into one arena with >= gap bytes between them.
        Returns (arena, ptrs uint64[n], lens int32[n])."""
        lens = np.array([len(a) for a in arrays], dtype=np.int32)
        strides = ((lens.astype(np.int64) + gap + 63) // 64) * 64
        offs = np.zeros(len(arrays) + 1, dtype=np.int64)
        np.cumsum(strides, out=offs[1:])
        arena = np.zeros(int(offs[-1]) + 64, dtype=np.uint8)
        for a, o, n in zip(arrays, offs[:-1], lens):
            if n:
                arena[o:o + n] = np.frombuffer(a, dtype=np.uint8) if not isinstance(a, np.ndarray) else a
        ptrs = (arena.ctypes.data + offs[:-1]).astype(np.uint64)
        return arena, ptrs, lens

    @staticmethod
    def slots(caps: np.ndarray, gap: int = 64):
        strides = ((caps.astype(np.int64) + gap + 63) // 64) * 64
        offs = np.zeros(len(caps) + 1, dtype=np.int64)
        np.cumsum(strides, out=offs[1:])
        arena = np.zeros(int(offs[-1]) + 64, dtype=np.uint8)
        ptrs = (arena.ctypes.data + offs[:-1]).astype(np.uint64)
        return arena, ptrs, offs

    @staticmethod
    def _streams(n: int, linked, stream_first):
        if stream_first is not None:
            return np.ascontiguousarray(stream_first, dtype=np.int32)
        if linked:
            return np.array([0, n], dtype=np.int32)
        return np.arange(n + 1, dtype=np.int32)

    # ---- array-level API mirroring compressChunks / decompressChunksRaw ----
    def compress_ptrs(self, ptrs, lens, dst_ptrs, dst_caps, out_len, accel, meta, stream_first,
                      max_block=0, threads=1) -> int:
        n = len(lens)
        return self.lib.drv_compress(ptrs.ctypes.data, lens.ctypes.data, n, stream_first.ctypes.data,
                                     len(stream_first) - 1, dst_ptrs.ctypes.data, dst_caps.ctypes.data,
                                     out_len.ctypes.data, accel, meta, max_block, threads)

    def decompress_ptrs(self, ptrs, lens, dst_ptrs, dst_caps, out_len, meta, stream_first,
                        max_block=0, threads=1) -> int:
        n = len(lens)
        return self.lib.drv_decompress(ptrs.ctypes.data, lens.ctypes.data, n, stream_first.ctypes.data,
                                       len(stream_first) - 1, dst_ptrs.ctypes.data, dst_caps.ctypes.data,
                                       out_len.ctypes.data, meta, max_block, threads)

    def compress_chunks(self, arrays: Sequence, accel: int = 1, block_size: str = "BlockHasSize",
                        linked: bool = True, stream_first=None, threads: int = 1) -> List[bytes]:
        """compressChunks cfg speed (LZ4.hs:353-394): one framed array per input array."""
        accel = max(accel, 0)                                   # LZ4.hs:364
        meta = meta_size(block_size)
        n = len(arrays)
        if n == 0:
            return []
        arena, ptrs, lens = self.lay_out(arrays)
        caps = (lens.astype(np.int64) + lens // 255 + 16 + meta).astype(np.int32)
        darena, dptrs, doffs = self.slots(caps)
        out_len = np.zeros(n, dtype=np.int32)
        sf = self._streams(n, linked, stream_first)
        rc = self.compress_ptrs(ptrs, lens, dptrs, caps, out_len, accel, meta, sf,
                                0 if block_size == "BlockHasSize" else BLOCK_MAX[block_size], threads)
        if rc:
            raise RuntimeError(f"compressChunk failed rc={rc}")
        return [darena[o:o + meta + l].tobytes() for o, l in zip(doffs[:-1], out_len)]

    def decompress_chunks_raw(self, framed: Sequence, block_size: str = "BlockHasSize",
                              linked: bool = True, stream_first=None, threads: int = 1) -> List[bytes]:
        """decompressChunksRawD (LZ4.hs:539-567): input arrays are exactly one framed block each."""
        meta = meta_size(block_size)
        n = len(framed)
        if n == 0:
            return []
        arena, ptrs, lens = self.lay_out(framed)
        if meta == 8:
            caps = np.array([int.from_bytes(bytes(f[4:8]), "little", signed=True) if len(f) >= 8 else 0
                             for f in framed], dtype=np.int32)
            caps = np.maximum(caps, 0)
        else:
            caps = np.full(n, BLOCK_MAX[block_size], dtype=np.int32)
        darena, dptrs, doffs = self.slots(caps)
        out_len = np.zeros(n, dtype=np.int32)
        sf = self._streams(n, linked, stream_first)
        rc = self.decompress_ptrs(ptrs, lens, dptrs, caps, out_len, meta, sf,
                                  BLOCK_MAX[block_size], threads)
        if rc:
            raise RuntimeError(f"decompressChunk failed rc={rc}")
        return [darena[o:o + l].tobytes() for o, l in zip(doffs[:-1], out_len)]


# ---------------------------------------------------------------------------
# resizeChunksD restated (LZ4.hs:413-523).  Pure function over a list of byte
# strings; raises RuntimeError with the reference's error texts.

def resize_chunks(chunks: Iterable[bytes], block_size: str = "BlockHasSize",
                  has_end_mark: bool = False) -> List[bytes]:
    meta = meta_size(block_size)
    footer = 4 if has_end_mark else 0                         # LZ4.hs:404-408
    out: List[bytes] = []
    it = iter(chunks)
    buf = None                                                # None == RInit
    while True:
        if buf is None:                                       # RInit, LZ4.hs:488-496
            try:
                buf = bytes(next(it))
            except StopIteration:
                if has_end_mark:
                    raise RuntimeError("resizeChunksD: No end mark found")
                return out
        # process, LZ4.hs:459-484
        ln = len(buf)
        need_more = False
        if ln < 4:
            need_more = True
        elif has_end_mark and buf[0:4] == b"\0\0\0\0":        # isEndMark, LZ4.hs:452-456
            while len(buf) < footer:                          # RFooter, LZ4.hs:506-522
                try:
                    buf = buf + bytes(next(it))
                except StopIteration:
                    raise RuntimeError("resizeChunksD: Incomplete footer")
            return out                                        # stream stops here
        elif ln <= meta:
            need_more = True
        else:
            comp = int.from_bytes(buf[0:4], "little", signed=True)
            required = comp + meta
            if ln == required:
                out.append(buf); buf = None; continue
            if ln < required:
                need_more = True
            else:
                out.append(buf[:required]); buf = buf[required:]; continue   # RProcess on the rest
        if need_more:                                         # RAccumulate, LZ4.hs:498-505
            try:
                buf = buf + bytes(next(it))
            except StopIteration:
                raise RuntimeError("resizeChunksD: Incomplete block")
