#!/usr/bin/env python
"""Driver of the CPU prototype tools/specparse_proto.c (speculate / verify / repair: a byte-exact LZ4 block parse split
across workers).  Compresses blocks of every data kind with the prototype, compares the bytes with the oracle's, and
prints how much of each block was left to the serial part.  Kept as profiles/r2_specparse_proto.txt.
usage: python tools/specparse_proto.py [--block 4194304] [--seg 262144] [--warm 131072] [--accels 1,400]"""
import argparse, ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so = os.path.join(ROOT, "build", "libspecparse.so")
    src = os.path.join(ROOT, "tools", "specparse_proto.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, src])
    lib = ctypes.CDLL(so)
    lib.specparse_compress.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p]
    lib.specparse_compress.restype = ctypes.c_int
    lib.specparse_compress_linked.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.specparse_compress_linked.restype = ctypes.c_int
    return lib


def compress(lib, block: np.ndarray, accel: int, seg: int, warm: int):
    n = int(block.size)
    dst = np.zeros(n + n // 255 + 64, dtype=np.uint8)
    stats = np.zeros(24, dtype=np.int64)
    blk = np.ascontiguousarray(block)
    m = lib.specparse_compress(blk.ctypes.data, n, accel, seg, warm, dst.ctypes.data, stats.ctypes.data)
    return dst[:m].tobytes(), stats


def compress_linked(lib, stream: np.ndarray, sizes, accel: int, warm_blocks: int):
    """one linked stream: `sizes` = the array lengths, `stream` = the arrays back to back"""
    off = np.zeros(len(sizes) + 1, dtype=np.int32)
    off[1:] = np.cumsum(sizes)
    assert int(off[-1]) == stream.size
    slot = [int(x) + int(x) // 255 + 64 for x in sizes]
    dst_off = np.zeros(len(sizes), dtype=np.int64)
    dst_off[1:] = np.cumsum(slot)[:-1]
    dst = np.zeros(sum(slot), dtype=np.uint8)
    out_len = np.zeros(len(sizes), dtype=np.int32)
    stats = np.zeros(24, dtype=np.int64)
    buf = np.ascontiguousarray(stream)
    lib.specparse_compress_linked(buf.ctypes.data, off.ctypes.data, len(sizes), accel, warm_blocks,
                                  dst.ctypes.data, dst_off.ctypes.data, out_len.ctypes.data, stats.ctypes.data)
    return [dst[int(dst_off[i]):int(dst_off[i]) + int(out_len[i])].tobytes() for i in range(len(sizes))], stats


def row(kind, accel, a, b, exact, k, tot):
    # critical path in table accesses (one access = one step of the finder's serial chain): the busiest worker, then the serial
    # part of phase 2; its data-parallel parts (log scans, replays, 4096-bucket compares) at 1/32 (one warp) each
    crit = tot[13] / k + tot[14] / k + (tot[7] + tot[8]) / k / 32 + (tot[5] + tot[20]) / k * 4096 / 32
    return (f"{kind:9s} {accel:5d} {a:7d} {b:7d} | {exact:>5s} | {tot[0] // k:7d} {tot[12] / max(1, tot[15]):6.2f} {tot[13] // k:14d} | {tot[3] // k:12d} {tot[14] // k:10d} {tot[6] // k:7d} "
            f"{tot[7] // k:11d} {tot[8] // k:12d} | {tot[15] // k:9d} {crit:9.0f} {tot[15] / k / crit:8.1f}")


def main():
    from oracle.oracle import Oracle
    from streamly_lz4_b200 import datagen
    ap = argparse.ArgumentParser()
    ap.add_argument("--block", type=int, default=4194304)
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--seg", default="262144,131072")
    ap.add_argument("--warm", default="65536,131072")
    ap.add_argument("--accels", default="1,400")
    ap.add_argument("--kinds", default="text,mixed,records,sparse01,bits01,biased01,random,zero")
    ap.add_argument("--linked-streams", type=int, default=2)
    ap.add_argument("--linked-block", type=int, default=65536)
    ap.add_argument("--linked-blocks", type=int, default=64)
    ap.add_argument("--warm-blocks", default="1,2,3")
    args = ap.parse_args()
    lib = build()
    ora = Oracle("auto")
    n = args.block
    print(f"blocks of {n} B, {args.blocks} per kind, independent (fresh state per block); every output compared with oracle[{ora.kind}]")
    print(f"{'kind':9s} {'accel':>5s} {'seg':>7s} {'warm':>7s} | {'exact':>5s} | {'workers':>7s} {'work x':>6s} {'busiest worker':>14s} | {'serial bytes':>12s} {'serial acc':>10s} {'diverg.':>7s} "
          f"{'log scanned':>11s} {'log replayed':>12s} | {'plain acc':>9s} {'crit path':>9s} {'speed-up':>8s}")
    for kind in args.kinds.split(","):
        data = datagen.make(kind, 11, n * args.blocks)
        for accel in [int(a) for a in args.accels.split(",")]:
            want = [ora.compress_chunks([data[b * n:(b + 1) * n].tobytes()], accel, linked=False)[0][8:] for b in range(args.blocks)]
            for seg in [int(x) for x in args.seg.split(",")]:
                for warm in [int(x) for x in args.warm.split(",")]:
                    tot = np.zeros(24, dtype=np.int64)
                    exact = 0
                    for b in range(args.blocks):
                        got, st = compress(lib, data[b * n:(b + 1) * n], accel, seg, warm)
                        exact += got == want[b]
                        tot += st
                    k = args.blocks
                    print(row(kind, accel, seg, warm, f"{exact}/{k}", k, tot))
    if args.linked_streams:
        bs, nb = args.linked_block, args.linked_blocks
        print()
        print(f"LINKED streams (one LZ4_stream_t, dictionary = previous block): {args.linked_streams} per kind, {nb} blocks of {bs} B, one worker per block, warm-up in blocks")
        print(f"{'kind':9s} {'accel':>5s} {'block':>7s} {'warm':>7s} | {'exact':>5s} | {'workers':>7s} {'work x':>6s} {'busiest worker':>14s} | {'serial bytes':>12s} {'serial acc':>10s} {'diverg.':>7s} "
              f"{'log scanned':>11s} {'log replayed':>12s} | {'plain acc':>9s} {'crit path':>9s} {'speed-up':>8s}")
        for kind in args.kinds.split(","):
            data = datagen.make(kind, 13, bs * nb * args.linked_streams)
            for accel in [int(a) for a in args.accels.split(",")]:
                for wb in [int(x) for x in args.warm_blocks.split(",")]:
                    tot = np.zeros(24, dtype=np.int64)
                    exact = 0
                    for sidx in range(args.linked_streams):
                        stream = data[sidx * bs * nb:(sidx + 1) * bs * nb]
                        arrays = [stream[i * bs:(i + 1) * bs].tobytes() for i in range(nb)]
                        want = [o[8:] for o in ora.compress_chunks(arrays, accel, linked=True)]
                        got, st = compress_linked(lib, stream, [bs] * nb, accel, wb)
                        exact += got == want
                        tot += st
                    k = args.linked_streams
                    print(row(kind, accel, bs, wb, f"{exact}/{k}", k, tot))
    print("per block: work x = table accesses of all workers / accesses of the plain serial parse (warm-ups are the overhead); busiest worker = its accesses;")
    print("serial bytes / acc = what the in-order phase had to parse itself (sync steps, repairs); diverg. = verifications that found a deciding difference;")
    print("crit path = busiest worker + serial acc + (log entries scanned + replayed + 4096 per full table compare and per final-table merge) / 32, in table accesses")
    print("(= steps of a finder's chain); log replayed counts only partial take-overs (up to a divergence): a unit that verifies to its end is merged from the worker's final table;")
    print("speed-up = accesses of the plain serial parse / crit path (match-length counting is not in this unit: it is warp-parallel already)")


if __name__ == "__main__":
    main()
