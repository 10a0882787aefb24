#!/usr/bin/env python
"""CPU-only study behind DESIGN.md section 7 ("why a block is not split across warps"): a speculative mid-block start of
the greedy parse judged by its OUTPUT.  Builds tools/spec_split_study.c into build/ and runs it over the bench's data
kinds; first validates the instrumented parser against the oracle's compressed bytes.  Writes a table to stdout
(kept as profiles/r2_spec_split_study.txt).  usage: python tools/spec_split_study.py [--block 640000] [--accels 1,400]"""
import argparse, ctypes, os, subprocess, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from streamly_lz4_b200 import datagen


def build():
    os.makedirs(os.path.join(ROOT, "build"), exist_ok=True)
    so = os.path.join(ROOT, "build", "libspecstudy.so")
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-o", so, os.path.join(ROOT, "tools", "spec_split_study.c")])
    lib = ctypes.CDLL(so)
    lib.spec_split_study.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p]
    lib.spec_true_parse.argtypes = [ctypes.c_void_p, ctypes.c_int32, ctypes.c_int, ctypes.c_void_p, ctypes.c_int]
    return lib


def sequences_of(payload):
    ip, n, seqs = 0, len(payload), []
    while ip < n:
        tok = payload[ip]; ip += 1
        lit = tok >> 4
        if lit == 15:
            while True:
                b = payload[ip]; ip += 1; lit += b
                if b != 255:
                    break
        ip += lit
        if ip >= n:
            break
        dist = payload[ip] | (payload[ip + 1] << 8); ip += 2
        ml = tok & 15
        if ml == 15:
            while True:
                b = payload[ip]; ip += 1; ml += b
                if b != 255:
                    break
        seqs.append((lit, ml + 4, dist))
    return seqs


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--block", type=int, default=640000)
    ap.add_argument("--blocks", type=int, default=4)
    ap.add_argument("--accels", default="1,400")
    ap.add_argument("--warms", default="4096,65536,131072")
    ap.add_argument("--kinds", default="text,mixed,records,sparse01,random")
    args = ap.parse_args()
    lib = build()
    ora = Oracle("auto")
    n = args.block
    print(f"block {n} B, {args.blocks} blocks per kind, split points at k/8 of the block (k = 2..7); oracle[{ora.kind}] validates the instrumented parser")
    print(f"{'kind':9s} {'accel':>5s} {'warm':>7s} | {'cases':>5s} {'sync':>5s} {'to sync':>8s} | {'live diff':>9s} {'exact':>6s} {'missing seq':>11s} {'runs':>6s} {'bytes apart':>11s} {'longest':>8s} | {'check: predicted':>16s} {'1st hit':>8s}")
    for kind in args.kinds.split(","):
        data = datagen.make(kind, 7, n * args.blocks)
        for accel in [int(a) for a in args.accels.split(",")]:
            # validation: the instrumented true parse == the sequences in the oracle's bytes
            for b in range(args.blocks):
                blk = np.ascontiguousarray(data[b * n:(b + 1) * n])
                out = np.zeros(3 * (n // 4 + 16), dtype=np.int32)
                cnt = lib.spec_true_parse(blk.ctypes.data, n, accel, out.ctypes.data, n // 4 + 16)
                mine = [tuple(int(x) for x in out[3 * i:3 * i + 3]) for i in range(cnt)]
                ref = sequences_of(ora.compress_chunks([blk.tobytes()], accel, linked=False)[0][8:])
                assert mine == ref, f"instrumented parser differs from the oracle on {kind} accel {accel} block {b}"
            for warm in [int(w) for w in args.warms.split(",")]:
                rows = []
                for b in range(args.blocks):
                    blk = np.ascontiguousarray(data[b * n:(b + 1) * n])
                    for k in range(2, 8):
                        split = n * k // 8
                        r = np.zeros(11, dtype=np.int64)
                        lib.spec_split_study(blk.ctypes.data, n, accel, split, warm, r.ctypes.data)
                        rows.append(r.copy())
                R = np.array(rows)
                ok = R[:, 0] >= 0
                S = R[ok]
                exact = int((S[:, 2] == 0).sum())
                predicted_ok = int(((S[:, 6] == 0) == (S[:, 2] == 0)).sum())
                # does the earliest deciding-differently access coincide with (or precede) the first missing sequence?
                hit = int(((S[:, 2] > 0) & (S[:, 7] >= 0) & (S[:, 7] <= S[:, 8] + 64)).sum())
                nz = max(1, int((S[:, 2] > 0).sum()))
                print(f"{kind:9s} {accel:5d} {warm:7d} | {len(R):5d} {int(ok.sum()):5d} {np.median(S[:, 9]) if len(S) else -1:8.0f} | "
                      f"{np.median(S[:, 5]) if len(S) else -1:9.0f} {exact:3d}/{len(S):<3d}"
                      f"{np.median(S[:, 2]) if len(S) else -1:11.0f} {np.median(S[:, 3]) if len(S) else -1:6.0f} {np.median(S[:, 10]) if len(S) else -1:11.0f} {np.median(S[:, 4]) if len(S) else -1:8.0f} | "
                      f"{predicted_ok:8d}/{len(S):<7d} {hit:4d}/{nz:<3d}")
    print("columns: sync = cases where the speculative parse (started `warm` bytes before the split with an empty table) shares a match end with the true parse at or after the split;")
    print("  to sync = median bytes from the split to that point; live diff = median number of table buckets that differ there and are still within reach;")
    print("  exact = cases whose speculative output after sync equals the true output; missing seq / runs / bytes apart / longest = medians over the cases of: true sequences the")
    print("  speculative parse lacks, separate stretches of them, bytes those stretches cover, the longest stretch;")
    print("  check = cases where 'no live-different bucket decides differently at its first access' agrees with 'exact'; 1st hit = inexact cases where the earliest")
    print("  deciding-differently access lies at or before the first missing sequence")


if __name__ == "__main__":
    main()
