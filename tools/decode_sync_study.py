#!/usr/bin/env python
"""CPU-only study for the decode side of "few large blocks" (DESIGN.md section 7): a block has ONE token chain, so one
parser warp per block is the limit (256 x 4 MiB: 40 GB/s).  Can several parsers share a block?  A parser started at an
arbitrary byte of the compressed block reads garbage as tokens at first; this measures how soon its chain of token
positions meets the TRUE chain (from then on it is the true parse: the token chain is a pure function of the position).
Per data kind: 4 MiB blocks compressed by the oracle (acceleration 1), speculative starts at every 1/16 of the
compressed size (+ small shifts), and for each the compressed bytes and sequences until the first common token position.
Writes a table to stdout (kept as profiles/r2_decode_sync_study.txt)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from streamly_lz4_b200 import datagen


def chain(payload: bytes, start: int, stop: int, true_set=None):
    """token positions visited from `start` (stops at `stop`, at the block's end, or at the first true token position)"""
    n, ip, steps = len(payload), start, 0
    pos = []
    while ip < n and ip < stop:
        if true_set is not None and ip in true_set:
            return pos, ip, steps
        pos.append(ip)
        tok = payload[ip]; ip += 1
        lit = tok >> 4
        if lit == 15:
            while ip < n:
                b = payload[ip]; ip += 1; lit += b
                if b != 255:
                    break
        ip += lit
        if ip + 2 > n:
            break
        ip += 2
        if (tok & 15) == 15:
            while ip < n:
                b = payload[ip]; ip += 1
                if b != 255:
                    break
        steps += 1
    return pos, -1, steps


def main():
    ora = Oracle("auto")
    n, nblocks, parts = 4 << 20, 2, 16
    print(f"blocks of {n} B compressed by oracle[{ora.kind}] at acceleration 1; speculative parser starts at k/{parts} of the compressed size, shifted by 0..6 bytes")
    print(f"{'kind':9s} {'ratio':>6s} {'B/seq':>6s} | {'starts':>6s} {'synced':>6s} | bytes to sync: {'median':>7s} {'p90':>7s} {'max':>8s} | sequences to sync: {'median':>6s} {'p90':>6s} {'max':>7s} | {'share of a 1/16 segment lost (p90)':>34s}")
    for kind in ("text", "mixed", "records", "sparse01", "bits01", "biased01"):
        data = datagen.make(kind, 17, n * nblocks)
        dist_b, dist_s, starts, synced, seg_len, ratio, bps = [], [], 0, 0, [], [], []
        for b in range(nblocks):
            payload = ora.compress_chunks([data[b * n:(b + 1) * n].tobytes()], 1, linked=False)[0][8:]
            true_pos, _, nseq = chain(payload, 0, len(payload))
            tset = set(true_pos)
            ratio.append(n / len(payload)); bps.append(len(payload) / max(1, nseq))
            for k in range(1, parts):
                for shift in range(7):
                    x = len(payload) * k // parts + shift
                    starts += 1
                    _, hit, steps = chain(payload, x, len(payload), tset)
                    if hit >= 0:
                        synced += 1
                        dist_b.append(hit - x); dist_s.append(steps)
                    seg_len.append(len(payload) / parts)
        db, ds = np.array(dist_b), np.array(dist_s)
        print(f"{kind:9s} {np.mean(ratio):6.2f} {np.mean(bps):6.1f} | {starts:6d} {synced:6d} |                {np.median(db):7.0f} {np.percentile(db, 90):7.0f} {db.max():8d} |"
              f"                    {np.median(ds):6.0f} {np.percentile(ds, 90):6.0f} {ds.max():7d} | {100 * np.percentile(db, 90) / np.mean(seg_len):33.2f}%")
    print("a speculative parser only has to record its token positions; the in-order stitcher (the wide decoder's dispatcher) takes over a parser's batches from the")
    print("first position the previous parser's true chain shares with it, and re-parses nothing: what is lost is the bytes to sync, once per segment")


if __name__ == "__main__":
    main()
