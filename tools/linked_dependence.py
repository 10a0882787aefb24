#!/usr/bin/env python
"""CPU-only analysis behind the wide decoder's design: how do the matches of LINKED 64 KiB blocks
depend on earlier output?  For each data kind (accel 1, the reference's linked mode) it reports, over the oracle's
compressed stream:
  * sequences per block, bytes per sequence, share of sequences the parser's lane-parallel window takes;
  * share of matches whose source starts within the last 350 / 700 / 4096 bytes (batches in flight in the copier pipeline);
  * share of matches that reach into the previous block (the dictionary), and the share of a block's BYTES that depend on the
    dictionary transitively (a block-level wavefront -- start block N's copies before block N-1 is complete -- could run
    only the rest ahead).
Writes a table to stdout (kept as profiles/r2_linked_dependence.txt)."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.oracle import Oracle
from streamly_lz4_b200 import datagen


def parse(payload):
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
            seqs.append((lit, 0, 0)); break
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
    ora = Oracle("auto")
    bs, nblocks = 65536, 24
    print(f"{'kind':9s} {'seq/blk':>8s} {'B/seq':>6s} {'window':>7s} {'<350B':>6s} {'<700B':>6s} {'<4KiB':>6s} {'dict':>6s} {'bytes dep. on dict':>19s}")
    for kind in ("text", "mixed", "sparse01", "records"):
        d = datagen.make(kind, 4, bs * nblocks)
        arrays = [d[i:i + bs].tobytes() for i in range(0, d.size, bs)]
        out = ora.compress_chunks(arrays, 1, linked=True)
        nseq = nm = win = n350 = n700 = n4k = ndict = 0
        dep_bytes = tot_bytes = 0
        for k, o in enumerate(out):
            seqs = parse(o[8:])
            nseq += len(seqs)
            tainted = np.zeros(bs + 1, dtype=np.uint8)            # byte depends (transitively) on the previous block
            op = 0
            for lit, ml, dist in seqs:
                if lit <= 32 and (ml == 0 or ml <= 64):
                    win += 1
                op += lit
                if ml:
                    nm += 1
                    n350 += dist < 350; n700 += dist < 700; n4k += dist < 4096
                    frm = op - dist
                    if frm < 0:
                        ndict += 1
                    for i in range(ml):                            # byte-exact propagation (overlapping matches included)
                        s = frm + i
                        tainted[op + i] = 1 if s < 0 else tainted[s]
                    op += ml
            if k:
                dep_bytes += int(tainted[:op].sum()); tot_bytes += op
        print(f"{kind:9s} {nseq / len(out):8.0f} {bs * len(out) / nseq:6.1f} {win / nseq:7.3f} {n350 / nm:6.3f} {n700 / nm:6.3f} {n4k / nm:6.3f} "
              f"{ndict / nm:6.3f} {dep_bytes / max(tot_bytes, 1):19.3f}")


if __name__ == "__main__":
    main()
