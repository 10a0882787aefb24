import sys, numpy as np
sys.path.insert(0, "/root/repo")
import streamly_lz4_b200 as lz
from streamly_lz4_b200 import datagen
from oracle.oracle import Oracle
ora = Oracle("auto")
ctx = lz.Context(0)
def parse(b):
    ip=0; n=len(b); seqs=[]; op=0
    while ip<n:
        t=b[ip]; ip+=1
        lit=t>>4
        if lit==15:
            while True:
                s=b[ip]; ip+=1; lit+=s
                if s!=255: break
        ip+=lit
        if ip>=n: seqs.append((op,lit,0,0)); break
        off=b[ip]|(b[ip+1]<<8); ip+=2
        ml=t&15
        if ml==15:
            while True:
                s=b[ip]; ip+=1; ml+=s
                if s!=255: break
        seqs.append((op,lit,off,ml+4)); op+=lit+ml+4
    return seqs
for ns, per in ((1, 8 << 20), (4, 8 << 20), (16, 16 << 20)):
    bs = 65536; total = ns * per
    data = datagen.make("mixed", 4, total)
    offs = np.arange(0, total, bs, dtype=np.int64); lens = np.full(len(offs), bs, dtype=np.int32)
    n = len(lens); bps = per // bs
    sf = (np.arange(ns + 1, dtype=np.int64) * bps).astype(np.int32)
    dst = ctx.pinned("dbg", int((lens.astype(np.int64) + lens // 255 + 24).sum()))
    rc, doff, olen = ctx.compress_batch(data, offs, lens, 1, 8, dst, stream_first=sf)
    assert rc == 0
    arrays = [data[o:o + bs].tobytes() for o in offs]
    want = ora.compress_chunks(arrays, 1, linked=True, stream_first=sf, threads=8)
    bad = [i for i in range(n) if dst[doff[i]:doff[i + 1]].tobytes() != want[i]]
    print("streams", ns, "blocks", n, "differ:", len(bad), bad[:8], "block-in-stream", [b % bps for b in bad[:8]])
    if bad:
        i = bad[0]; g = parse(dst[doff[i] + 8:doff[i + 1]].tobytes()); w = parse(want[i][8:])
        for k, (a, b) in enumerate(zip(g, w)):
            if a != b:
                print("  first differing sequence", k, "gpu", a, "ref", b, "prev", g[k - 1] if k else None); break
