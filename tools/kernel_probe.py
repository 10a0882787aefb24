#!/usr/bin/env python
"""Kernel-resident timing of the codec kernels per data kind / block size / acceleration.
Development aid (not the graded benchmark): python tools/kernel_probe.py [--mib 256] [--kinds ...]"""
import argparse
import ctypes
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from streamly_lz4_b200 import _lib, datagen
    ap = argparse.ArgumentParser()
    ap.add_argument("--mib", type=int, default=256)
    ap.add_argument("--kinds", default="mixed,text,random,sparse01,records,biased01")
    ap.add_argument("--blocks", default="640000")
    ap.add_argument("--accels", default="400,1")
    ap.add_argument("--reps", type=int, default=3)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream()
    sh = ctypes.c_void_p(stream.cuda_stream)
    total = args.mib << 20
    scratch = torch.zeros(lib.b200lz4_scratch_bytes(), dtype=torch.uint8, device=dev)

    def p(t):
        return ctypes.c_void_p(t.data_ptr())

    print(f"{'kind':10s} {'block':>8s} {'accel':>6s} {'ratio':>6s} {'comp ms':>9s} {'comp GB/s':>10s} {'dec ms':>8s} {'dec GB/s':>9s}")
    for kind in args.kinds.split(","):
        host = datagen.make(kind, 2, total)
        d_src = torch.from_numpy(host).to(dev)
        for bs in [int(x) for x in args.blocks.split(",")]:
            offs = np.arange(0, total, bs, dtype=np.int64)
            lens = np.minimum(bs, total - offs).astype(np.int32)
            n = len(lens)
            bound = lens.astype(np.int64) + lens // 255 + 16
            ss = (bound + 8 + 16 + 15) // 16 * 16
            so = np.zeros(n, dtype=np.int64); so[1:] = np.cumsum(ss[:-1])
            d_off = torch.from_numpy(offs).to(dev); d_len = torch.from_numpy(lens).to(dev)
            d_so = torch.from_numpy(so).to(dev)
            d_slots = torch.empty(int(ss.sum()) + 64, dtype=torch.uint8, device=dev)
            d_out = torch.empty(int(ss.sum()) + 64, dtype=torch.uint8, device=dev)
            d_olen = torch.zeros(n, dtype=torch.int32, device=dev)
            d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            d_back = torch.empty(total + 64, dtype=torch.uint8, device=dev)
            d_blen = torch.zeros(n, dtype=torch.int32, device=dev)
            for accel in [int(x) for x in args.accels.split(",")]:
                def comp():
                    rc = lib.b200lz4_compress_dev(p(d_src), p(d_off), p(d_len), n, None, 0, None, p(d_slots), p(d_so), None,
                                                  p(d_olen), accel, 8, p(scratch), sh)
                    assert rc == 0
                comp(); torch.cuda.synchronize()
                e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
                e[0].record(stream)
                for _ in range(args.reps):
                    comp()
                e[1].record(stream); torch.cuda.synchronize()
                cms = e[0].elapsed_time(e[1]) / args.reps
                rc = lib.b200lz4_compact_dev(p(d_slots), p(d_so), p(d_olen), n, 8, p(d_out), p(d_ooff), p(scratch), sh)
                assert rc == 0
                torch.cuda.synchronize()
                ctot = int(d_ooff[-1].item())
                c_off = d_ooff[:-1].contiguous(); c_len = (d_ooff[1:] - d_ooff[:-1]).to(torch.int32).contiguous()

                def dec():
                    rc = lib.b200lz4_decompress_dev(p(d_out), p(c_off), p(c_len), n, None, 0, None, p(d_back), p(d_off), p(d_len),
                                                    p(d_blen), 8, 0, p(scratch), sh)
                    assert rc == 0
                dec(); torch.cuda.synchronize()
                e[0].record(stream)
                for _ in range(args.reps):
                    dec()
                e[1].record(stream); torch.cuda.synchronize()
                dms = e[0].elapsed_time(e[1]) / args.reps
                ok = bool(torch.equal(d_back[:total], d_src))
                print(f"{kind:10s} {bs:8d} {accel:6d} {total / ctot:6.2f} {cms:9.3f} {total / cms / 1e6:10.1f} {dms:8.3f} "
                      f"{total / dms / 1e6:9.1f} {'ok' if ok else 'ROUNDTRIP MISMATCH'}", flush=True)


if __name__ == "__main__":
    main()
