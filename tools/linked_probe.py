#!/usr/bin/env python
"""Kernel-resident throughput of LINKED streams (BASELINE config 4 shape: S concurrent streams of 64 KiB blocks,
previous block = dictionary).  python tools/linked_probe.py [--streams 128] [--mib-per-stream 16] [--kind mixed] [--accel 1]"""
import argparse, ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    from streamly_lz4_b200 import _lib, datagen
    ap = argparse.ArgumentParser()
    ap.add_argument("--streams", default="128")
    ap.add_argument("--mib-per-stream", type=int, default=16)
    ap.add_argument("--kinds", default="mixed,text")
    ap.add_argument("--accel", type=int, default=1)
    ap.add_argument("--block", type=int, default=65536)
    ap.add_argument("--stats", action="store_true", help="print the wide decoder's cycle counters (needs the stats build)")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda", 0)
    stream = torch.cuda.current_stream(); sh = ctypes.c_void_p(stream.cuda_stream)
    scratch = torch.zeros(lib.b200lz4_scratch_bytes(), dtype=torch.uint8, device=dev)
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    print(f"{'kind':8s} {'streams':>7s} {'MiB/str':>7s} {'ratio':>6s} {'comp ms':>9s} {'comp GB/s':>10s} {'MB/s/stream':>11s} {'dec ms':>8s} {'dec GB/s':>9s}")
    for kind in args.kinds.split(","):
        for ns in [int(x) for x in args.streams.split(",")]:
            per = args.mib_per_stream << 20
            total = ns * per
            host = datagen.make(kind, 4, total)
            bs = args.block
            offs = np.arange(0, total, bs, dtype=np.int64); lens = np.minimum(bs, total - offs).astype(np.int32)
            n = len(lens); bps = per // bs
            sf = (np.arange(ns + 1, dtype=np.int64) * bps).astype(np.int32)
            bound = lens.astype(np.int64) + lens // 255 + 16
            ss = (bound + 8 + 16 + 15) // 16 * 16
            so = np.zeros(n, dtype=np.int64); so[1:] = np.cumsum(ss[:-1])
            d_src = torch.from_numpy(host).to(dev)
            d_off = torch.from_numpy(offs).to(dev); d_len = torch.from_numpy(lens).to(dev); d_so = torch.from_numpy(so).to(dev)
            d_sf = torch.from_numpy(sf).to(dev)
            d_slots = torch.empty(int(ss.sum()) + 64, dtype=torch.uint8, device=dev)
            d_out = torch.empty(int(ss.sum()) + 64, dtype=torch.uint8, device=dev)
            d_olen = torch.zeros(n, dtype=torch.int32, device=dev); d_ooff = torch.zeros(n + 1, dtype=torch.int64, device=dev)
            d_back = torch.empty(total + 64, dtype=torch.uint8, device=dev); d_blen = torch.zeros(n, dtype=torch.int32, device=dev)

            def comp():
                rc = lib.b200lz4_compress_dev(p(d_src), p(d_off), p(d_len), n, p(d_sf), ns, None, p(d_slots), p(d_so), None,
                                              p(d_olen), args.accel, 8, p(scratch), sh)
                assert rc == 0
            comp(); torch.cuda.synchronize()
            e = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            e[0].record(stream); comp(); e[1].record(stream); torch.cuda.synchronize()
            cms = e[0].elapsed_time(e[1])
            assert lib.b200lz4_compact_dev(p(d_slots), p(d_so), p(d_olen), n, 8, p(d_out), p(d_ooff), p(scratch), sh) == 0
            torch.cuda.synchronize()
            ctot = int(d_ooff[-1].item())
            c_off = d_ooff[:-1].contiguous(); c_len = (d_ooff[1:] - d_ooff[:-1]).to(torch.int32).contiguous()

            def dec():
                rc = lib.b200lz4_decompress_dev(p(d_out), p(c_off), p(c_len), n, p(d_sf), ns, None, p(d_back), p(d_off), p(d_len),
                                                p(d_blen), 8, 0, p(scratch), sh)
                assert rc == 0
            dec(); torch.cuda.synchronize()
            scratch[:256].zero_()
            e[0].record(stream); dec(); e[1].record(stream); torch.cuda.synchronize()
            if args.stats:
                st = scratch[16:256].cpu().numpy().view(np.uint64)
                names = ["parser total", "parser ring wait", "batches", "-", "disp other", "disp wait parser", "disp wait flow", "disp open/close",
                         "L wait", "L loads", "L copy", "L tickets", "F wait L", "F wait N", "F copy", "N wait", "N near", "N long", "N near count",
                         "G wait", "G flush"]
                print("   stats (Mcycles summed over warps): " + ", ".join(f"{n}={int(v) / 1e6:.1f}" for n, v in zip(names, st)))
            dms = e[0].elapsed_time(e[1])
            ok = bool(torch.equal(d_back[:total], d_src))
            print(f"{kind:8s} {ns:7d} {args.mib_per_stream:7d} {total / ctot:6.2f} {cms:9.2f} {total / cms / 1e6:10.2f} {per / cms / 1e3:11.1f} "
                  f"{dms:8.2f} {total / dms / 1e6:9.2f} {'ok' if ok else 'ROUNDTRIP MISMATCH'}", flush=True)


if __name__ == "__main__":
    main()
