"""Development aid: the streamed host call (b200lz4_compress_batch on equally long blocks) under different copy
schedules.  usage: python tools/streamed_probe.py [--mib 1024] [--kind mixed] [--accel 400] [--block 640000]"""
import argparse, os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import streamly_lz4_b200 as lz
from streamly_lz4_b200 import datagen

ap = argparse.ArgumentParser()
ap.add_argument("--mib", type=int, default=1024)
ap.add_argument("--kind", default="mixed")
ap.add_argument("--accel", type=int, default=400)
ap.add_argument("--block", type=int, default=640000)
ap.add_argument("--reps", type=int, default=4)
ap.add_argument("--sweep", default="auto,0,0.25,0.5,0.75,1.0,1.5,1000")
ap.add_argument("--groups", default="")
ap.add_argument("--debug", action="store_true")
ap.add_argument("--matrix", action="store_true")
args = ap.parse_args()
total = args.mib << 20
host = datagen.make(args.kind, 2, total)
offs = np.arange(0, total, args.block, dtype=np.int64)
lens = np.minimum(args.block, total - offs).astype(np.int32)
ctx = lz.Context(0)
src = ctx.pinned("s", total); src[:total] = host
dst = ctx.pinned("d", int((lens.astype(np.int64) + lens // 255 + 24).sum()))


def run(label, env):
    for k in ("B200LZ4_STREAM_W", "B200LZ4_STREAM_G", "B200LZ4_STREAM_S", "B200LZ4_DEBUG", "B200LZ4_STREAM_FLAG", "B200LZ4_DECODE_OUT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    if args.debug:
        os.environ["B200LZ4_DEBUG"] = "1"
    best = 1e9
    for i in range(args.reps):
        if i == args.reps - 1 and label in ("auto",):
            os.environ["B200LZ4_DEBUG"] = "1"
        t = time.perf_counter()
        rc, doff, ol = ctx.compress_batch(src[:total], offs, lens, args.accel, 8, dst)
        dt = time.perf_counter() - t
        assert rc == 0, ctx.last_error()
        best = min(best, dt)
    tm = ctx.timing()
    print(f"{label:>28}  best wall {best*1e3:7.2f} ms  {total/best/1e9:6.2f} GB/s   h2d {tm['h2d_ms']:.2f} kernel {tm['kernel_ms']:.2f} d2h {tm['d2h_ms']:.2f}", flush=True)


if args.matrix:
    run("auto", {})
    for G, S, W in ((12, 4, None), (12, 8, None), (12, 4, "0.5"), (12, 4, "1.0"), (12, 4, "1.5")):
        env = {"B200LZ4_STREAM_S": str(S), "B200LZ4_STREAM_G": str(G)}
        if W:
            env["B200LZ4_STREAM_W"] = W
        run(f"G={G} S={S} W={W or 'auto'}", env)
for w in ([] if args.matrix else args.sweep.split(",")):
    if w == "auto":
        run("auto", {})
    else:
        run(f"W={w}", {"B200LZ4_STREAM_W": w})
        if args.groups:
            for g in args.groups.split(","):
                run(f"W={w} G={g}", {"B200LZ4_STREAM_W": w, "B200LZ4_STREAM_G": g})
run("auto again", {})

# the same stream back: streamed decompress call (output leaves segment by segment)
rc, doff, ol = ctx.compress_batch(src[:total], offs, lens, args.accel, 8, dst)
c_off = doff[:-1].copy(); c_len = (ol + 8).astype(np.int32)
back = ctx.pinned("b", total + 64)
dcases = [("decompress mirror (default)", {}), ("decompress pieces", {"B200LZ4_DECODE_OUT": "pieces"}), ("decompress copy", {"B200LZ4_DECODE_OUT": "copy"}),
          ("decompress mirror again", {})]
for label, env in dcases:
    for k in ("B200LZ4_STREAM_W", "B200LZ4_STREAM_G", "B200LZ4_STREAM_S", "B200LZ4_DEBUG", "B200LZ4_STREAM_FLAG", "B200LZ4_DECODE_OUT"):
        os.environ.pop(k, None)
    os.environ.update(env)
    best = 1e9
    for i in range(args.reps):
        if i == args.reps - 1 and label == "decompress mirror (default)":
            os.environ["B200LZ4_DEBUG"] = "1"
        t = time.perf_counter()
        rc2, boff, blen = ctx.decompress_batch(dst, c_off, c_len, 8, 0, back)
        dt = time.perf_counter() - t
        assert rc2 == 0, ctx.last_error()
        best = min(best, dt)
    tm = ctx.timing()
    ok = bool((back[:total] == host).all())
    print(f"{label:>28}  best wall {best*1e3:7.2f} ms  {total/best/1e9:6.2f} GB/s   h2d {tm['h2d_ms']:.2f} kernel {tm['kernel_ms']:.2f} d2h {tm['d2h_ms']:.2f}  {'ok' if ok else 'MISMATCH'}", flush=True)
