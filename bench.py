#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 LZ4 block codec (contract: see DESIGN.md section "Measurement").

Workload (BASELINE.json configs[1]): 1 GiB synthetic mixed-entropy stream per GPU, strategy
`c+400+640000` (acceleration 400, 640000-byte arrays -> 1678 blocks), independent blocks.
A step = one pass of the compress path (codec kernel + compaction pass) over that batch.

  value     uncompressed GB/s, inputs resident in HBM, CUDA-event timed on the launching stream
  e2e       same metric through b200lz4_compress_batch with pinned HOST buffers (H2D + D2H inside)
  roofline  dominant kernel (compress_kernel): (U + C) bytes / its own event-timed duration vs measured HBM peak
  cpu_baseline / --impl reference: the reference's own cbits/lz4.c (oracle/_ref) on the box's host cores
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BLOCK = 640000
ACCEL = 400
HEADER = 8
METRIC = "LZ4 compress GB/s (uncompressed), c+400+640000 independent blocks"
UNIT = "GB/s"


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def make_blocks(total: int):
    offs = np.arange(0, total, BLOCK, dtype=np.int64)
    lens = np.minimum(BLOCK, total - offs).astype(np.int32)
    return offs, lens


# --------------------------------------------------------------------------- reference arm

def cpu_reference(data: np.ndarray, offs, lens, threads: int, reps: int, sample_blocks: int | None = None):
    """The reference's lz4.c driven like compressChunk (fresh LZ4_stream_t per block), `threads` pthreads."""
    from oracle.oracle import Oracle
    ora = Oracle("auto")
    if sample_blocks is not None:
        offs, lens = offs[:sample_blocks], lens[:sample_blocks]
    n = len(lens)
    ptrs = (data.ctypes.data + offs).astype(np.uint64)
    caps = (lens.astype(np.int64) + lens // 255 + 16 + HEADER).astype(np.int32)
    arena, dptrs, _ = ora.slots(caps)
    out_len = np.zeros(n, dtype=np.int32)
    sf = np.arange(n + 1, dtype=np.int32)
    best = None
    for _ in range(reps):
        t0 = time.perf_counter()
        rc = ora.compress_ptrs(ptrs, lens, dptrs, caps, out_len, ACCEL, HEADER, sf, 0, threads)
        dt = time.perf_counter() - t0
        assert rc == 0
        best = dt if best is None else min(best, dt)
    u = int(lens.astype(np.int64).sum())
    return {"gbps": u / best / 1e9, "seconds": best, "bytes": u, "kind": "reference" if ora.kind == "reference" else "port",
            "comp_bytes": int(out_len.astype(np.int64).sum())}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from streamly_lz4_b200 import datagen
    total = args.size_mib << 20
    data = datagen.make("mixed", 2, total)
    offs, lens = make_blocks(total)
    cores = os.cpu_count() or 1
    threads = min(cores, 256)
    for _ in range(args.warmup):
        cpu_reference(data, offs, lens, threads, 1)
    t = []
    for _ in range(args.steps):
        r = cpu_reference(data, offs, lens, threads, 1)
        t.append(r["seconds"])
    ms = 1e3 * sum(t) / len(t)
    val = total / (ms / 1e3) / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{args.size_mib} MiB mixed-entropy stream, c+{ACCEL}+{BLOCK}, independent blocks "
                                   f"({len(lens)} blocks), reference cbits/lz4.c (LZ4 1.9.3) on host cores"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": r["kind"],
                             "sample": f"whole {args.size_mib} MiB batch per step, {threads} pthreads, one block range per thread"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------- GPU arm

def run_gpu(args):
    import torch
    import torch.distributed as dist
    from streamly_lz4_b200 import _lib, datagen
    import streamly_lz4_b200 as lz

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the GPU arm has no CPU fallback")
    from streamly_lz4_b200 import stripe
    numa = stripe.bind_host_to_device(local)               # pinned staging buffers next to this rank's GPU
    log(f"[rank {rank}] host binding: {numa}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()
    stream = torch.cuda.current_stream()
    sh = ctypes.c_void_p(stream.cuda_stream)

    total = args.size_mib << 20
    t0 = time.time()
    host = datagen.make("mixed", 2 + rank, total)          # every rank its own stripe of the stream (weak scaling)
    offs, lens = make_blocks(total)
    n = len(lens)
    log(f"[rank {rank}] generated {total >> 20} MiB in {time.time() - t0:.1f}s, {n} blocks")

    bound = lens.astype(np.int64) + lens // 255 + 16
    slot_sizes = (bound + HEADER + 16 + 15) // 16 * 16
    slot_off = np.zeros(n, dtype=np.int64)
    slot_off[1:] = np.cumsum(slot_sizes[:-1])
    slots_total = int(slot_sizes.sum())

    d_src = torch.from_numpy(host).to(dev)
    d_src_off = torch.from_numpy(offs).to(dev)
    d_src_len = torch.from_numpy(lens).to(dev)
    d_slot_off = torch.from_numpy(slot_off).to(dev)
    d_slots = torch.empty(slots_total + 64, dtype=torch.uint8, device=dev)
    d_out = torch.empty(slots_total + 64, dtype=torch.uint8, device=dev)
    d_out_len = torch.zeros(n, dtype=torch.int32, device=dev)
    d_out_off = torch.zeros(n + 1, dtype=torch.int64, device=dev)
    d_scratch = torch.zeros(lib.b200lz4_scratch_bytes(), dtype=torch.uint8, device=dev)

    def p(t):
        return ctypes.c_void_p(t.data_ptr())

    def compress_only():
        rc = lib.b200lz4_compress_dev(p(d_src), p(d_src_off), p(d_src_len), n, None, 0, None,
                                      p(d_slots), p(d_slot_off), None, p(d_out_len), ACCEL, HEADER, p(d_scratch), sh)
        assert rc == 0, _lib.last_error()

    def compact_only():
        rc = lib.b200lz4_compact_dev(p(d_slots), p(d_slot_off), p(d_out_len), n, HEADER, p(d_out), p(d_out_off),
                                     p(d_scratch), sh)
        assert rc == 0, _lib.last_error()

    def step():
        compress_only()
        compact_only()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    comp_total = int(d_out_off[-1].item())
    assert int((d_out_len <= 0).sum().item()) == 0, "a block failed to compress"

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # ---- timed region: exactly K steps, device-timed, barrier + synchronize on both sides
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
    barrier()
    ev[0].record(stream)
    for k in range(args.steps):
        compress_only()
        ev[2 * k + 1].record(stream)
        compact_only()
        ev[2 * k + 2].record(stream)
    barrier()
    total_ms = ev[0].elapsed_time(ev[-1])
    kern_ms = [ev[2 * k].elapsed_time(ev[2 * k + 1]) for k in range(args.steps)]
    compact_ms = [ev[2 * k + 1].elapsed_time(ev[2 * k + 2]) for k in range(args.steps)]
    gpu_launches = args.steps * 3

    # ---- e2e: the C-ABI host call with pinned host buffers (H2D + kernels + D2H + sync inside)
    ctx = lz.Context(local)
    pin_src = ctx.pinned("b_src", total, write_combined=bool(os.environ.get("B200LZ4_BENCH_WC")))
    pin_src[:total] = host
    pin_dst = ctx.pinned("b_dst", int((bound + HEADER).sum()))
    src_view = pin_src[:total]
    for _ in range(min(args.warmup, 2)):
        rc, doff, olen = ctx.compress_batch(src_view, offs, lens, ACCEL, HEADER, pin_dst)
        assert rc == 0, _lib.last_error()
    e2e_t = []
    split = {"h2d_ms": [], "kernel_ms": [], "d2h_ms": []}
    barrier()
    launches0 = ctx.launch_count()
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        t1 = time.perf_counter()
        rc, doff, olen = ctx.compress_batch(src_view, offs, lens, ACCEL, HEADER, pin_dst)
        e2e_t.append(time.perf_counter() - t1)
        tm = ctx.timing()
        for k in split:
            split[k].append(tm[k])
    barrier()
    e2e_wall = time.perf_counter() - t_all0
    e2e_launches = ctx.launch_count() - launches0          # counted by the library: chunks x (codec + scan + gather)
    assert rc == 0 and int(doff[-1]) == comp_total
    clocks = sampler.stop() if rank == 0 else None

    # (the cpu_baseline leg below also spot-checks a few of these GPU blocks against the oracle's bytes)
    parity = None

    # ---- decompress of the same stream (extra, not the headline): d+640000 on the compressed output
    extra = {}
    if not args.no_extras:
        c_off = d_out_off[:-1].contiguous()
        c_len = (d_out_off[1:] - d_out_off[:-1]).to(torch.int32).contiguous()
        d_back = torch.empty(total + 64, dtype=torch.uint8, device=dev)
        d_back_len = torch.zeros(n, dtype=torch.int32, device=dev)

        def decomp():
            rc = lib.b200lz4_decompress_dev(p(d_out), p(c_off), p(c_len), n, None, 0, None, p(d_back), p(d_src_off),
                                            p(d_src_len), p(d_back_len), HEADER, 0, p(d_scratch), sh)
            assert rc == 0, _lib.last_error()
        for _ in range(args.warmup):
            decomp()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(args.steps):
            decomp()
        e1.record(stream)
        barrier()
        d_ms = e0.elapsed_time(e1) / args.steps
        same = bool(torch.equal(d_back[:total], d_src)) and bool((d_back_len == d_src_len).all().item())
        extra["decompress"] = {"ms_per_step": d_ms, "round_trip_identical": same}
        gpu_launches_extra = args.steps
        # the same through b200lz4_decompress_batch (pinned host in / out)
        comp_host = pin_dst[:comp_total]
        pin_back = ctx.pinned("b_back", total + 64)
        c_off_h = np.ascontiguousarray(doff[:-1]); c_len_h = np.diff(doff).astype(np.int32)
        for _ in range(2):
            rc2, boff, blen = ctx.decompress_batch(comp_host, c_off_h, c_len_h, HEADER, 0, pin_back)
            assert rc2 == 0, _lib.last_error()
        barrier()
        t_d0 = time.perf_counter()
        for _ in range(args.steps):
            rc2, boff, blen = ctx.decompress_batch(comp_host, c_off_h, c_len_h, HEADER, 0, pin_back)
        barrier()
        d_e2e = (time.perf_counter() - t_d0) / args.steps
        tmd = ctx.timing()
        extra["decompress"]["e2e"] = {"wall_ms_per_step": 1e3 * d_e2e, "h2d_ms": tmd["h2d_ms"], "kernel_ms": tmd["kernel_ms"],
                                      "d2h_ms": tmd["d2h_ms"], "identical": bool((pin_back[:total] == host).all())}
        gpu_launches_extra += ctx.launch_count() - launches0 - e2e_launches

    # ---- reduce over ranks (max time)
    def rmax(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    total_ms = rmax(total_ms)
    e2e_wall = rmax(e2e_wall)
    ms_per_step = total_ms / args.steps
    value = world * total / (ms_per_step / 1e3) / 1e9
    e2e_value = world * total / (e2e_wall / args.steps) / 1e9
    if "decompress" in extra:
        dms = rmax(extra["decompress"]["ms_per_step"])
        extra["decompress"].update({"ms_per_step": dms, "value": world * total / (dms / 1e3) / 1e9, "unit": UNIT})
        de = rmax(extra["decompress"]["e2e"]["wall_ms_per_step"])
        extra["decompress"]["e2e"].update({"wall_ms_per_step": de, "value": world * total / (de / 1e3) / 1e9, "unit": UNIT})

    if rank == 0:
        peak, peak_src = measured_peak()
        k_ms = statistics.mean(kern_ms)
        algo = total + comp_total                      # U + C per launch
        achieved = algo / (k_ms / 1e3) / 1e9
        roof = {"bound": "hbm", "kernel": "compress_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": algo, "kernel_ms": k_ms, "compact_ms": statistics.mean(compact_ms),
                "kernel_share_of_step": k_ms / ms_per_step}
        traffic = {}
        tp = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tp):
            try:
                tj = json.load(open(tp))
                if tj.get("size_mib") == args.size_mib:
                    traffic = tj
                    roof["traffic"] = tj.get("compress_kernel_dram_bytes")
            except Exception:
                pass
        if "decompress" in extra:
            dv = extra["decompress"]
            dv["roofline"] = {"bound": "hbm", "kernel": "decompress_kernel",
                              "achieved": algo / (dv["ms_per_step"] / 1e3) / 1e9 * 1.0, "peak": peak, "unit": "GB/s"}
            dv["roofline"]["frac"] = dv["roofline"]["achieved"] / peak
            dv["roofline"]["traffic"] = traffic.get("decompress_kernel_dram_bytes")
        cpu = None
        if not args.no_cpu:
            # ---- cpu_baseline leg: the only place of this arm that touches oracle/ (timing + a parity spot check)
            if not args.no_check:
                from oracle.oracle import Oracle
                ora = Oracle("auto")
                idx = sorted(set([0, n // 2, n - 1]))
                ok = True
                for i in idx:
                    a = host[offs[i]:offs[i] + lens[i]].tobytes()
                    ok &= pin_dst[doff[i]:doff[i + 1]].tobytes() == ora.compress_chunks([a], ACCEL, linked=False)[0]
                parity = {"blocks_checked": idx, "byte_identical_to_oracle": bool(ok), "oracle": ora.kind}
            cores = os.cpu_count() or 1
            threads = min(cores, 256)
            r_all = cpu_reference(host, offs, lens, threads, 3)
            r_one = cpu_reference(host, offs, lens, 1, 1, sample_blocks=min(n, 420))
            cpu = {"value": r_all["gbps"], "unit": UNIT, "cores": threads, "kind": r_all["kind"],
                   "sample": f"all {n} blocks ({total >> 20} MiB), best of 3, {threads} pthreads; single thread on first "
                             f"{min(n, 420)} blocks: {r_one['gbps']:.3f} GB/s",
                   "single_thread_value": r_one["gbps"]}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: {args.size_mib} MiB mixed-entropy stream per GPU, c+{ACCEL}+{BLOCK}, "
                                   f"independent blocks ({n} blocks/GPU), BlockHasSize headers; step = codec kernel + compaction",
                       "ratio": total / comp_total, "l2": "inputs_exceed_l2 (1 GiB in, 0.6+ GiB out per step)",
                       "parallelism": f"block stripes over {world} GPU(s), no collective", "host_numa": numa},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": total + 20 * n,
                    "d2h_bytes_per_step": comp_total + 12 * n + 8,
                    "api": "b200lz4_compress_batch (pinned host in/out)",
                    "h2d_ms": statistics.mean(split["h2d_ms"]), "kernel_ms": statistics.mean(split["kernel_ms"]),
                    "d2h_ms": statistics.mean(split["d2h_ms"]), "wall_ms_per_step": 1e3 * e2e_wall / args.steps},
            "gpu_launches": gpu_launches + e2e_launches + (gpu_launches_extra if "decompress" in extra else 0),
            "gpu_launches_detail": {"timed_value_region": gpu_launches, "e2e_region": e2e_launches,
                                    "per_step": "compress_kernel + scan_kernel + gather_kernel (e2e: per pipeline chunk)"},
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks, "parity": parity, "extra": extra,
        }
        print(json.dumps(line), flush=True)
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--size-mib", type=int, default=1024, help="stream size per GPU (default: the 1 GiB of configs[1])")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-extras", action="store_true", help="skip the decompress extra")
    ap.add_argument("--no-check", action="store_true", help="skip the in-bench parity spot check")
    args = ap.parse_args()
    import __graft_entry__ as ge
    if int(os.environ.get("LOCAL_RANK", "0")) == 0:
        ge.build_cpu_side()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
