"""Striping of blocks / streams over the GPUs of one box.

The codec path has no exchange step: independent blocks shard by contiguous block
range, linked streams shard by whole streams, results are concatenated in block
order.  The only communication is bookkeeping (per-block output lengths, so that
every rank -- or just rank 0 -- knows where each block lands in the global output
stream); it runs over whatever torch.distributed backend the launcher initialised
(NCCL on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def stripe_range(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of unit indices owned by `rank`; sizes differ by at most one."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("rank/world")
    q, r = divmod(max(n_units, 0), world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def stripe_streams(stream_first: Optional[np.ndarray], n_blocks: int, rank: int, world: int):
    """Shard whole streams.  Returns (block_lo, block_hi, local_stream_first or None).

    stream_first None means independent blocks (every block its own unit)."""
    if stream_first is None:
        lo, hi = stripe_range(n_blocks, rank, world)
        return lo, hi, None
    sf = np.asarray(stream_first, dtype=np.int64)
    s_lo, s_hi = stripe_range(len(sf) - 1, rank, world)
    b_lo, b_hi = int(sf[s_lo]), int(sf[s_hi])
    return b_lo, b_hi, (sf[s_lo:s_hi + 1] - b_lo).astype(np.int32)


def gather_lengths(local_lens: np.ndarray, n_total: int, rank: int, world: int, device=None,
                   block_range: Optional[Tuple[int, int]] = None) -> np.ndarray:
    """All ranks' per-block output lengths in global block order (int32[n_total]).

    block_range = (lo, hi): the global block indices this rank owns.  Default: stripe_range(n_total, rank, world), the
    independent-blocks stripe; linked streams (stripe_streams) own uneven block ranges and MUST pass theirs.  The
    ranges of all ranks are exchanged first, so a mismatch between len(local_lens) and the owned range is an error
    instead of lengths landing at the wrong block indices."""
    local = np.ascontiguousarray(local_lens, dtype=np.int32)
    lo, hi = block_range if block_range is not None else stripe_range(n_total, rank, world)
    if hi - lo != len(local):
        raise ValueError(f"rank {rank} owns blocks [{lo}, {hi}) but reports {len(local)} lengths")
    if world == 1:
        if (lo, hi) != (0, n_total):
            raise ValueError("a single rank must own every block")
        return local.copy()
    import torch
    import torch.distributed as dist
    r_t = torch.tensor([lo, hi], dtype=torch.int64, device=device)
    r_all = [torch.zeros_like(r_t) for _ in range(world)]
    dist.all_gather(r_all, r_t)
    ranges = [(int(t[0]), int(t[1])) for t in r_all]
    width = max(max(b - a for a, b in ranges), 1)
    t = torch.zeros(width, dtype=torch.int32, device=device)
    t[:len(local)] = torch.from_numpy(local).to(t.device)
    parts = [torch.zeros_like(t) for _ in range(world)]
    dist.all_gather(parts, t)
    out = np.zeros(n_total, dtype=np.int32)
    covered = 0
    for (a, b), p in zip(ranges, parts):
        out[a:b] = p[:b - a].cpu().numpy()
        covered += b - a
    if covered != n_total:
        raise ValueError(f"ranks cover {covered} of {n_total} blocks")
    return out


def global_offsets(all_lens: np.ndarray, header: int) -> np.ndarray:
    """Start of every framed block in the concatenated output stream (int64[n + 1]);
    failed blocks (len <= 0) occupy nothing, as in b200lz4_compact_dev."""
    sizes = np.where(all_lens > 0, all_lens.astype(np.int64) + header, 0)
    off = np.zeros(len(all_lens) + 1, dtype=np.int64)
    np.cumsum(sizes, out=off[1:])
    return off


def bind_host_to_device(device: int) -> dict:
    """Best effort: run this process on the CPUs of the NUMA node the GPU hangs off and prefer that node for new
    pages, so that pinned staging buffers (allocated afterwards) are node-local and H2D / D2H copies do not cross
    the socket interconnect.  One process per GPU calls this before allocating.  Returns what was done."""
    import ctypes
    import os
    import subprocess
    info = {"device": device, "node": None, "cpus": None, "mempolicy": None}
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(device)],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip().lower()
        if not bus:
            return info
        if len(bus.split(":")[0]) == 8:          # nvidia-smi prints an 8-digit domain, sysfs uses 4
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read().strip())
        info["node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        target = cpus & os.sched_getaffinity(0)
        if target:
            os.sched_setaffinity(0, target)
            info["cpus"] = len(target)
        # set_mempolicy(MPOL_PREFERRED, {node}): syscall 238 on x86-64, 237 on aarch64
        nr = {"x86_64": 238, "aarch64": 237}.get(os.uname().machine)
        if nr is not None:
            mask = ctypes.c_ulong(1 << node)
            rc = ctypes.CDLL(None, use_errno=True).syscall(nr, 1, ctypes.byref(mask), ctypes.c_ulong(64))
            info["mempolicy"] = "preferred" if rc == 0 else f"errno {ctypes.get_errno()}"
    except Exception as e:                      # sysfs not visible, no nvidia-smi, ...: stay as we are
        info["error"] = repr(e)
    return info
