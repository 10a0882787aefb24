"""N > 1 host logic on CPU: world_size-2 gloo processes stripe a batch, run the codec on their stripe
(the CPU oracle stands in for the GPU here: this test is about the striping and bookkeeping, not the
kernels) and the gathered block table must equal the single-process result."""
import os
import socket

import numpy as np
import pytest


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, linked, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle.oracle import Oracle
        from streamly_lz4_b200 import datagen, stripe
        ora = Oracle("port")
        data = datagen.make("mixed", 11, 37 * 50000 + 123)
        arrays = [data[i:i + 50000].tobytes() for i in range(0, data.size, 50000)]
        n = len(arrays)
        sf = np.array([0, 5, 5, 17, 30, n], dtype=np.int32) if linked else None      # includes an empty stream
        lo, hi, local_sf = stripe.stripe_streams(sf, n, rank, world)
        mine = ora.compress_chunks(arrays[lo:hi], 3, linked=linked, stream_first=local_sf)
        lens = np.array([len(m) - 8 for m in mine], dtype=np.int32)
        # independent blocks own stripe_range(); linked streams own uneven block ranges and say so
        all_lens = stripe.gather_lengths(lens, n, rank, world, block_range=(lo, hi) if linked else None)
        off = stripe.global_offsets(all_lens, 8)
        blob = [None] * world
        dist.all_gather_object(blob, b"".join(mine))
        if rank == 0:
            whole = ora.compress_chunks(arrays, 3, linked=linked, stream_first=sf)
            ok = b"".join(blob) == b"".join(whole)
            ok &= [int(x) for x in np.diff(off)] == [len(w) for w in whole]
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("linked", [False, True])
def test_two_rank_striping_matches_single_process(built, linked):
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, linked, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_stripe_ranges_partition():
    from streamly_lz4_b200 import stripe
    for n in (0, 1, 7, 1678, 13422):
        for world in (1, 2, 4, 8):
            spans = [stripe.stripe_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    sf = np.array([0, 3, 3, 10, 12], dtype=np.int32)
    got = [stripe.stripe_streams(sf, 12, r, 2) for r in range(2)]
    assert got[0][:2] == (0, 3) and list(got[0][2]) == [0, 3, 3]
    assert got[1][:2] == (3, 12) and list(got[1][2]) == [0, 7, 9]
