"""world_size-2 gloo run of the N > 1 host logic on CPU: the shard rule (dvp_shard_range), the exchange pattern of a
sharded MSM (all-gather of the ranks' partial sums, fold in rank order) and of a batch of them (dvp_msm_sharded_batch:
ONE all-gather of nb partial sums per rank, folded per MSM in rank order) with the oracle standing in for the device MSM,
and the index ranges of the sharded g_k."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, n, q):
    try:
        _worker_body(rank, world, port, n, q)
    except Exception as e:  # report instead of letting the parent wait for its time-out
        q.put((rank, False, repr(e), 0))


def _worker_body(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
    import dvpari
    from oracle import oracle as O

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sc = dvpari.random_fr_mont(n, 21)
    pts = O.chain_points(n, O.pt_mul(O.generator(), 12345), O.pt_mul(O.generator(), 777))
    lo, hi = dvpari.shard_range(n, rank, world)
    mine_pts = (O.Pt * (hi - lo))(*pts[lo:hi])
    part = O.pt_encode(O.msm(sc[lo:hi], mine_pts, 1))  # this rank's partial sum (the device MSM's role)
    t = torch.frombuffer(bytearray(part) + b"\0\0", dtype=torch.uint8)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    acc = O.pt_decode(bytes(outs[0].numpy()[:30]))[0]
    for o in outs[1:]:
        acc = O.pt_add(acc, O.pt_decode(bytes(o.numpy()[:30]))[0])
    whole = O.pt_encode(O.msm(sc, pts, 1))
    ok = O.pt_encode(acc) == whole
    # every rank must hold the same bytes
    mine = torch.frombuffer(bytearray(O.pt_encode(acc)) + b"\0\0", dtype=torch.uint8)
    ref = mine.clone()
    dist.broadcast(ref, src=0)
    ok = ok and bool((ref == mine).all())
    # the batched exchange: nb partial sums per rank travel together, rank r's block is [r][b]
    nb = 3
    scs = [sc, dvpari.random_fr_mont(n, 22), dvpari.random_fr_mont(n, 23)]
    parts = b"".join(O.pt_encode(O.msm(s[lo:hi], mine_pts, 1)) + b"\0\0" for s in scs)
    tb = torch.frombuffer(bytearray(parts), dtype=torch.uint8)
    outs_b = [torch.empty_like(tb) for _ in range(world)]
    dist.all_gather(outs_b, tb)
    for b in range(nb):
        acc_b = O.pt_decode(bytes(outs_b[0].numpy()[32 * b:32 * b + 30]))[0]
        for o in outs_b[1:]:
            acc_b = O.pt_add(acc_b, O.pt_decode(bytes(o.numpy()[32 * b:32 * b + 30]))[0])
        ok = ok and O.pt_encode(acc_b) == O.pt_encode(O.msm(scs[b], pts, 1))
    q.put((rank, ok, lo, hi))
    dist.destroy_process_group()


def test_shard_range_partitions():
    sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
    import dvpari

    for total in (0, 1, 7, 4096, 4717470, (1 << 24) + 3):
        for world in (1, 2, 3, 4, 8):
            prev = 0
            for r in range(world):
                lo, hi = dvpari.shard_range(total, r, world)
                assert lo == prev and hi >= lo and hi - lo <= total // world + 1
                prev = hi
            assert prev == total
    # g_k is cut by the index range of D: every point of the 4n vector belongs to exactly one rank
    n = 1 << 10
    for world in (1, 2, 4, 8):
        idx = np.concatenate([dvpari.gk_shard_indices(n, r, world) for r in range(world)])
        assert idx.size == 4 * n and np.array_equal(np.sort(idx), np.arange(4 * n))
        assert all(dvpari.gk_shard_indices(n, r, world).size == 4 * n // world for r in range(world))


def test_sharded_msm_exchange_gloo():
    world, n = 2, 600
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + os.getpid() % 500
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
    assert all(ok for _, ok, _, _ in res), res
    spans = sorted((lo, hi) for _, _, lo, hi in res)
    assert spans[0][0] == 0 and spans[-1][1] == n and spans[0][1] == spans[1][0]
