"""One rank of the multi-GPU parity check (launched by tests/test_gpu_multi.py through torch.distributed.run).

Every rank holds its contiguous range of the SRS (generated on its own GPU from the oracle's setup scalars), the
MSMs / rows / extends are sharded as dvp_comm_init documents, and every rank must end up with the 118 bytes the
oracle's single-process prover produces; the oracle's verifier accepts them."""
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))

import numpy as np
import torch
import torch.distributed as dist

import dvpari
import synth
from oracle import oracle as O

P = dvpari.P


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    lg = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    dist.init_process_group("gloo")
    ctx = dvpari.Context(local)
    ids = [dvpari.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx.comm_init(ids[0], rank, world)

    # sharded MSM against the oracle's MSM over the whole vector
    n = 5000
    sc = dvpari.random_fr_mont(n, 11)
    pts = O.mul_batch(O.generator(), dvpari.random_fr_mont(n, 12))
    enc = O.encode_batch(pts)
    lo, hi = dvpari.shard_range(n, rank, world)
    ctx.srs_load(0, enc[lo:hi])
    got = ctx.msm_sharded(sc[lo:hi], 0)
    want_msm = O.pt_encode(O.msm(sc, pts, 0))
    assert got == want_msm, "sharded MSM differs from the oracle"
    # the batched form: pipelined local MSMs, one NCCL all-gather for all partial sums
    sc2 = dvpari.random_fr_mont(n, 13)
    got_b = ctx.msm_sharded_batch([sc[lo:hi], sc2[lo:hi], sc[lo:hi]], 0)
    assert got_b == [want_msm, O.pt_encode(O.msm(sc2, pts, 0)), want_msm], "batched sharded MSM differs from the oracle"

    # sharded prove
    circ = synth.synth_r1cs(lg, seed=0xD5A10003 + lg)
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"],
                               circ["coeff"], circ["coeffs_mont"])
    w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
    r1cs = O.R1CS.from_arrays(circ["coeffs_mont"], circ["rowptr"], circ["wire"], circ["coeff"], circ["nrows"],
                              circ["k"], circ["nwires"])
    od = O.Domain(lg + 1)
    rnd = random.Random(lg)
    tdi = [rnd.randrange(1, P) for _ in range(3)]
    td = O.trapdoor(*tdi)
    scs = O.setup_scalars(r1cs, od, td)
    n = circ["n"]
    for slot, s in enumerate(scs):
        if slot < 2:
            lo, hi = dvpari.shard_range(s.shape[0], rank, world)
            ctx.srs_mulgen(slot, s[lo:hi])
        else:  # g_k: the three parts cut by this rank's index range of D
            ctx.srs_mulgen(slot, s[dvpari.gk_shard_indices(n, rank, world)])
    gd = dvpari.Domain(ctx, lg + 1)
    prover = dvpari.Prover(ctx, gd, inst, 0, 1, 2)
    k = circ["k"]
    proof = prover.prove(w[1:1 + k], w[1 + k:])
    assert O.verify(td, dvpari.fr_from_mont(w[1:1 + k]), proof), "verifier rejects the sharded proof"
    if lg <= 14:
        srs = O.Srs(r1cs, od, td)
        want, rc, _ = O.prove(r1cs, od, srs, w)
        assert rc == 0 and proof == want, "sharded proof differs from the oracle's"
    # a bad witness is reported on every rank (no rank may hang in a collective)
    w_bad = w.copy()
    w_bad[1 + k + 5] = w_bad[1 + k + 6]
    try:
        prover.prove(w_bad[1:1 + k], w_bad[1 + k:])
        raise AssertionError("bad witness accepted")
    except dvpari.DvpError as e:
        assert e.code == 6
    # commit_p as two MSMs (g_m from the helper thread, then g_q) and as one MSM over g_m | g_q: the same bytes
    for joint in (0, 1):
        ctx.set("prove_joint", joint)
        assert prover.prove(w[1:1 + k], w[1 + k:]) == proof, f"prove_joint={joint}"
    ctx.set("prove_joint", -1)
    outs = [None] * world
    dist.all_gather_object(outs, proof)
    assert all(o == outs[0] for o in outs)
    # the same proof from an SRS made by the sharded device setup (dvp_setup fills every rank's slots itself)
    dvpari.setup(inst, gd, tdi, 3, 4, 5)
    prover2 = dvpari.Prover(ctx, gd, inst, 3, 4, 5)
    assert prover2.prove(w[1:1 + k], w[1 + k:]) == proof
    prover2.close()
    print(f"rank {rank}/{world}: sharded msm + prove 2^{lg} OK {prover.last_times()}", flush=True)
    prover.close(); inst.close(); gd.close()
    ctx.comm_destroy()
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
