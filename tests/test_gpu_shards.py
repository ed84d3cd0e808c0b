"""The multi-GPU partition logic on ONE GPU: world in {2, 4, 8} contexts of this process joined by
the in-process communicator (dvp_comm_init_local, no NCCL), one host thread per rank.  Point-range-sharded MSMs,
row-range R1CS evaluation, sharded extends, index-range K scalars: every rank must return the 118 bytes of the
oracle's single-process prover (proving.rs:426-688), and the oracle's verifier accepts them."""
import random

import numpy as np
import pytest

import dvpari
import synth

pytestmark = pytest.mark.gpu
P = dvpari.P


def _circuit(O, lg):
    circ = synth.synth_r1cs(lg, seed=0xD5A10003 + lg)
    r1cs = O.R1CS.from_arrays(circ["coeffs_mont"], circ["rowptr"], circ["wire"], circ["coeff"], circ["nrows"],
                              circ["k"], circ["nwires"])
    return circ, r1cs


@pytest.mark.parametrize("world,lg", [(2, 10), (4, 12), (8, 12), (4, 14), (8, 14)])
def test_sharded_prove_on_one_gpu_is_bit_exact(oracle, world, lg):
    O = oracle
    circ, r1cs = _circuit(O, lg)
    n, k = circ["n"], circ["k"]
    od = O.Domain(lg + 1)
    rnd = random.Random(1000 * world + lg)
    td = O.trapdoor(*[rnd.randrange(1, P) for _ in range(3)])
    scs = O.setup_scalars(r1cs, od, td)
    srs = O.Srs(r1cs, od, td)
    ctxs = [dvpari.Context(0) for _ in range(world)]
    dvpari.comm_init_local(ctxs)
    # witness from rank 0's device solver (single-rank call: no collective inside)
    inst0 = dvpari.R1CSInstance(ctxs[0], circ["nrows"], k, circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"],
                                circ["coeffs_mont"])
    w = inst0.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
    inst0.close()
    want, rc, want_st = O.prove(r1cs, od, srs, w, want_stages=True)
    assert rc == 0
    # sharded MSM over a vector that does not divide evenly
    m = 5003
    sc = dvpari.random_fr_mont(m, 11 + world)
    pts = O.mul_batch(O.generator(), dvpari.random_fr_mont(m, 12))
    enc = O.encode_batch(pts)
    want_msm = O.pt_encode(O.msm(sc, pts, 0))
    sc2 = dvpari.random_fr_mont(m, 13 + world)
    want_msm2 = O.pt_encode(O.msm(sc2, pts, 0))

    def rank_body(r):
        ctx = ctxs[r]
        lo, hi = dvpari.shard_range(m, r, world)
        ctx.srs_load(6, enc[lo:hi])
        got_msm = ctx.msm_sharded(sc[lo:hi], 6)
        # the batched form: pipelined local MSMs, one all-gather for the three partial sums of every rank
        got_batch = ctx.msm_sharded_batch([sc[lo:hi], sc2[lo:hi], sc[lo:hi]], 6)
        inst = dvpari.R1CSInstance(ctx, circ["nrows"], k, circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"],
                                   circ["coeffs_mont"])
        for slot, s in enumerate(scs):
            if slot < 2:
                lo, hi = dvpari.shard_range(s.shape[0], r, world)
                ctx.srs_mulgen(slot, s[lo:hi])
            else:
                ctx.srs_mulgen(slot, s[dvpari.gk_shard_indices(n, r, world)])
        gd = dvpari.Domain(ctx, lg + 1)
        prover = dvpari.Prover(ctx, gd, inst, 0, 1, 2)
        proof = prover.prove(w[1:1 + k], w[1 + k:])
        # an unsatisfied witness is reported on every rank with the same row (nobody hangs in a collective)
        w_bad = w.copy()
        w_bad[1 + k + 5] = w_bad[1 + k + 6]
        try:
            prover.prove(w_bad[1:1 + k], w_bad[1 + k:])
            bad = None
        except dvpari.DvpError as e:
            bad = (e.code, str(e))
        proof2 = prover.prove(w[1:1 + k], w[1 + k:])  # and the prover is still usable
        # the stage dump runs the sharded extend with whole-vector outputs: a b c i a' b' c' i' must be the oracle's
        proof3, st = prover.prove(w[1:1 + k], w[1 + k:], want_stages=True)
        stages_ok = proof3 == proof and st[:8 * n].tobytes() == want_st[:8 * n].tobytes()
        prover.close(); inst.close(); gd.close()
        return got_msm, proof, bad, proof2, stages_ok, got_batch

    try:
        res = dvpari.run_ranks(rank_body, world)
    finally:
        for c in ctxs:
            c.comm_destroy()
            c.close()
    for r, (got_msm, proof, bad, proof2, stages_ok, got_batch) in enumerate(res):
        assert got_msm == want_msm, r
        assert got_batch == [want_msm, want_msm2, want_msm], r
        assert stages_ok, f"rank {r}/{world}: stage vectors of the sharded prove differ from the oracle's"
        assert proof == want, f"rank {r}/{world}: sharded proof differs from the oracle's"
        assert proof2 == want
        assert bad is not None and bad[0] == 6 and bad[1] == res[0][2][1]
    assert O.verify(td, dvpari.fr_from_mont(w[1:1 + k]), res[0][1])
