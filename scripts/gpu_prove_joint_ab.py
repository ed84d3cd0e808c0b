"""A/B of dvp_prove with commit_p as two MSMs (g_m beside the Fr-side work, then g_q) or as one MSM over g_m | g_q,
on one GPU.  2^19 constraints = the MSM sizes of one rank of an 8-way sharded 2^22-constraint proof.
Usage (on a GPU box): python scripts/gpu_prove_joint_ab.py [lg ...]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari  # noqa: E402
import synth  # noqa: E402
import torch  # noqa: E402


def main():
    lgs = [int(a) for a in sys.argv[1:]] or [16, 19, 20]
    ctx = dvpari.Context(0)
    for lg in lgs:
        circ = synth.synth_r1cs(lg)
        inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"],
                                   circ["coeff"], circ["coeffs_mont"])
        w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
        dom = dvpari.Domain(ctx, lg + 1)
        k = circ["k"]
        dvpari.setup(inst, dom, [0xD5A10005, 0xD5A10006, 0xD5A10007], 1, 2, 3)
        pinned = torch.empty((circ["nwires"], 4), dtype=torch.int64).pin_memory()
        wp = pinned.numpy().view(np.uint64)
        wp[:] = w
        pub, priv = wp[1:1 + k], wp[1 + k:]
        proofs = {}
        for rnd in range(2):
            for joint in (0, 1):
                ctx.set("prove_joint", joint)
                prover = dvpari.Prover(ctx, dom, inst, 1, 2, 3)
                proofs[joint] = prover.prove(pub, priv)
                prover.prove(pub, priv)
                reps = 5
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                for _ in range(reps):
                    prover.prove(pub, priv)
                torch.cuda.synchronize()
                ms = 1e3 * (time.perf_counter() - t0) / reps
                st = {a: round(b, 3) for a, b in prover.last_times().items()}
                print(f"2^{lg} constraints, joint={joint}: {ms:.3f} ms per proof  {st}", flush=True)
                prover.close()
        assert proofs[0] == proofs[1], "the joint and the two-MSM commitment differ"
        ctx.set("prove_joint", -1)
        inst.close()
        dom.close()
        for sl in (1, 2, 3):
            ctx.srs_free(sl)
    ctx.close()


if __name__ == "__main__":
    main()
