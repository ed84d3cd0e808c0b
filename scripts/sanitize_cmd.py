"""Small end-to-end run (written for compute-sanitizer memcheck, which is closed on this pool; kept as a quick smoke): MSM plain + table layouts, mulgen, device setup, prove,
enter / exit, all at sizes a sanitizer run finishes in a minute."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import numpy as np
import dvpari, synth
ctx = dvpari.Context(0)
n = 3000
ctx.srs_random(0, n, 5)
sc = dvpari.random_fr_mont(n, 6)
a = ctx.multi_scalar_mul(sc, 0)
ctx.set("msm_tables_min", 64)
b = ctx.multi_scalar_mul(sc, 0)
assert a == b and ctx.msm_stats()["tables"] == 1
c = ctx.multi_scalar_mul(sc[: n - 100], 0, offset=50)
ctx.set("msm_tables_min", 1 << 15)
circ = synth.synth_r1cs(13, seed=3, nlevels=8)
inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"], circ["coeffs_mont"])
w = inst.synth_solve(synth.synth_assignment(circ), circ["nlevels"])
dom = dvpari.Domain(ctx, 14)
dvpari.setup(inst, dom, [3, 5, 7], 1, 2, 3)
prover = dvpari.Prover(ctx, dom, inst, 1, 2, 3)
k = circ["k"]
p1 = prover.prove(w[1:1 + k], w[1 + k:])
p2 = prover.prove(w[1:1 + k], w[1 + k:])
assert p1 == p2
plan = dvpari.EcfftPlan(ctx, 7)
cf = dvpari.random_fr_mont(128, 1)
assert plan.exit(plan.enter(cf)).tobytes() == cf.tobytes()
print("sanitize run ok", p1[:8].hex())
