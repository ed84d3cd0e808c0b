"""Regenerates tests/golden/golden_v2.json (run from the repo root, CPU only): vectors for the components added
after v1 -- setup scalars, FFTree::enter / exit, the batched fixed-base multiplication, the SP1 public-input
scalar and a synthetic-circuit proof -- from the oracle and from big-integer arithmetic in Python.

Like v1 these are NOT upstream outputs (the Rust reference cannot run here), except `sp1_public_input`, which is
the reference's own known answer (gnark_r1cs.rs:494-504)."""
import json
import os
import random
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "dv-pari_b200"))
import dvpari
import synth
from oracle import oracle as O

P = O.P
rnd = random.Random(20261019)
out = {"note": "oracle / big-integer generated; encodings are candidates (parity unpinned)"}

# enter / exit on the 8-leaf tree: coefficients <-> values on the leaves x(C + i G_8) (Horner with big integers)
dom = O.Domain(3)
leaves = dom.leaves()
coeffs = [rnd.randrange(P) for _ in range(8)]
vals = []
for s in leaves:
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * s + c) % P
    vals.append(acc)
out["enter_exit"] = {"log_n": 3, "coeffs": [hex(c) for c in coeffs], "values_on_leaves": [hex(v) for v in vals]}

# mulgen: k * generator
G = O.generator()
ks = [0, 1, 255, 256, 0xDEADBEEFCAFE, P - 1, rnd.randrange(P)]
out["mulgen"] = [{"k": hex(k), "xsk233_candidate": O.pt_encode(O.pt_mul(G, k)).hex()} for k in ks]

# setup scalars of the toy circuit (srs.rs:53-167)
r1cs, pub, priv = O.toy_r1cs()
od = O.Domain(r1cs.n.bit_length())
tdi = [rnd.randrange(1, P) for _ in range(3)]
sc_m, sc_q, sc_k = O.setup_scalars(r1cs, od, O.trapdoor(*tdi))
out["toy_setup"] = {"trapdoor": [hex(v) for v in tdi], "sc_m": [hex(v) for v in dvpari.fr_from_mont(sc_m)],
                    "sc_q": [hex(v) for v in dvpari.fr_from_mont(sc_q)], "sc_k": [hex(v) for v in dvpari.fr_from_mont(sc_k)]}

# the reference's own known answer
out["sp1_public_input"] = {"raw_le_bytes": [55, 0, 0, 0, 89, 0, 0, 0],
                           "fr": "19542051593079647282099705468191403958371264520862632234952945594121"}

# a synthetic SP1-shaped circuit (synth.py, 2^6 constraints), witness solved level by level with big integers
circ = synth.synth_r1cs(6, seed=5, nlevels=4)
w = dvpari.fr_from_mont(synth.synth_assignment(circ, seed=9))
cf = dvpari.fr_from_mont(circ["coeffs_mont"])
k, nrows, nl = circ["k"], circ["nrows"], circ["nlevels"]


def dot(which, r):
    lo, hi = int(circ["rowptr"][which][r]), int(circ["rowptr"][which][r + 1])
    return sum(cf[int(circ["coeff"][which][p])] * w[int(circ["wire"][which][p])] for p in range(lo, hi)) % P


for lvl in range(nl):
    for r in range(lvl, nrows, nl):
        w[1 + k + r] = (w[1 + k + r] + dot(0, r) * dot(1, r) - dot(2, r)) % P
r2 = O.R1CS.from_arrays(circ["coeffs_mont"], circ["rowptr"], circ["wire"], circ["coeff"], nrows, k, circ["nwires"])
od2 = O.Domain(7)
td2 = [rnd.randrange(1, P) for _ in range(3)]
srs = O.Srs(r2, od2, O.trapdoor(*td2))
proof, rc, _ = O.prove(r2, od2, srs, dvpari.fr_to_mont(w))
assert rc == 0 and O.verify(O.trapdoor(*td2), w[1:1 + k], proof)
out["synth_2_6"] = {"seed": 5, "nlevels": 4, "assignment_seed": 9, "trapdoor": [hex(v) for v in td2],
                    "witness_tail": [hex(v) for v in w[-4:]], "fresh_wire_0": hex(w[1 + k]), "proof118": proof.hex()}
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "golden_v2.json"), "w"), indent=1)
print("written")
