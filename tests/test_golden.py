"""Committed fixtures (tests/golden/golden_v1.json, made by tests/golden/make_golden.py): the oracle must keep
reproducing them (CPU), and the CUDA path must reproduce the toy proof from the committed SRS bytes (GPU)."""
import json
import os

import numpy as np
import pytest

import dvpari

GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.json")))
P = dvpari.P


def test_oracle_reproduces_golden(oracle):
    O = oracle
    G = O.generator()
    for g in GOLD["gf"]:
        a, b = int(g["a"], 16), int(g["b"], 16)
        assert O.gf_mul(a, b) == int(g["mul"], 16) and O.gf_sqr(a) == int(g["sqr_a"], 16) and O.gf_inv(a) == int(g["inv_a"], 16)
    for pt in GOLD["points"]:
        p = O.pt_mul(G, int(pt["k"], 16))
        assert O.pt_encode(p).hex() == pt["xsk233_candidate"]
        if pt["x"] is not None:
            assert O.pt_xy(p) == (int(pt["x"], 16), int(pt["y"], 16))
    e = GOLD["ecfft"]
    dom = O.Domain(e["log_n2"])
    assert [hex(v) for v in dom.leaves()] == e["leaves"]
    assert [hex(v) for v in dom.extend([int(v, 16) for v in e["evals_on_D"]])] == e["extend_to_Dprime"]
    assert hex(dom.vanish_at(0, 12345)) == e["Z_D_at_12345"]
    t = GOLD["toy"]
    r1cs, pub, priv = O.toy_r1cs()
    td = O.trapdoor(*[int(v, 16) for v in t["trapdoor"]])
    srs = O.Srs(r1cs, dom, td)
    assert srs.g_k30().tobytes().hex() == t["g_k"]
    proof, rc, _ = O.prove(r1cs, dom, srs, O.mont_array([1] + pub + priv))
    assert rc == 0 and proof.hex() == t["proof118"] and O.verify(td, pub, proof)
    assert hex(O.transcript_alpha(proof[:30], pub)) == t["alpha"]


@pytest.mark.gpu
def test_gpu_reproduces_golden_toy_proof(oracle):
    O = oracle
    t = GOLD["toy"]
    ctx = dvpari.Context(0)
    for slot, key in enumerate(["g_m", "g_q", "g_k"]):
        ctx.srs_load(slot, bytes.fromhex(t[key]))
    r1cs, pub, priv = O.toy_r1cs()
    dom = dvpari.Domain(ctx, GOLD["ecfft"]["log_n2"])
    assert [hex(v) for v in dvpari.fr_from_mont(dom.leaves())] == GOLD["ecfft"]["leaves"]
    ev = dvpari.fr_to_mont([int(v, 16) for v in GOLD["ecfft"]["evals_on_D"]])
    assert [hex(v) for v in dvpari.fr_from_mont(dom.extend(ev))] == GOLD["ecfft"]["extend_to_Dprime"]
    inst = dvpari.R1CSInstance(ctx, r1cs.nrows, r1cs.k, r1cs.nwires, r1cs.rowptr, r1cs.wire, r1cs.coeff, r1cs.coeffs)
    proof = dvpari.Prover(ctx, dom, inst, 0, 1, 2).prove(dvpari.fr_to_mont(pub), dvpari.fr_to_mont(priv))
    assert proof.hex() == t["proof118"]
    ctx.close()
