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


GOLD2 = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.json")))


def test_oracle_and_host_code_reproduce_golden_v2(oracle):
    """tests/golden/golden_v2.json (make_golden_v2.py): setup scalars, mulgen points, SP1 public input, synthetic proof."""
    import artifacts
    import synth

    O = oracle
    G = O.generator()
    for e in GOLD2["mulgen"]:
        assert O.pt_encode(O.pt_mul(G, int(e["k"], 16))).hex() == e["xsk233_candidate"]
    t = GOLD2["toy_setup"]
    r1cs, _, _ = O.toy_r1cs()
    od = O.Domain(r1cs.n.bit_length())
    scs = O.setup_scalars(r1cs, od, O.trapdoor(*[int(v, 16) for v in t["trapdoor"]]))
    for key, sc in zip(("sc_m", "sc_q", "sc_k"), scs):
        assert [hex(v) for v in dvpari.fr_from_mont(sc)] == t[key]
    kat = GOLD2["sp1_public_input"]
    raw = int.from_bytes(bytes(kat["raw_le_bytes"]), "little")
    assert dvpari.fr_from_mont(artifacts.sp1_generate_scalar_from_raw_public_input(raw))[0] == int(kat["fr"])
    s = GOLD2["synth_2_6"]
    circ = synth.synth_r1cs(6, seed=s["seed"], nlevels=s["nlevels"])
    assert dvpari.fr_from_mont(synth.synth_assignment(circ, seed=s["assignment_seed"]))[-4:] == [int(v, 16) for v in s["witness_tail"]]


@pytest.mark.gpu
def test_gpu_reproduces_golden_v2(oracle):
    import synth

    ctx = dvpari.Context(0)
    # FFTree::enter / exit on the 8-leaf tree
    e = GOLD2["enter_exit"]
    plan = dvpari.EcfftPlan(ctx, e["log_n"])
    coeffs = dvpari.fr_to_mont([int(v, 16) for v in e["coeffs"]])
    vals = dvpari.fr_to_mont([int(v, 16) for v in e["values_on_leaves"]])
    assert plan.enter(coeffs).tobytes() == vals.tobytes()
    assert plan.exit(vals).tobytes() == coeffs.tobytes()
    plan.close()
    # batched fixed-base multiplication
    ks = [int(m["k"], 16) for m in GOLD2["mulgen"]]
    ctx.srs_mulgen(0, dvpari.fr_to_mont(ks))
    got = ctx.srs_read(0, 0, len(ks))
    assert [bytes(r).hex() for r in got] == [m["xsk233_candidate"] for m in GOLD2["mulgen"]]
    # setup scalars of the toy circuit on the device
    O = oracle
    r1cs, _, _ = O.toy_r1cs()
    t = GOLD2["toy_setup"]
    dom = dvpari.Domain(ctx, r1cs.n.bit_length())
    inst = dvpari.R1CSInstance(ctx, r1cs.nrows, r1cs.k, r1cs.nwires, r1cs.rowptr, r1cs.wire, r1cs.coeff, r1cs.coeffs)
    scs = dvpari.setup_scalars(inst, dom, [int(v, 16) for v in t["trapdoor"]])
    for key, sc in zip(("sc_m", "sc_q", "sc_k"), scs):
        assert [hex(v) for v in dvpari.fr_from_mont(sc)] == t[key]
    inst.close()
    dom.close()
    # synthetic circuit: device-solved witness, device setup, device proof == the committed 118 bytes
    s = GOLD2["synth_2_6"]
    circ = synth.synth_r1cs(6, seed=s["seed"], nlevels=s["nlevels"])
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"],
                               circ["coeffs_mont"])
    w = inst.synth_solve(synth.synth_assignment(circ, seed=s["assignment_seed"]), circ["nlevels"])
    k = circ["k"]
    assert hex(dvpari.fr_from_mont(w[1 + k:2 + k])[0]) == s["fresh_wire_0"]
    dom = dvpari.Domain(ctx, 7)
    dvpari.setup(inst, dom, [int(v, 16) for v in s["trapdoor"]], 0, 1, 2)
    prover = dvpari.Prover(ctx, dom, inst, 0, 1, 2)
    assert prover.prove(w[1:1 + k], w[1 + k:]).hex() == s["proof118"]
    prover.close()
    inst.close()
    dom.close()
    ctx.close()
