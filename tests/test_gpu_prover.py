"""Fr-side hot path on the device against the oracle: domain, extend, R1CS rows, full prove.

Mirrors the reference's tests: leaf order (ec_fft.rs:633-645), extend == interpolation (ec_fft.rs:883-907),
and setup -> prove -> verify on the toy circuit (dvsnark_test.rs:131-180) with byte-equal proofs."""
import ctypes as C
import random

import numpy as np
import pytest

import dvpari
from test_oracle_protocol import random_r1cs

pytestmark = pytest.mark.gpu
P = dvpari.P


@pytest.fixture(scope="module")
def ctx():
    c = dvpari.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("log_n2", [2, 3, 5, 8, 13])
def test_domain_and_precomputes(ctx, oracle, log_n2):
    O = oracle
    od = O.Domain(log_n2)
    gd = dvpari.Domain(ctx, log_n2)
    assert gd.leaves().tobytes() == od.leaves_mont().tobytes()
    z, w = gd.precomputes()
    # bar_wts = 1/Z'_D(d_i), z_vals2inv = 1/Z_D(d'_i)
    want_w = od.vanish_derivative_on_roots_mont(0)
    want_z = od.vanish_on_other_mont(0)
    O.lib().fr_batch_inv(want_w.ctypes.data_as(C.c_void_p), C.c_size_t(want_w.shape[0]))
    O.lib().fr_batch_inv(want_z.ctypes.data_as(C.c_void_p), C.c_size_t(want_z.shape[0]))
    assert w.tobytes() == want_w.tobytes()
    assert z.tobytes() == want_z.tobytes()
    x = 0x1234567890ABCDEF1234567890ABCDEF
    for shift in (0, 1):
        got = dvpari.fr_from_mont(gd.vanish_at(shift, dvpari.fr_to_mont([x])[0]))[0]
        assert got == od.vanish_at(shift, x)
    gd.close()


@pytest.mark.parametrize("log_n2", [2, 4, 9, 15])
def test_extend_matches_oracle(ctx, oracle, log_n2):
    O = oracle
    od = O.Domain(log_n2)
    gd = dvpari.Domain(ctx, log_n2)
    n = od.n
    ev = np.stack([dvpari.random_fr_mont(n, 300 + log_n2 + p) for p in range(3)])
    got = gd.extend(ev)
    for p in range(3):
        assert got[p].tobytes() == od.extend_mont(ev[p]).tobytes()
    # zeros and a constant stay what they are
    const = np.tile(dvpari.fr_to_mont([7]), (n, 1))
    assert gd.extend(const).tobytes() == const.tobytes()
    with pytest.raises(dvpari.DvpError):
        gd.extend(ev[0][: n // 2])
    gd.close()


def _load_circuit(ctx, r1cs):
    return dvpari.R1CSInstance(ctx, r1cs.nrows, r1cs.k, r1cs.nwires, r1cs.rowptr, r1cs.wire, r1cs.coeff, r1cs.coeffs)


def test_r1cs_rows_match_oracle(ctx, oracle):
    O = oracle
    rnd = random.Random(41)
    r1cs, w = random_r1cs(O, rnd, 200, 1 + 2 + 230, 2)
    od = O.Domain(r1cs.n.bit_length())
    gd = dvpari.Domain(ctx, r1cs.n.bit_length())
    inst = _load_circuit(ctx, r1cs)
    wm = O.mont_array(w)
    want, bad = O.r1cs_eval(r1cs, od, wm)
    assert bad == -1
    got = inst.eval(gd, wm)
    for g, x in zip(got, want):
        assert g.tobytes() == x.tobytes()
    # an unsatisfied row is reported with its index (reference: assert_eq! panic, proving.rs:389-395)
    w_bad = list(w)
    w_bad[1 + 2 + 30 + 17] = (w_bad[1 + 2 + 30 + 17] + 1) % P
    _, want_bad = O.r1cs_eval(r1cs, od, O.mont_array(w_bad))
    with pytest.raises(dvpari.DvpError) as e:
        inst.eval(gd, O.mont_array(w_bad))
    assert e.value.code == 6 and str(want_bad) in str(e.value)
    inst.close()
    gd.close()


def _prove_both(ctx, O, r1cs, w, td):
    log_n2 = r1cs.n.bit_length()
    od = O.Domain(log_n2)
    srs = O.Srs(r1cs, od, td)
    wm = O.mont_array(w)
    want_proof, rc, want_st = O.prove(r1cs, od, srs, wm, want_stages=True)
    assert rc == 0
    gd = dvpari.Domain(ctx, log_n2)
    inst = _load_circuit(ctx, r1cs)
    ctx.srs_load(0, srs.g_m30())
    ctx.srs_load(1, srs.g_q30())
    ctx.srs_load(2, srs.g_k30())
    prover = dvpari.Prover(ctx, gd, inst, 0, 1, 2)
    k = r1cs.k
    got_proof, got_st = prover.prove(wm[1:1 + k], wm[1 + k:], want_stages=True)
    # commit_p = msm(w, g_m) + msm(q, g_q) (proving.rs:463-515) as two MSMs and as one MSM over g_m | g_q: same bytes
    for joint in (0, 1):
        ctx.set("prove_joint", joint)
        assert prover.prove(wm[1:1 + k], wm[1 + k:]) == got_proof, f"prove_joint={joint}"
    ctx.set("prove_joint", -1)
    # the joint copy follows a reloaded slot (same points here, new version): still the same proof
    ctx.srs_load(1, srs.g_q30())
    assert prover.prove(wm[1:1 + k], wm[1 + k:]) == got_proof
    n = r1cs.n
    names = ["a", "b", "c", "i", "a'", "b'", "c'", "i'", "q", "k_a", "k_b", "k_r", "k_r(2)"]
    for s, name in enumerate(names):
        assert got_st[s * n:(s + 1) * n].tobytes() == want_st[s * n:(s + 1) * n].tobytes(), name
    assert got_proof == want_proof
    assert O.verify(td, w[1:1 + k], got_proof)
    prover.close()
    inst.close()
    gd.close()
    return got_proof


def test_toy_prove_is_bit_exact_and_verifies(ctx, oracle):
    """dvsnark_test.rs:131-180 through the C ABI: same 118 bytes as the oracle, verifier accepts."""
    O = oracle
    r1cs, pub, priv = O.toy_r1cs()
    rnd = random.Random(43)
    td = O.trapdoor(rnd.randrange(1, P), rnd.randrange(1, P), rnd.randrange(1, P))
    proof = _prove_both(ctx, O, r1cs, [1] + pub + priv, td)
    assert dvpari.proof_from_bits(dvpari.proof_to_bits(proof)) == proof
    assert len(dvpari.proof_to_bits(proof)) == 944


@pytest.mark.parametrize("nrows,k", [(13, 2), (200, 3), (1000, 2), (300, 64)])  # k = 64: a two-chunk public-input hash
def test_random_circuit_prove_is_bit_exact(ctx, oracle, nrows, k):
    O = oracle
    rnd = random.Random(500 + nrows)
    r1cs, w = random_r1cs(O, rnd, nrows, 1 + k + nrows + 7, k)
    td = O.trapdoor(rnd.randrange(1, P), rnd.randrange(1, P), rnd.randrange(1, P))
    _prove_both(ctx, O, r1cs, w, td)


def test_prove_rejects_bad_witness_and_lengths(ctx, oracle):
    O = oracle
    r1cs, pub, priv = O.toy_r1cs()
    td = O.trapdoor(3, 5, 7)
    od = O.Domain(4)
    srs = O.Srs(r1cs, od, td)
    gd = dvpari.Domain(ctx, 4)
    inst = _load_circuit(ctx, r1cs)
    ctx.srs_load(0, srs.g_m30())
    ctx.srs_load(1, srs.g_q30())
    ctx.srs_load(2, srs.g_k30())
    prover = dvpari.Prover(ctx, gd, inst, 0, 1, 2)
    bad_priv = list(priv)
    bad_priv[0] += 1
    with pytest.raises(dvpari.DvpError) as e:
        prover.prove(O.mont_array(pub), O.mont_array(bad_priv))
    assert e.value.code == 6
    with pytest.raises(dvpari.DvpError) as e:
        prover.prove(O.mont_array(pub), O.mont_array(priv[:-1]))
    assert e.value.code == 5
    ctx.srs_load(1, srs.g_q30()[:4])
    with pytest.raises(dvpari.DvpError) as e:
        dvpari.Prover(ctx, gd, inst, 0, 1, 2)
    assert e.value.code == 5
    prover.close()
    inst.close()
    gd.close()


@pytest.mark.parametrize("log_n", [1, 2, 3, 6, 10])
def test_enter_evaluates_on_the_leaves(ctx, log_n):
    """FFTree::enter (ec_fft.rs:317,411): coefficients -> values on the n leaves, against Horner with big integers."""
    n = 1 << log_n
    rnd = random.Random(70 + log_n)
    coeffs = [rnd.randrange(P) for _ in range(n)]
    plan = dvpari.EcfftPlan(ctx, log_n)
    got = dvpari.fr_from_mont(plan.enter(dvpari.fr_to_mont(coeffs)))
    gd = dvpari.Domain(ctx, max(2, log_n))
    leaves = dvpari.fr_from_mont(gd.leaves())
    if log_n == 1:
        leaves = leaves[0::2]  # the 2-leaf tree is the even half of the 4-leaf tree
    for i, s in enumerate(leaves):
        acc = 0
        for c in reversed(coeffs):
            acc = (acc * s + c) % P
        assert got[i] == acc, i
    with pytest.raises(dvpari.DvpError):
        plan.enter(dvpari.fr_to_mont(coeffs[:-1] if n > 1 else []))
    plan.close()
    gd.close()


def test_enter_2_16_sparse_polynomial(ctx):
    """Full-size property: a polynomial with a few non-zero coefficients is evaluated directly at sampled leaves."""
    log_n = 16
    n = 1 << log_n
    rnd = random.Random(71)
    idx = sorted(rnd.sample(range(n), 6) + [0, n - 1])
    vals = {i: rnd.randrange(1, P) for i in idx}
    coeffs = np.zeros((n, 4), dtype=np.uint64)
    coeffs[idx] = dvpari.fr_to_mont([vals[i] for i in idx])
    plan = dvpari.EcfftPlan(ctx, log_n)
    got = plan.enter(coeffs)
    gd = dvpari.Domain(ctx, log_n)
    leaves = gd.leaves()
    for j in [0, 1, 2, n // 2, n - 1] + [rnd.randrange(n) for _ in range(20)]:
        s = dvpari.fr_from_mont(leaves[j:j + 1])[0]
        want = sum(v * pow(s, i, P) for i, v in vals.items()) % P
        assert dvpari.fr_from_mont(got[j:j + 1])[0] == want, j
    plan.close()
    gd.close()


@pytest.mark.parametrize("log_n", [1, 2, 3, 4, 7, 11])
def test_exit_inverts_enter(ctx, log_n):
    """FFTree::exit (ec_fft.rs:266,897): values on the leaves -> coefficients.  exit(enter(c)) == c for random c, and
    the values of an explicitly evaluated polynomial interpolate back to its coefficients (ec_fft.rs:883-907)."""
    n = 1 << log_n
    rnd = random.Random(80 + log_n)
    coeffs = [rnd.randrange(P) for _ in range(n)]
    plan = dvpari.EcfftPlan(ctx, log_n)
    cm = dvpari.fr_to_mont(coeffs)
    ev = plan.enter(cm)
    assert plan.exit(ev).tobytes() == cm.tobytes()
    # independent of enter: Horner values of a low-degree polynomial on the leaves
    gd = dvpari.Domain(ctx, max(2, log_n))
    leaves = dvpari.fr_from_mont(gd.leaves())
    if log_n == 1:
        leaves = leaves[0::2]
    deg = min(n, 5)
    small = [rnd.randrange(P) for _ in range(deg)]
    vals = []
    for s in leaves:
        acc = 0
        for c in reversed(small):
            acc = (acc * s + c) % P
        vals.append(acc)
    got = dvpari.fr_from_mont(plan.exit(dvpari.fr_to_mont(vals)))
    assert got == small + [0] * (n - deg)
    plan.close()
    gd.close()


def test_exit_gives_the_vanishing_polynomial(ctx):
    """vanish via exit, as the reference computes z_poly (ec_fft.rs:241-283): the values of Z_D on the 2n-leaf tree
    (zero on D, the chain-rule values on D') interpolate to a monic degree-n polynomial that vanishes on D."""
    log_n2 = 9
    n2, n = 1 << log_n2, 1 << (log_n2 - 1)
    gd = dvpari.Domain(ctx, log_n2)
    zinv, _ = gd.precomputes()                      # 1 / Z_D(d'_i)
    z_on_dprime = [pow(v, P - 2, P) for v in dvpari.fr_from_mont(zinv)]
    vals = [0] * n2
    vals[1::2] = z_on_dprime
    plan = dvpari.EcfftPlan(ctx, log_n2)
    c = dvpari.fr_from_mont(plan.exit(dvpari.fr_to_mont(vals)))
    assert c[n] == 1 and not any(c[n + 1:])         # monic of degree n
    leaves = dvpari.fr_from_mont(gd.leaves())
    for s in leaves[0::2][:16] + leaves[0::2][-4:]:
        acc = 0
        for co in reversed(c[:n + 1]):
            acc = (acc * s + co) % P
        assert acc == 0
    # and z_poly.evaluate(x) agrees with the chain rule at an arbitrary point (ec_fft.rs:475)
    x = 0x123456789ABCDEF
    acc = 0
    for co in reversed(c[:n + 1]):
        acc = (acc * x + co) % P
    assert acc == dvpari.fr_from_mont(gd.vanish_at(0, dvpari.fr_to_mont([x])[0]))[0]
    plan.close()
    gd.close()


def test_enter_exit_round_trip_2_16(ctx):
    """BASELINE config #3 shape at a size the suite can afford: 2^16 random coefficients -> leaves -> coefficients."""
    log_n = 16
    plan = dvpari.EcfftPlan(ctx, log_n)
    c = dvpari.random_fr_mont(1 << log_n, 99)
    ev = plan.enter(c)
    assert plan.exit(ev).tobytes() == c.tobytes()
    # linearity of enter: enter(c1 + c2) = enter(c1) + enter(c2) on a sample
    c2 = dvpari.random_fr_mont(1 << log_n, 100)
    s = dvpari.fr_to_mont([(a + b) % P for a, b in zip(dvpari.fr_from_mont(c[:64]), dvpari.fr_from_mont(c2[:64]))])
    csum = c.copy()
    csum[:64] = s
    csum[64:] = c[64:]
    c2z = c2.copy()
    c2z[64:] = 0
    e1, e2 = plan.enter(csum), plan.enter(c2z)
    for j in (0, 1, 12345, (1 << log_n) - 1):
        lhs = dvpari.fr_from_mont(e1[j:j + 1])[0]
        rhs = (dvpari.fr_from_mont(ev[j:j + 1])[0] + dvpari.fr_from_mont(e2[j:j + 1])[0]) % P
        assert lhs == rhs
    plan.close()
