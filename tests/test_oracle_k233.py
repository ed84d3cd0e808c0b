"""Oracle K-233 group law + xsk233 codec (oracle/k233.c).

Group law is pinned against OpenSSL NID_sect233k1.  The reference's own tests for this
layer are algebraic only (curve.rs:198-248, io_utils.rs:253-267); they are restated here.
"""
import ctypes as C
import random

TWO_G_X = 0x01A96A52534C02824C92539163F2ED13243FEB57B45ADBE4CF7EC61957F6


def _be30(v):
    return v.to_bytes(30, "big")


def _ossl_mul(ossl, O, k, p, q=None):
    ox, oy = (C.c_ubyte * 30)(), (C.c_ubyte * 30)()
    x, y = O.pt_xy(p)
    if q is None:
        r = ossl.ossl_ec_mul_add(k.to_bytes(32, "big"), 32, _be30(x), _be30(y), None, None, ox, oy)
    else:
        qx, qy = O.pt_xy(q)
        r = ossl.ossl_ec_mul_add(k.to_bytes(32, "big"), 32, _be30(x), _be30(y), _be30(qx), _be30(qy), ox, oy)
    assert r in (1, 2)
    return None if r == 2 else (int.from_bytes(bytes(ox), "big"), int.from_bytes(bytes(oy), "big"))


def test_group_law_vs_openssl(oracle, ossl):
    O = oracle
    rnd = random.Random(4)
    G = O.generator()
    assert O.lib().k233_on_curve(C.byref(G))
    assert O.pt_xy(O.pt_mul(G, 2))[0] == TWO_G_X
    for _ in range(12):
        k, k2 = rnd.randrange(O.P), rnd.randrange(O.P)
        pk = O.pt_mul(G, k)
        assert O.pt_xy(pk) == _ossl_mul(ossl, O, k, G)
        p2 = O.pt_mul(pk, k2)
        assert O.pt_xy(p2) == _ossl_mul(ossl, O, k2, pk)
        assert O.pt_xy(O.pt_add(pk, p2)) == _ossl_mul(ossl, O, k2, pk, pk)
    assert O.pt_mul(G, O.P).inf
    assert O.pt_xy(O.pt_mul(G, O.P - 1)) == O.pt_xy(O.pt_neg(G))
    assert O.pt_add(G, O.pt_neg(G)).inf
    assert O.pt_xy(O.pt_add(G, G)) == O.pt_xy(O.pt_mul(G, 2))


def test_psm_homomorphism(oracle):
    """curve.rs:198-215: k1*G + k2*G == (k1+k2)*G"""
    O = oracle
    rnd = random.Random(5)
    G = O.generator()
    for _ in range(5):
        k1, k2 = rnd.randrange(O.P), rnd.randrange(O.P)
        lhs = O.pt_add(O.pt_mul(G, k1), O.pt_mul(G, k2))
        assert O.pt_encode(lhs) == O.pt_encode(O.pt_mul(G, (k1 + k2) % O.P))


def test_msm_all_generator(oracle):
    """curve.rs:218-232: msm(scalars, [G]*n) == (sum scalars)*G"""
    O = oracle
    rnd = random.Random(6)
    n = 300
    ks = [rnd.randrange(O.P) for _ in range(n)]
    pts = O.points_to_array([O.generator()] * n)
    got = O.msm(O.mont_array(ks), pts, 2)
    assert O.pt_encode(got) == O.pt_encode(O.pt_mul(O.generator(), sum(ks) % O.P))


def test_codec_round_trip(oracle):
    """curve.rs:236-248 and io_utils.rs:253-267: G and the neutral round-trip and decode as valid."""
    O = oracle
    rnd = random.Random(7)
    for p in [O.generator(), O.pt()] + [O.pt_mul(O.generator(), rnd.randrange(1, O.P)) for _ in range(10)]:
        b = O.pt_encode(p)
        q, ok = O.pt_decode(b)
        assert ok and O.pt_xy(q) == O.pt_xy(p)
    assert O.pt_encode(O.pt()) == bytes(30)


def test_decode_rejects_points_outside_the_group(oracle):
    O = oracle
    rnd = random.Random(8)
    valid = 0
    for _ in range(120):
        w = rnd.getrandbits(233).to_bytes(30, "little")
        q, ok = O.pt_decode(w)
        if ok:
            valid += 1
            assert O.lib().k233_on_curve(C.byref(q)) and O.pt_mul(q, O.P).inf and O.pt_encode(q) == w
    assert 10 < valid < 60  # one w in four names a group element
    assert not O.pt_decode((1).to_bytes(30, "little"))[1]  # w = 1: x = 1, an order-4 point
    assert not O.pt_decode(bytes(29) + b"\x02")[1]  # bit 233 set


def test_tau_adic_scalar_mul(oracle, ossl):
    """k233_mul_fr_tau (width-4 TNAF over the Frobenius map, the algorithm of the reference's xsk233_mul_frob,
    curve.rs:118) against the double-and-add wNAF and against OpenSSL: edge scalars, random scalars, several points,
    the point at infinity, and a point outside E[r] (the ladder's table degenerates there and falls back)."""
    O = oracle
    L = O.lib()
    rnd = random.Random(41)
    G = O.generator()
    P = O.P
    edge = [0, 1, 2, 3, 5, 7, 9, 10, 15, 16, P - 1, P - 2, (P - 1) // 2, 1 << 231, (1 << 231) - 1, 1 << 116, (1 << 117) - 1,
            (1 << 128) - 1, 1 << 128]
    pts = [G, O.pt_mul(G, rnd.randrange(1, P)), O.pt_mul(G, P - 1), O.pt()]
    for p in pts:
        for k in edge + [rnd.randrange(P) for _ in range(150)]:
            a, b = O.pt(), O.pt()
            km = O.fr_mont(k)
            L.k233_mul_fr_tau(C.byref(a), C.byref(p), C.byref(km))
            L.k233_mul_fr_wnaf(C.byref(b), C.byref(p), C.byref(km))
            assert O.pt_xy(a) == O.pt_xy(b), k
    for k in [rnd.randrange(P) for _ in range(10)] + [P - 1, 1 << 231]:
        a = O.pt()
        km = O.fr_mont(k)
        L.k233_mul_fr_tau(C.byref(a), C.byref(G), C.byref(km))
        assert O.pt_xy(a) == _ossl_mul(ossl, O, k, G)
    # N = (0, 1), the point of order 2: tau N = N, the table's denominators vanish
    N = O.pt(0, 1)
    for k in (1, 2, 3, P - 1):
        a, b = O.pt(), O.pt()
        km = O.fr_mont(k)
        L.k233_mul_fr_tau(C.byref(a), C.byref(N), C.byref(km))
        L.k233_mul_fr_wnaf(C.byref(b), C.byref(N), C.byref(km))
        assert O.pt_xy(a) == O.pt_xy(b)


def test_msm_projective_accumulation_degenerate_cases(oracle):
    """k233_msm keeps the per-point products projective and adds them with the general Lopez-Dahab addition: equal
    operands (doubling branch), opposite operands (infinity), infinity on either side, in a fixed order (one thread),
    and a random set on all cores against the sum of single multiplications (curve.rs:218-232 restated)."""
    O = oracle
    P = O.P
    rnd = random.Random(43)
    G = O.generator()
    G2, nG, inf = O.pt_mul(G, 2), O.pt_neg(G), O.pt()
    cases = [
        ([G, G], [1, 1]),                       # acc = G, + G: doubling
        ([G, nG, G2], [1, 1, 1]),               # G - G = infinity, then + 2G
        ([inf, G, inf, G], [5, 1, 7, 1]),       # infinity operands on both sides
        ([G, G, G, G], [3, 3, P - 6, 0]),       # 3G + 3G (doubling) + (-6G) = infinity, + 0
        ([G2, G], [1, 2]),                      # 2G + 2G from different inputs
        ([G, G], [P - 1, 1]),                   # -G + G
    ]
    for pts, ks in cases:
        want = O.pt()
        for p, k in zip(pts, ks):
            want = O.pt_add(want, O.pt_mul(p, k))
        got = O.msm(O.mont_array(ks), O.points_to_array(pts), 1)
        assert O.pt_xy(got) == O.pt_xy(want), ks
    n = 600
    ks = [rnd.randrange(P) for _ in range(n)]
    pts = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(8)]
    plist = [pts[i % 8] for i in range(n)]
    want = O.pt()
    for p, k in zip(plist, ks):
        want = O.pt_add(want, O.pt_mul(p, k))
    assert O.pt_xy(O.msm(O.mont_array(ks), O.points_to_array(plist), 0)) == O.pt_xy(want)
