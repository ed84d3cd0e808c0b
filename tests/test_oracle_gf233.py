"""Oracle GF(2^233) (oracle/gf233.c) pinned against OpenSSL BN_GF2m_* with {233,74,0}."""
import ctypes as C
import random


def _be30(v):
    return v.to_bytes(30, "big")


def _ossl(ossl, op, a, b=0):
    out = (C.c_ubyte * 30)()
    assert ossl.ossl_gf_op(op, _be30(a), _be30(b), out)
    return int.from_bytes(bytes(out), "big")


def test_mul_sqr_inv_sqrt_vs_openssl(oracle, ossl):
    rnd = random.Random(1)
    edge = [1, 2, (1 << 233) - 1, 1 << 232, (1 << 74) | 1]
    vals = edge + [rnd.getrandbits(233) for _ in range(150)]
    for i in range(len(vals) - 1):
        a, b = vals[i], vals[i + 1]
        assert oracle.gf_mul(a, b) == _ossl(ossl, 0, a, b)
        assert oracle.gf_sqr(a) == _ossl(ossl, 1, a)
        assert oracle.gf_inv(a) == _ossl(ossl, 2, a)
        assert oracle.gf_sqrt(a) == _ossl(ossl, 3, a)
    assert oracle.gf_mul(0, vals[7]) == 0 and oracle.gf_inv(0) == 0


def test_portable_multiplier_agrees(oracle, ossl):
    rnd = random.Random(2)
    oracle.lib().gf_set_portable(1)
    try:
        for _ in range(40):
            a, b = rnd.getrandbits(233), rnd.getrandbits(233)
            assert oracle.gf_mul(a, b) == _ossl(ossl, 0, a, b)
    finally:
        oracle.lib().gf_set_portable(0)


def test_trace_and_halftrace(oracle):
    rnd = random.Random(3)
    for _ in range(40):
        a = rnd.getrandbits(233)
        # for x^233 + x^74 + 1 the trace is bit0 + bit159
        assert oracle.gf_trace(a) == ((a & 1) ^ ((a >> 159) & 1))
        if oracle.gf_trace(a) == 0:
            h = oracle.gf_halftrace(a)
            assert oracle.gf_sqr(h) ^ h == a
