"""The kernels' own __host__ __device__ source (gf233.cuh, fr.cuh, k233*.cuh, host_gf.hpp), evaluated on
the CPU through dvp_hostcheck_op and compared with the oracle.  No GPU needed; the GPU runs of the
same source are in test_gpu_primitives.py."""
import random

import numpy as np

import dvpari

P = dvpari.P


def gf_arr(vals):
    return np.array([[(v >> (32 * k)) & 0xFFFFFFFF for k in range(8)] for v in vals], dtype=np.uint32)


def gf_ints(arr):
    return [sum(int(r[k]) << (32 * k) for k in range(8)) for r in arr]


def pt_arr(O, pts):
    out = np.zeros((len(pts), 16), dtype=np.uint32)
    for i, p in enumerate(pts):
        xy = O.pt_xy(p)
        if xy is not None:
            out[i, :8] = gf_arr([xy[0]])[0]
            out[i, 8:] = gf_arr([xy[1]])[0]
    return out


def pt_from_row(O, row):
    x, y = gf_ints([row[:8]])[0], gf_ints([row[8:]])[0]
    return O.pt() if x == 0 else O.pt(x, y)


def test_gf_mul_sqr_inv(oracle):
    rnd = random.Random(11)
    edge = [0, 1, (1 << 233) - 1, 1 << 232, 0x1111111111111111111111111111111111111111111111111111111111, (1 << 233) - 2]
    a = edge + [rnd.getrandbits(233) for _ in range(300)]
    b = a[1:] + a[:1]
    A, B = gf_arr(a), gf_arr(b)
    for op in (0, 6):  # device multiplier, PCLMUL host-tail multiplier
        got = gf_ints(dvpari.hostcheck_op(op, A, B))
        assert got == [oracle.gf_mul(x, y) for x, y in zip(a, b)]
    assert gf_ints(dvpari.hostcheck_op(1, A)) == [oracle.gf_sqr(x) for x in a]
    for op in (2, 7):
        assert gf_ints(dvpari.hostcheck_op(op, A[:60])) == [oracle.gf_inv(x) for x in a[:60]]


def test_fr_ops(oracle):
    rnd = random.Random(12)
    edge = [0, 1, P - 1, P - 2, 1 << 231, (1 << 231) - 1, 2]
    a = edge + [rnd.randrange(P) for _ in range(300)]
    b = a[3:] + a[:3]
    A = dvpari.fr_to_mont(a).view(np.uint32).reshape(-1, 8)
    B = dvpari.fr_to_mont(b).view(np.uint32).reshape(-1, 8)
    mul = dvpari.hostcheck_op(3, A, B).view(np.uint64).reshape(-1, 4)
    assert dvpari.fr_from_mont(mul) == [x * y % P for x, y in zip(a, b)]
    can = dvpari.hostcheck_op(4, A)
    assert gf_ints(can) == a
    add = dvpari.hostcheck_op(12, A, B).view(np.uint64).reshape(-1, 4)
    assert dvpari.fr_from_mont(add) == [(x + y) % P for x, y in zip(a, b)]
    sub = dvpari.hostcheck_op(13, A, B).view(np.uint64).reshape(-1, 4)
    assert dvpari.fr_from_mont(sub) == [(x - y) % P for x, y in zip(a, b)]
    inv = dvpari.hostcheck_op(14, A[:40], B[:40]).view(np.uint64).reshape(-1, 4)
    assert dvpari.fr_from_mont(inv) == [pow(x, -1, P) if x else 0 for x in a[:40]]
    # the bytes are exactly ark's Montgomery limbs, fully reduced
    for row in mul:
        assert sum(int(row[k]) << (64 * k) for k in range(4)) < P


def test_point_add_complete(oracle):
    O = oracle
    rnd = random.Random(13)
    G = O.generator()
    pts = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(12)]
    inf = O.pt()
    lhs = pts + [pts[0], pts[1], inf, pts[2], inf]
    rhs = pts[1:] + pts[:1] + [pts[0], O.pt_neg(pts[1]), pts[3], inf, inf]  # generic, double, negation, infinities
    got = dvpari.hostcheck_op(5, pt_arr(O, lhs), pt_arr(O, rhs))
    for row, p, q in zip(got, lhs, rhs):
        assert O.pt_xy(pt_from_row(O, row)) == O.pt_xy(O.pt_add(p, q))
    # host LD accumulator used by the MSM tail: 2a + b
    got = dvpari.hostcheck_op(11, pt_arr(O, lhs), pt_arr(O, rhs))
    for row, p, q in zip(got, lhs, rhs):
        assert O.pt_xy(pt_from_row(O, row)) == O.pt_xy(O.pt_add(O.pt_add(p, p), q))


def test_codec_matches_oracle(oracle):
    O = oracle
    rnd = random.Random(14)
    G = O.generator()
    pts = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(10)] + [O.pt(), G]
    enc = np.frombuffer(b"".join(O.pt_encode(p) for p in pts), dtype=np.uint8).reshape(-1, 30)
    for op in (9, 10):
        got = dvpari.hostcheck_op(op, pt_arr(O, pts), out_stride=30)
        assert got.tobytes() == enc.tobytes()
    dec = dvpari.hostcheck_op(8, enc, out_stride=65)
    for row, p in zip(dec, pts):
        assert row[64] == 1
        assert O.pt_xy(pt_from_row(O, row[:64].copy().view(np.uint32))) == O.pt_xy(p)
    # random byte strings: validity and value agree with the oracle
    raw = np.array([list(rnd.getrandbits(233).to_bytes(30, "little")) for _ in range(60)], dtype=np.uint8)
    raw[0, 29] = 0x02  # bit 233 set
    raw[1] = 0
    raw[1, 0] = 1  # w = 1
    dec = dvpari.hostcheck_op(8, raw, out_stride=65)
    nvalid = 0
    for row, b in zip(dec, raw):
        q, ok = O.pt_decode(b.tobytes())
        assert bool(row[64]) == ok
        if ok:
            nvalid += 1
            assert O.pt_xy(pt_from_row(O, row[:64].copy().view(np.uint32))) == O.pt_xy(q)
    assert 0 < nvalid < 40


def test_fr29_butterfly_row():
    """fr29_dot2 (29-bit limbs, one reduction for two products, pre-scaled matrix entries) gives exactly
    m0 x0 + m1 x1 in ark's Montgomery form -- the row of an ECFFT butterfly (proving.rs:410-422 via ecfft extend)."""
    rnd = random.Random(29)
    edge = [0, 1, P - 1, P - 2, 1 << 231, (1 << 231) - 1, (1 << 203) - 1, (1 << 29) - 1, 1 << 29]
    vals = edge + [rnd.randrange(P) for _ in range(400)]
    n = len(vals)
    m0, x0 = vals, vals[5:] + vals[:5]
    m1, x1 = vals[11:] + vals[:11], vals[2:] + vals[:2]
    # worst case for the column sums: every operand p - 1
    m0, x0, m1, x1 = [P - 1] + m0, [P - 1] + x0, [P - 1] + m1, [P - 1] + x1
    A = np.concatenate([dvpari.fr_to_mont(m0), dvpari.fr_to_mont(x0)], axis=1).view(np.uint32).reshape(-1, 16)
    B = np.concatenate([dvpari.fr_to_mont(m1), dvpari.fr_to_mont(x1)], axis=1).view(np.uint32).reshape(-1, 16)
    out = np.zeros((n + 1, 8), dtype=np.uint32)
    dvpari._ck(dvpari.lib().dvp_hostcheck_op(15, dvpari._ptr(A), dvpari._ptr(B), dvpari._ptr(out), n + 1))
    got = dvpari.fr_from_mont(out.view(np.uint64).reshape(-1, 4))
    assert got == [(a * b + c * d) % P for a, b, c, d in zip(m0, x0, m1, x1)]
    # and the limbs are canonical (fully reduced): the bytes equal the canonical Montgomery encoding
    want = dvpari.fr_to_mont([(a * b + c * d) % P for a, b, c, d in zip(m0, x0, m1, x1)])
    assert out.view(np.uint64).reshape(-1, 4).tobytes() == want.tobytes()


def test_fr29_normalised_butterfly_row():
    """fr29_muladd: x0 + m x1 with the addend injected into the accumulator above the reduction (one product, one
    reduction, result < 2.5 p before the conditional subtractions) -- the row (1, m) of the normalised ECFFT butterflies."""
    rnd = random.Random(30)
    edge = [0, 1, P - 1, P - 2, 1 << 231, (1 << 231) - 1, (1 << 203) - 1, (1 << 29) - 1, 1 << 29]
    vals = edge + [rnd.randrange(P) for _ in range(400)]
    m, x0, x1 = [P - 1] + vals, [P - 1] + vals[5:] + vals[:5], [P - 1] + vals[2:] + vals[:2]
    # every combination of the edge values as well
    for a in edge:
        for b in edge:
            for c in (0, 1, P - 1):
                m.append(a), x1.append(b), x0.append(c)
    n = len(m)
    A = np.concatenate([dvpari.fr_to_mont(m), dvpari.fr_to_mont(x0)], axis=1).view(np.uint32).reshape(-1, 16)
    B = np.concatenate([dvpari.fr_to_mont(m), dvpari.fr_to_mont(x1)], axis=1).view(np.uint32).reshape(-1, 16)
    out = np.zeros((n, 8), dtype=np.uint32)
    dvpari._ck(dvpari.lib().dvp_hostcheck_op(19, dvpari._ptr(A), dvpari._ptr(B), dvpari._ptr(out), n))
    want = dvpari.fr_to_mont([(c + a * b) % P for a, b, c in zip(m, x1, x0)])
    assert out.view(np.uint64).reshape(-1, 4).tobytes() == want.tobytes()


def test_fr29_semi_reduced_rows():
    """fr29_muladd_semi / fr29_dot2_semi: the rows between the levels of an extend stay below 2^232 with normalised limbs
    and are congruent to the exact row; operands anywhere in [0, 2^232), including the extremes."""
    rnd = random.Random(32)
    top = (1 << 232) - 1
    xs = [0, 1, P - 1, P, P + 1, top, top - 1, 1 << 231, (1 << 231) - 1, top - 12345]
    ms = [0, 1, P - 1, P - 2, 1 << 230, (1 << 29) - 1]
    m, m1, x0, x1 = [], [], [], []
    for a in ms:
        for b in xs:
            for c in xs:
                m.append(a), m1.append(ms[(len(m) * 7) % len(ms)]), x0.append(b), x1.append(c)
    for _ in range(3000):
        m.append(rnd.randrange(P)), m1.append(rnd.randrange(P))
        x0.append(rnd.randrange(1 << 232)), x1.append(rnd.randrange(1 << 232))
    n = len(m)
    raw = lambda v: np.array([[(x >> (64 * i)) & (2**64 - 1) for i in range(4)] for x in v], dtype=np.uint64)
    A = np.concatenate([dvpari.fr_to_mont(m), raw(x0)], axis=1).view(np.uint32).reshape(-1, 16)
    B = np.concatenate([dvpari.fr_to_mont(m1), raw(x1)], axis=1).view(np.uint32).reshape(-1, 16)
    for op in (20, 21):
        out = np.zeros((n, 4), dtype=np.uint64)
        dvpari._ck(dvpari.lib().dvp_hostcheck_op(op, dvpari._ptr(A), dvpari._ptr(B), dvpari._ptr(out), n))
        got = [sum(int(out[i, j]) << (64 * j) for j in range(4)) for i in range(n)]
        assert max(got) < 1 << 232
        if op == 20:
            want = [(c + a * b) % P for a, b, c in zip(m, x1, x0)]
        else:
            want = [(a * c + a1 * b) % P for a, a1, b, c in zip(m, m1, x1, x0)]
        assert [g % P for g in got] == want
        assert any(g >= P for g in got)  # the rows really are only semi-reduced


def test_ld_projective_addition_is_complete(oracle):
    """k233_ld.cuh: the inversion-free addition used by the MSM's reduction trees, on projective operands with
    Z != 1, including equal operands in different representations (doubling branch), opposite operands and infinity."""
    O = oracle
    rnd = random.Random(31)
    G = O.generator()
    pts = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(40)]
    a = pts + [O.pt(), pts[0], pts[1], O.pt(), pts[2]]
    b = pts[7:] + pts[:7] + [pts[3], O.pt(), O.pt_neg(pts[1]), O.pt(), pts[2]]
    A, B = pt_arr(O, a), pt_arr(O, b)
    got = dvpari.hostcheck_op(16, A, B)
    for i, (p, q) in enumerate(zip(a, b)):
        want = O.pt_add(O.pt_add(O.pt_add(p, p), p), q)
        assert O.pt_encode(pt_from_row(O, got[i])) == O.pt_encode(want), i
    out = np.zeros((len(a), 32), dtype=np.uint32)
    dvpari._ck(dvpari.lib().dvp_hostcheck_op(17, dvpari._ptr(A), dvpari._ptr(B), dvpari._ptr(out), len(a)))
    for i, (p, q) in enumerate(zip(a, b)):
        s = O.pt_add(p, q)
        assert O.pt_encode(pt_from_row(O, out[i, :16])) == O.pt_encode(O.pt_add(s, s)), i
        assert not out[i, 16:].any(), i


def test_fr29_three_term_dot_and_add():
    """fr29_dotn<3> + fr29_add: the accumulation step of the R1CS row products (proving.rs:382-396)."""
    rnd = random.Random(30)
    vals = [0, 1, P - 1, P - 2, 1 << 231, (1 << 231) - 1] + [rnd.randrange(P) for _ in range(300)]
    m0, x0 = [P - 1] + vals, [P - 1] + vals[3:] + vals[:3]
    m1, x1 = [P - 1] + vals[7:] + vals[:7], [P - 1] + vals[1:] + vals[:1]
    n = len(m0)
    A = np.concatenate([dvpari.fr_to_mont(m0), dvpari.fr_to_mont(x0)], axis=1).view(np.uint32).reshape(-1, 16)
    B = np.concatenate([dvpari.fr_to_mont(m1), dvpari.fr_to_mont(x1)], axis=1).view(np.uint32).reshape(-1, 16)
    out = np.zeros((n, 8), dtype=np.uint32)
    dvpari._ck(dvpari.lib().dvp_hostcheck_op(18, dvpari._ptr(A), dvpari._ptr(B), dvpari._ptr(out), n))
    # the addend m1 is in Montgomery form like the products, i.e. the value added is m1
    want = [(a * b + c * d + a * d + c) % P for a, b, c, d in zip(m0, x0, m1, x1)]
    assert out.view(np.uint64).reshape(-1, 4).tobytes() == dvpari.fr_to_mont(want).tobytes()
