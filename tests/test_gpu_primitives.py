"""Device runs of the field / group primitives through the C ABI (dvp_selftest_op) against the oracle."""
import random

import numpy as np
import pytest

import dvpari
from test_hostcheck_primitives import gf_arr, gf_ints, pt_arr, pt_from_row

pytestmark = pytest.mark.gpu
P = dvpari.P


@pytest.fixture(scope="module")
def ctx():
    c = dvpari.Context(0)
    yield c
    c.close()


def test_gf_ops_on_device(ctx, oracle):
    rnd = random.Random(21)
    a = [0, 1, (1 << 233) - 1, 1 << 232] + [rnd.getrandbits(233) for _ in range(2000)]
    b = a[1:] + a[:1]
    A, B = gf_arr(a), gf_arr(b)
    assert gf_ints(ctx.selftest_op(0, A, B)) == [oracle.gf_mul(x, y) for x, y in zip(a, b)]
    assert gf_ints(ctx.selftest_op(1, A)) == [oracle.gf_sqr(x) for x in a]
    assert gf_ints(ctx.selftest_op(2, A[:300])) == [oracle.gf_inv(x) for x in a[:300]]
    # the warp-cooperative forms (gf233_warp.cuh): one warp per product / inverse
    assert gf_ints(ctx.selftest_op(6, A, B)) == [oracle.gf_mul(x, y) for x, y in zip(a, b)]
    assert gf_ints(ctx.selftest_op(7, A[:600])) == [oracle.gf_inv(x) for x in a[:600]]


def test_fr_ops_on_device(ctx):
    rnd = random.Random(22)
    a = [0, 1, P - 1, P - 2, 1 << 231] + [rnd.randrange(P) for _ in range(2000)]
    b = a[3:] + a[:3]
    A = dvpari.fr_to_mont(a).view(np.uint32).reshape(-1, 8)
    B = dvpari.fr_to_mont(b).view(np.uint32).reshape(-1, 8)
    mul = ctx.selftest_op(3, A, B).view(np.uint64).reshape(-1, 4)
    assert dvpari.fr_from_mont(mul) == [x * y % P for x, y in zip(a, b)]
    assert gf_ints(ctx.selftest_op(4, A)) == a


def test_point_add_on_device(ctx, oracle):
    O = oracle
    rnd = random.Random(23)
    G = O.generator()
    pts = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(40)]
    inf = O.pt()
    lhs = pts + [pts[0], pts[1], inf, pts[2], inf]
    rhs = pts[1:] + pts[:1] + [pts[0], O.pt_neg(pts[1]), pts[3], inf, inf]
    got = ctx.selftest_op(5, pt_arr(O, lhs), pt_arr(O, rhs))
    for row, p, q in zip(got, lhs, rhs):
        assert O.pt_xy(pt_from_row(O, row)) == O.pt_xy(O.pt_add(p, q))


def test_codec_round_trip_on_device(ctx, oracle):
    """io_utils.rs:253-267 restated: [G, neutral, G, neutral] and random points survive load -> read."""
    O = oracle
    rnd = random.Random(24)
    G = O.generator()
    pts = [G, O.pt(), G, O.pt()] + [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(200)]
    enc = O.encode_batch(O.points_to_array(pts))
    ctx.srs_load(7, enc)
    assert ctx.srs_size(7) == len(pts)
    assert ctx.srs_read(7, 0, len(pts)).tobytes() == enc.tobytes()
    # CurvePoint::add on encodings (curve.rs:76-82)
    for i in range(0, 20, 2):
        want = O.pt_encode(O.pt_add(pts[i], pts[i + 1]))
        assert ctx.point_add(enc[i].tobytes(), enc[i + 1].tobytes()) == want
    ctx.srs_free(7)


def test_invalid_point_is_reported(ctx, oracle):
    O = oracle
    enc = O.encode_batch(O.points_to_array([O.generator()] * 5)).copy()
    enc[3] = 0
    enc[3, 0] = 1  # w = 1 is not a group element
    with pytest.raises(dvpari.DvpError) as e:
        ctx.srs_load(6, enc)
    assert e.value.code == 4 and "3" in str(e.value)
