"""Full-size configurations of BASELINE.json that the oracle cannot recompute outright, checked through
size-independent properties against oracle-side arithmetic:

  config #5  MSM over 2^24 points P_i = A + i Q: sum k_i P_i = (sum k_i) A + (sum i k_i) Q  (closed form; the two
             scalar multiplications and the point chain are the oracle's)
  config #3  extend at n = 2^20 (and 2^22, the prove size): the device's D -> D' extension of random evaluations is
             compared at 64 random indices of D' with the oracle's O(n) barycentric evaluation of the same interpolant
             (evaluate_poly_at_alpha_using_barycentric_weights, /root/reference/src/ec_fft.rs:455-491), the check
             SURVEY section 8d prescribes; the 2n leaves are compared byte for byte on the way."""
import random

import numpy as np
import pytest

import dvpari

pytestmark = pytest.mark.gpu
P = dvpari.P


@pytest.fixture(scope="module")
def ctx():
    c = dvpari.Context(0)
    yield c
    c.close()


def test_msm_2_24_closed_form(ctx, oracle):
    O = oracle
    n = 1 << 24
    G = O.generator()
    a, q = O.pt_mul(G, 0x2468ACE13579BDF), O.pt_mul(G, 0x1F2E3D4C5B6A7988)
    pts = O.chain_points(n, a, q)
    enc = O.encode_batch(pts)
    del pts
    ctx.srs_load(0, enc)
    del enc
    sc = dvpari.random_fr_mont(n, 0xD5A10024)
    s0, s1 = O.sum_weighted(sc)
    want = O.pt_encode(O.pt_add(O.pt_mul(a, s0), O.pt_mul(q, s1)))
    try:
        assert ctx.multi_scalar_mul(sc, 0) == want
        assert ctx.msm_stats()["tables"] == 1
        # the batched call at this size (separate-launch path: a read-back in the middle of every MSM) and on a 2^20
        # prefix of the slot (plain layout: the sub-range is too short for the tables; persistent path, sort ahead)
        sc2 = dvpari.random_fr_mont(n, 0xD5A10025)
        t0, t1 = O.sum_weighted(sc2)
        want2 = O.pt_encode(O.pt_add(O.pt_mul(a, t0), O.pt_mul(q, t1)))
        assert ctx.multi_scalar_mul_batch([sc, sc2, sc], 0) == [want, want2, want]
        m = 1 << 20
        u0, u1 = O.sum_weighted(sc[:m])
        v0, v1 = O.sum_weighted(sc2[:m])
        wm = [O.pt_encode(O.pt_add(O.pt_mul(a, u0), O.pt_mul(q, u1))), O.pt_encode(O.pt_add(O.pt_mul(a, v0), O.pt_mul(q, v1)))]
        assert ctx.multi_scalar_mul_batch([sc[:m], sc2[:m], sc[:m], sc2[:m], sc[:m]], 0) == [wm[0], wm[1], wm[0], wm[1], wm[0]]
        # the plain layout (one bucket set per window, no precomputed multiples) gives the same group element
        ctx.set("msm_tables", 0)
        assert ctx.multi_scalar_mul(sc, 0) == want
        assert ctx.msm_stats()["tables"] == 0
    finally:
        ctx.set("msm_tables", 1)
        ctx.srs_free(0)


@pytest.mark.parametrize("log_n", [20, 22])
def test_extend_fullsize_against_barycentric(ctx, oracle, log_n):
    O = oracle
    od = O.Domain(log_n + 1, light=True)
    gd = dvpari.Domain(ctx, log_n + 1)
    n = od.n
    leaves = gd.leaves()
    assert leaves.tobytes() == od.leaves_mont().tobytes()
    npoly = 3 if log_n <= 20 else 2
    ev = np.stack([dvpari.random_fr_mont(n, 0xD5A10006 + 16 * log_n + p) for p in range(npoly)])
    got = gd.extend(ev)
    rnd = random.Random(log_n)
    idx = sorted({0, 1, n // 2 - 1, n // 2, n - 1} | {rnd.randrange(n) for _ in range(59)})
    xs = leaves[1::2][idx]
    for p in range(npoly):
        want = od.bary_eval_mont(ev[p], xs)
        for j, i in enumerate(idx):
            assert got[p][i].tobytes() == want[j].tobytes(), (p, i)
    gd.close()
