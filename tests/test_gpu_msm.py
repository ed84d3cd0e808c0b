"""multi_scalar_mul parity (curve.rs:141-158): the CUDA path through the C ABI against the oracle.

Cases follow the reference's tests (curve.rs:198-232) plus the degenerate inputs of SURVEY section 8d:
all points = G, zero / one / p-1 scalars, duplicate and negated points, neutral points, ragged sizes."""
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


def _oracle_msm(O, vals, pts_arr):
    return O.pt_encode(O.msm(O.mont_array(vals), pts_arr, 0))


def test_msm_matches_oracle_random(ctx, oracle):
    O = oracle
    G = O.generator()
    for n, seed in [(1, 1), (2, 2), (3, 3), (33, 4), (1000, 5), (4099, 6)]:
        rnd = random.Random(seed)
        ks = [rnd.randrange(P) for _ in range(n)]
        pts = O.mul_batch(G, O.mont_array([rnd.randrange(1, P) for _ in range(n)]))
        ctx.srs_load(0, O.encode_batch(pts))
        got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks), 0)
        assert got == _oracle_msm(O, ks, pts), n


def test_msm_all_points_generator(ctx, oracle):
    """curve.rs:218-232 (n = 10_000, every point = G): msm == (sum k) * G"""
    O = oracle
    rnd = random.Random(7)
    n = 10_000
    ks = [rnd.randrange(P) for _ in range(n)]
    enc = np.tile(np.frombuffer(O.pt_encode(O.generator()), dtype=np.uint8), (n, 1))
    ctx.srs_load(0, enc)
    got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks), 0)
    assert got == O.pt_encode(O.pt_mul(O.generator(), sum(ks) % P))


def test_msm_degenerate_scalars_and_points(ctx, oracle):
    O = oracle
    rnd = random.Random(8)
    G = O.generator()
    base = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(64)]
    pts = base + base + [O.pt_neg(p) for p in base] + [O.pt()] * 16  # duplicates, negations, neutral points
    n = len(pts)
    arr = O.points_to_array(pts)
    enc = O.encode_batch(arr)
    ctx.srs_load(1, enc)
    cases = {
        "zeros": [0] * n,
        "ones": [1] * n,
        "p-1": [P - 1] * n,
        "same": [rnd.randrange(P)] * n,
        "small": [rnd.randrange(4) for _ in range(n)],
        "mixed": [rnd.choice([0, 1, P - 1, rnd.randrange(P), 1 << 231, (1 << 16) - 1, 1 << 15]) for _ in range(n)],
    }
    for name, ks in cases.items():
        got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks), 1)
        assert got == _oracle_msm(O, ks, arr), name
    # cancelling pairs: k*P + k*(-P) = neutral = 30 zero bytes
    ks = [5] * 64 + [0] * 64 + [5] * 64 + [9] * 16
    assert ctx.multi_scalar_mul(dvpari.fr_to_mont(ks), 1) == bytes(30)


def test_msm_every_window_size(ctx, oracle):
    O = oracle
    rnd = random.Random(9)
    G = O.generator()
    n = 700
    ks = [rnd.randrange(P) for _ in range(n)]
    pts = O.mul_batch(G, O.mont_array([rnd.randrange(1, P) for _ in range(n)]))
    ctx.srs_load(0, O.encode_batch(pts))
    want = _oracle_msm(O, ks, pts)
    for c in (4, 5, 7, 8, 10, 13, 16):
        ctx.set("msm_window_bits", c)
        assert ctx.multi_scalar_mul(dvpari.fr_to_mont(ks), 0) == want, c
    ctx.set("msm_window_bits", 0)


def test_msm_offset_empty_and_length_errors(ctx, oracle):
    O = oracle
    rnd = random.Random(10)
    G = O.generator()
    n = 300
    ks = [rnd.randrange(P) for _ in range(n)]
    pts = O.mul_batch(G, O.mont_array([rnd.randrange(1, P) for _ in range(n)]))
    enc = O.encode_batch(pts)
    ctx.srs_load(2, enc[:100])
    ctx.srs_load(2, enc[100:], append=True)  # g_k_0 | g_k_1 | g_k_2 style concatenation
    got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks[50:250]), 2, offset=50)
    assert got == O.pt_encode(O.msm(O.mont_array(ks[50:250]), O.points_to_array([pts[i] for i in range(50, 250)]), 0))
    assert ctx.multi_scalar_mul(np.zeros((0, 4), dtype=np.uint64), 2) == bytes(30)
    with pytest.raises(dvpari.DvpError) as e:  # curve.rs:142 assert_eq!(scalars.len(), points.len())
        ctx.multi_scalar_mul(dvpari.fr_to_mont([1] * 301), 2)
    assert e.value.code == 5
    # ad-hoc form (srs.rs:422: a 2-term msm with a fresh point)
    got = ctx.multi_scalar_mul_adhoc(dvpari.fr_to_mont(ks[:2]), enc[:2])
    assert got == O.pt_encode(O.msm(O.mont_array(ks[:2]), O.points_to_array([pts[0], pts[1]]), 0))


def test_msm_2_16_config(ctx, oracle):
    """BASELINE config: sect233k1 MSM 2^16 points, random scalars, true random subgroup points."""
    O = oracle
    n = 1 << 16
    sc = dvpari.random_fr_mont(n, 0xD5A10001)
    pts = O.mul_batch(O.generator(), dvpari.random_fr_mont(n, 0xD5A10002))
    ctx.srs_load(0, O.encode_batch(pts))
    got = ctx.multi_scalar_mul(sc, 0)
    assert got == O.pt_encode(O.msm(sc, pts, 0))


def test_msm_2_20_closed_form(ctx, oracle):
    """Full-size property: for points P_i = A + i*Q the sum is (sum k_i) A + (sum i k_i) Q."""
    O = oracle
    n = 1 << 20
    G = O.generator()
    a, q = O.pt_mul(G, 0x1234567), O.pt_mul(G, 0x7654321)
    pts = O.chain_points(n, a, q)
    ctx.srs_load(0, O.encode_batch(pts))
    sc = dvpari.random_fr_mont(n, 0xD5A10007)
    ks = dvpari.fr_from_mont(sc)
    s0 = sum(ks) % P
    s1 = sum(i * k for i, k in enumerate(ks)) % P
    want = O.pt_encode(O.pt_add(O.pt_mul(a, s0), O.pt_mul(q, s1)))
    assert ctx.multi_scalar_mul(sc, 0) == want
    ctx.srs_free(0)


def test_msm_with_precomputed_tables(ctx, oracle):
    """The shared-bucket-set layout over the slot's window multiples T[j] = 2^(j c) P gives the same group element:
    forced on small slots, with degenerate points, sub-ranges that still use the tables and ones that do not."""
    O = oracle
    rnd = random.Random(21)
    G = O.generator()
    ctx.set("msm_tables_min", 32)
    try:
        for n in (40, 700, 5000):
            base = [O.pt_mul(G, rnd.randrange(1, P)) for _ in range(n - 8)]
            pts = base + [base[0], O.pt_neg(base[1]), O.pt(), base[2], base[2], O.pt(), O.pt_neg(base[0]), base[3]]
            arr = O.points_to_array(pts)
            ctx.srs_load(2, O.encode_batch(arr))
            ks = [rnd.randrange(P) for _ in range(n)]
            ks[0], ks[1], ks[2], ks[3] = 0, 1, P - 1, (1 << 231) + 12345
            got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks), 2)
            assert ctx.msm_stats()["tables"] == 1
            assert got == _oracle_msm(O, ks, arr), n
            # a sub-range covering most of the slot keeps the tables, a short one falls back to the plain layout
            off, m = n // 5, n - n // 5 - 3
            got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks[:m]), 2, offset=off)
            assert ctx.msm_stats()["tables"] == 1
            assert got == _oracle_msm(O, ks[:m], O.points_to_array(pts[off:off + m]))
            got = ctx.multi_scalar_mul(dvpari.fr_to_mont(ks[:7]), 2, offset=3)
            assert ctx.msm_stats()["tables"] == 0
            assert got == _oracle_msm(O, ks[:7], O.points_to_array(pts[3:10]))
        # every point = G and equal scalars: one bucket carries everything, the tangent case in every round
        n = 3000
        enc = np.tile(np.frombuffer(O.pt_encode(G), dtype=np.uint8), (n, 1))
        ctx.srs_load(2, enc)
        got = ctx.multi_scalar_mul(dvpari.fr_to_mont([5] * n), 2)
        assert got == O.pt_encode(O.pt_mul(G, 5 * n % P))
        # appending to the slot drops the tables
        ctx.srs_append(2, enc[:10])
        got = ctx.multi_scalar_mul(dvpari.fr_to_mont([7] * (n + 10)), 2)
        assert got == O.pt_encode(O.pt_mul(G, 7 * (n + 10) % P))
        ctx.set("msm_tables", 0)
        got = ctx.multi_scalar_mul(dvpari.fr_to_mont([7] * (n + 10)), 2)
        assert ctx.msm_stats()["tables"] == 0 and got == O.pt_encode(O.pt_mul(G, 7 * (n + 10) % P))
    finally:
        ctx.set("msm_tables", 1)
        ctx.set("msm_tables_min", 1 << 15)
        ctx.srs_free(2)


def test_msm_beside_saturating_work_on_the_same_gpu(oracle):
    """Scheduling stress (k_plan look-back): MSMs whose kernels wait on each other across blocks -- the
    persistent k_accumulate (grid barriers, cooperative launch) at 2^17 points and the ticketed k_plan look-back of the
    separate-launch path at 2^15 and 2^22/8 -- must give the same bytes while two other contexts keep the same GPU
    saturated with extends and with MSMs of their own.  A scheduling assumption that does not hold shows up as a
    time-out (the barriers give up after 4 s and the call returns an error) or as a wrong result, not as a hang."""
    import threading

    O = oracle
    G = O.generator()
    ctxs = [dvpari.Context(0) for _ in range(3)]
    try:
        sizes = [1 << 17, 1 << 15, 1 << 19]
        jobs = []
        for i, n in enumerate(sizes):
            ctx = ctxs[i % 2]
            slot = i
            ctx.srs_random(slot, n, 40 + i)
            sc = dvpari.random_fr_mont(n, 50 + i)
            d = ctx.dev_alloc(n * 32)
            ctx.dev_upload(d, sc)
            ref = ctx.multi_scalar_mul_device(d, n, slot)  # quiet GPU
            jobs.append((ctx, slot, d, n, ref))
        # one of the references against the oracle (the others are compared with themselves under load)
        ctx0, slot0, d0, n0, ref0 = jobs[1]
        pts, bad = O.decode_batch(ctx0.srs_read(slot0, 0, n0))
        assert bad < 0 and ref0 == O.pt_encode(O.msm(dvpari.random_fr_mont(n0, 51), pts, 0))
        stop = threading.Event()
        errors = []

        def extend_load():
            try:
                dom = dvpari.Domain(ctxs[2], 19)
                dd = ctxs[2].dev_alloc(3 * (1 << 18) * 32)
                ctxs[2].dev_upload(dd, dvpari.random_fr_mont(3 << 18, 7))
                while not stop.is_set():
                    dom.extend_device(dd, 3)
                ctxs[2].dev_free(dd)
                dom.close()
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        def msm_load(job, reps):
            ctx, slot, d, n, ref = job
            try:
                for _ in range(reps):
                    got = ctx.multi_scalar_mul_device(d, n, slot)
                    if got != ref:
                        errors.append(AssertionError(f"MSM of 2^{n.bit_length() - 1} points changed under load"))
                        return
            except Exception as e:  # noqa: BLE001
                errors.append(e)

        t_ext = threading.Thread(target=extend_load)
        t_ext.start()
        # contexts 0 and 1 run their MSMs at the same time (two engines, four persistent kernels wanting the grid)
        ts = [threading.Thread(target=msm_load, args=(jobs[0], 12)), threading.Thread(target=msm_load, args=(jobs[1], 25))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        ts = [threading.Thread(target=msm_load, args=(jobs[2], 6)), threading.Thread(target=msm_load, args=(jobs[1], 25))]
        for t in ts:
            t.start()
        for t in ts:
            t.join()
        stop.set()
        t_ext.join()
        assert not errors, errors[0]
        for ctx, slot, d, n, ref in jobs:
            ctx.dev_free(d)
    finally:
        for c in ctxs:
            c.close()


@pytest.mark.parametrize("n", [0, 1, 777, 5003, (1 << 17) + 13])
def test_msm_batch_equals_single_calls_and_oracle(ctx, oracle, n):
    """dvp_msm_batch (pipelined calls of multi_scalar_mul over one point vector, curve.rs:141-158): same bytes as one
    dvp_msm per vector and as the oracle -- host vectors and device vectors, batches of 1, 2, 3 and 5, on the
    separate-launch path (small n, mid-MSM read-back) and on the persistent path (2^17 + 13 points)."""
    O = oracle
    pts = O.mul_batch(O.generator(), dvpari.random_fr_mont(n, 900 + n % 7)) if n else None
    ctx.srs_load(3, O.encode_batch(pts) if n else np.zeros((0, 30), dtype=np.uint8))
    vecs = [dvpari.random_fr_mont(n, 1000 + 31 * b + n % 11) for b in range(5)]
    if n:
        vecs[1][: min(n, 40)] = 0  # a vector with zero scalars, one with repeated scalars
        vecs[2][:] = vecs[2][0]
    single = [ctx.multi_scalar_mul(v, 3) for v in vecs]
    nchk = 5 if n <= 5003 else 2
    for b in range(nchk):
        assert single[b] == (O.pt_encode(O.msm(vecs[b], pts, 0)) if n else bytes(30)), b
    for nb in (1, 2, 3, 5):
        assert ctx.multi_scalar_mul_batch(vecs[:nb], 3) == single[:nb], (n, nb)
    d = [ctx.dev_alloc(max(32, n * 32)) for _ in vecs]
    try:
        for p, v in zip(d, vecs):
            if n:
                ctx.dev_upload(p, v)
        for nb in (1, 4, 5):
            assert ctx.multi_scalar_mul_batch(d[:nb], 3, on_device=True, n=n) == single[:nb], (n, nb)
        # a batch followed by a plain call and another batch: the staging buffers and landing zones are reusable
        assert ctx.multi_scalar_mul(vecs[4], 3) == single[4]
        assert ctx.multi_scalar_mul_batch(list(reversed(vecs)), 3) == list(reversed(single))
        # with the stage timers on the batch runs unpipelined and still returns the same bytes
        ctx.set("timing", 1)
        assert ctx.multi_scalar_mul_batch(vecs[:3], 3) == single[:3]
        ctx.set("timing", 0)
    finally:
        for p in d:
            ctx.dev_free(p)
        ctx.srs_free(3)


def test_msm_batch_rejects_bad_arguments(ctx):
    ctx.srs_random(3, 100, 5)
    v = dvpari.random_fr_mont(100, 1)
    with pytest.raises(dvpari.DvpError):
        ctx.multi_scalar_mul_batch([v, v[:50]], 3)  # ragged batch
    with pytest.raises(dvpari.DvpError):
        ctx.multi_scalar_mul_batch([dvpari.random_fr_mont(101, 1)], 3)  # longer than the slot (curve.rs:142)
    assert ctx.multi_scalar_mul_batch([], 3) == []
    ctx.srs_free(3)
