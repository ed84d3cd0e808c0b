"""Oracle ECFFT (oracle/ecfft.c): domain constants, leaf order and extend against first principles.

Restates the reference's own checks: extend == interpolate-then-evaluate (ec_fft.rs:883-907), even
leaves of the 2n tree are the n tree (ec_fft.rs:633-645), vanishing polynomial == prod (X - r)
(ec_fft.rs:820-880); the curve constants are those hard-coded at ec_fft.rs:209-229."""
import random

import pytest

P = 3450873173395281893717377931138512760570940988862252126328087024741343
A = 2125753088427212854352924174339172498722499297750753614229533284661082
B = 3303427382072851929105738691313541325219445842218525662544269869787589
GX = 1969398527398874941115360315313056361667745675958024267654083765592400
GY = 917696706299601920847965073366118878832337776859300472447868491055982
CX = 1557215852494830750811239888869886110709986867282698163663807961412586
CY = 2302954593454110051167704558708330032236229062988890422530712548754008


def sw_add(p, q):
    if p is None:
        return q
    if q is None:
        return p
    (x1, y1), (x2, y2) = p, q
    if x1 == x2:
        if (y1 + y2) % P == 0:
            return None
        lam = (3 * x1 * x1 + A) * pow(2 * y1, -1, P) % P
    else:
        lam = (y2 - y1) * pow(x2 - x1, -1, P) % P
    x3 = (lam * lam - x1 - x2) % P
    return x3, (lam * (x1 - x3) - y1) % P


def lagrange_eval(xs, ys, x):
    tot = 0
    for i, (xi, yi) in enumerate(zip(xs, ys)):
        num, den = 1, 1
        for j, xj in enumerate(xs):
            if i != j:
                num = num * (x - xj) % P
                den = den * (xi - xj) % P
        tot = (tot + yi * num * pow(den, -1, P)) % P
    return tot


def test_curve_constants():
    assert (GY * GY - (GX**3 + A * GX + B)) % P == 0
    assert (CY * CY - (CX**3 + A * CX + B)) % P == 0
    g = (GX, GY)
    for _ in range(27):
        g = sw_add(g, g)
    assert g is not None and g[1] == 0  # 2^27 * G has order 2: G has exact order 2^28
    assert sw_add(g, g) is None


@pytest.mark.parametrize("log_n2", [1, 2, 3, 4, 5, 6])
def test_leaves_are_coset_x_coordinates(oracle, log_n2):
    dom = oracle.Domain(log_n2)
    g = (GX, GY)
    for _ in range(28 - log_n2):
        g = sw_add(g, g)
    pt, want = (CX, CY), []
    for _ in range(1 << log_n2):
        want.append(pt[0])
        pt = sw_add(pt, g)
    assert dom.leaves() == want


def test_even_leaves_of_2n_tree_are_the_n_tree(oracle):
    """ec_fft.rs:633-645 / :1043-1054: subtree = even leaves; shifted tree = odd leaves."""
    big, small = oracle.Domain(6).leaves(), oracle.Domain(5).leaves()
    assert big[0::2] == small


@pytest.mark.parametrize("log_n2", [2, 3, 4, 5, 6, 7])
def test_extend_is_interpolation(oracle, log_n2):
    rnd = random.Random(100 + log_n2)
    dom = oracle.Domain(log_n2)
    lv = dom.leaves()
    d, d2 = lv[0::2], lv[1::2]
    ev = [rnd.randrange(P) for _ in range(dom.n)]
    got = dom.extend(ev)
    assert got == [lagrange_eval(d, ev, x) for x in d2]
    # a low-degree polynomial extends to its own values (i(X) in the prover has degree k-1)
    poly = lambda x: (7 + 11 * x + 13 * x * x) % P
    if dom.n >= 4:
        assert dom.extend([poly(x) for x in d]) == [poly(x) for x in d2]


@pytest.mark.parametrize("log_n2", [2, 4, 6])
def test_vanishing_polynomial_chain_rule(oracle, log_n2):
    rnd = random.Random(200 + log_n2)
    dom = oracle.Domain(log_n2)
    lv = dom.leaves()
    halves = [lv[0::2], lv[1::2]]
    for shift in (0, 1):
        s, other = halves[shift], halves[1 - shift]
        x = rnd.randrange(P)
        want = 1
        for r in s:
            want = want * (x - r) % P
        assert dom.vanish_at(shift, x) == want
        assert dom.vanish_at(shift, s[1]) == 0
        deriv = oracle.mont_array_to_ints(dom.vanish_derivative_on_roots_mont(shift))
        for i, si in enumerate(s):
            w = 1
            for j, sj in enumerate(s):
                if i != j:
                    w = w * (si - sj) % P
            assert deriv[i] == w
        oth = oracle.mont_array_to_ints(dom.vanish_on_other_mont(shift))
        for i, t in enumerate(other):
            w = 1
            for r in s:
                w = w * (t - r) % P
            assert oth[i] == w
