"""Host-side checks of the synthetic SP1-shaped circuit generator (no GPU): shape, the level discipline that the
device solver relies on, and -- with a pure-Python solve of a small instance -- that the oracle accepts the rows."""
import numpy as np

import dvpari
import synth

P = dvpari.P


def _rows(circ, which):
    rp = circ["rowptr"][which].astype(np.int64)
    return np.repeat(np.arange(circ["nrows"]), np.diff(rp))


def test_shape_and_levels():
    circ = synth.synth_r1cs(16)
    n, nrows, k, nl = circ["n"], circ["nrows"], circ["k"], circ["nlevels"]
    assert n == 1 << 16 and nrows == n - 1000 and circ["nwires"] == 1 + k + nrows + circ["nfree"]
    terms = sum(len(w) for w in circ["wire"])
    assert 9.5 < terms / nrows < 11.5  # SP1: ~10.4 terms per row (src/artifacts.rs:100)
    for which in range(3):
        w = circ["wire"][which].astype(np.int64)
        rows = _rows(circ, which)
        assert w.max() < circ["nwires"] and circ["coeff"][which].max() < circ["coeffs_mont"].shape[0]
        fresh = (w >= 1 + k) & (w < 1 + k + nrows)
        src = w - 1 - k
        if which == 2:
            own = fresh & (src == rows)
            assert own.sum() == nrows
            last = circ["rowptr"][2][1:].astype(np.int64) - 1
            assert (w[last] == 1 + k + np.arange(nrows)).all() and (circ["coeff"][2][last] == 0).all()
            fresh &= ~own
        assert ((src[fresh] % nl) < (rows[fresh] % nl)).all()
        assert (rows[fresh] - src[fresh]).max() <= 1000 + nl
    assert dvpari.fr_from_mont(circ["coeffs_mont"][:3]) == [1, P - 1, 2]


def test_small_instance_is_satisfiable(oracle):
    O = oracle
    circ = synth.synth_r1cs(6, seed=5, nlevels=4)
    w = dvpari.fr_from_mont(synth.synth_assignment(circ, seed=9))
    coeffs = dvpari.fr_from_mont(circ["coeffs_mont"])
    k, nrows, nl = circ["k"], circ["nrows"], circ["nlevels"]

    def dot(which, r):
        lo, hi = int(circ["rowptr"][which][r]), int(circ["rowptr"][which][r + 1])
        return sum(coeffs[int(circ["coeff"][which][p])] * w[int(circ["wire"][which][p])] for p in range(lo, hi)) % P

    for lvl in range(nl):  # what k_r1cs_solve_level does on the device
        for r in range(lvl, nrows, nl):
            w[1 + k + r] = (w[1 + k + r] + dot(0, r) * dot(1, r) - dot(2, r)) % P
    r1cs = O.R1CS.from_arrays(circ["coeffs_mont"], circ["rowptr"], circ["wire"], circ["coeff"], nrows, k, circ["nwires"])
    od = O.Domain(7)
    _, bad = O.r1cs_eval(r1cs, od, dvpari.fr_to_mont(w))
    assert bad == -1
