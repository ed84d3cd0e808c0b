"""Synthetic SP1-shaped sparse R1CS for benchmarks and full-size tests.

The reference proves dumps of SP1's Groth16-wrapper circuit (gnark_r1cs.rs:1-20; ~10.4 terms per row,
src/artifacts.rs:100).  No dump exists offline, so this generator reproduces its shape: n_rows just below a
power of two (padding is exercised), k = 2 public inputs, L : R : O term counts ~ 4 : 3 : 3.4 with geometric
tails, a 4096-entry coefficient table dominated by {1, -1, 2}, 70 % of the wire references local (one of the
~1000 rows before the referencing row), 30 % to free inputs.  Every row's O side ends with the row's own
fresh wire (coefficient one), so a satisfying witness exists for any values of the other wires; rows are
spread over `nlevels` interleaved levels and a row only references fresh wires of lower levels, which lets
dvp_r1cs_synth_solve fill the witness in nlevels passes on the device.
"""
import numpy as np

import dvpari

P = dvpari.P


def _counts(rng, nrows, mean, cap):
    c = rng.geometric(1.0 / mean, size=nrows)
    return np.minimum(c, cap).astype(np.int64)


def synth_r1cs(lg_n, seed=0xD5A10003, k=2, nlevels=16, pad=1000):
    """Returns dict(nrows, n, k, nwires, rowptr[3], wire[3], coeff[3], coeffs_mont, nlevels)."""
    n = 1 << lg_n
    nrows = n - pad if n >= 4 * pad else n - max(1, n // 8)
    rng = np.random.Generator(np.random.PCG64(seed))
    nfree = max(16, nrows // 8)
    nwires = 1 + k + nrows + nfree
    free0 = 1 + k + nrows
    ncoef = 4096
    # coefficient table: 1, -1, 2, powers of two, then uniform field elements
    coeffs = dvpari.random_fr_mont(ncoef, seed ^ 0x5151)
    small = [1, P - 1, 2] + [pow(2, j, P) for j in range(2, 34)]
    coeffs[:len(small)] = dvpari.fr_to_mont(small)
    rowptr, wire, coeff = [], [], []
    rows = np.arange(nrows, dtype=np.int64)
    for which, (mean, cap) in enumerate(((4.0, 24), (3.0, 16), (2.4, 12))):
        cnt = _counts(rng, nrows, mean, cap)
        rp = np.zeros(nrows + 1, dtype=np.int64)
        np.cumsum(cnt, out=rp[1:])
        nt = int(rp[-1])
        r = np.repeat(rows, cnt)
        lvl = r % nlevels
        u = rng.random(nt)
        # local reference: a fresh wire of a lower level, at most ~1000 rows back
        tgt = (rng.random(nt) * lvl).astype(np.int64)           # target level in [0, lvl)
        back = (lvl - tgt) + nlevels * rng.integers(0, max(1, 1000 // nlevels), size=nt)
        src_row = r - back
        local_ok = (lvl > 0) & (src_row >= 0) & (u < 0.70)
        w = free0 + rng.integers(0, nfree, size=nt)             # free input wires
        w = np.where(local_ok, 1 + k + src_row, w)
        pubsel = (~local_ok) & (u > 0.99)                       # a few references to 1 and the public inputs
        w = np.where(pubsel, rng.integers(0, 1 + k, size=nt), w)
        v = rng.random(nt)
        c = np.where(v < 0.60, 0, np.where(v < 0.80, 1, np.where(v < 0.90, 2, rng.integers(3, ncoef, size=nt))))
        if which == 2:
            # append the row's fresh wire with coefficient one as the last O term
            cnt2 = cnt + 1
            rp2 = np.zeros(nrows + 1, dtype=np.int64)
            np.cumsum(cnt2, out=rp2[1:])
            w2 = np.empty(int(rp2[-1]), dtype=np.int64)
            c2 = np.empty_like(w2)
            pos = np.arange(nt, dtype=np.int64) + r             # every earlier row has inserted one extra term
            w2[pos] = w
            c2[pos] = c
            last = rp2[1:] - 1
            w2[last] = 1 + k + rows
            c2[last] = 0
            rp, w, c = rp2, w2, c2
        rowptr.append(rp.astype(np.uint32))
        wire.append(w.astype(np.uint32))
        coeff.append(c.astype(np.uint32))
    return dict(nrows=nrows, n=n, k=k, nwires=nwires, rowptr=rowptr, wire=wire, coeff=coeff, coeffs_mont=coeffs,
                nlevels=nlevels, nfree=nfree)


def synth_assignment(circ, seed=0xD5A10004):
    """[1, public.., private..] with random public / free wires and zeros where the fresh wires go."""
    w = np.zeros((circ["nwires"], 4), dtype=np.uint64)
    w[0] = dvpari.fr_to_mont([1])[0]
    k, nrows = circ["k"], circ["nrows"]
    w[1:1 + k] = dvpari.random_fr_mont(k, seed)
    w[1 + k + nrows:] = dvpari.random_fr_mont(circ["nfree"], seed + 1)
    return w
