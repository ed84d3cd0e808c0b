"""ORACLE (test infrastructure only): ctypes view of oracle/liboracle.so.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (dv-pari_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
P = 3450873173395281893717377931138512760570940988862252126328087024741343  # curve.rs:17
R = 1 << 256


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "all"])
    return so


class Fr(C.Structure):
    _fields_ = [("l", C.c_uint64 * 4)]


class Gf(C.Structure):
    _fields_ = [("w", C.c_uint64 * 4)]


class Pt(C.Structure):
    _fields_ = [("x", Gf), ("y", Gf), ("inf", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.xsk233_decode.restype = C.c_int
        _lib.fr_from_le29.restype = C.c_int
        _lib.gf_from_le30.restype = C.c_int
        _lib.gf_trace.restype = C.c_int
        _lib.k233_on_curve.restype = C.c_int
        _lib.k233_eq.restype = C.c_int
        _lib.k233_mul_bytes.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_size_t]
        _lib.k233_msm.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
        # callers pass numpy addresses: without argtypes ctypes would truncate a Python int to 32 bits
        _lib.fr_batch_inv.argtypes = [C.c_void_p, C.c_size_t]
    return _lib


# ---------------------------------------------------------------- Fr helpers
def int_to_limbs(v):
    return [(v >> (64 * i)) & (2**64 - 1) for i in range(4)]


def limbs_to_int(l):
    return sum(int(l[i]) << (64 * i) for i in range(4))


def fr_mont(v):
    """canonical int -> Fr struct (Montgomery limbs)"""
    f = Fr()
    f.l[:] = int_to_limbs((v % P) * R % P)
    return f


def fr_int(f):
    """Fr struct -> canonical int"""
    return limbs_to_int(f.l) * pow(R, -1, P) % P


def mont_array(vals):
    """list of canonical ints -> (n,4) uint64 array of Montgomery limbs (ark in-memory layout)"""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        out[i] = int_to_limbs((v % P) * R % P)
    return out


def mont_array_to_ints(arr):
    rinv = pow(R, -1, P)
    return [limbs_to_int(row) * rinv % P for row in arr]


# ---------------------------------------------------------------- GF(2^233) helpers
def gf_from_int(v):
    g = Gf()
    g.w[:] = int_to_limbs(v)
    return g


def gf_int(g):
    return limbs_to_int(g.w)


def gf_mul(a, b):
    r = Gf()
    lib().gf_mul(C.byref(r), C.byref(gf_from_int(a)), C.byref(gf_from_int(b)))
    return gf_int(r)


def gf_sqr(a):
    r = Gf()
    lib().gf_sqr(C.byref(r), C.byref(gf_from_int(a)))
    return gf_int(r)


def gf_inv(a):
    r = Gf()
    lib().gf_inv(C.byref(r), C.byref(gf_from_int(a)))
    return gf_int(r)


def gf_sqrt(a):
    r = Gf()
    lib().gf_sqrt(C.byref(r), C.byref(gf_from_int(a)))
    return gf_int(r)


def gf_trace(a):
    return lib().gf_trace(C.byref(gf_from_int(a)))


def gf_halftrace(a):
    r = Gf()
    lib().gf_halftrace(C.byref(r), C.byref(gf_from_int(a)))
    return gf_int(r)


# ---------------------------------------------------------------- curve helpers
def pt(x=None, y=None):
    p = Pt()
    if x is None:
        p.inf = 1
    else:
        p.x.w[:] = int_to_limbs(x)
        p.y.w[:] = int_to_limbs(y)
        p.inf = 0
    return p


def pt_xy(p):
    return None if p.inf else (gf_int(p.x), gf_int(p.y))


def generator():
    g = Pt()
    lib().k233_generator(C.byref(g))
    return g


def pt_add(a, b):
    r = Pt()
    lib().k233_add(C.byref(r), C.byref(a), C.byref(b))
    return r


def pt_neg(a):
    r = Pt()
    lib().k233_neg(C.byref(r), C.byref(a))
    return r


def pt_mul(a, k):
    """k * a for a non-negative Python int k (little-endian bytes, as xsk233_mul_frob takes)"""
    r = Pt()
    kb = k.to_bytes(32, "little").rstrip(b"\0")
    lib().k233_mul_bytes(C.byref(r), C.byref(a), kb, len(kb))
    return r


def pt_encode(a):
    out = (C.c_uint8 * 30)()
    lib().xsk233_encode(out, C.byref(a))
    return bytes(out)


def pt_decode(b):
    r = Pt()
    ok = lib().xsk233_decode(C.byref(r), (C.c_uint8 * 30).from_buffer_copy(b))
    return r, bool(ok)


def points_to_array(pts):
    """list of Pt -> structured numpy array usable as k233_pt[]"""
    arr = (Pt * len(pts))(*pts)
    return arr


def msm(scalars_mont, pts_arr, nthreads=0):
    """scalars_mont: (n,4) uint64 Montgomery limbs; pts_arr: ctypes Pt array.  multi_scalar_mul, curve.rs:141-158."""
    n = len(pts_arr)
    assert scalars_mont.shape == (n, 4)
    sc = np.ascontiguousarray(scalars_mont, dtype=np.uint64)
    r = Pt()
    lib().k233_msm(C.byref(r), sc.ctypes.data_as(C.c_void_p), C.cast(pts_arr, C.c_void_p), n, nthreads)
    return r


def sum_weighted(scalars_mont):
    """(sum k_i, sum i k_i) mod p as ints, for (n,4) uint64 Montgomery scalars"""
    sc = np.ascontiguousarray(scalars_mont, dtype=np.uint64)
    s0, s1 = Fr(), Fr()
    lib().fr_sum_weighted(sc.ctypes.data_as(C.c_void_p), C.c_size_t(sc.shape[0]), C.byref(s0), C.byref(s1))
    return fr_int(s0), fr_int(s1)


# ---------------------------------------------------------------- bulk fixtures
def chain_points(n, p0, q):
    """ctypes Pt array with out[i] = p0 + i*q"""
    arr = (Pt * n)()
    lib().k233_chain_points(arr, C.c_size_t(n), C.byref(p0), C.byref(q))
    return arr


def mul_batch(p, scalars_mont):
    """out[i] = k_i * p for (n,4) uint64 Montgomery scalars"""
    sc = np.ascontiguousarray(scalars_mont, dtype=np.uint64)
    n = sc.shape[0]
    arr = (Pt * n)()
    lib().k233_mul_batch(arr, C.byref(p), sc.ctypes.data_as(C.c_void_p), C.c_size_t(n))
    return arr


def encode_batch(pts_arr):
    n = len(pts_arr)
    out = np.zeros((n, 30), dtype=np.uint8)
    lib().xsk233_encode_batch(out.ctypes.data_as(C.c_void_p), pts_arr, C.c_size_t(n))
    return out


def decode_batch(enc):
    enc = np.ascontiguousarray(enc, dtype=np.uint8).reshape(-1, 30)
    arr = (Pt * enc.shape[0])()
    lib().xsk233_decode_batch.restype = C.c_long
    bad = lib().xsk233_decode_batch(arr, enc.ctypes.data_as(C.c_void_p), C.c_size_t(enc.shape[0]))
    return arr, bad


def random_fr_mont(n, seed):
    """n uniform Fr elements as (n,4) uint64 Montgomery limbs, from numpy's PCG64 with rejection"""
    rng = np.random.Generator(np.random.PCG64(seed))
    vals = []
    while len(vals) < n:
        raw = rng.integers(0, 1 << 64, size=(n, 4), dtype=np.uint64)
        raw[:, 3] &= np.uint64((1 << 40) - 1)  # 232 bits
        for row in raw:
            v = limbs_to_int(row)
            if v < P:
                vals.append(v)
                if len(vals) == n:
                    break
    return vals, mont_array(vals)


# ---------------------------------------------------------------- ECFFT domain
class Domain:
    """ecfft_domain wrapper: leaves, extend (D -> D'), chain-rule vanishing-polynomial helpers."""

    class _S(C.Structure):
        _fields_ = [("log_n2", C.c_int), ("n2", C.c_size_t), ("leaves", C.c_void_p), ("levels", C.c_int),
                    ("dec", C.c_void_p), ("rec", C.c_void_p), ("x0", C.c_void_p), ("t", C.c_void_p),
                    ("last", C.c_void_p)]

    def __init__(self, log_n2, light=False):
        """light: leaves + isogeny chain only (no extend matrices) -- for the chain-rule helpers and bary_eval at full size"""
        L = lib()
        L.ecfft_domain_new.restype = C.POINTER(Domain._S)
        L.ecfft_domain_new_light.restype = C.POINTER(Domain._S)
        self._d = (L.ecfft_domain_new_light if light else L.ecfft_domain_new)(log_n2)
        self.light = light
        self.log_n2 = log_n2
        self.n2 = 1 << log_n2
        self.n = self.n2 >> 1

    def __del__(self):
        try:
            lib().ecfft_domain_free(self._d)
        except Exception:
            pass

    def _arr(self, ptr, count):
        buf = (C.c_uint64 * (4 * count)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint64).reshape(count, 4).copy()

    def leaves_mont(self):
        return self._arr(self._d.contents.leaves, self.n2)

    def leaves(self):
        return mont_array_to_ints(self.leaves_mont())

    def isogenies(self):
        return (mont_array_to_ints(self._arr(self._d.contents.x0, self.log_n2)),
                mont_array_to_ints(self._arr(self._d.contents.t, self.log_n2)))

    def bary_eval_mont(self, evals_mont, xs_mont):
        """interpolant of evals (on D) at the points xs: O(n) each, ec_fft.rs:455-491"""
        a = np.ascontiguousarray(evals_mont, dtype=np.uint64).reshape(self.n, 4)
        xs = np.ascontiguousarray(xs_mont, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros_like(xs)
        lib().ecfft_bary_eval(self._d, a.ctypes.data_as(C.c_void_p), xs.ctypes.data_as(C.c_void_p),
                              C.c_size_t(xs.shape[0]), out.ctypes.data_as(C.c_void_p))
        return out

    def extend_mont(self, evals_mont):
        assert not self.light
        a = np.ascontiguousarray(evals_mont, dtype=np.uint64).reshape(self.n, 4)
        out = np.zeros_like(a)
        lib().ecfft_extend(self._d, a.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p))
        return out

    def extend(self, evals):
        return mont_array_to_ints(self.extend_mont(mont_array(evals)))

    def vanish_at(self, shift, x):
        r = Fr()
        lib().ecfft_vanish_at(self._d, shift, C.byref(fr_mont(x)), C.byref(r))
        return fr_int(r)

    def vanish_derivative_on_roots_mont(self, shift):
        out = np.zeros((self.n, 4), dtype=np.uint64)
        lib().ecfft_vanish_derivative_on_roots(self._d, shift, out.ctypes.data_as(C.c_void_p))
        return out

    def vanish_on_other_mont(self, shift):
        out = np.zeros((self.n, 4), dtype=np.uint64)
        lib().ecfft_vanish_on_other(self._d, shift, out.ctypes.data_as(C.c_void_p))
        return out


# ---------------------------------------------------------------- R1CS + protocol
class _R1csS(C.Structure):
    _fields_ = [("nrows", C.c_size_t), ("n", C.c_size_t), ("k", C.c_size_t), ("nwires", C.c_size_t),
                ("rowptr", C.c_void_p * 3), ("wire", C.c_void_p * 3), ("coeff", C.c_void_p * 3),
                ("coeffs", C.c_void_p), ("ncoeffs", C.c_size_t)]


class _TrapdoorS(C.Structure):
    _fields_ = [("tau", Fr), ("delta", Fr), ("epsilon", Fr)]


class _SrsS(C.Structure):
    _fields_ = [("g_m", C.c_void_p), ("g_q", C.c_void_p), ("g_k", C.c_void_p), ("z_vals2inv", C.c_void_p),
                ("bar_wts", C.c_void_p)]


class R1CS:
    """Sparse R1CS in the dump's own order (gnark_r1cs.rs:1-20).

    rows: list of (L, R, O), each a list of (wire_id, coeff_id); coeffs: canonical ints."""

    def __init__(self, coeffs, rows, num_public, nwires):
        self.nrows = len(rows)
        self.n = 1
        while self.n < max(self.nrows, 1):
            self.n *= 2
        self.n = max(self.n, 2)
        self.k = num_public
        self.nwires = nwires
        self.coeffs_int = list(coeffs)
        self.coeffs = mont_array(coeffs) if len(coeffs) else np.zeros((1, 4), dtype=np.uint64)
        self.rowptr, self.wire, self.coeff = [], [], []
        for which in range(3):
            rp, wi, ci = [0], [], []
            for row in rows:
                for (w, c) in row[which]:
                    wi.append(w)
                    ci.append(c)
                rp.append(len(wi))
            self.rowptr.append(np.array(rp, dtype=np.uint32))
            self.wire.append(np.array(wi if wi else [0], dtype=np.uint32))
            self.coeff.append(np.array(ci if ci else [0], dtype=np.uint32))
        self._s = _R1csS()
        self._s.nrows, self._s.n, self._s.k, self._s.nwires = self.nrows, self.n, self.k, self.nwires
        for which in range(3):
            self._s.rowptr[which] = self.rowptr[which].ctypes.data
            self._s.wire[which] = self.wire[which].ctypes.data
            self._s.coeff[which] = self.coeff[which].ctypes.data
        self._s.coeffs = self.coeffs.ctypes.data
        self._s.ncoeffs = len(coeffs)

    @classmethod
    def from_arrays(cls, coeffs_mont, rowptr, wire, coeff, nrows, num_public, nwires):
        """Adopt prebuilt CSR arrays (large synthetic circuits)."""
        self = cls.__new__(cls)
        self.nrows, self.k, self.nwires = nrows, num_public, nwires
        self.n = 2
        while self.n < nrows:
            self.n *= 2
        self.coeffs = np.ascontiguousarray(coeffs_mont, dtype=np.uint64)
        self.coeffs_int = None
        self.rowptr = [np.ascontiguousarray(x, dtype=np.uint32) for x in rowptr]
        self.wire = [np.ascontiguousarray(x, dtype=np.uint32) for x in wire]
        self.coeff = [np.ascontiguousarray(x, dtype=np.uint32) for x in coeff]
        self._s = _R1csS()
        self._s.nrows, self._s.n, self._s.k, self._s.nwires = self.nrows, self.n, self.k, self.nwires
        for which in range(3):
            self._s.rowptr[which] = self.rowptr[which].ctypes.data
            self._s.wire[which] = self.wire[which].ctypes.data
            self._s.coeff[which] = self.coeff[which].ctypes.data
        self._s.coeffs = self.coeffs.ctypes.data
        self._s.ncoeffs = self.coeffs.shape[0]
        return self


def trapdoor(tau, delta, epsilon):
    t = _TrapdoorS()
    t.tau, t.delta, t.epsilon = fr_mont(tau), fr_mont(delta), fr_mont(epsilon)
    return t


def r1cs_eval(r1cs, dom, assignment_mont):
    """get_matrix_evaluations_from_witness (proving.rs:348-403): returns (a, b, c, i) Montgomery arrays, bad_row"""
    n = r1cs.n
    w = np.ascontiguousarray(assignment_mont, dtype=np.uint64)
    outs = [np.zeros((n, 4), dtype=np.uint64) for _ in range(4)]
    L = lib()
    L.r1cs_eval.restype = C.c_long
    bad = L.r1cs_eval(C.byref(r1cs._s), dom._d, w.ctypes.data_as(C.c_void_p), *[o.ctypes.data_as(C.c_void_p) for o in outs])
    return outs, bad


class Srs:
    def __init__(self, r1cs, dom, td):
        L = lib()
        L.dv_setup.restype = C.POINTER(_SrsS)
        self._p = L.dv_setup(C.byref(r1cs._s), dom._d, C.byref(td))
        self.n, self.nwires = r1cs.n, r1cs.nwires

    def __del__(self):
        try:
            lib().dv_srs_free(self._p)
        except Exception:
            pass

    def _enc(self, ptr, count):
        arr = C.cast(ptr, C.POINTER(Pt * count)).contents
        return encode_batch(arr)

    def g_m30(self):
        return self._enc(self._p.contents.g_m, self.nwires)

    def g_q30(self):
        return self._enc(self._p.contents.g_q, self.n)

    def g_k30(self):
        return self._enc(self._p.contents.g_k, 4 * self.n)

    def _frs(self, ptr, count):
        buf = (C.c_uint64 * (4 * count)).from_address(ptr)
        return np.frombuffer(buf, dtype=np.uint64).reshape(count, 4).copy()

    def z_vals2inv_mont(self):
        return self._frs(self._p.contents.z_vals2inv, self.n)

    def bar_wts_mont(self):
        return self._frs(self._p.contents.bar_wts, self.n)


def setup_scalars(r1cs, dom, td):
    """Discrete logs of g_m / g_q / g_k (srs.rs:112-167 before the fixed-base multiplications), Montgomery."""
    sc_m = np.zeros((r1cs.nwires, 4), dtype=np.uint64)
    sc_q = np.zeros((r1cs.n, 4), dtype=np.uint64)
    sc_k = np.zeros((4 * r1cs.n, 4), dtype=np.uint64)
    L = lib()
    L.dv_setup_scalars.restype = None
    L.dv_setup_scalars(C.byref(r1cs._s), dom._d, C.byref(td), sc_m.ctypes.data_as(C.c_void_p),
                       sc_q.ctypes.data_as(C.c_void_p), sc_k.ctypes.data_as(C.c_void_p), None, None)
    return sc_m, sc_q, sc_k


def prove(r1cs, dom, srs, assignment_mont, want_stages=False, nthreads=0):
    """Proof::prove (proving.rs:426-688).  Returns (118 proof bytes, status, stages or None)."""
    w = np.ascontiguousarray(assignment_mont, dtype=np.uint64)
    proof = (C.c_uint8 * 118)()
    st = np.zeros((13 * r1cs.n, 4), dtype=np.uint64) if want_stages else None
    L = lib()
    L.dv_prove.restype = C.c_long
    rc = L.dv_prove(C.byref(r1cs._s), dom._d, srs._p, w.ctypes.data_as(C.c_void_p), proof,
                    st.ctypes.data_as(C.c_void_p) if want_stages else None, nthreads)
    return bytes(proof), rc, st


def verify(td, public_ints, proof118):
    """SRS::verify (srs.rs:374-428)"""
    pub = mont_array(public_ints)
    return bool(lib().dv_verify(C.byref(td), pub.ctypes.data_as(C.c_void_p), C.c_size_t(len(public_ints)),
                                (C.c_uint8 * 118).from_buffer_copy(proof118)))


def transcript_alpha(commit_p30, public_ints):
    pub = mont_array(public_ints)
    r = Fr()
    lib().dv_transcript_alpha((C.c_uint8 * 30).from_buffer_copy(commit_p30), pub.ctypes.data_as(C.c_void_p),
                              C.c_size_t(len(public_ints)), C.byref(r))
    return fr_int(r)


def toy_r1cs():
    """The five-constraint circuit of dvsnark_test.rs:34-128; returns (R1CS, public ints, private ints)."""
    ONE, O_, W, Y, Z, X, T, S = range(8)
    c1 = lambda w: (w, 0)
    c2 = lambda w: (w, 1)
    rows = [
        ([c1(X)], [c1(X)], [c1(Y)]),
        ([c1(Y), c1(Z)], [c1(ONE)], [c1(W)]),
        ([c2(Z)], [c1(ONE)], [c1(T)]),
        ([c1(X), c1(T)], [c1(ONE)], [c1(S)]),
        ([c1(W), c1(S)], [c1(ONE)], [c1(O_)]),
    ]
    x, z = 3, 4
    y = x * x
    w = y + z
    t = z + z
    s = x + t
    o = w + s
    return R1CS([1, 2], rows, 2, 8), [o, w], [y, z, x, t, s]
