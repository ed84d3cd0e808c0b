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
