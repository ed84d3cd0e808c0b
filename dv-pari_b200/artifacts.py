"""The reference's artifact files read straight into the layouts of the C ABI.

Mirrors, by name, the readers of /root/reference/src/io_utils.rs (point / Fr vector files) and
/root/reference/src/gnark_r1cs.rs (SP1 sparse-R1CS dump, gnark witness file, SP1 public input); the writers exist so
that tests can round-trip.  Byte-level formats: include/dvpari.h, "Artifact formats"."""
import ctypes as C
import struct

import numpy as np

import dvpari
from dvpari import _ck, _ptr, lib


def _bind():
    L = lib()
    vp, sz = C.c_void_p, C.c_size_t
    L.dvp_fr_from_le29.argtypes = [vp, sz, vp]
    L.dvp_fr_to_le29.argtypes = [vp, sz, vp]
    L.dvp_fr_from_be32_mod_order.argtypes = [vp, sz, vp]
    L.dvp_sp1_public_input.argtypes = [C.c_uint64, vp]
    L.dvp_r1cs_dump_sizes.argtypes = [vp, sz, C.POINTER(sz), C.POINTER(sz), C.POINTER(sz * 3), C.POINTER(sz)]
    L.dvp_r1cs_dump_parse.argtypes = [vp, sz, vp, vp, vp, vp]
    return L


# ---------------------------------------------------------------- io_utils.rs
def read_point_vec_from_file(path):
    """-> (n, 30) uint8 payload for Context.srs_load (read_point_vec_from_file, io_utils.rs:187-239)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 8:
        raise ValueError("File too short for length prefix")
    n = int(raw[:8].view("<u8")[0])
    if raw.size < 8 + 30 * n:
        raise ValueError("File too short for expected point data")
    return raw[8:8 + 30 * n].reshape(n, 30)


def write_point_vec_to_file(path, pts30):
    a = np.ascontiguousarray(pts30, dtype=np.uint8).reshape(-1, 30)
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", a.shape[0]))
        f.write(a.tobytes())


def read_fr_vec_from_file(path):
    """-> (n, 4) uint64 Montgomery limbs = Vec<Fr> (read_fr_vec_from_file, io_utils.rs:113-165)."""
    raw = np.fromfile(path, dtype=np.uint8)
    if raw.size < 8:
        raise ValueError("File too short for length prefix")
    n = int(raw[:8].view("<u8")[0])
    if raw.size < 8 + 29 * n:
        raise ValueError("File too short for expected point data")
    out = np.zeros((n, 4), dtype=np.uint64)
    payload = np.ascontiguousarray(raw[8:8 + 29 * n])
    _ck(_bind().dvp_fr_from_le29(_ptr(payload), n, _ptr(out)), "dvp_fr_from_le29")
    return out


def write_fr_vec_to_file(path, fr_mont):
    a = np.ascontiguousarray(fr_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(29 * a.shape[0], dtype=np.uint8)
    _ck(_bind().dvp_fr_to_le29(_ptr(a), a.shape[0], _ptr(out)))
    with open(path, "wb") as f:
        f.write(struct.pack("<Q", a.shape[0]))
        f.write(out.tobytes())


# ---------------------------------------------------------------- gnark_r1cs.rs
def load_sparse_r1cs_from_file(path, num_public):
    """SP1 dump -> dict(nrows, n, k, nwires, rowptr[3], wire[3], coeff[3], coeffs_mont), the arguments of
    dvpari.R1CSInstance (load_sparse_r1cs_from_file, gnark_r1cs.rs:121-185).  nwires = highest wire id + 1."""
    raw = np.fromfile(path, dtype=np.uint8)
    L = _bind()
    nc, nr, mw = C.c_size_t(), C.c_size_t(), C.c_size_t()
    nnz = (C.c_size_t * 3)()
    _ck(L.dvp_r1cs_dump_sizes(_ptr(raw), raw.size, C.byref(nc), C.byref(nr), C.byref(nnz), C.byref(mw)),
        "malformed R1CS dump")
    rowptr = [np.zeros(nr.value + 1, dtype=np.uint32) for _ in range(3)]
    wire = [np.zeros(max(1, nnz[w]), dtype=np.uint32) for w in range(3)]
    coeff = [np.zeros(max(1, nnz[w]), dtype=np.uint32) for w in range(3)]
    coeffs = np.zeros((max(1, nc.value), 4), dtype=np.uint64)
    arr = lambda grp: (C.c_void_p * 3)(*[x.ctypes.data for x in grp])
    _ck(L.dvp_r1cs_dump_parse(_ptr(raw), raw.size, arr(rowptr), arr(wire), arr(coeff), _ptr(coeffs)))
    n = 2
    while n < nr.value:
        n *= 2
    return dict(nrows=nr.value, n=n, k=num_public, nwires=max(mw.value + 1, 1 + num_public), rowptr=rowptr,
                wire=wire, coeff=coeff, coeffs_mont=coeffs[:nc.value])


def write_sparse_r1cs_to_file(path, circ):
    """Inverse of load_sparse_r1cs_from_file for a circuit in CSR form (tests, synthetic dumps)."""
    coeffs = dvpari.fr_from_mont(circ["coeffs_mont"])
    nrows = circ["nrows"]
    rp = [np.asarray(x, dtype=np.int64) for x in circ["rowptr"]]
    cnt = np.stack([np.diff(r) for r in rp], axis=1).astype("<u4")  # (nrows, 3)
    # per row: 3 counts then the L, R, O terms, each term (wire, coeff)
    terms = [np.stack([np.asarray(circ["wire"][w][:rp[w][-1]], dtype="<u4"),
                       np.asarray(circ["coeff"][w][:rp[w][-1]], dtype="<u4")], axis=1) for w in range(3)]
    row_words = 3 + 2 * cnt.sum(axis=1).astype(np.int64)
    off = np.zeros(nrows + 1, dtype=np.int64)
    np.cumsum(row_words, out=off[1:])
    body = np.zeros(int(off[-1]), dtype="<u4")
    base = off[:-1]
    for w in range(3):
        body[base + w] = cnt[:, w]
    cursor = base + 3
    for w in range(3):
        c = cnt[:, w].astype(np.int64)
        rows = np.repeat(np.arange(nrows), c)
        within = np.arange(int(c.sum())) - np.repeat(rp[w][:-1], c)
        pos = cursor[rows] + 2 * within
        body[pos] = terms[w][:, 0]
        body[pos + 1] = terms[w][:, 1]
        cursor = cursor + 2 * c
    with open(path, "wb") as f:
        f.write(struct.pack("<I", len(coeffs)))
        for v in coeffs:
            f.write(int(v).to_bytes(32, "big"))
        f.write(struct.pack("<I", nrows))
        f.write(body.tobytes())


def load_witness_from_file(path):
    """gnark witness vector -> (n, 4) Montgomery limbs (load_witness_from_file, gnark_r1cs.rs:188-199)."""
    raw = np.fromfile(path, dtype=np.uint8)
    n = int.from_bytes(raw[:4].tobytes(), "big")
    if raw.size < 4 + 32 * n:
        raise ValueError("witness file too short")
    out = np.zeros((n, 4), dtype=np.uint64)
    payload = np.ascontiguousarray(raw[4:4 + 32 * n])
    _ck(_bind().dvp_fr_from_be32_mod_order(_ptr(payload), n, _ptr(out)))
    return out


def write_witness_to_file(path, values):
    with open(path, "wb") as f:
        f.write(len(values).to_bytes(4, "big"))
        for v in values:
            f.write(int(v).to_bytes(32, "big"))


def sp1_generate_scalar_from_raw_public_input(raw_pub_input):
    """gnark_r1cs.rs:218-236 -> (4,) Montgomery limbs."""
    out = np.zeros(4, dtype=np.uint64)
    _ck(_bind().dvp_sp1_public_input(raw_pub_input, _ptr(out)))
    return out
