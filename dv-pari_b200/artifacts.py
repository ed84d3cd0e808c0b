"""The reference's artifact files read straight into the layouts of the C ABI.

Mirrors, by name, the readers of /root/reference/src/io_utils.rs (point / Fr vector files),
/root/reference/src/gnark_r1cs.rs (SP1 sparse-R1CS dump, gnark witness file, SP1 public input) and
/root/reference/src/tree_io.rs (sectioned FFTree files); the writers exist so that tests can round-trip.  Byte-level formats: include/dvpari.h, "Artifact formats"."""
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
    L.dvp_r1cs_dump_parse.argtypes = [vp, sz, sz, sz, C.POINTER(sz * 3), vp, vp, vp, vp]
    L.dvp_fftree_file_sections.argtypes = [vp, sz, sz, vp, vp]
    L.dvp_fftree_file_leaves.argtypes = [vp, sz, sz, C.POINTER(sz), vp]
    L.dvp_fftree_file_matrices.argtypes = [vp, sz, sz, C.c_int, C.POINTER(sz), vp]
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
    _ck(L.dvp_r1cs_dump_parse(_ptr(raw), raw.size, nc.value, nr.value, C.byref(nnz), arr(rowptr), arr(wire), arr(coeff),
                              _ptr(coeffs)), "R1CS dump changed between the two passes")
    # rows.next_power_of_two() (gnark_r1cs.rs:291); the smallest domain the library builds has n = 2 (a circuit of
    # one row pads to two rows instead of the reference's one: extra rows are zero rows, the proof format is unchanged)
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


# ---------------------------------------------------------------- tree_io.rs
FFTR_MAGIC = b"FFTR\0\0\0\0"
# SectionId, tree_io.rs:32-48
(SEC_F_LEAVES, SEC_RECOMBINE, SEC_DECOMPOSE, SEC_RATIONAL_MAPS, SEC_XNN_S, SEC_XNN_S_INV, SEC_Z0_S1, SEC_Z1_S0,
 SEC_Z0_INV_S1, SEC_Z1_INV_S0, SEC_Z0Z0_REM_XNN_S, SEC_Z1Z1_REM_XNN_S, SEC_SUBTREE) = range(13)


def fftree_sections(path, depth=0):
    """{section id: (offset, length)} of the node `depth` subtrees below the root (header walk of tree_io.rs:243-261)."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    off, ln = np.zeros(13, dtype=np.uint64), np.zeros(13, dtype=np.uint64)
    _ck(_bind().dvp_fftree_file_sections(raw.ctypes.data, raw.size, depth, _ptr(off), _ptr(ln)), "not an FFTR file")
    return {i: (int(off[i]), int(ln[i])) for i in range(13) if ln[i]}


def read_minimal_fftree_from_file(path, depth=0, matrices=True):
    """The three sections FFTree::extend needs (read_minimal_fftree_from_file, tree_io.rs:353-433):
    dict(leaves = f.leaves() as (m, 4) Montgomery limbs, recombine / decompose = (count, 2, 2, 4) heap arrays of Mat2x2).
    depth > 0 reads the same sections of a nested subtree (tree.subtree_with_size)."""
    raw = np.memmap(path, dtype=np.uint8, mode="r")
    L = _bind()
    m = C.c_size_t()
    _ck(L.dvp_fftree_file_leaves(raw.ctypes.data, raw.size, depth, C.byref(m), None), "not an FFTR file")
    leaves = np.zeros((m.value, 4), dtype=np.uint64)
    _ck(L.dvp_fftree_file_leaves(raw.ctypes.data, raw.size, depth, C.byref(m), _ptr(leaves)), "FLeaves")
    out = dict(leaves=leaves)
    if matrices:
        for which, name in ((1, "recombine"), (2, "decompose")):
            cnt = C.c_size_t()
            _ck(L.dvp_fftree_file_matrices(raw.ctypes.data, raw.size, depth, which, C.byref(cnt), None), name)
            mats = np.zeros((max(1, cnt.value), 2, 2, 4), dtype=np.uint64)
            _ck(L.dvp_fftree_file_matrices(raw.ctypes.data, raw.size, depth, which, C.byref(cnt), _ptr(mats)), name)
            out[name] = mats[:cnt.value]
    return out


def _ark_vec(fr_mont, per_elem=1):
    """ark-serialize compressed Vec<T>: u64 LE count, then 29-byte Fr (per_elem of them per T)."""
    a = np.ascontiguousarray(fr_mont, dtype=np.uint64).reshape(-1, 4)
    out = np.zeros(29 * a.shape[0], dtype=np.uint8)
    if a.shape[0]:
        _ck(_bind().dvp_fr_to_le29(_ptr(a), a.shape[0], _ptr(out)))
    return struct.pack("<Q", a.shape[0] // per_elem) + out.tobytes()


def _fftree_node(tree):
    blobs = [(SEC_F_LEAVES, _ark_vec(tree["f"])), (SEC_RECOMBINE, _ark_vec(tree["recombine"], 4)),
             (SEC_DECOMPOSE, _ark_vec(tree["decompose"], 4))]
    for sid in range(SEC_RATIONAL_MAPS, SEC_SUBTREE):
        blobs.append((sid, tree.get("raw", {}).get(sid, struct.pack("<Q", 0))))
    if tree.get("subtree") is not None:
        blobs.append((SEC_SUBTREE, _fftree_node(tree["subtree"])))
    head = struct.pack("<II", len(blobs), 0)
    cur = 8 + 24 * len(blobs)
    for sid, b in blobs:
        head += struct.pack("<B7xQQ", sid, cur, len(b))
        cur += len(b)
    return head + b"".join(b for _, b in blobs)


def write_fftree_to_file(path, tree):
    """write_fftree_to_file (tree_io.rs:144-214).  tree = dict(f = (2m, 4) heap array of the layers with the m leaves in
    the upper half, recombine / decompose = (m, 2, 2, 4) heap arrays, optional raw = {section id: blob} for sections
    3..11 (written as empty vectors otherwise), optional subtree = a dict of the same shape)."""
    body = _fftree_node(tree)
    with open(path, "wb") as f:
        f.write(FFTR_MAGIC)
        f.write(struct.pack("<Q", len(body)))
        f.write(body)


# ---------------------------------------------------------------- artifacts.rs: the cache directory
# file names of /root/reference/src/artifacts.rs:18-83
SRS_G_M, SRS_G_Q, SRS_G_K = "g_m", "g_q", ("g_k_0", "g_k_1", "g_k_2")
BAR_WTS, Z_VALS2_INV = "bar_wts", "z_vals2inv"
TREE_2N, TREE_2ND, TREE_N, TREE_ND = "tree2n", "tree2nd", "treen", "treend"
R1CS_CONSTRAINTS_FILE, R1CS_WITNESS_FILE = "r1cs_to_dvsnark", "witness_to_dvsnark"


def write_cache_dir(cache_dir, circ, g_m30, g_q30, g_k30, witness_ints=None, bar_wts=None, z_vals2inv=None):
    """Lay out a cache directory the way SRS::verifier_runs_setup / prover_prepares_precomputes leave it
    (srs.rs:330-361, proving.rs:225-325): point-vector files, the R1CS dump, optionally the witness and precomputes."""
    import os

    os.makedirs(cache_dir, exist_ok=True)
    n = circ["n"]
    write_sparse_r1cs_to_file(os.path.join(cache_dir, R1CS_CONSTRAINTS_FILE), circ)
    write_point_vec_to_file(os.path.join(cache_dir, SRS_G_M), g_m30)
    write_point_vec_to_file(os.path.join(cache_dir, SRS_G_Q), g_q30)
    gk = np.ascontiguousarray(g_k30, dtype=np.uint8).reshape(-1, 30)
    for name, (lo, hi) in zip(SRS_G_K, ((0, n), (n, 2 * n), (2 * n, 4 * n))):
        write_point_vec_to_file(os.path.join(cache_dir, name), gk[lo:hi])
    if witness_ints is not None:
        write_witness_to_file(os.path.join(cache_dir, R1CS_WITNESS_FILE), witness_ints)
    if bar_wts is not None:
        write_fr_vec_to_file(os.path.join(cache_dir, BAR_WTS), bar_wts)
    if z_vals2inv is not None:
        write_fr_vec_to_file(os.path.join(cache_dir, Z_VALS2_INV), z_vals2inv)


def load_prover_from_cache_dir(ctx, cache_dir, num_public, slots=(0, 1, 2)):
    """What Proof::prove(cache_dir, ..) reads on every call (proving.rs:426-470, 511, 666-673), read ONCE: the R1CS dump
    and the SRS point files go to the device, the domain is rebuilt from its constants (a tree2n file, if the cache has
    one, is read for its leaves only and must agree) and, if the cache holds bar_wts / z_vals2inv, they are compared with the device's own.
    Returns (prover, r1cs_instance, domain); prover.prove(public, private) then needs no file."""
    import os

    circ = load_sparse_r1cs_from_file(os.path.join(cache_dir, R1CS_CONSTRAINTS_FILE), num_public)
    inst = dvpari.R1CSInstance(ctx, circ["nrows"], circ["k"], circ["nwires"], circ["rowptr"], circ["wire"], circ["coeff"],
                               circ["coeffs_mont"])
    tree = os.path.join(cache_dir, TREE_2N)
    if os.path.exists(tree):  # proving.rs:436: sized and checked by the file, tables rebuilt on the device
        dom = dvpari.Domain.from_fftree_file(ctx, tree)
        if dom.n != circ["n"]:
            dom.close()
            inst.close()
            raise ValueError("tree2n does not have 2n leaves for this circuit")
    else:
        dom = dvpari.Domain(ctx, circ["n"].bit_length())
    ctx.srs_load(slots[0], read_point_vec_from_file(os.path.join(cache_dir, SRS_G_M)))
    ctx.srs_load(slots[1], read_point_vec_from_file(os.path.join(cache_dir, SRS_G_Q)))
    ctx.srs_load(slots[2], read_point_vec_from_file(os.path.join(cache_dir, SRS_G_K[0])))
    for name in SRS_G_K[1:]:
        ctx.srs_append(slots[2], read_point_vec_from_file(os.path.join(cache_dir, name)))
    z, wts = None, None
    for name in (BAR_WTS, Z_VALS2_INV):
        path = os.path.join(cache_dir, name)
        if os.path.exists(path):
            if z is None:
                z, wts = dom.precomputes()
            ours = wts if name == BAR_WTS else z
            if read_fr_vec_from_file(path).tobytes() != ours.tobytes():
                raise ValueError(f"{name} in the cache directory differs from the device's precompute")
    return dvpari.Prover(ctx, dom, inst, *slots), inst, dom
