"""Artifact formats either side of the path (no GPU): the reference's file layouts
(io_utils.rs:1-7, gnark_r1cs.rs:1-20,58-77) read into the C ABI's layouts, with the reference's own known answer."""
import ctypes as C
import os
import random
import struct

import numpy as np
import pytest

import artifacts
import dvpari
import synth
from guarded import Guarded

P = dvpari.P


def test_sp1_public_input_known_answer():
    """gnark_r1cs.rs:494-504: raw input LE [55,0,0,0,89,0,0,0] must equal witness[2] of the SP1 fibonacci proof."""
    raw = int.from_bytes(bytes([55, 0, 0, 0, 89, 0, 0, 0]), "little")
    got = dvpari.fr_from_mont(artifacts.sp1_generate_scalar_from_raw_public_input(raw))[0]
    assert got == 19542051593079647282099705468191403958371264520862632234952945594121


def test_fr_vector_file_round_trip(tmp_path):
    rnd = random.Random(3)
    vals = [0, 1, P - 1, 1 << 231] + [rnd.randrange(P) for _ in range(200)]
    path = tmp_path / "z_vals2inv"
    artifacts.write_fr_vec_to_file(path, dvpari.fr_to_mont(vals))
    raw = path.read_bytes()
    # u64 LE count, then 29-byte LE canonical elements (io_utils.rs:27-66, FR_UNCOMPRESSED_SIZE = 29)
    assert struct.unpack("<Q", raw[:8])[0] == len(vals) and len(raw) == 8 + 29 * len(vals)
    assert [int.from_bytes(raw[8 + 29 * i:8 + 29 * (i + 1)], "little") for i in range(len(vals))] == vals
    assert dvpari.fr_from_mont(artifacts.read_fr_vec_from_file(path)) == vals
    # a non-canonical element is refused, a truncated file too
    bad = bytearray(raw)
    bad[8:8 + 29] = (P).to_bytes(29, "little")
    (tmp_path / "bad").write_bytes(bytes(bad))
    with pytest.raises(dvpari.DvpError):
        artifacts.read_fr_vec_from_file(tmp_path / "bad")
    (tmp_path / "short").write_bytes(raw[:-5])
    with pytest.raises(ValueError):
        artifacts.read_fr_vec_from_file(tmp_path / "short")


def test_witness_file_reduces_mod_order(tmp_path):
    rnd = random.Random(4)
    vals = [0, 1, P - 1, P, P + 5, (1 << 256) - 1, 1 << 255] + [rnd.getrandbits(256) for _ in range(100)]
    path = tmp_path / "witness"
    artifacts.write_witness_to_file(path, vals)
    got = dvpari.fr_from_mont(artifacts.load_witness_from_file(path))
    assert got == [v % P for v in vals]  # Fr::from_be_bytes_mod_order, gnark_r1cs.rs:201-212


def test_point_vector_file_round_trip(tmp_path, oracle):
    O = oracle
    pts = O.mul_batch(O.generator(), dvpari.random_fr_mont(20, 5))
    enc = O.encode_batch(pts)
    enc[3] = 0  # the neutral element encodes as zeros (io_utils.rs:253-267)
    path = tmp_path / "g_q"
    artifacts.write_point_vec_to_file(path, enc)
    assert path.stat().st_size == 8 + 30 * 20
    assert artifacts.read_point_vec_from_file(path).tobytes() == enc.tobytes()


def test_sparse_r1cs_dump_round_trip(tmp_path, oracle):
    O = oracle
    circ = synth.synth_r1cs(7, seed=11, nlevels=4)
    path = tmp_path / "r1cs"
    artifacts.write_sparse_r1cs_to_file(path, circ)
    raw = path.read_bytes()
    # layout by hand (gnark_r1cs.rs:1-20): u32 nbCoeffs, 32-byte BE coefficients, u32 nbRows, first row's counts
    nc = struct.unpack("<I", raw[:4])[0]
    assert nc == circ["coeffs_mont"].shape[0]
    assert int.from_bytes(raw[4 + 32:4 + 64], "big") == P - 1  # coefficient 1 is -1
    pos = 4 + 32 * nc
    assert struct.unpack("<I", raw[pos:pos + 4])[0] == circ["nrows"]
    nL, nR, nO = struct.unpack("<III", raw[pos + 4:pos + 16])
    assert (nL, nR, nO) == tuple(int(circ["rowptr"][w][1]) for w in range(3))
    assert struct.unpack("<II", raw[pos + 16:pos + 24]) == (int(circ["wire"][0][0]), int(circ["coeff"][0][0]))
    back = artifacts.load_sparse_r1cs_from_file(path, circ["k"])
    assert back["nrows"] == circ["nrows"] and back["n"] == circ["n"] and back["nwires"] == circ["nwires"]
    assert back["coeffs_mont"].tobytes() == circ["coeffs_mont"].tobytes()
    for w in range(3):
        nnz = int(circ["rowptr"][w][-1])
        assert back["rowptr"][w].tobytes() == circ["rowptr"][w].tobytes()
        assert back["wire"][w][:nnz].tobytes() == circ["wire"][w][:nnz].tobytes()
        assert back["coeff"][w][:nnz].tobytes() == circ["coeff"][w][:nnz].tobytes()
    # the oracle evaluates the re-read circuit (pads rows to a power of two, gnark_r1cs.rs:291)
    r1cs = O.R1CS.from_arrays(back["coeffs_mont"], back["rowptr"], back["wire"], back["coeff"], back["nrows"],
                              back["k"], back["nwires"])
    assert r1cs.n == circ["n"]
    # truncated dump
    (tmp_path / "short").write_bytes(raw[:-3])
    with pytest.raises(dvpari.DvpError):
        artifacts.load_sparse_r1cs_from_file(tmp_path / "short", 2)


def test_r1cs_dump_parser_survives_mutations(tmp_path):
    """Truncated or corrupted dumps are refused (or parsed, when the damage leaves a well-formed file): never a read
    outside the buffer.  The reference indexes its mmap with bounds checks and panics (gnark_r1cs.rs:121-185)."""
    circ = synth.synth_r1cs(6, seed=3, nlevels=3)
    path = tmp_path / "r1cs"
    artifacts.write_sparse_r1cs_to_file(path, circ)
    raw = path.read_bytes()
    rnd = random.Random(6)
    guard, L = Guarded(len(raw)), artifacts._bind()
    ok = err = 0
    for trial in range(400):
        data = bytearray(raw)
        kind = trial % 3
        if kind == 0:
            data = data[:rnd.randrange(len(data))]
        elif kind == 1:
            pos = rnd.randrange(0, len(data) - 4) & ~3
            data[pos:pos + 4] = struct.pack("<I", rnd.choice([0, 1, 0xFFFFFFFF, 1 << 31, len(raw), 1 << 20]))
        else:
            for _ in range(6):
                data[rnd.randrange(len(data))] ^= 1 << rnd.randrange(8)
        # the image ends at an inaccessible page: an over-read is a crash, not a pass
        addr = guard.put(data)
        nc, nr, mw = C.c_size_t(), C.c_size_t(), C.c_size_t()
        nnz = (C.c_size_t * 3)()
        if L.dvp_r1cs_dump_sizes(addr, len(data), C.byref(nc), C.byref(nr), C.byref(nnz), C.byref(mw)) != 0:
            err += 1
            continue
        ok += 1
        rowptr = [np.zeros(nr.value + 1, dtype=np.uint32) for _ in range(3)]
        wire = [np.zeros(max(1, nnz[w]), dtype=np.uint32) for w in range(3)]
        coeff = [np.zeros(max(1, nnz[w]), dtype=np.uint32) for w in range(3)]
        coeffs = np.zeros((max(1, nc.value), 4), dtype=np.uint64)
        arr = lambda grp: (C.c_void_p * 3)(*[x.ctypes.data for x in grp])
        assert L.dvp_r1cs_dump_parse(addr, len(data), nc.value, nr.value, C.byref(nnz), arr(rowptr), arr(wire), arr(coeff),
                                     dvpari._ptr(coeffs)) == 0
        assert all(int(rowptr[w][-1]) == nnz[w] for w in range(3))
        # the second pass validates the walk itself: a shorter image, or sizes that are not this image's, are errors
        # (never an out-of-bounds read or write; the arrays above are exactly as large as the first pass said)
        if nnz[0] > 0:
            less = (C.c_size_t * 3)(nnz[0] - 1, nnz[1], nnz[2])
            assert L.dvp_r1cs_dump_parse(addr, len(data), nc.value, nr.value, C.byref(less), arr(rowptr), arr(wire),
                                         arr(coeff), dvpari._ptr(coeffs)) != 0
        if len(data) > 8:
            short = guard.put(data[:len(data) - 5])
            assert L.dvp_r1cs_dump_parse(short, len(data) - 5, nc.value, nr.value, C.byref(nnz), arr(rowptr), arr(wire),
                                         arr(coeff), dvpari._ptr(coeffs)) != 0
    assert ok and err
