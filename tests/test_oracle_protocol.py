"""Oracle protocol layer (oracle/blake3.c, oracle/dvsnark.c): transcript, R1CS evaluation, and the
reference's acceptance test -- setup -> prove -> verify on the toy circuit (dvsnark_test.rs:131-180)."""
import ctypes as C
import os
import random

import blake3 as pyblake3
import numpy as np
import pytest

P = 3450873173395281893717377931138512760570940988862252126328087024741343


def _b3(oracle, data):
    out = (C.c_uint8 * 32)()
    assert oracle.lib().blake3_hash_small(data, len(data), out) == 0
    return bytes(out)


def test_blake3_matches_reference_implementation(oracle):
    rnd = random.Random(31)
    import dvpari

    L = dvpari.lib()
    L.dvp_blake3.argtypes = [C.c_char_p, C.c_size_t, C.c_void_p]
    # one block, one chunk, chunk boundaries, every tree shape up to 9 chunks, 29 * 64 bytes (64 public inputs)
    sizes = [0, 1, 29, 30, 58, 63, 64, 65, 127, 128, 129, 1000, 1023, 1024, 1025, 29 * 64, 2047, 2048, 2049, 3072, 3073,
             4096, 4097, 5000, 6144, 7168, 8192, 8193, 9000, 31 * 1024 + 5]
    for n in sizes:
        data = bytes(rnd.getrandbits(8) for _ in range(n))
        want = pyblake3.blake3(data).digest()
        assert _b3(oracle, data) == want, n
        out = (C.c_uint8 * 32)()
        assert L.dvp_blake3(data, n, out) == 0 and bytes(out) == want, n  # the product's own (independent) BLAKE3


def test_public_input_hash_kat(oracle):
    """gnark_r1cs.rs:497-504: sp1_generate_scalar_from_raw_public_input(LE u64 [55,0,0,0,89,0,0,0])"""
    h = _b3(oracle, bytes([55, 0, 0, 0, 89, 0, 0, 0]))
    v = 0
    for idx, byte in enumerate(h):
        v = v * 256 + (0 if idx < 4 else byte)
    assert v % P == 19542051593079647282099705468191403958371264520862632234952945594121


def test_transcript_alpha(oracle):
    """proving.rs:164-197 restated with the Python blake3 package."""
    rnd = random.Random(32)
    commit = bytes(rnd.getrandbits(8) for _ in range(30))
    pub = [rnd.randrange(P), 24]
    e = pyblake3.blake3(b"").digest()
    ct = pyblake3.blake3(e + e).digest()
    hw = pyblake3.blake3(commit).digest()
    hp = pyblake3.blake3(b"".join(x.to_bytes(29, "little") for x in pub)).digest()
    rt = pyblake3.blake3(hw + hp).digest()
    root = bytearray(pyblake3.blake3(ct + rt).digest())
    root[28:] = b"\0\0\0\0"
    assert oracle.transcript_alpha(commit, pub) == int.from_bytes(root, "little")


def test_transcript_alpha_many_public_inputs(oracle):
    """k = 64 public inputs hash 1856 bytes (two chunks): the reference has no limit (proving.rs:149-161); product
    (dvp_transcript_alpha) and oracle against the Python blake3 package."""
    import dvpari

    rnd = random.Random(33)
    L = dvpari.lib()
    L.dvp_transcript_alpha.argtypes = [C.c_char_p, C.c_void_p, C.c_size_t, C.c_void_p]
    for k in (0, 1, 35, 36, 64, 200):
        commit = bytes(rnd.getrandbits(8) for _ in range(30))
        pub = [rnd.randrange(P) for _ in range(k)]
        e = pyblake3.blake3(b"").digest()
        ct = pyblake3.blake3(e + e).digest()
        hp = pyblake3.blake3(b"".join(x.to_bytes(29, "little") for x in pub)).digest()
        rt = pyblake3.blake3(pyblake3.blake3(commit).digest() + hp).digest()
        root = bytearray(pyblake3.blake3(ct + rt).digest())
        root[28:] = b"\0\0\0\0"
        want = int.from_bytes(root, "little")
        assert oracle.transcript_alpha(commit, pub) == want, k
        pm = dvpari.fr_to_mont(pub)
        out = np.zeros(4, dtype=np.uint64)
        assert L.dvp_transcript_alpha(commit, pm.ctypes.data if k else None, k, out.ctypes.data) == 0
        assert dvpari.fr_from_mont(out)[0] == want, k


def test_toy_r1cs_rows(oracle):
    """a, b, c, i of the toy circuit against direct evaluation; c = Cw - i (gnark_r1cs.rs:333-386)."""
    O = oracle
    r1cs, pub, priv = O.toy_r1cs()
    assert r1cs.n == 8  # 5 rows -> next power of two (gnark_r1cs.rs:291)
    dom = O.Domain(4)
    d = dom.leaves()[0::2]
    w = [1] + pub + priv
    (a, b, c, i), bad = O.r1cs_eval(r1cs, dom, O.mont_array(w))
    assert bad == -1
    a, b, c, i = (O.mont_array_to_ints(v) for v in (a, b, c, i))
    x, z = 3, 4
    assert a[:5] == [x, x * x + z, 2 * z, x + 2 * z, (x * x + z) + (x + 2 * z)]
    assert b[:5] == [x, 1, 1, 1, 1]
    for row in range(8):
        ival = (pub[0] + pub[1] * d[row]) % P
        assert i[row] == ival
        assert (a[row] * b[row] - c[row] - ival) % P == 0
    assert a[5:] == [0, 0, 0] and b[5:] == [0, 0, 0]
    # an unsatisfying witness reports the first bad row (reference: assert_eq! panic, proving.rs:389-395)
    w_bad = list(w)
    w_bad[6] += 1  # t
    assert O.r1cs_eval(r1cs, dom, O.mont_array(w_bad))[1] == 2


def test_toy_setup_prove_verify(oracle):
    """dvsnark_test.rs:131-180: Verification should succeed for a valid multi-constraint witness."""
    O = oracle
    r1cs, pub, priv = O.toy_r1cs()
    dom = O.Domain(4)
    rnd = random.Random(43)
    td = O.trapdoor(rnd.randrange(1, P), rnd.randrange(1, P), rnd.randrange(1, P))
    srs = O.Srs(r1cs, dom, td)
    proof, rc, _ = O.prove(r1cs, dom, srs, O.mont_array([1] + pub + priv))
    assert rc == 0 and len(proof) == 118
    assert O.verify(td, pub, proof)
    # every part of the proof and the statement is bound
    for pos in (0, 31, 61, 90):
        bad = bytearray(proof)
        bad[pos] ^= 1
        assert not O.verify(td, pub, bytes(bad))
    assert not O.verify(td, [pub[0] + 1, pub[1]], proof)
    td2 = O.trapdoor(5, 6, 7)
    assert not O.verify(td2, pub, proof)
    # a0, b0 are canonical (FrBits::to_fr rejects >= p, curve.rs:42-59)
    bad = bytearray(proof)
    bad[60:89] = P.to_bytes(29, "little")
    assert not O.verify(td, pub, bytes(bad))


def random_r1cs(O, rnd, nrows, nwires, k, ncoeff=6):
    """Satisfiable random circuit: every row's O side defines a fresh wire."""
    coeffs = [1, P - 1, 2] + [rnd.randrange(P) for _ in range(ncoeff - 3)]
    npriv = nwires - 1 - k
    assert npriv >= nrows
    w = [1] + [rnd.randrange(P) for _ in range(k)] + [0] * npriv
    first_free = 1 + k + (npriv - nrows)
    for j in range(1 + k, first_free):
        w[j] = rnd.randrange(P)
    rows = []
    for r in range(nrows):
        avail = first_free + r  # wires defined so far
        lt = [(rnd.randrange(avail), rnd.randrange(ncoeff)) for _ in range(rnd.randint(1, 4))]
        rt = [(rnd.randrange(avail), rnd.randrange(ncoeff)) for _ in range(rnd.randint(1, 3))]
        a = sum(coeffs[c] * w[x] for x, c in lt) % P
        b = sum(coeffs[c] * w[x] for x, c in rt) % P
        extra = [(rnd.randrange(avail), rnd.randrange(ncoeff)) for _ in range(rnd.randint(0, 2))]
        e = sum(coeffs[c] * w[x] for x, c in extra) % P
        cid = rnd.randrange(ncoeff)
        # coeffs[cid] * w[new] + e = a*b
        w[avail] = (a * b - e) * pow(coeffs[cid], -1, P) % P
        rows.append((lt, rt, extra + [(avail, cid)]))
    return O.R1CS(coeffs, rows, k, nwires), w


@pytest.mark.parametrize("nrows,k", [(13, 2), (50, 3), (64, 1)])
def test_random_circuit_with_padding(oracle, nrows, k):
    O = oracle
    rnd = random.Random(1000 + nrows)
    r1cs, w = random_r1cs(O, rnd, nrows, 1 + k + nrows + 5, k)
    log_n2 = r1cs.n.bit_length()  # 2n leaves
    dom = O.Domain(log_n2)
    td = O.trapdoor(rnd.randrange(1, P), rnd.randrange(1, P), rnd.randrange(1, P))
    srs = O.Srs(r1cs, dom, td)
    proof, rc, _ = O.prove(r1cs, dom, srs, O.mont_array(w))
    assert rc == 0
    assert O.verify(td, w[1:1 + k], proof)
