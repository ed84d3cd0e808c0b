"""Oracle Fr (oracle/fr.c) against Python big integers and the SURVEY section 4.3 constants."""
import ctypes as C
import random

P = 3450873173395281893717377931138512760570940988862252126328087024741343
R = 1 << 256


def test_constants(oracle):
    O = oracle
    assert O.P == P == 0x8000000000000000000000000000069D5BB915BCD46EFB1AD5F173ABDF
    assert P.bit_length() == 232
    one = O.Fr.in_dll(O.lib(), "FR_ONE")
    assert O.limbs_to_int(one.l) == R % P == 0x7FFFFFFFFFFFFFFFFFFFFFF2C5489471E21037C69EC318337E3373ABDF
    assert (-pow(P, -1, 1 << 64)) % (1 << 64) == 0xA2918B898C382FE1


def test_field_ops(oracle):
    O = oracle
    L = O.lib()
    rnd = random.Random(7)
    edge = [0, 1, 2, P - 1, P - 2, (1 << 231), (1 << 231) - 1]
    vals = edge + [rnd.randrange(P) for _ in range(200)]
    for i in range(0, len(vals) - 1):
        a, b = vals[i], vals[i + 1]
        fa, fb, r = O.fr_mont(a), O.fr_mont(b), O.Fr()
        L.fr_mul(C.byref(r), C.byref(fa), C.byref(fb))
        assert O.fr_int(r) == a * b % P
        assert O.limbs_to_int(r.l) < P
        L.fr_add(C.byref(r), C.byref(fa), C.byref(fb))
        assert O.fr_int(r) == (a + b) % P
        L.fr_sub(C.byref(r), C.byref(fa), C.byref(fb))
        assert O.fr_int(r) == (a - b) % P
        L.fr_inv(C.byref(r), C.byref(fa))
        assert O.fr_int(r) == (pow(a, -1, P) if a else 0)


def test_codecs(oracle):
    O = oracle
    L = O.lib()
    rnd = random.Random(8)
    for _ in range(50):
        a = rnd.randrange(P)
        out = (C.c_uint8 * 29)()
        L.fr_to_le29(out, C.byref(O.fr_mont(a)))
        assert bytes(out) == a.to_bytes(29, "little")
        r = O.Fr()
        assert L.fr_from_le29(C.byref(r), out) == 1 and O.fr_int(r) == a
        big = rnd.getrandbits(256)
        L.fr_from_be32_mod(C.byref(r), (C.c_uint8 * 32).from_buffer_copy(big.to_bytes(32, "big")))
        assert O.fr_int(r) == big % P
    assert L.fr_from_le29(C.byref(O.Fr()), (C.c_uint8 * 29).from_buffer_copy(P.to_bytes(29, "little"))) == 0


def test_batch_inversion_keeps_zeros(oracle):
    O = oracle
    rnd = random.Random(9)
    vals = [rnd.randrange(P) for _ in range(33)]
    vals[5] = 0
    vals[32] = 0
    arr = O.mont_array(vals)
    O.lib().fr_batch_inv(arr.ctypes.data_as(C.c_void_p), C.c_size_t(len(vals)))
    got = O.mont_array_to_ints(arr)
    assert got == [pow(v, -1, P) if v else 0 for v in vals]
