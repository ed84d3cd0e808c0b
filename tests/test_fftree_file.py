"""FFTree files: the sectioned "FFTR" container of tree_io.rs:1-15 parsed into the C ABI's layouts.

The container layout (magic, node header, section metas, nesting) is restated from tree_io.rs itself; the blobs are
ark-serialize compressed vectors (u64 count | 29-byte Fr), the same element format io_utils.rs:127 pins.  The test trees
are built from the oracle's domain: f = heap array of the isogeny layers, matrices in the form (v0, s0 v0; v1, s1 v1)."""
import ctypes as C
import struct

import numpy as np
import pytest

import artifacts
import dvpari
from guarded import Guarded

P = dvpari.P


def oracle_tree(O, log_n2):
    """FFTree dict for artifacts.write_fftree_to_file: the tree on the 2^log_n2-leaf domain of the oracle."""
    od = O.Domain(log_n2)
    x0, t = od.isogenies()
    layer = od.leaves()
    k = 0
    m = len(layer)
    f = [0] * (2 * m)
    f[m:] = layer
    rec = [[1, 0, 0, 1]] * m
    dec = [[1, 0, 0, 1]] * m
    size = m
    while size > 1:
        half = size // 2
        cur = f[size:2 * size]
        # psi_k(x) = x + t_k / (x - x0_k) pairs leaf j with leaf j + half (oracle/ecfft.h)
        nxt = [(cur[j] + t[k] * pow(cur[j] - x0[k], -1, P)) % P for j in range(half)]
        assert nxt == [(cur[j + half] + t[k] * pow(cur[j + half] - x0[k], -1, P)) % P for j in range(half)]
        f[half:size] = nxt
        e = max(half // 2 - 1, 0)
        for j in range(half):
            s0, s1 = cur[j], cur[j + half]
            v0, v1 = pow(s0 - x0[k], e, P), pow(s1 - x0[k], e, P)
            r = [v0, s0 * v0 % P, v1, s1 * v1 % P]
            det_inv = pow((r[0] * r[3] - r[1] * r[2]) % P, -1, P)
            rec[half + j] = r
            dec[half + j] = [r[3] * det_inv % P, -r[1] * det_inv % P, -r[2] * det_inv % P, r[0] * det_inv % P]
        size = half
        k += 1
    tree = dict(f=dvpari.fr_to_mont(f), recombine=dvpari.fr_to_mont([x for mm in rec for x in mm]),
                decompose=dvpari.fr_to_mont([x for mm in dec for x in mm]))
    return tree, od


def test_fftr_container_round_trip(tmp_path, oracle):
    tree, od = oracle_tree(oracle, 6)
    # the child tree (FFTree.subtree, section 12): half the leaves; only its position in the container matters here
    tree["subtree"] = dict(f=tree["f"][:64], recombine=tree["recombine"][:32 * 4], decompose=tree["decompose"][:32 * 4])
    path = tmp_path / "tree2n"
    artifacts.write_fftree_to_file(path, tree)
    raw = path.read_bytes()
    # layout by hand (tree_io.rs:1-15, 74-118, 144-214)
    assert raw[:8] == b"FFTR\0\0\0\0" and struct.unpack("<Q", raw[8:16])[0] == len(raw) - 16
    count, pad = struct.unpack("<II", raw[16:24])
    assert (count, pad) == (13, 0)
    sid, off, ln = struct.unpack("<B7xQQ", raw[24:48])
    assert (sid, off, ln) == (0, 8 + 24 * 13, 8 + 29 * 128)  # FLeaves first, offsets relative to the node
    assert struct.unpack("<Q", raw[16 + off:16 + off + 8])[0] == 128
    leaves = od.leaves()
    assert int.from_bytes(raw[16 + off + 8 + 29 * 64:16 + off + 8 + 29 * 65], "little") == leaves[0]
    secs = artifacts.fftree_sections(path)
    assert sorted(secs) == list(range(13)) and secs[0] == (16 + off, ln)
    assert secs[1][1] == 8 + 116 * 64 and secs[3][1] == 8  # Mat2x2 = 4 Fr; untouched sections are empty vectors
    got = artifacts.read_minimal_fftree_from_file(path)
    assert got["leaves"].tobytes() == od.leaves_mont().tobytes()
    assert got["recombine"].shape == (64, 2, 2, 4)
    assert got["recombine"].tobytes() == tree["recombine"].tobytes()
    assert got["decompose"].tobytes() == tree["decompose"].tobytes()
    # decompose = recombine^-1 survives the trip (entry 0 of the heap is the identity filler)
    r = dvpari.fr_from_mont(got["recombine"].reshape(-1, 4))
    d = dvpari.fr_from_mont(got["decompose"].reshape(-1, 4))
    for j in (1, 17, 63):
        a, b = r[4 * j:4 * j + 4], d[4 * j:4 * j + 4]
        prod = [(a[0] * b[0] + a[1] * b[2]) % P, (a[0] * b[1] + a[1] * b[3]) % P,
                (a[2] * b[0] + a[3] * b[2]) % P, (a[2] * b[1] + a[3] * b[3]) % P]
        assert prod == [1, 0, 0, 1]
    # nested node
    sub = artifacts.read_minimal_fftree_from_file(path, depth=1)
    assert sub["leaves"].shape == (32, 4) and sub["leaves"].tobytes() == tree["f"][32:64].tobytes()
    assert 12 not in artifacts.fftree_sections(path, depth=1)
    with pytest.raises(dvpari.DvpError):
        artifacts.fftree_sections(path, depth=2)


def test_fftr_malformed_files_are_refused(tmp_path, oracle):
    tree, _ = oracle_tree(oracle, 4)
    path = tmp_path / "tree"
    artifacts.write_fftree_to_file(path, tree)
    raw = bytearray(path.read_bytes())

    def refused(data, name):
        q = tmp_path / name
        q.write_bytes(bytes(data))
        with pytest.raises(dvpari.DvpError) as e:
            artifacts.read_minimal_fftree_from_file(q)
        assert e.value.code == 1

    refused(b"FFTX" + bytes(raw[4:]), "magic")  # "not an FFTR file", tree_io.rs:225
    refused(raw[:-7], "short")  # total length beyond the file
    bad = bytearray(raw)
    bad[24] = 13
    refused(bad, "section")  # "unknown section id", tree_io.rs:69
    bad = bytearray(raw)
    bad[24] = 3
    bad[24 + 24 * 3] = 4
    refused(bad, "noleaves")  # "missing section", tree_io.rs:139
    bad = bytearray(raw)
    struct.pack_into("<Q", bad, 24 + 16, 1 << 40)
    refused(bad, "length")  # a section running past its node
    bad = bytearray(raw)
    off = struct.unpack_from("<Q", raw, 24 + 8)[0]
    bad[16 + off + 8:16 + off + 8 + 29] = P.to_bytes(29, "little")
    q = tmp_path / "noncanonical"
    q.write_bytes(bytes(bad))
    assert artifacts.fftree_sections(q)  # container fine; entry 0 of the heap is not a leaf
    bad[16 + off + 8 + 29 * 16:16 + off + 8 + 29 * 17] = P.to_bytes(29, "little")
    refused(bad, "noncanonical2")  # a leaf >= p


@pytest.mark.gpu
def test_domain_from_fftree_file(tmp_path, oracle):
    """tree2n read for the prover (proving.rs:436): the extend over the file's domain equals the oracle's (the reference's
    test_verify_that_extend_works_over_minimal_tree, tree_io.rs:482-502, uses 64 constraints = 2^7 leaves); a tree with
    other leaves is refused."""
    O = oracle
    ctx = dvpari.Context(0)
    for lg in (4, 7, 11):
        tree, od = oracle_tree(O, lg)
        path = tmp_path / f"tree2n_{lg}"
        artifacts.write_fftree_to_file(path, tree)
        dom = dvpari.Domain.from_fftree_file(ctx, path)
        assert dom.n2 == 1 << lg and dom.leaves().tobytes() == od.leaves_mont().tobytes()
        ev = dvpari.random_fr_mont(dom.n, 40 + lg)
        assert dom.extend(ev).tobytes() == od.extend_mont(ev).tobytes()
        dom.close()
    tree, _ = oracle_tree(O, 6)
    tree["f"][64 + 5, 0] ^= np.uint64(2)
    path = tmp_path / "other"
    artifacts.write_fftree_to_file(path, tree)
    with pytest.raises(dvpari.DvpError) as e:
        dvpari.Domain.from_fftree_file(ctx, path)
    assert e.value.code == 11  # DVP_ERR_DOMAIN_MISMATCH
    (tmp_path / "junk").write_bytes(b"FFTR\0\0\0\0" + bytes(40))
    with pytest.raises(dvpari.DvpError) as e:
        dvpari.Domain.from_fftree_file(ctx, tmp_path / "junk")
    assert e.value.code == 1
    ctx.close()


def test_fftr_parser_survives_mutations(tmp_path, oracle):
    """Truncations and random byte flips of a valid file: the parser answers with a leaf set or an error, never reads
    outside the image (the reference slices with bounds checks and panics, tree_io.rs:254-259)."""
    import random

    tree, od = oracle_tree(oracle, 5)
    tree["subtree"] = dict(f=tree["f"][:32], recombine=tree["recombine"][:16 * 4], decompose=tree["decompose"][:16 * 4])
    path = tmp_path / "tree"
    artifacts.write_fftree_to_file(path, tree)
    raw = path.read_bytes()
    rnd = random.Random(5)
    guard = Guarded(len(raw))
    outcomes = {"ok": 0, "err": 0}
    head = 16 + 8 + 24 * 13  # magic, length, node header, section table
    for trial in range(400):
        data = bytearray(raw)
        kind = trial % 4
        if kind == 0:
            data = data[:rnd.randrange(len(data))]
        elif kind == 1:
            for _ in range(rnd.randrange(1, 4)):
                data[rnd.randrange(head)] = rnd.randrange(256)  # header / section table
        elif kind == 2:
            pos = rnd.randrange(8, head - 8)
            data[pos:pos + 8] = struct.pack("<Q", rnd.choice([0, 1, 7, len(raw), len(raw) + 1, 1 << 31, 1 << 63, (1 << 64) - 1]))
        else:
            for _ in range(8):
                data[rnd.randrange(len(data))] ^= 1 << rnd.randrange(8)
        # the image ends at an inaccessible page: an over-read is a crash, not a pass
        addr, L = guard.put(data), artifacts._bind()
        for depth in (0, 1):
            m = C.c_size_t()
            rc = L.dvp_fftree_file_leaves(addr, len(data), depth, C.byref(m), None)
            if rc == 0:
                leaves = np.zeros((m.value, 4), dtype=np.uint64)
                rc = L.dvp_fftree_file_leaves(addr, len(data), depth, C.byref(m), dvpari._ptr(leaves))
            for which in (1, 2):
                cnt = C.c_size_t()
                if L.dvp_fftree_file_matrices(addr, len(data), depth, which, C.byref(cnt), None) == 0:
                    mats = np.zeros((max(1, cnt.value), 16), dtype=np.uint64)
                    L.dvp_fftree_file_matrices(addr, len(data), depth, which, C.byref(cnt), dvpari._ptr(mats))
            outcomes["ok" if rc == 0 else "err"] += 1
    assert outcomes["ok"] and outcomes["err"]
