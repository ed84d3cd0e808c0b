"""Host-side mirror of the reference's prover-facing interface over libdvpari's C ABI.

Names follow /root/reference/src/curve.rs and src/proving.rs (`multi_scalar_mul`, `CurvePoint`,
`Fr`), so a parity test reads like the reference's own test.  Everything here is plumbing: ctypes
calls into libdvpari.so (hand-written sm_100a kernels).  There is no CPU fallback -- without the
built library, or without a CUDA device, construction fails loudly.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DVP_LIB") or os.path.join(_HERE, "libdvpari.so")  # DVP_LIB: development A/B builds

P = 3450873173395281893717377931138512760570940988862252126328087024741343  # src/curve.rs:17
R = 1 << 256
R_INV = pow(R, -1, P)

OK = 0
ERR_NAMES = {
    1: "BAD_ARG", 2: "CUDA", 3: "OOM", 4: "INVALID_POINT", 5: "LENGTH_MISMATCH", 6: "UNSATISFIED",
    7: "ALPHA_IN_DOMAIN", 8: "INTERNAL", 9: "NO_DEVICE", 10: "NCCL", 11: "DOMAIN_MISMATCH",
}


class DvpError(RuntimeError):
    def __init__(self, code, what=""):
        self.code = code
        super().__init__(f"libdvpari: {ERR_NAMES.get(code, code)} {what}".strip())


class MsmStats(C.Structure):
    _fields_ = [("window_bits", C.c_int), ("windows", C.c_int), ("rounds_main", C.c_int), ("rounds_a", C.c_int),
                ("rounds_b", C.c_int), ("launches", C.c_ulonglong), ("ms_recode_sort", C.c_float),
                ("ms_accumulate", C.c_float), ("ms_reduce", C.c_float), ("ms_tail", C.c_float),
                ("ms_pass2_round0", C.c_float), ("adds_round0", C.c_ulonglong), ("ms_device", C.c_float), ("lanes", C.c_int), ("tables", C.c_int), ("ms_tail_host", C.c_float)]


def build(force=False):
    """Compile libdvpari.so in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    srcs = []
    for root, _, files in os.walk(os.path.join(_HERE, "csrc")):
        srcs += [os.path.join(root, f) for f in files if f.endswith((".cu", ".cuh", ".hpp"))]
    srcs.append(os.path.join(_HERE, "..", "include", "dvpari.h"))
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s", "-j4", "libdvpari.so"])
    return LIB_PATH


_lib = None


def lib():
    """Load libdvpari.so.  Raises if it has not been built: the product path has no fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'`")
        L = C.CDLL(LIB_PATH)
        L.dvp_strerror.restype = C.c_char_p
        vp, sz, i32 = C.c_void_p, C.c_size_t, C.c_int
        L.dvp_ctx_create.argtypes = [i32, C.POINTER(vp)]
        L.dvp_ctx_destroy.argtypes = [vp]
        L.dvp_ctx_destroy.restype = None
        L.dvp_ctx_set.argtypes = [vp, C.c_char_p, C.c_long]
        L.dvp_srs_load.argtypes = [vp, i32, vp, sz, C.POINTER(C.c_int64)]
        L.dvp_srs_append.argtypes = [vp, i32, vp, sz, C.POINTER(C.c_int64)]
        L.dvp_srs_size.argtypes = [vp, i32, C.POINTER(sz)]
        L.dvp_srs_random.argtypes = [vp, i32, sz, C.c_uint64]
        L.dvp_srs_mulgen.argtypes = [vp, i32, vp, sz]
        L.dvp_srs_free.argtypes = [vp, i32]
        L.dvp_srs_read.argtypes = [vp, i32, sz, sz, vp]
        L.dvp_msm.argtypes = [vp, i32, sz, vp, sz, vp]
        L.dvp_msm_device.argtypes = [vp, i32, sz, vp, sz, vp]
        L.dvp_msm_adhoc.argtypes = [vp, vp, vp, sz, vp]
        L.dvp_msm_batch.argtypes = [vp, i32, sz, vp, sz, sz, i32, vp]
        L.dvp_msm_sharded_batch.argtypes = [vp, i32, vp, sz, sz, i32, vp]
        L.dvp_msm_last_stats.argtypes = [vp, C.POINTER(MsmStats)]
        L.dvp_msm_last_profile.argtypes = [vp, vp, vp]
        L.dvp_msm_last_timeline.argtypes = [vp, vp, sz, C.POINTER(sz)]
        L.dvp_point_add.argtypes = [vp, vp, vp, vp]
        L.dvp_comm_unique_id.argtypes = [vp]
        L.dvp_comm_init.argtypes = [vp, vp, i32, i32]
        L.dvp_comm_init_local.argtypes = [vp, i32]
        L.dvp_comm_destroy.argtypes = [vp]
        L.dvp_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
        L.dvp_shard_range.argtypes = [sz, i32, i32, C.POINTER(sz), C.POINTER(sz)]
        L.dvp_shard_range.restype = None
        L.dvp_msm_sharded.argtypes = [vp, i32, vp, sz, i32, vp]
        L.dvp_dev_alloc.argtypes = [vp, sz, C.POINTER(vp)]
        L.dvp_dev_free.argtypes = [vp, vp]
        L.dvp_dev_upload.argtypes = [vp, vp, vp, sz]
        L.dvp_dev_download.argtypes = [vp, vp, vp, sz]
        L.dvp_selftest_op.argtypes = [vp, i32, vp, vp, vp, sz]
        L.dvp_microbench.argtypes = [vp, i32, i32, C.POINTER(C.c_double)]
        L.dvp_hostcheck_op.argtypes = [i32, vp, vp, vp, sz]
        L.dvp_latency_probe.argtypes = [vp, i32, i32, C.POINTER(C.c_float)]
        L.dvp_pipebench.argtypes = [vp, i32, i32, i32, C.POINTER(C.c_double)]
        L.dvp_domain_create.argtypes = [vp, C.c_uint, C.POINTER(vp)]
        L.dvp_domain_from_fftree.argtypes = [vp, vp, sz, C.POINTER(vp)]
        L.dvp_domain_destroy.argtypes = [vp]
        L.dvp_domain_destroy.restype = None
        L.dvp_domain_leaves.argtypes = [vp, vp]
        L.dvp_domain_precomputes.argtypes = [vp, vp, vp]
        L.dvp_domain_vanish_at.argtypes = [vp, i32, vp, vp]
        L.dvp_ecfft_extend.argtypes = [vp, vp, vp, i32]
        L.dvp_ecfft_extend_device.argtypes = [vp, vp, i32]
        L.dvp_ecfft_plan_create.argtypes = [vp, C.c_uint, C.POINTER(vp)]
        L.dvp_ecfft_plan_destroy.argtypes = [vp]
        L.dvp_ecfft_plan_destroy.restype = None
        L.dvp_ecfft_enter.argtypes = [vp, vp, vp]
        L.dvp_ecfft_exit.argtypes = [vp, vp, vp]
        L.dvp_r1cs_load.argtypes = [vp, sz, sz, sz, vp, vp, vp, vp, sz, C.POINTER(vp)]
        L.dvp_r1cs_destroy.argtypes = [vp]
        L.dvp_r1cs_destroy.restype = None
        L.dvp_r1cs_eval.argtypes = [vp, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_int64)]
        L.dvp_r1cs_synth_solve.argtypes = [vp, vp, C.c_uint]
        L.dvp_r1cs_eval_time.argtypes = [vp, vp, vp, i32, C.POINTER(C.c_float)]
        L.dvp_setup_scalars.argtypes = [vp, vp, vp, vp, vp, vp]
        L.dvp_setup.argtypes = [vp, vp, vp, i32, i32, i32]
        L.dvp_prover_create.argtypes = [vp, vp, vp, i32, i32, i32, C.POINTER(vp)]
        L.dvp_prover_destroy.argtypes = [vp]
        L.dvp_prover_destroy.restype = None
        L.dvp_prove.argtypes = [vp, vp, sz, vp, sz, vp]
        L.dvp_prove_stages.argtypes = [vp, vp, sz, vp, sz, vp, vp]
        L.dvp_prove_last_times.argtypes = [vp, vp]
        L.dvp_verify.argtypes = [vp, vp, vp, sz, vp, C.POINTER(i32)]
        _lib = L
    return _lib


def _ck(rc, what=""):
    if rc != OK:
        raise DvpError(rc, what)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------------------------- Fr helpers
def fr_to_mont(vals):
    """canonical ints -> (n,4) uint64 Montgomery limbs: the in-memory layout of the reference's Vec<Fr>."""
    out = np.zeros((len(vals), 4), dtype=np.uint64)
    for i, v in enumerate(vals):
        m = (int(v) % P) * R % P
        out[i] = [(m >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
    return out


def fr_from_mont(arr):
    arr = np.asarray(arr, dtype=np.uint64).reshape(-1, 4)
    return [sum(int(row[k]) << (64 * k) for k in range(4)) * R_INV % P for row in arr]


class Context:
    """One CUDA device + resident SRS slots (replaces the per-prove artifact re-reads)."""

    def __init__(self, device=0):
        self._h = C.c_void_p()
        _ck(lib().dvp_ctx_create(device, C.byref(self._h)), "dvp_ctx_create")

    def close(self):
        if self._h:
            lib().dvp_ctx_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set(self, name, value):
        _ck(lib().dvp_ctx_set(self._h, name.encode(), int(value)), name)

    # -- SRS slots ---------------------------------------------------------------------------
    def srs_load(self, slot, pts30, append=False):
        """pts30: bytes or (n,30) uint8 -- the payload of a point-vector file (io_utils.rs:187-239)."""
        a = np.frombuffer(pts30, dtype=np.uint8) if isinstance(pts30, (bytes, bytearray)) else np.ascontiguousarray(pts30, dtype=np.uint8)
        assert a.size % 30 == 0
        bad = C.c_int64(-1)
        fn = lib().dvp_srs_append if append else lib().dvp_srs_load
        rc = fn(self._h, slot, _ptr(a), a.size // 30, C.byref(bad))
        if rc != OK:
            raise DvpError(rc, f"point {bad.value}")

    def srs_append(self, slot, pts30):
        """g_k_0 | g_k_1 | g_k_2 concatenation (proving.rs:666-673)."""
        self.srs_load(slot, pts30, append=True)

    def srs_random(self, slot, n, seed):
        """n uniformly random group elements, deterministic in (seed, index)."""
        _ck(lib().dvp_srs_random(self._h, slot, n, seed), "dvp_srs_random")

    def srs_mulgen(self, slot, scalars_mont):
        """slot[i] = scalars[i] * generator (compute_srs_matrices, srs.rs:126-160), batched on the device."""
        s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
        _ck(lib().dvp_srs_mulgen(self._h, slot, _ptr(s), s.shape[0]), "dvp_srs_mulgen")

    def srs_size(self, slot):
        n = C.c_size_t()
        _ck(lib().dvp_srs_size(self._h, slot, C.byref(n)))
        return n.value

    def srs_read(self, slot, offset, n):
        out = np.zeros((n, 30), dtype=np.uint8)
        _ck(lib().dvp_srs_read(self._h, slot, offset, n, _ptr(out)))
        return out

    def srs_free(self, slot):
        _ck(lib().dvp_srs_free(self._h, slot))

    # -- multi_scalar_mul (curve.rs:141-158) -------------------------------------------------
    def multi_scalar_mul(self, scalars_mont, slot, offset=0):
        """scalars_mont: (n,4) uint64 Montgomery limbs.  Returns the 30-byte CurvePoint::to_bytes of the sum."""
        s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
        out = np.zeros(30, dtype=np.uint8)
        _ck(lib().dvp_msm(self._h, slot, offset, _ptr(s), s.shape[0], _ptr(out)), "dvp_msm")
        return out.tobytes()

    def multi_scalar_mul_device(self, d_scalars, n, slot, offset=0):
        out = np.zeros(30, dtype=np.uint8)
        _ck(lib().dvp_msm_device(self._h, slot, offset, d_scalars, n, _ptr(out)), "dvp_msm_device")
        return out.tobytes()

    @staticmethod
    def _batch_ptrs(vectors, on_device, n):
        """(array of nb pointers, n, keep-alive list) for the batched calls: host vectors (n,4) uint64 each, or device
        pointers with on_device=True and n given."""
        if on_device:
            keep = []
            ptrs = (C.c_void_p * len(vectors))(*[C.c_void_p(int(getattr(v, "value", v))) for v in vectors])
            return ptrs, n, keep
        keep = [np.ascontiguousarray(v, dtype=np.uint64).reshape(-1, 4) for v in vectors]
        if any(k.shape[0] != keep[0].shape[0] for k in keep):
            raise DvpError(5, "multi_scalar_mul_batch: scalar vectors of different lengths")
        ptrs = (C.c_void_p * len(keep))(*[k.ctypes.data for k in keep])
        return ptrs, (keep[0].shape[0] if keep else 0), keep

    def multi_scalar_mul_batch(self, vectors, slot, offset=0, on_device=False, n=None):
        """len(vectors) calls of multi_scalar_mul over the same points, pipelined (dvp_msm_batch): the upload and the
        device side of the next MSM overlap the current one.  Returns the list of 30-byte sums."""
        ptrs, n, keep = self._batch_ptrs(vectors, on_device, n)
        out = np.zeros((max(1, len(vectors)), 30), dtype=np.uint8)
        _ck(lib().dvp_msm_batch(self._h, slot, offset, ptrs, n, len(vectors), 1 if on_device else 0, _ptr(out)),
            "dvp_msm_batch")
        return [out[b].tobytes() for b in range(len(vectors))]

    def multi_scalar_mul_adhoc(self, scalars_mont, pts30):
        s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
        a = np.ascontiguousarray(pts30, dtype=np.uint8).reshape(-1, 30)
        if a.shape[0] != s.shape[0]:
            raise DvpError(5, "multi_scalar_mul: scalars.len() != points.len()")  # curve.rs:142
        out = np.zeros(30, dtype=np.uint8)
        _ck(lib().dvp_msm_adhoc(self._h, _ptr(a), _ptr(s), s.shape[0], _ptr(out)), "dvp_msm_adhoc")
        return out.tobytes()

    # -- multi-GPU (one process per GPU, NCCL) ----------------------------------------------------
    def comm_init(self, unique_id, rank, world):
        """Join the NCCL communicator made from rank 0's comm_unique_id() (shipped by the host's own channel)."""
        buf = np.frombuffer(bytes(unique_id), dtype=np.uint8).copy()
        assert buf.size == 128
        _ck(lib().dvp_comm_init(self._h, _ptr(buf), rank, world), "dvp_comm_init")

    def comm_destroy(self):
        _ck(lib().dvp_comm_destroy(self._h))

    def msm_sharded(self, scalars_mont, slot=0, on_device=False, n=None):
        """multi_scalar_mul over a point vector sharded by contiguous range: the slot holds this rank's points,
        scalars_mont its scalars (or a device pointer with on_device=True and n); same result on every rank."""
        out = np.zeros(30, dtype=np.uint8)
        if on_device:
            _ck(lib().dvp_msm_sharded(self._h, slot, scalars_mont, n, 1, _ptr(out)), "dvp_msm_sharded")
        else:
            s = np.ascontiguousarray(scalars_mont, dtype=np.uint64).reshape(-1, 4)
            _ck(lib().dvp_msm_sharded(self._h, slot, _ptr(s), s.shape[0], 0, _ptr(out)), "dvp_msm_sharded")
        return out.tobytes()

    def msm_sharded_batch(self, vectors, slot=0, on_device=False, n=None):
        """msm_sharded for a batch of scalar vectors: pipelined local MSMs, one all-gather for all partial sums."""
        ptrs, n, keep = self._batch_ptrs(vectors, on_device, n)
        out = np.zeros((max(1, len(vectors)), 30), dtype=np.uint8)
        _ck(lib().dvp_msm_sharded_batch(self._h, slot, ptrs, n, len(vectors), 1 if on_device else 0, _ptr(out)),
            "dvp_msm_sharded_batch")
        return [out[b].tobytes() for b in range(len(vectors))]

    def msm_stats(self):
        st = MsmStats()
        _ck(lib().dvp_msm_last_stats(self._h, C.byref(st)))
        return {f[0]: getattr(st, f[0]) for f in MsmStats._fields_}

    def msm_profile(self):
        """Per-category kernel ms / launch counts of the last MSM run with set("msm_profile", 1)."""
        ms = np.zeros(8, dtype=np.float32)
        cnt = np.zeros(8, dtype=np.uint32)
        _ck(lib().dvp_msm_last_profile(self._h, _ptr(ms), _ptr(cnt)))
        names = ["sort", "plan", "pass1", "binv_up", "binv_direct", "binv_down", "pass2", "misc"]
        return {nm: (float(ms[i]), int(cnt[i])) for i, nm in enumerate(names)}

    def msm_timeline(self):
        """(lane, category name, start ms, end ms) per launch bracket of the last profiled MSM."""
        cnt = C.c_size_t()
        _ck(lib().dvp_msm_last_timeline(self._h, None, 0, C.byref(cnt)))
        rows = np.zeros((max(1, cnt.value), 4), dtype=np.float32)
        _ck(lib().dvp_msm_last_timeline(self._h, _ptr(rows), cnt.value, C.byref(cnt)))
        names = ["sort", "plan", "pass1", "binv_up", "binv_direct", "binv_down", "pass2", "misc"]
        return [(int(r[0]), names[int(r[1])], float(r[2]), float(r[3])) for r in rows[:cnt.value]]

    def point_add(self, a30, b30):
        """CurvePoint::add on encodings (curve.rs:76-82)."""
        a = np.frombuffer(bytes(a30), dtype=np.uint8).copy()
        b = np.frombuffer(bytes(b30), dtype=np.uint8).copy()
        out = np.zeros(30, dtype=np.uint8)
        _ck(lib().dvp_point_add(self._h, _ptr(a), _ptr(b), _ptr(out)), "dvp_point_add")
        return out.tobytes()

    # -- device memory -----------------------------------------------------------------------
    def dev_alloc(self, nbytes):
        p = C.c_void_p()
        _ck(lib().dvp_dev_alloc(self._h, nbytes, C.byref(p)))
        return p

    def dev_free(self, p):
        _ck(lib().dvp_dev_free(self._h, p))

    def dev_upload(self, p, arr):
        arr = np.ascontiguousarray(arr)
        _ck(lib().dvp_dev_upload(self._h, p, _ptr(arr), arr.nbytes))

    def dev_download(self, p, arr):
        _ck(lib().dvp_dev_download(self._h, _ptr(arr), p, arr.nbytes))

    # -- self tests ---------------------------------------------------------------------------
    def selftest_op(self, op, a, b=None):
        a = np.ascontiguousarray(a, dtype=np.uint32)
        out = np.zeros_like(a)
        bb = np.ascontiguousarray(b, dtype=np.uint32) if b is not None else None
        _ck(lib().dvp_selftest_op(self._h, op, _ptr(a), _ptr(bb), _ptr(out), a.shape[0]))
        return out

    def latency_probe(self, mode, iters):
        v = C.c_float()
        _ck(lib().dvp_latency_probe(self._h, mode, iters, C.byref(v)))
        return v.value

    def pipebench(self, mode, iters, blocks_per_sm):
        v = C.c_double()
        _ck(lib().dvp_pipebench(self._h, mode, iters, blocks_per_sm, C.byref(v)))
        return v.value

    def microbench(self, op, iters):
        v = C.c_double()
        _ck(lib().dvp_microbench(self._h, op, iters, C.byref(v)))
        return v.value


def comm_unique_id():
    """128-byte NCCL unique id (rank 0 makes it, every rank passes it to Context.comm_init)."""
    buf = np.zeros(128, dtype=np.uint8)
    _ck(lib().dvp_comm_unique_id(_ptr(buf)), "dvp_comm_unique_id")
    return buf.tobytes()


def comm_init_local(contexts):
    """Join contexts of THIS process (one device or several) as ranks 0..len-1 without NCCL.  Every rank must then be
    driven by its own host thread: the collectives rendezvous (ctypes releases the GIL inside the calls)."""
    arr = (C.c_void_p * len(contexts))(*[c._h for c in contexts])
    _ck(lib().dvp_comm_init_local(C.cast(arr, C.c_void_p), len(contexts)), "dvp_comm_init_local")


def run_ranks(fn, world):
    """fn(rank) on one host thread per rank; returns the results in rank order, re-raises the first exception."""
    import threading

    out, err = [None] * world, [None] * world

    def body(r):
        try:
            out[r] = fn(r)
        except BaseException as e:  # noqa: BLE001 - handed to the caller
            err[r] = e

    ts = [threading.Thread(target=body, args=(r,)) for r in range(world)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    for e in err:
        if e is not None:
            raise e
    return out


def shard_range(total, rank, world):
    """[lo, hi) of `total` items owned by `rank` (the rule the library shards points, rows and scalars by)."""
    lo, hi = C.c_size_t(), C.c_size_t()
    lib().dvp_shard_range(total, rank, world, C.byref(lo), C.byref(hi))
    return lo.value, hi.value


def gk_shard_indices(n, rank, world):
    """Indices into g_k = g_k_0 | g_k_1 | g_k_2 (4n points, proving.rs:666-673) of the points rank `rank` holds:
    g_k_0[ilo, ihi) | g_k_1[ilo, ihi) | g_k_2[2 ilo, 2 ihi) for its range [ilo, ihi) of D."""
    ilo, ihi = shard_range(n, rank, world)
    return np.concatenate([np.arange(ilo, ihi), n + np.arange(ilo, ihi), 2 * n + np.arange(2 * ilo, 2 * ihi)])


def hostcheck_op(op, a, b=None, out_stride=None):
    """Same __host__ __device__ source as the kernels, evaluated on the CPU (tests only)."""
    a = np.ascontiguousarray(a)
    n = a.shape[0]
    if out_stride is None:
        out = np.zeros_like(a)
    else:
        out = np.zeros((n, out_stride), dtype=np.uint8)
    bb = np.ascontiguousarray(b) if b is not None else None
    _ck(lib().dvp_hostcheck_op(op, _ptr(a), _ptr(bb), _ptr(out), n))
    return out


def random_fr_mont(n, seed):
    """n uniform Fr elements directly as (n,4) uint64 Montgomery limbs (vectorised rejection sampling).

    A uniform residue's Montgomery form is again uniform in [0,p), so uniform limbs below p are a
    uniform Vec<Fr>.  Used for synthetic scalars in benchmarks and large tests."""
    rng = np.random.Generator(np.random.PCG64(seed))
    p_limbs = [(P >> (64 * k)) & 0xFFFFFFFFFFFFFFFF for k in range(4)]
    out = np.zeros((0, 4), dtype=np.uint64)
    while out.shape[0] < n:
        raw = rng.integers(0, 1 << 64, size=(n + 1024, 4), dtype=np.uint64)
        raw[:, 3] &= np.uint64((1 << 40) - 1)
        # lexicographic raw < p from the top limb down
        lt = np.zeros(raw.shape[0], dtype=bool)
        eq = np.ones(raw.shape[0], dtype=bool)
        for k in (3, 2, 1, 0):
            lt |= eq & (raw[:, k] < np.uint64(p_limbs[k]))
            eq &= raw[:, k] == np.uint64(p_limbs[k])
        out = np.concatenate([out, raw[lt]])
    return np.ascontiguousarray(out[:n])


class Domain:
    """D / D' of the 2n-leaf ECFFT tree (build_sect_ecfft_tree + get_both_domains, ec_fft.rs:93-239)."""

    def __init__(self, ctx, log2_2n):
        self.ctx = ctx
        self.n2 = 1 << log2_2n
        self.n = self.n2 >> 1
        self._h = C.c_void_p()
        _ck(lib().dvp_domain_create(ctx._h, log2_2n, C.byref(self._h)), "dvp_domain_create")

    @classmethod
    def from_fftree_file(cls, ctx, path):
        """read_minimal_fftree_from_file(cache_dir/tree2n) (tree_io.rs:419-433, proving.rs:436): the size comes from the
        file, the tables are rebuilt on the device, the leaves must equal the file's (else DvpError DOMAIN_MISMATCH)."""
        raw = np.memmap(path, dtype=np.uint8, mode="r")
        self = cls.__new__(cls)
        self.ctx = ctx
        self._h = C.c_void_p()
        _ck(lib().dvp_domain_from_fftree(ctx._h, raw.ctypes.data, raw.size, C.byref(self._h)), "dvp_domain_from_fftree")
        self.n2 = 0
        n2 = C.c_size_t()
        import artifacts

        _ck(artifacts._bind().dvp_fftree_file_leaves(raw.ctypes.data, raw.size, 0, C.byref(n2), None))
        self.n2 = n2.value
        self.n = self.n2 >> 1
        return self

    def close(self):
        if self._h:
            lib().dvp_domain_destroy(self._h)
            self._h = C.c_void_p()

    def leaves(self):
        out = np.zeros((self.n2, 4), dtype=np.uint64)
        _ck(lib().dvp_domain_leaves(self._h, _ptr(out)))
        return out

    def precomputes(self):
        z, w = np.zeros((self.n, 4), dtype=np.uint64), np.zeros((self.n, 4), dtype=np.uint64)
        _ck(lib().dvp_domain_precomputes(self._h, _ptr(z), _ptr(w)))
        return z, w

    def vanish_at(self, shift, x_mont):
        x = np.ascontiguousarray(x_mont, dtype=np.uint64).reshape(4)
        out = np.zeros(4, dtype=np.uint64)
        _ck(lib().dvp_domain_vanish_at(self._h, shift, _ptr(x), _ptr(out)))
        return out

    def extend_device(self, d_data, npoly):
        """In place on device memory: npoly contiguous vectors of n Fr on D -> their values on D'."""
        _ck(lib().dvp_ecfft_extend_device(self._h, d_data, npoly), "dvp_ecfft_extend_device")

    def extend(self, evals_mont):
        """tree2n.extend(evals, Moiety::S1) for one (n,4) or several (p,n,4) vectors (proving.rs:410-422)."""
        a = np.ascontiguousarray(evals_mont, dtype=np.uint64)
        npoly = 1 if a.ndim == 2 else a.shape[0]
        if a.size != npoly * self.n * 4:
            raise DvpError(5, "extend: evals.len() != n")
        out = np.zeros_like(a)
        _ck(lib().dvp_ecfft_extend(self._h, _ptr(a), _ptr(out), npoly), "dvp_ecfft_extend")
        return out


class EcfftPlan:
    """The trees with 4 .. n leaves: FFTree::enter for polynomials of degree < n (ec_fft.rs:317,411)."""

    def __init__(self, ctx, log2_n):
        self.n = 1 << log2_n
        self._h = C.c_void_p()
        _ck(lib().dvp_ecfft_plan_create(ctx._h, log2_n, C.byref(self._h)), "dvp_ecfft_plan_create")

    def close(self):
        if self._h:
            lib().dvp_ecfft_plan_destroy(self._h)
            self._h = C.c_void_p()

    def enter(self, coeffs_mont):
        """coefficients (n,4), low degree first -> values on the n leaves of the n-leaf tree, natural order"""
        a = np.ascontiguousarray(coeffs_mont, dtype=np.uint64).reshape(-1, 4)
        if a.shape[0] != self.n:
            raise DvpError(5, "enter: coeffs.len() != n")
        out = np.zeros_like(a)
        _ck(lib().dvp_ecfft_enter(self._h, _ptr(a), _ptr(out)), "dvp_ecfft_enter")
        return out

    def exit(self, evals_mont):
        """values on the n leaves (natural order) -> coefficients (n,4), low degree first (FFTree::exit)"""
        a = np.ascontiguousarray(evals_mont, dtype=np.uint64).reshape(-1, 4)
        if a.shape[0] != self.n:
            raise DvpError(5, "exit: evals.len() != n")
        out = np.zeros_like(a)
        _ck(lib().dvp_ecfft_exit(self._h, _ptr(a), _ptr(out)), "dvp_ecfft_exit")
        return out


class R1CSInstance:
    """R1CSInstance (gnark_r1cs.rs:261-267) resident on the device, CSR per matrix in dump order."""

    def __init__(self, ctx, nrows, num_public, nwires, rowptr, wire, coeff, coeffs_mont):
        self.ctx = ctx
        self.nrows, self.k, self.nwires = nrows, num_public, nwires
        self.n = 2
        while self.n < nrows:
            self.n *= 2
        self._keep = [[np.ascontiguousarray(x, dtype=np.uint32) for x in grp] for grp in (rowptr, wire, coeff)]
        cm = np.ascontiguousarray(coeffs_mont, dtype=np.uint64).reshape(-1, 4)
        arrs = []
        for grp in self._keep:
            arr = (C.c_void_p * 3)(*[x.ctypes.data for x in grp])
            arrs.append(arr)
        self._h = C.c_void_p()
        _ck(lib().dvp_r1cs_load(ctx._h, nrows, num_public, nwires, arrs[0], arrs[1], arrs[2], _ptr(cm), cm.shape[0],
                                C.byref(self._h)), "dvp_r1cs_load")

    def close(self):
        if self._h:
            lib().dvp_r1cs_destroy(self._h)
            self._h = C.c_void_p()

    def eval_time(self, dom, assignment_mont, reps=5):
        """Milliseconds of one row evaluation with the outputs left on the device (CUDA events, average of reps)."""
        w = np.ascontiguousarray(assignment_mont, dtype=np.uint64).reshape(-1, 4)
        ms = C.c_float()
        _ck(lib().dvp_r1cs_eval_time(self._h, dom._h, _ptr(w), reps, C.byref(ms)), "dvp_r1cs_eval_time")
        return ms.value

    def synth_solve(self, assignment_mont, nlevels):
        """Fill the fresh wires of a synth.py circuit in place (device passes, one per level)."""
        w = np.ascontiguousarray(assignment_mont, dtype=np.uint64).reshape(-1, 4)
        if w.shape[0] != self.nwires:
            raise DvpError(5, "assignment length")
        _ck(lib().dvp_r1cs_synth_solve(self._h, _ptr(w), nlevels), "dvp_r1cs_synth_solve")
        return w

    def eval(self, dom, assignment_mont):
        """get_matrix_evaluations_from_witness (proving.rs:348-403) -> (a, b, c, i); raises on a bad row."""
        w = np.ascontiguousarray(assignment_mont, dtype=np.uint64).reshape(-1, 4)
        if w.shape[0] != self.nwires:
            raise DvpError(5, "assignment length")
        outs = [np.zeros((self.n, 4), dtype=np.uint64) for _ in range(4)]
        bad = C.c_int64(-1)
        rc = lib().dvp_r1cs_eval(self._h, dom._h, _ptr(w), *[_ptr(o) for o in outs], C.byref(bad))
        if rc != OK:
            raise DvpError(rc, f"constraint {bad.value}")
        return outs


def verify(ctx, trapdoor_ints, public_mont, proof118):
    """SRS::verify (srs.rs:374-428) with the trapdoor (tau, delta, epsilon): True iff the proof is accepted."""
    td = fr_to_mont(list(trapdoor_ints))
    pub = np.ascontiguousarray(public_mont, dtype=np.uint64).reshape(-1, 4)
    pr = np.frombuffer(bytes(proof118), dtype=np.uint8).copy()
    if pr.size != 118:
        raise DvpError(1, "proof must be 118 bytes")
    ok = C.c_int(0)
    _ck(lib().dvp_verify(ctx._h, _ptr(td), _ptr(pub), pub.shape[0], _ptr(pr), C.byref(ok)), "dvp_verify")
    return bool(ok.value)


def setup_scalars(r1cs, dom, trapdoor_ints):
    """Discrete logs of g_m / g_q / g_k for the trapdoor (tau, delta, epsilon) (srs.rs:53-167), computed on the device."""
    td = fr_to_mont(list(trapdoor_ints))
    sc_m = np.zeros((r1cs.nwires, 4), dtype=np.uint64)
    sc_q = np.zeros((r1cs.n, 4), dtype=np.uint64)
    sc_k = np.zeros((4 * r1cs.n, 4), dtype=np.uint64)
    _ck(lib().dvp_setup_scalars(r1cs._h, dom._h, _ptr(td), _ptr(sc_m), _ptr(sc_q), _ptr(sc_k)), "dvp_setup_scalars")
    return sc_m, sc_q, sc_k


def setup(r1cs, dom, trapdoor_ints, slot_gm=0, slot_gq=1, slot_gk=2):
    """SRS::verifier_runs_setup (srs.rs:177-361) with the SRS left resident in the slots (this rank's ranges)."""
    td = fr_to_mont(list(trapdoor_ints))
    _ck(lib().dvp_setup(r1cs._h, dom._h, _ptr(td), slot_gm, slot_gq, slot_gk), "dvp_setup")


class Prover:
    """Proof::prove with resident artifacts: SRS slots g_m / g_q / g_k, domain, R1CS (proving.rs:426-688)."""

    def __init__(self, ctx, dom, r1cs, slot_gm=0, slot_gq=1, slot_gk=2):
        self.ctx, self.dom, self.r1cs = ctx, dom, r1cs
        self._h = C.c_void_p()
        _ck(lib().dvp_prover_create(ctx._h, dom._h, r1cs._h, slot_gm, slot_gq, slot_gk, C.byref(self._h)),
            "dvp_prover_create")

    def close(self):
        if self._h:
            lib().dvp_prover_destroy(self._h)
            self._h = C.c_void_p()

    def prove(self, public_mont, private_mont, want_stages=False):
        pub = np.ascontiguousarray(public_mont, dtype=np.uint64).reshape(-1, 4)
        priv = np.ascontiguousarray(private_mont, dtype=np.uint64).reshape(-1, 4)
        proof = np.zeros(118, dtype=np.uint8)
        if want_stages:
            st = np.zeros((13 * self.dom.n, 4), dtype=np.uint64)
            rc = lib().dvp_prove_stages(self._h, _ptr(pub), pub.shape[0], _ptr(priv), priv.shape[0], _ptr(proof), _ptr(st))
        else:
            st = None
            rc = lib().dvp_prove(self._h, _ptr(pub), pub.shape[0], _ptr(priv), priv.shape[0], _ptr(proof))
        _ck(rc, "dvp_prove")
        return (proof.tobytes(), st) if want_stages else proof.tobytes()

    def last_times(self):
        ms = np.zeros(8, dtype=np.float32)
        _ck(lib().dvp_prove_last_times(self._h, _ptr(ms)))
        return dict(zip(["r1cs", "msm_gm", "extend_quotient", "msm_gq", "challenge_kscalars", "msm_gk", "witness_h2d"],
                        ms.tolist()))


def proof_to_bits(proof118):
    """Proof::to_bits (proving.rs:691-718): 240 + 240 + 232 + 232 little-endian bits."""
    bits = []
    for lo, hi, nbits in ((0, 30, 240), (30, 60, 240), (60, 89, 232), (89, 118, 232)):
        chunk = proof118[lo:hi]
        bits += [(chunk[i // 8] >> (i % 8)) & 1 for i in range(nbits)]
    return bits


def proof_from_bits(bits):
    """Proof::from_bits (proving.rs:721-770)."""
    assert len(bits) == 944
    out = bytearray()
    pos = 0
    for nbytes in (30, 30, 29, 29):
        for _ in range(nbytes):
            out.append(sum(bits[pos + i] << i for i in range(8)))
            pos += 8
    return bytes(out)
