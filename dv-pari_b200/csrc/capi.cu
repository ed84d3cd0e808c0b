// C ABI of libdvpari (include/dvpari.h): contexts, SRS slots, MSM entry points, self-tests.
#include <cstdio>
#include <cstring>
#include <vector>
#include "../../include/dvpari.h"
#include "ctx.cuh"
#include "fr.cuh"
#include "host_gf.hpp"
#include "k233_codec.cuh"

using namespace dvp;

#define CKC(x)                                                                                           \
    do {                                                                                                 \
        cudaError_t e_ = (x);                                                                            \
        if (e_ != cudaSuccess) {                                                                         \
            fprintf(stderr, "[dvpari] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DVP_ERR_CUDA;                                                                         \
        }                                                                                                \
    } while (0)

static inline uint32_t cdivu(size_t a, size_t b) { return (uint32_t)((a + b - 1) / b); }

// ---------------------------------------------------------------------------------------- kernels
__global__ void k_decode30(const uint8_t *__restrict__ in, size_t n, AffPt *__restrict__ out,
                           unsigned long long *__restrict__ first_bad) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[30];
    for (int k = 0; k < 30; k++) b[k] = in[i * 30 + k];
    AffPt p;
    if (!xsk233_decode_pt(b, p)) atomicMin(first_bad, (unsigned long long)i);
    pt_store(&out[i], p);
}
__global__ void k_encode30(const AffPt *__restrict__ in, size_t n, uint8_t *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[30];
    xsk233_encode_pt(b, pt_load(&in[i]));
    for (int k = 0; k < 30; k++) out[i * 30 + k] = b[k];
}
// out = a (+) b on encodings; ok[0] = both decoded
__global__ void k_point_add30(const uint8_t *__restrict__ ab, uint8_t *__restrict__ out, int *__restrict__ ok) {
    uint8_t a[30], b[30], r[30];
    for (int k = 0; k < 30; k++) {
        a[k] = ab[k];
        b[k] = ab[30 + k];
    }
    AffPt p, q;
    const bool oa = xsk233_decode_pt(a, p), ob = xsk233_decode_pt(b, q);
    xsk233_encode_pt(r, pt_add_slow(p, q));
    for (int k = 0; k < 30; k++) out[k] = r[k];
    ok[0] = oa && ob;
}

// Synthetic SRS: uniformly random group elements by rejection sampling of the 233-bit encoding
// (one w in four names an element of the group).  Deterministic in (seed, index).
__device__ __forceinline__ uint64_t splitmix64(uint64_t &x) {
    uint64_t z = (x += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}
__global__ void k_random_points(uint64_t seed, size_t n, AffPt *__restrict__ out) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t st = seed ^ (0xd1b54a32d192ed03ull * (i + 1));
    AffPt p;
    for (;;) {
        uint8_t b[32];
        for (int k = 0; k < 4; k++) {
            const uint64_t v = splitmix64(st);
            for (int j = 0; j < 8; j++) b[8 * k + j] = (uint8_t)(v >> (8 * j));
        }
        b[29] &= 0x01; // 233 bits
        if (xsk233_decode_pt(b, p) && !pt_is_inf(p)) break;
    }
    pt_store(&out[i], p);
}

__global__ void k_selftest(int op, const uint32_t *__restrict__ a, const uint32_t *__restrict__ b,
                           uint32_t *__restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (op <= 2) {
        gf x, y, r;
        for (int k = 0; k < 8; k++) {
            x.v[k] = a[i * 8 + k];
            y.v[k] = b ? b[i * 8 + k] : 0;
        }
        r = op == 0 ? gf_mul(x, y) : op == 1 ? gf_sqr(x) : gf_inv(x);
        for (int k = 0; k < 8; k++) out[i * 8 + k] = r.v[k];
    } else if (op <= 4) {
        fr x, y;
        for (int k = 0; k < 8; k++) {
            x.v[k] = a[i * 8 + k];
            y.v[k] = b ? b[i * 8 + k] : 0;
        }
        if (op == 3) {
            fr r = fr_mul(x, y);
            for (int k = 0; k < 8; k++) out[i * 8 + k] = r.v[k];
        } else {
            uint32_t c[8];
            fr_to_canonical(c, x);
            for (int k = 0; k < 8; k++) out[i * 8 + k] = c[k];
        }
    } else {
        AffPt p, q;
        for (int k = 0; k < 8; k++) {
            p.x.v[k] = a[i * 16 + k];
            p.y.v[k] = a[i * 16 + 8 + k];
            q.x.v[k] = b[i * 16 + k];
            q.y.v[k] = b[i * 16 + 8 + k];
        }
        AffPt r = pt_add_slow(p, q);
        for (int k = 0; k < 8; k++) {
            out[i * 16 + k] = r.x.v[k];
            out[i * 16 + 8 + k] = r.y.v[k];
        }
    }
}

// dependent chains of the field primitives, 2 independent chains per thread.
// MINB = minimum resident blocks per SM (caps the register budget: 1 -> 255, 2 -> 128, 3 -> 80).
template <int OP, int MINB>
__global__ void __launch_bounds__(256, MINB) k_microbench(int iters, uint32_t *__restrict__ sink) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t s = 0;
    if (OP < 2) {
        gf a, b;
        for (int k = 0; k < 8; k++) {
            a.v[k] = t * 2654435761u + k * 40503u + 1;
            b.v[k] = t * 2246822519u + k * 9176u + 7;
        }
        a.v[7] &= 0x1ff;
        b.v[7] &= 0x1ff;
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            if (OP == 0) {
                a = gf_mul(a, b);
                b = gf_mul(b, a);
            } else {
                a = gf_sqr(a);
                b = gf_sqr(b);
            }
        }
        for (int k = 0; k < 8; k++) s ^= a.v[k] ^ b.v[k];
    } else {
        fr a, b;
        for (int k = 0; k < 8; k++) {
            a.v[k] = t * 2654435761u + k * 40503u + 1;
            b.v[k] = t * 2246822519u + k * 9176u + 7;
        }
        a.v[7] &= 0x7f;
        b.v[7] &= 0x7f;
#pragma unroll 1
        for (int i = 0; i < iters; i++) {
            a = fr_mul(a, b);
            b = fr_mul(b, a);
        }
        for (int k = 0; k < 8; k++) s ^= a.v[k] ^ b.v[k];
    }
    if (s == 0x12345678u) sink[0] = s;
}

template <int MINB> static void launch_microbench(int op, int blocks, cudaStream_t st, int iters, uint32_t *sink) {
    if (op == 0) k_microbench<0, MINB><<<blocks, 256, 0, st>>>(iters, sink);
    else if (op == 1) k_microbench<1, MINB><<<blocks, 256, 0, st>>>(iters, sink);
    else k_microbench<2, MINB><<<blocks, 256, 0, st>>>(iters, sink);
}

// ---------------------------------------------------------------------------------------- context
int ctx_decode_into(dvp_ctx *ctx, const uint8_t *pts30, size_t n, AffPt *d_out, int64_t *first_invalid) {
    int rc;
    if ((rc = ctx->bytes.reserve(n * 30 + 16)) != 0) return rc;
    if ((rc = ctx->small.reserve(64)) != 0) return rc;
    unsigned long long init = ~0ull;
    CKC(cudaMemcpyAsync(ctx->bytes.p, pts30, n * 30, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaMemcpyAsync(ctx->small.p, &init, 8, cudaMemcpyHostToDevice, ctx->stream));
    k_decode30<<<cdivu(n, 128), 128, 0, ctx->stream>>>((const uint8_t *)ctx->bytes.p, n, d_out,
                                                      (unsigned long long *)ctx->small.p);
    CKC(cudaGetLastError());
    unsigned long long bad = 0;
    CKC(cudaMemcpyAsync(&bad, ctx->small.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    if (bad != ~0ull) {
        if (first_invalid) *first_invalid = (int64_t)bad;
        return DVP_ERR_INVALID_POINT;
    }
    return DVP_OK;
}

extern "C" {

const char *dvp_strerror(int code) {
    switch (code) {
    case DVP_OK: return "ok";
    case DVP_ERR_BAD_ARG: return "bad argument";
    case DVP_ERR_CUDA: return "CUDA error";
    case DVP_ERR_OOM: return "out of device memory";
    case DVP_ERR_INVALID_POINT: return "invalid point encoding";
    case DVP_ERR_LENGTH_MISMATCH: return "scalar/point length mismatch";
    case DVP_ERR_UNSATISFIED: return "R1CS row not satisfied";
    case DVP_ERR_ALPHA_IN_DOMAIN: return "challenge lies in the evaluation domain";
    case DVP_ERR_INTERNAL: return "internal error";
    case DVP_ERR_NO_DEVICE: return "no CUDA device";
    case DVP_ERR_NCCL: return "NCCL error";
    case DVP_ERR_DOMAIN_MISMATCH: return "FFTree file is not the tree of the reference's domain constants";
    }
    return "unknown error";
}
int dvp_abi_version(void) { return 1; }

int dvp_ctx_create(int device, dvp_ctx **out) {
    if (!out) return DVP_ERR_BAD_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return DVP_ERR_NO_DEVICE;
    if (device < 0 || device >= ndev) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(device));
    dvp_ctx *c = new dvp_ctx();
    c->device = device;
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return DVP_ERR_CUDA;
    }
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (cudaStreamCreateWithPriority(&c->aux_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&c->ev_aux, cudaEventDisableTiming) != cudaSuccess) {
        dvp_ctx_destroy(c);
        return DVP_ERR_CUDA;
    }
    int rc = c->msm.init(c->stream);
    if (rc) {
        dvp_ctx_destroy(c);
        return rc;
    }
    *out = c;
    return DVP_OK;
}
void dvp_ctx_destroy(dvp_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    dvp_comm_destroy(ctx);
    ctx->msm.destroy();
    ctx->commbuf.release();
    for (auto &s : ctx->slots) {
        s.buf.release();
        s.table.release();
    }
    ctx->bytes.release();
    ctx->small.release();
    ctx->scal.release();
    ctx->scal2.release();
    ctx->adhoc.release();
    for (int k = 0; k < 2; k++) {
        if (ctx->ev_up[k]) cudaEventDestroy(ctx->ev_up[k]);
        if (ctx->ev_free[k]) cudaEventDestroy(ctx->ev_free[k]);
    }
    if (ctx->ev_batch) cudaEventDestroy(ctx->ev_batch);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->ev_aux) cudaEventDestroy(ctx->ev_aux);
    if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}
int dvp_ctx_set(dvp_ctx *ctx, const char *name, long value) {
    if (!ctx || !name) return DVP_ERR_BAD_ARG;
    if (!strcmp(name, "msm_window_bits")) {
        if (value != 0 && (value < 4 || value > 20)) return DVP_ERR_BAD_ARG;
        ctx->msm.force_window_bits = (int)value;
        return DVP_OK;
    }
    if (!strcmp(name, "msm_lanes")) {
        if (value < 0 || value > 16) return DVP_ERR_BAD_ARG;
        ctx->msm.force_lanes = (int)value;
        return DVP_OK;
    }
    if (!strcmp(name, "msm_tables")) {
        ctx->msm_tables = value != 0;
        return DVP_OK;
    }
    if (!strcmp(name, "msm_table_windows")) {
        if (value != 0 && (value < 9 || value > 59)) return DVP_ERR_BAD_ARG;
        ctx->msm_table_windows = (int)value;
        for (auto &sl : ctx->slots) sl.invalidate();
        return DVP_OK;
    }
    if (!strcmp(name, "msm_tables_min")) {
        if (value < 1) return DVP_ERR_BAD_ARG;
        ctx->msm_tables_min = (size_t)value;
        return DVP_OK;
    }
    if (!strcmp(name, "prio_split")) {
        ctx->msm.prio_split = value != 0;
        return DVP_OK;
    }
    if (!strcmp(name, "pass_b_max")) {
        if (value != 1 && value != 4 && value != 16 && value != 64) return DVP_ERR_BAD_ARG;
        ctx->msm.pass_b_max = (int)value;
        return DVP_OK;
    }
    if (!strcmp(name, "ld_tree_max")) {
        if (value != 0 && value < 64) return DVP_ERR_BAD_ARG;
        ctx->msm.ld_tree_max = (size_t)value;
        return DVP_OK;
    }
    if (!strcmp(name, "b16_min")) {
        ctx->msm.b16_min = value > 0 ? (size_t)value : ((size_t)1 << 21);
        return DVP_OK;
    }
    if (!strcmp(name, "b64_min")) {
        ctx->msm.b64_min = value > 0 ? (size_t)value : ((size_t)1 << 23);
        return DVP_OK;
    }
    if (!strcmp(name, "use_accumulate")) {
        if (value < 0 || value > 2) return DVP_ERR_BAD_ARG;
        ctx->msm.use_accumulate = (int)value;
        return DVP_OK;
    }
    if (!strcmp(name, "binv_coop_warps")) {
        if (value < 1) return DVP_ERR_BAD_ARG;
        ctx->msm.binv_coop_warps = (uint32_t)value;
        return DVP_OK;
    }
    if (!strcmp(name, "round_warp_max")) {
        if (value < 0) return DVP_ERR_BAD_ARG;
        ctx->msm.round_warp_max = (size_t)value;
        return DVP_OK;
    }
    if (!strcmp(name, "binv_direct")) {
        if (value < 64 || value > (1 << 20)) return DVP_ERR_BAD_ARG;
        ctx->msm.binv_direct = (uint32_t)value;
        return DVP_OK;
    }
    if (!strcmp(name, "msm_profile")) {
        ctx->msm.profile = value != 0;
        return DVP_OK;
    }
    if (!strcmp(name, "fused_rounds")) {
        ctx->msm.fused_rounds = value ? 1 : 0;
        return DVP_OK;
    }
    if (!strcmp(name, "ld_tree_warp_a")) {
        ctx->msm.ld_tree_warp_a = value ? 1 : 0;
        return DVP_OK;
    }
    if (!strcmp(name, "pass2_minb")) {
        if (value < 1 || value > 3) return DVP_ERR_BAD_ARG;
        ctx->msm.pass2_minb = (int)value;
        return DVP_OK;
    }
    if (!strcmp(name, "timing")) {
        ctx->msm.timing = value != 0;
        return DVP_OK;
    }
    if (!strcmp(name, "prove_joint")) { // -1 automatic, 0 never, 1 always: commit_p as one MSM over g_m | g_q
        if (value < -1 || value > 1) return DVP_ERR_BAD_ARG;
        ctx->prove_joint = (int)value;
        return DVP_OK;
    }
    if (!strcmp(name, "msm_preplan")) { // persistent path: plan all rounds before the launch (default 1)
        ctx->msm.preplan = value != 0;
        return DVP_OK;
    }
    if (!strcmp(name, "msm_sort_ahead")) { // batches: sort MSM b+1 on a side stream while MSM b runs (default 1)
        ctx->msm.sort_ahead = value != 0;
        return DVP_OK;
    }
    return DVP_ERR_BAD_ARG;
}

static int slot_ok(dvp_ctx *ctx, int slot) { return ctx && slot >= 0 && slot < DVP_MAX_SRS_SLOTS; }

extern "C++" {
// Where the points of slot[offset, offset + n) are for an MSM.  Large slots get W tables T[j] = 2^(j c) P once, on
// first use (W x the slot's memory), after which all windows share one bucket set; small slots, small sub-ranges and
// tight memory use the plain vector.
static int slot_points(dvp_ctx *ctx, SrsSlot &s, size_t offset, size_t n, const AffPt **pts, MsmTable *tab, bool *use_tab) {
    const bool want = ctx->msm_tables && s.n >= ctx->msm_tables_min && n >= s.n / 2 && !ctx->msm.force_window_bits;
    if (want && !s.table_ok && !s.table_failed) {
        const int W = ctx->msm_table_windows ? ctx->msm_table_windows : choose_table_windows(s.n);
        const size_t bytes = (size_t)W * s.n * sizeof(AffPt);
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        // keep room for the MSM scratch (about as large again) and the caller's own vectors
        if ((size_t)W * s.n < (1ull << 31) && bytes + bytes / 2 + (4ull << 30) < free_b && s.table.reserve(bytes) == 0) {
            int rc = ctx->msm.build_table(s.buf.as<AffPt>(), s.n, W, s.table.as<AffPt>());
            if (rc) {
                s.table.release();
                return rc;
            }
            s.tab.W = W;
            s.tab.stride = s.n;
            s.table_ok = true;
        } else {
            s.table.release();
            s.table_failed = true;
        }
    }
    *use_tab = want && s.table_ok;
    if (*use_tab) {
        *tab = s.tab;
        tab->offset = offset;
        *pts = s.table.as<AffPt>();
    } else {
        *pts = s.buf.as<AffPt>() + offset;
    }
    return 0;
}

// MSM over slot[offset, offset + n).
int slot_msm(dvp_ctx *ctx, int slot, size_t offset, const uint32_t *d_scalars, size_t n, AffPt *out, cudaStream_t on) {
    return slot_msm_at(ctx, ctx->slots[slot], offset, d_scalars, n, out, on);
}
int slot_msm_at(dvp_ctx *ctx, SrsSlot &s, size_t offset, const uint32_t *d_scalars, size_t n, AffPt *out, cudaStream_t on) {
    struct StreamSwap { // the engine's main stream for this call
        MsmEngine &e;
        cudaStream_t saved;
        StreamSwap(MsmEngine &eng, cudaStream_t s_) : e(eng), saved(eng.stream) {
            if (s_) e.stream = s_;
        }
        ~StreamSwap() { e.stream = saved; }
    } swap(ctx->msm, on);
    const AffPt *pts = nullptr;
    MsmTable t;
    bool use_tab = false;
    int rc = slot_points(ctx, s, offset, n, &pts, &t, &use_tab);
    if (rc) return rc;
    return ctx->msm.run(pts, d_scalars, n, out, use_tab ? &t : nullptr);
}

int slot_msm_batch(dvp_ctx *ctx, int slot, size_t offset, const uint64_t *const *scalars, size_t n, size_t nb,
                   bool on_device, AffPt *out) {
    for (size_t b = 0; b < nb; b++) out[b] = pt_inf();
    if (n == 0 || nb == 0) return DVP_OK;
    const AffPt *pts = nullptr;
    MsmTable t;
    bool use_tab = false;
    int rc = slot_points(ctx, ctx->slots[slot], offset, n, &pts, &t, &use_tab);
    if (rc) return rc;
    DevBuf *stage[2] = {&ctx->scal, &ctx->scal2};
    if (!on_device)
        for (int k = 0; k < (nb > 1 ? 2 : 1); k++)
            if ((rc = stage[k]->reserve(n * 32 + 32)) != 0) return rc;
    if (!ctx->copy_stream) {
        CKC(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
        for (int k = 0; k < 2; k++) {
            CKC(cudaEventCreateWithFlags(&ctx->ev_up[k], cudaEventDisableTiming));
            CKC(cudaEventCreateWithFlags(&ctx->ev_free[k], cudaEventDisableTiming));
        }
        CKC(cudaEventCreateWithFlags(&ctx->ev_batch, cudaEventDisableTiming));
    }
    MsmEngine &E = ctx->msm;
    // With the engine's per-stage timers on, the MSMs run one after the other (the timers are shared).
    const bool pipelined = !E.timing && !E.profile;
    MsmEngine::Pending pend[2];
    size_t pend_b[2] = {0, 0};
    auto upload = [&](size_t b) -> int {
        const int k = (int)(b & 1);
        if (b >= 2) CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_free[k], 0)); // MSM b-2 has read stage[k]
        CKC(cudaMemcpyAsync(stage[k]->p, scalars[b], n * 32, cudaMemcpyHostToDevice, ctx->copy_stream));
        CKC(cudaEventRecord(ctx->ev_up[k], ctx->copy_stream));
        return 0;
    };
    int first_rc = 0;
    // everything enqueued so far on the context stream comes first (device vectors may have been written on it, an
    // earlier call may still read the staging buffers); after that the scalars of MSM b are valid once ev_up fires
    // (host vectors) or at once (device vectors), whatever else the batch has queued on the context stream -- which
    // is what lets the engine sort MSM b+1 while MSM b is still running
    CKC(cudaEventRecord(ctx->ev_batch, ctx->stream));
    if (!on_device) {
        CKC(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_batch, 0));
        if ((rc = upload(0)) != 0) return rc;
    }
    for (size_t b = 0; b < nb && !first_rc; b++) {
        const int k = (int)(b & 1);
        const uint32_t *d_sc = on_device ? (const uint32_t *)scalars[b] : stage[k]->as<uint32_t>();
        rc = E.enqueue(pts, d_sc, n, use_tab ? &t : nullptr, k, &pend[k], pipelined, on_device ? ctx->ev_batch : ctx->ev_up[k]);
        pend_b[k] = b;
        if (rc) first_rc = rc;
        if (!on_device) {
            CKC(cudaEventRecord(ctx->ev_free[k], ctx->stream));
            if (b + 1 < nb && !first_rc && (rc = upload(b + 1)) != 0) first_rc = rc;
        }
        if (!pipelined) {
            if (!first_rc && (rc = E.finish(pend[k], &out[b])) != 0) first_rc = rc;
        } else if (b >= 1 && pend[k ^ 1].active) {
            if ((rc = E.finish(pend[k ^ 1], &out[b - 1])) != 0 && !first_rc) first_rc = rc;
        }
    }
    // drain whatever is still in flight, older first (also on an error: the landing zones must be quiet on return)
    for (int i = 0; i < 2; i++) {
        const int k = (pend[0].active && pend[1].active) ? (pend_b[0] < pend_b[1] ? i : i ^ 1) : i;
        if (!pend[k].active) continue;
        rc = E.finish(pend[k], &out[pend_b[k]]);
        if (rc && !first_rc) first_rc = rc;
    }
    if (!on_device) cudaStreamSynchronize(ctx->copy_stream);
    return first_rc;
}
} // extern "C++"

int dvp_srs_append(dvp_ctx *ctx, int slot, const uint8_t *pts30, size_t n, int64_t *first_invalid) {
    if (!slot_ok(ctx, slot) || (!pts30 && n)) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    SrsSlot &s = ctx->slots[slot];
    s.invalidate();
    const size_t need = (s.n + n) * sizeof(AffPt);
    if (need > s.buf.cap) {
        DevBuf nb;
        int rc = nb.reserve(need);
        if (rc) return rc;
        if (s.n) CKC(cudaMemcpyAsync(nb.p, s.buf.p, s.n * sizeof(AffPt), cudaMemcpyDeviceToDevice, ctx->stream));
        CKC(cudaStreamSynchronize(ctx->stream));
        s.buf.release();
        s.buf = nb;
    }
    if (n == 0) return DVP_OK;
    int rc = ctx_decode_into(ctx, pts30, n, s.buf.as<AffPt>() + s.n, first_invalid);
    if (rc) return rc;
    s.n += n;
    return DVP_OK;
}
int dvp_srs_load(dvp_ctx *ctx, int slot, const uint8_t *pts30, size_t n, int64_t *first_invalid) {
    if (!slot_ok(ctx, slot)) return DVP_ERR_BAD_ARG;
    ctx->slots[slot].n = 0;
    ctx->slots[slot].invalidate();
    return dvp_srs_append(ctx, slot, pts30, n, first_invalid);
}
int dvp_srs_random(dvp_ctx *ctx, int slot, size_t n, uint64_t seed) {
    if (!slot_ok(ctx, slot)) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    SrsSlot &s = ctx->slots[slot];
    s.invalidate();
    int rc;
    if ((rc = s.buf.reserve((n ? n : 1) * sizeof(AffPt))) != 0) return rc;
    s.n = n;
    if (!n) return DVP_OK;
    k_random_points<<<cdivu(n, 64), 64, 0, ctx->stream>>>(seed, n, s.buf.as<AffPt>());
    CKC(cudaGetLastError());
    CKC(cudaStreamSynchronize(ctx->stream));
    return DVP_OK;
}
int dvp_srs_mulgen(dvp_ctx *ctx, int slot, const uint64_t *scalars_mont, size_t n) {
    if (!slot_ok(ctx, slot) || (!scalars_mont && n)) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    SrsSlot &s = ctx->slots[slot];
    s.invalidate();
    int rc;
    if ((rc = s.buf.reserve((n ? n : 1) * sizeof(AffPt))) != 0) return rc;
    s.n = n;
    if (!n) return DVP_OK;
    if ((rc = ctx->scal.reserve(n * 32 + 32)) != 0) return rc;
    CKC(cudaMemcpyAsync(ctx->scal.p, scalars_mont, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return ctx->msm.mulgen((const uint32_t *)ctx->scal.p, n, s.buf.as<AffPt>());
}
int dvp_srs_size(dvp_ctx *ctx, int slot, size_t *n) {
    if (!slot_ok(ctx, slot) || !n) return DVP_ERR_BAD_ARG;
    *n = ctx->slots[slot].n;
    return DVP_OK;
}
int dvp_srs_free(dvp_ctx *ctx, int slot) {
    if (!slot_ok(ctx, slot)) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    ctx->slots[slot].buf.release();
    ctx->slots[slot].invalidate();
    ctx->slots[slot].n = 0;
    return DVP_OK;
}
int dvp_srs_read(dvp_ctx *ctx, int slot, size_t offset, size_t n, uint8_t *pts30) {
    if (!slot_ok(ctx, slot) || (!pts30 && n)) return DVP_ERR_BAD_ARG;
    SrsSlot &s = ctx->slots[slot];
    if (offset > s.n || n > s.n - offset) return DVP_ERR_LENGTH_MISMATCH;
    if (!n) return DVP_OK;
    CKC(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->bytes.reserve(n * 30)) != 0) return rc;
    k_encode30<<<cdivu(n, 128), 128, 0, ctx->stream>>>(s.buf.as<AffPt>() + offset, n, (uint8_t *)ctx->bytes.p);
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(pts30, ctx->bytes.p, n * 30, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return DVP_OK;
}

int dvp_msm_device(dvp_ctx *ctx, int slot, size_t offset, const void *d_scalars, size_t n, uint8_t out30[30]) {
    if (!slot_ok(ctx, slot) || !out30 || (!d_scalars && n)) return DVP_ERR_BAD_ARG;
    SrsSlot &s = ctx->slots[slot];
    if (offset > s.n || n > s.n - offset) return DVP_ERR_LENGTH_MISMATCH;
    CKC(cudaSetDevice(ctx->device));
    AffPt r;
    int rc = slot_msm(ctx, slot, offset, (const uint32_t *)d_scalars, n, &r);
    if (rc) return rc;
    host::encode30(out30, r);
    return DVP_OK;
}
int dvp_msm(dvp_ctx *ctx, int slot, size_t offset, const uint64_t *scalars_mont, size_t n, uint8_t out30[30]) {
    if (!slot_ok(ctx, slot) || !out30 || (!scalars_mont && n)) return DVP_ERR_BAD_ARG;
    SrsSlot &s = ctx->slots[slot];
    if (offset > s.n || n > s.n - offset) return DVP_ERR_LENGTH_MISMATCH;
    CKC(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->scal.reserve(n * 32 + 32)) != 0) return rc;
    if (n) CKC(cudaMemcpyAsync(ctx->scal.p, scalars_mont, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    return dvp_msm_device(ctx, slot, offset, ctx->scal.p, n, out30);
}
int dvp_msm_batch(dvp_ctx *ctx, int slot, size_t offset, const uint64_t *const *scalars_mont, size_t n, size_t nb,
                  int scalars_on_device, uint8_t *out30) {
    if (!slot_ok(ctx, slot) || (nb && (!out30 || !scalars_mont))) return DVP_ERR_BAD_ARG;
    SrsSlot &s = ctx->slots[slot];
    if (offset > s.n || n > s.n - offset) return DVP_ERR_LENGTH_MISMATCH;
    for (size_t b = 0; b < nb; b++)
        if (!scalars_mont[b] && n) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    std::vector<AffPt> r(nb);
    int rc = slot_msm_batch(ctx, slot, offset, scalars_mont, n, nb, scalars_on_device != 0, r.data());
    if (rc) return rc;
    for (size_t b = 0; b < nb; b++) host::encode30(out30 + 30 * b, r[b]);
    return DVP_OK;
}
int dvp_msm_adhoc(dvp_ctx *ctx, const uint8_t *pts30, const uint64_t *scalars_mont, size_t n, uint8_t out30[30]) {
    if (!ctx || !out30 || ((!pts30 || !scalars_mont) && n)) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->adhoc.reserve(n * sizeof(AffPt) + 64)) != 0) return rc;
    if ((rc = ctx->scal.reserve(n * 32 + 32)) != 0) return rc;
    AffPt r = pt_inf();
    if (n) {
        if ((rc = ctx_decode_into(ctx, pts30, n, ctx->adhoc.as<AffPt>(), nullptr)) != 0) return rc;
        CKC(cudaMemcpyAsync(ctx->scal.p, scalars_mont, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = ctx->msm.run(ctx->adhoc.as<AffPt>(), ctx->scal.as<uint32_t>(), n, &r)) != 0) return rc;
    }
    host::encode30(out30, r);
    return DVP_OK;
}
int dvp_msm_last_stats(dvp_ctx *ctx, dvp_msm_stats *out) {
    if (!ctx || !out) return DVP_ERR_BAD_ARG;
    const MsmStats &s = ctx->msm.last;
    out->window_bits = s.window_bits;
    out->windows = s.windows;
    out->rounds_main = s.rounds_main;
    out->rounds_a = s.rounds_a;
    out->rounds_b = s.rounds_b;
    out->launches = s.launches;
    out->ms_recode_sort = s.ms_recode_sort;
    out->ms_accumulate = s.ms_accumulate;
    out->ms_reduce = s.ms_reduce;
    out->ms_tail = s.ms_tail;
    out->ms_pass2_round0 = s.ms_pass2_round0;
    out->adds_round0 = s.adds_round0;
    out->ms_device = s.ms_device;
    out->lanes = s.lanes;
    out->tables = s.tables;
    out->ms_tail_host = s.ms_tail_host;
    return DVP_OK;
}

/* development: per-category kernel time of the last MSM run with the "msm_profile" knob set.
 * categories: 0 sort 1 plan 2 pass1 3 binv_up 4 binv_direct 5 binv_down 6 pass2 7 misc */
int dvp_msm_last_profile(dvp_ctx *ctx, float ms[8], unsigned count[8]) {
    if (!ctx || !ms || !count) return DVP_ERR_BAD_ARG;
    for (int i = 0; i < 8; i++) {
        ms[i] = i < dvp::PC_COUNT ? ctx->msm.prof_ms[i] : 0.f;
        count[i] = i < dvp::PC_COUNT ? ctx->msm.prof_n[i] : 0u;
    }
    return DVP_OK;
}

/* development: (lane, category, start ms, end ms) of every launch bracket of the last profiled MSM */
int dvp_msm_last_timeline(dvp_ctx *ctx, float *rows4, size_t cap_rows, size_t *count) {
    if (!ctx || !count) return DVP_ERR_BAD_ARG;
    const auto &t = ctx->msm.timeline;
    *count = t.size();
    if (rows4)
        for (size_t i = 0; i < t.size() && i < cap_rows; i++) {
            rows4[4 * i] = t[i].lane;
            rows4[4 * i + 1] = t[i].cat;
            rows4[4 * i + 2] = t[i].t0;
            rows4[4 * i + 3] = t[i].t1;
        }
    return DVP_OK;
}

int dvp_point_add(dvp_ctx *ctx, const uint8_t a30[30], const uint8_t b30[30], uint8_t out30[30]) {
    if (!ctx || !a30 || !b30 || !out30) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->small.reserve(128)) != 0) return rc;
    uint8_t h[60];
    memcpy(h, a30, 30);
    memcpy(h + 30, b30, 30);
    uint8_t *d = (uint8_t *)ctx->small.p;
    CKC(cudaMemcpyAsync(d, h, 60, cudaMemcpyHostToDevice, ctx->stream));
    k_point_add30<<<1, 1, 0, ctx->stream>>>(d, d + 64, (int *)(d + 96));
    CKC(cudaGetLastError());
    uint8_t res[36];
    CKC(cudaMemcpyAsync(res, d + 64, 36, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    int ok;
    memcpy(&ok, res + 32, 4);
    if (!ok) return DVP_ERR_INVALID_POINT;
    memcpy(out30, res, 30);
    return DVP_OK;
}

int dvp_dev_alloc(dvp_ctx *ctx, size_t bytes, void **dptr) {
    if (!ctx || !dptr) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    if (cudaMalloc(dptr, bytes ? bytes : 1) != cudaSuccess) return DVP_ERR_OOM;
    return DVP_OK;
}
int dvp_dev_free(dvp_ctx *ctx, void *dptr) {
    if (!ctx) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaFree(dptr));
    return DVP_OK;
}
int dvp_dev_upload(dvp_ctx *ctx, void *dptr, const void *host, size_t bytes) {
    if (!ctx || !dptr || !host) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaMemcpyAsync(dptr, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return DVP_OK;
}
int dvp_dev_download(dvp_ctx *ctx, void *host, const void *dptr, size_t bytes) {
    if (!ctx || !dptr || !host) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    CKC(cudaMemcpyAsync(host, dptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    return DVP_OK;
}

int dvp_selftest_op(dvp_ctx *ctx, int op, const void *a, const void *b, void *out, size_t n) {
    if (!ctx || !a || !out || op < 0 || op > 7) return DVP_ERR_BAD_ARG; // 6, 7: warp-cooperative gf_mul / inverse
    if ((op == 0 || op == 3 || op == 5 || op == 6) && !b) return DVP_ERR_BAD_ARG;
    if (!n) return DVP_OK;
    CKC(cudaSetDevice(ctx->device));
    const size_t esz = op == 5 ? 64 : 32;
    void *da = nullptr, *db = nullptr, *dout = nullptr;
    CKC(cudaMalloc(&da, n * esz));
    CKC(cudaMalloc(&db, n * esz));
    CKC(cudaMalloc(&dout, n * esz));
    CKC(cudaMemcpyAsync(da, a, n * esz, cudaMemcpyHostToDevice, ctx->stream));
    if (b) CKC(cudaMemcpyAsync(db, b, n * esz, cudaMemcpyHostToDevice, ctx->stream));
    if (op >= 6) {
        int rcw = selftest_warp(ctx->msm, op - 6, da, db, dout, n);
        if (rcw) return rcw;
    } else {
        k_selftest<<<cdivu(n, 128), 128, 0, ctx->stream>>>(op, (const uint32_t *)da, b ? (const uint32_t *)db : nullptr,
                                                          (uint32_t *)dout, n);
    }
    CKC(cudaGetLastError());
    CKC(cudaMemcpyAsync(out, dout, n * esz, cudaMemcpyDeviceToHost, ctx->stream));
    CKC(cudaStreamSynchronize(ctx->stream));
    cudaFree(da);
    cudaFree(db);
    cudaFree(dout);
    return DVP_OK;
}

int dvp_latency_probe(dvp_ctx *ctx, int mode, int iters, float *us_per_op) {
    if (!ctx || !us_per_op || mode < 0 || mode > 4 || iters <= 0) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    return latency_probe(ctx->msm, mode, iters, us_per_op);
}
int dvp_microbench(dvp_ctx *ctx, int op, int iters, double *ops_per_sec) {
    // op = primitive (0 gf_mul, 1 gf_sqr, 2 fr_mul) + 10 * (resident blocks per SM - 1)
    if (!ctx || !ops_per_sec || op < 0 || op % 10 > 2 || op / 10 > 2 || iters <= 0) return DVP_ERR_BAD_ARG;
    CKC(cudaSetDevice(ctx->device));
    int rc;
    if ((rc = ctx->small.reserve(128)) != 0) return rc;
    cudaDeviceProp prop;
    CKC(cudaGetDeviceProperties(&prop, ctx->device));
    const int minb = op / 10 + 1, prim = op % 10;
    const int blocks = prop.multiProcessorCount * minb * 2;
    cudaEvent_t e0, e1;
    CKC(cudaEventCreate(&e0));
    CKC(cudaEventCreate(&e1));
    uint32_t *sink = (uint32_t *)ctx->small.p;
    for (int rep = 0; rep < 2; rep++) {
        const int it = rep ? iters : 4;
        if (rep) CKC(cudaEventRecord(e0, ctx->stream));
        if (minb == 1) launch_microbench<1>(prim, blocks, ctx->stream, it, sink);
        else if (minb == 2) launch_microbench<2>(prim, blocks, ctx->stream, it, sink);
        else launch_microbench<3>(prim, blocks, ctx->stream, it, sink);
    }
    CKC(cudaEventRecord(e1, ctx->stream));
    CKC(cudaEventSynchronize(e1));
    float ms = 0;
    CKC(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *ops_per_sec = (double)blocks * 256 * 2.0 * iters / (ms * 1e-3);
    return DVP_OK;
}

} // extern "C"
