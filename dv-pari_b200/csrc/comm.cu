// Multi-GPU plumbing of the C ABI: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The reference is single-process (rayon inside one prove).  Here the three MSMs of Proof::prove
// (/root/reference/src/proving.rs:463,512,680) shard by contiguous point range, the row evaluation
// (proving.rs:382-396) by row range and the extends (proving.rs:410-422) by polynomial.  The only
// exchanges are: an all-gather of the row-range outputs, a broadcast of every extended polynomial from
// its owner, and an all-gather of one 64-byte partial sum per rank and MSM, folded identically on every
// rank (GF(2^233) point addition is not an NCCL reduction operator).
// A second backend joins several contexts of ONE process (each driven by its own host thread, on one device or
// several) without NCCL: the same collectives as device-to-device copies between the ranks' buffers behind a
// host-side rendezvous (dvp_comm_init_local).  It runs the whole partition logic on a single GPU.
// libnccl is bound at run time (dlopen) so that the library loads on hosts without it; the unique id
// travels by whatever channel the host has (torch.distributed store, MPI, a file).
#include <dlfcn.h>
#include <nccl.h>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <vector>
#include <mutex>
#include "ctx.cuh"
#include "host_gf.hpp"

using namespace dvp;

namespace {
struct NcclApi {
    void *h = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    bool ok = false;
};
NcclApi g_nccl;

bool nccl_load() {
    if (g_nccl.ok) return true;
    if (!g_nccl.h) {
        const char *names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char *nm : names) {
            g_nccl.h = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
            if (g_nccl.h) break;
        }
    }
    if (!g_nccl.h) return false;
#define SYM(field, name)                                            \
    *(void **)(&g_nccl.field) = dlsym(g_nccl.h, name);              \
    if (!g_nccl.field) return false
    SYM(GetUniqueId, "ncclGetUniqueId");
    SYM(CommInitRank, "ncclCommInitRank");
    SYM(CommDestroy, "ncclCommDestroy");
    SYM(AllGather, "ncclAllGather");
    SYM(Broadcast, "ncclBroadcast");
    SYM(GroupStart, "ncclGroupStart");
    SYM(GroupEnd, "ncclGroupEnd");
    SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
    g_nccl.ok = true;
    return true;
}
} // namespace

#define NCK(x)                                                                                                   \
    do {                                                                                                         \
        ncclResult_t r_ = (x);                                                                                   \
        if (r_ != ncclSuccess) {                                                                                 \
            fprintf(stderr, "[dvpari] NCCL error %s at %s:%d\n", g_nccl.GetErrorString(r_), __FILE__, __LINE__); \
            return DVP_ERR_NCCL;                                                                                 \
        }                                                                                                        \
    } while (0)
#define CKN(x)                                                                                                \
    do {                                                                                                      \
        cudaError_t e_ = (x);                                                                                 \
        if (e_ != cudaSuccess) {                                                                              \
            fprintf(stderr, "[dvpari] CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); \
            return DVP_ERR_CUDA;                                                                              \
        }                                                                                                     \
    } while (0)

// ------------------------------------------------------------------------------------------------
// in-process backend: the ranks are contexts of this process; a collective is a rendezvous of their host threads
// around plain device-to-device copies (cudaMemcpyDefault: same device, peer or staged)
// ------------------------------------------------------------------------------------------------
struct dvp_local_group {
    int world = 0, refs = 0;
    std::mutex mu;
    std::condition_variable cv;
    int arrived = 0;
    unsigned long long gen = 0;
    const void *ptr[64] = {nullptr};
    int status[64] = {0};
    void barrier() {
        std::unique_lock<std::mutex> lk(mu);
        const unsigned long long g = gen;
        if (++arrived == world) {
            arrived = 0;
            gen++;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return gen != g; });
        }
    }
};

static int local_all_gather(dvp_ctx *ctx, const void *send, void *recv, size_t bytes) {
    dvp_local_group *g = ctx->local;
    CKN(cudaStreamSynchronize(ctx->stream)); // my part is complete before anybody reads it
    g->ptr[ctx->rank] = send;
    g->barrier();
    for (int r = 0; r < g->world; r++) {
        void *dst = (char *)recv + (size_t)r * bytes;
        if (dst != g->ptr[r]) CKN(cudaMemcpyAsync(dst, g->ptr[r], bytes, cudaMemcpyDefault, ctx->stream));
    }
    CKN(cudaStreamSynchronize(ctx->stream));
    g->barrier(); // every rank has read: the send buffers may change again
    return DVP_OK;
}
static int local_broadcast(dvp_ctx *ctx, void *buf, size_t bytes, int root) {
    dvp_local_group *g = ctx->local;
    CKN(cudaStreamSynchronize(ctx->stream));
    g->ptr[ctx->rank] = buf;
    g->barrier();
    if (ctx->rank != root) CKN(cudaMemcpyAsync(buf, g->ptr[root], bytes, cudaMemcpyDefault, ctx->stream));
    CKN(cudaStreamSynchronize(ctx->stream));
    g->barrier();
    return DVP_OK;
}

// all-gather `bytes` per rank (device buffers) on the context stream
int comm_all_gather(dvp_ctx *ctx, const void *send, void *recv, size_t bytes) {
    if (ctx->local) return local_all_gather(ctx, send, recv, bytes);
    NCK(g_nccl.AllGather(send, recv, bytes, ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
    return DVP_OK;
}
int comm_broadcast(dvp_ctx *ctx, void *buf, size_t bytes, int root) {
    if (ctx->local) return local_broadcast(ctx, buf, bytes, root);
    NCK(g_nccl.Broadcast(buf, buf, bytes, ncclUint8, root, (ncclComm_t)ctx->comm, ctx->stream));
    return DVP_OK;
}
// several collectives as one NCCL group (the in-process backend runs them one by one)
int comm_group(dvp_ctx *ctx, bool start) {
    if (ctx->local || ctx->world <= 1) return DVP_OK;
    NCK(start ? g_nccl.GroupStart() : g_nccl.GroupEnd());
    return DVP_OK;
}
// The first non-zero status of any rank, on every rank: a rank-local failure (out of memory, a bad handle) must not
// leave the peers waiting in the next collective.  Host-side for the in-process backend, an 4-byte all-gather otherwise.
int comm_agree(dvp_ctx *ctx, int rc) {
    if (ctx->world <= 1) return rc;
    if (ctx->local) {
        dvp_local_group *g = ctx->local;
        g->status[ctx->rank] = rc;
        g->barrier();
        int all = 0;
        for (int r = 0; r < g->world && !all; r++) all = g->status[r];
        g->barrier();
        return all;
    }
    int rc2;
    const size_t W = (size_t)ctx->world;
    if ((rc2 = ctx->commbuf.reserve(4096)) != 0) return rc ? rc : rc2; // sized in dvp_comm_init: cannot fail here
    int *d = ctx->commbuf.as<int>() + 512;
    int all[64];
    CKN(cudaMemcpyAsync(d + W, &rc, sizeof(int), cudaMemcpyHostToDevice, ctx->stream));
    NCK(g_nccl.AllGather(d + W, d, sizeof(int), ncclUint8, (ncclComm_t)ctx->comm, ctx->stream));
    CKN(cudaMemcpyAsync(all, d, W * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CKN(cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < W; r++)
        if (all[r]) return all[r];
    return DVP_OK;
}

// Sum of the ranks' partial MSM results, identical on every rank: all-gather 64-byte affine points, fold in
// rank order on the host (CurvePoint::add, curve.rs:76-82).  The status of the local MSM travels with the point
// (80 bytes per rank): a rank whose MSM failed still takes part, and every rank returns the first failure -- nobody
// is left waiting in the collective.
int comm_fold_points(dvp_ctx *ctx, const AffPt &mine, int rc_local, AffPt *total) {
    return comm_fold_points_n(ctx, &mine, rc_local, 1, total);
}

int comm_fold_points_n(dvp_ctx *ctx, const AffPt *mine, int rc_local, size_t nb, AffPt *total) {
    if (ctx->world <= 1) {
        for (size_t b = 0; b < nb; b++) total[b] = mine[b];
        return rc_local;
    }
    struct Slot {
        AffPt p;
        int rc, pad[3];
    };
    static_assert(sizeof(Slot) == 80, "point + status");
    int rc;
    const size_t W = (size_t)ctx->world;
    if (nb == 0) nb = 1; // the status still travels
    if ((rc = ctx->commbuf.reserve(4096 + (W + 1) * nb * sizeof(Slot))) != 0) return rc_local ? rc_local : rc;
    Slot *d = reinterpret_cast<Slot *>(ctx->commbuf.as<char>() + 4096);
    std::vector<Slot> me(nb), all(W * nb);
    for (size_t b = 0; b < nb; b++) {
        me[b].p = (rc_local || !mine) ? pt_inf() : mine[b];
        me[b].rc = rc_local;
        me[b].pad[0] = me[b].pad[1] = me[b].pad[2] = 0;
    }
    CKN(cudaMemcpyAsync(d + W * nb, me.data(), nb * sizeof(Slot), cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = comm_all_gather(ctx, d + W * nb, d, nb * sizeof(Slot))) != 0) return rc;
    CKN(cudaMemcpyAsync(all.data(), d, W * nb * sizeof(Slot), cudaMemcpyDeviceToHost, ctx->stream));
    CKN(cudaStreamSynchronize(ctx->stream));
    for (size_t r = 0; r < W; r++)
        if (all[r * nb].rc) return all[r * nb].rc;
    for (size_t b = 0; b < nb; b++) {
        AffPt acc = all[b].p;
        for (size_t r = 1; r < W; r++) acc = host::aff_add(acc, all[r * nb + b].p);
        if (total) total[b] = acc;
    }
    return DVP_OK;
}

extern "C" {

int dvp_comm_unique_id(uint8_t id[128]) {
    if (!id) return DVP_ERR_BAD_ARG;
    if (!nccl_load()) return DVP_ERR_NCCL;
    ncclUniqueId u;
    NCK(g_nccl.GetUniqueId(&u));
    static_assert(sizeof(u) == 128, "ncclUniqueId is 128 bytes");
    memcpy(id, &u, 128);
    return DVP_OK;
}

int dvp_comm_init(dvp_ctx *ctx, const uint8_t id[128], int rank, int world) {
    if (!ctx || !id || world < 1 || world > 64 || rank < 0 || rank >= world) return DVP_ERR_BAD_ARG;
    if (ctx->comm) return DVP_ERR_BAD_ARG;
    if (!nccl_load()) return DVP_ERR_NCCL;
    CKN(cudaSetDevice(ctx->device));
    ncclUniqueId u;
    memcpy(&u, id, 128);
    ncclComm_t c;
    NCK(g_nccl.CommInitRank(&c, world, u, rank));
    ctx->comm = c;
    ctx->rank = rank;
    ctx->world = world;
    return ctx->commbuf.reserve(16384); // the small exchanges (partial sums, status words) never allocate later
}

// `world` contexts of this process become ranks 0 .. world-1 of one group without NCCL.  Every rank must then be
// driven by its own host thread (the collectives rendezvous); the contexts may share a device.
int dvp_comm_init_local(dvp_ctx *const *ctxs, int world) {
    if (!ctxs || world < 1 || world > 64) return DVP_ERR_BAD_ARG;
    for (int r = 0; r < world; r++)
        if (!ctxs[r] || ctxs[r]->comm || ctxs[r]->local) return DVP_ERR_BAD_ARG;
    dvp_local_group *g = new dvp_local_group();
    g->world = world;
    g->refs = world;
    for (int r = 0; r < world; r++) {
        ctxs[r]->local = g;
        ctxs[r]->rank = r;
        ctxs[r]->world = world;
        if (cudaSetDevice(ctxs[r]->device) != cudaSuccess || ctxs[r]->commbuf.reserve(16384)) return DVP_ERR_CUDA;
    }
    return DVP_OK;
}

int dvp_comm_destroy(dvp_ctx *ctx) {
    if (!ctx) return DVP_ERR_BAD_ARG;
    if (ctx->comm) {
        cudaSetDevice(ctx->device);
        cudaStreamSynchronize(ctx->stream);
        g_nccl.CommDestroy((ncclComm_t)ctx->comm);
    }
    if (ctx->local) {
        dvp_local_group *g = ctx->local;
        bool last;
        {
            std::lock_guard<std::mutex> lk(g->mu);
            last = --g->refs == 0;
        }
        if (last) delete g;
        ctx->local = nullptr;
    }
    ctx->comm = nullptr;
    ctx->rank = 0;
    ctx->world = 1;
    return DVP_OK;
}

int dvp_comm_info(dvp_ctx *ctx, int *rank, int *world) {
    if (!ctx) return DVP_ERR_BAD_ARG;
    if (rank) *rank = ctx->rank;
    if (world) *world = ctx->world;
    return DVP_OK;
}

void dvp_shard_range(size_t total, int rank, int world, size_t *lo, size_t *hi) {
    if (world < 1) world = 1;
    const size_t a = (size_t)((unsigned __int128)total * (unsigned)rank / (unsigned)world);
    const size_t b = (size_t)((unsigned __int128)total * (unsigned)(rank + 1) / (unsigned)world);
    if (lo) *lo = a;
    if (hi) *hi = b;
}

// multi_scalar_mul over a point vector sharded by contiguous range: this rank's slot holds its range, `scalars_mont`
// are the scalars of that range.  Every rank receives the encoding of the whole sum.
int dvp_msm_sharded(dvp_ctx *ctx, int slot, const uint64_t *scalars_mont, size_t n, int scalars_on_device, uint8_t out30[30]) {
    if (!ctx || slot < 0 || slot >= DVP_MAX_SRS_SLOTS || !out30 || (!scalars_mont && n)) return DVP_ERR_BAD_ARG;
    SrsSlot &s = ctx->slots[slot];
    if (n != s.n) return DVP_ERR_LENGTH_MISMATCH;
    CKN(cudaSetDevice(ctx->device));
    int rc;
    const void *d_sc = scalars_mont;
    if (!scalars_on_device) {
        if ((rc = ctx->scal.reserve(n * 32 + 32)) != 0) return rc;
        if (n) CKN(cudaMemcpyAsync(ctx->scal.p, scalars_mont, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        d_sc = ctx->scal.p;
    }
    AffPt mine = pt_inf(), total;
    rc = slot_msm(ctx, slot, 0, (const uint32_t *)d_sc, n, &mine);
    if ((rc = comm_fold_points(ctx, mine, rc, &total)) != 0) return rc;
    host::encode30(out30, total);
    return DVP_OK;
}

int dvp_msm_sharded_batch(dvp_ctx *ctx, int slot, const uint64_t *const *scalars_mont, size_t n, size_t nb,
                          int scalars_on_device, uint8_t *out30) {
    if (!ctx || slot < 0 || slot >= DVP_MAX_SRS_SLOTS || (nb && (!out30 || !scalars_mont))) return DVP_ERR_BAD_ARG;
    SrsSlot &s = ctx->slots[slot];
    if (n != s.n) return DVP_ERR_LENGTH_MISMATCH;
    for (size_t b = 0; b < nb; b++)
        if (!scalars_mont[b] && n) return DVP_ERR_BAD_ARG;
    CKN(cudaSetDevice(ctx->device));
    std::vector<AffPt> mine(nb), total(nb);
    int rc = slot_msm_batch(ctx, slot, 0, scalars_mont, n, nb, scalars_on_device != 0, mine.data());
    // one all-gather for the whole batch: nb x (64-byte partial sum + status) per rank
    if ((rc = comm_fold_points_n(ctx, mine.data(), rc, nb, total.data())) != 0) return rc;
    for (size_t b = 0; b < nb; b++) host::encode30(out30 + 30 * b, total[b]);
    return DVP_OK;
}

} // extern "C"
