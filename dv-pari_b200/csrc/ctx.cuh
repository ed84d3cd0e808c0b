// The opaque context behind dvp_ctx: one device, one stream, resident SRS slots, grow-only scratch.
#pragma once
#include "../../include/dvpari.h"
#include "msm.cuh"

struct dvp_local_group;
struct SrsSlot {
    dvp::DevBuf buf; // AffPt[n], decoded once (replaces read_point_vec_from_file per prove)
    size_t n = 0;
    // precomputed window multiples of the slot (built on first use, dropped when the slot changes)
    dvp::DevBuf table;
    dvp::MsmTable tab;
    bool table_ok = false, table_failed = false;
    uint64_t version = 0; // bumped whenever the points change (a prover's joint g_m | g_q copy checks it)
    void invalidate() {
        table.release();
        table_ok = table_failed = false;
        version++;
    }
};

struct dvp_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr; // an MSM that overlaps work on `stream` runs here (dvp_prove)
    cudaEvent_t ev_aux = nullptr;
    dvp::MsmEngine msm;
    SrsSlot slots[DVP_MAX_SRS_SLOTS];
    dvp::DevBuf bytes, small, scal, adhoc, commbuf;
    // batches of MSMs (dvp_msm_batch): the scalars of the next MSM are uploaded on their own stream into a second
    // staging buffer while the current one runs
    dvp::DevBuf scal2;
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t ev_up[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr}, ev_batch = nullptr;
    // multi-GPU: NCCL communicator (ncclComm_t) of this rank, see comm.cu
    void *comm = nullptr;
    struct dvp_local_group *local = nullptr; // or: a group of contexts of this process (dvp_comm_init_local)
    int rank = 0, world = 1;
    int msm_table_windows = 0;             // 0: choose_table_windows(slot size)
    int msm_tables = 1;                    // 0: never, 1: when the slot is large enough and the memory is there
    size_t msm_tables_min = (size_t)1 << 15; // smallest slot that gets tables
    // dvp_prove: commit_p = msm(w, g_m) + msm(q, g_q) (proving.rs:463-515) as ONE MSM over g_m | g_q.
    // -1: whenever the joint vector got its window tables (it wins at every size measured: 2^16 constraints 6.85 ->
    // 5.78 ms, 2^19 23.9 -> 23.0, 2^21 77.2 -> 74.8, 2^22 141.3 -> 139.5), 0 never, 1 always
    int prove_joint = -1;
};

// sum_i scalars[i] * slot[offset + i] on the device of ctx (device scalars); uses the slot's tables when it has them
// `on` = the stream the scalars were produced on and the MSM is ordered after (default: the context's stream)
int slot_msm(dvp_ctx *ctx, int slot, size_t offset, const uint32_t *d_scalars, size_t n, dvp::AffPt *out,
             cudaStream_t on = nullptr);
// the same over a point vector that is not one of the context's numbered slots (a prover's joint g_m | g_q copy)
int slot_msm_at(dvp_ctx *ctx, SrsSlot &s, size_t offset, const uint32_t *d_scalars, size_t n, dvp::AffPt *out,
                cudaStream_t on = nullptr);

// comm.cu
int comm_all_gather(dvp_ctx *ctx, const void *send, void *recv, size_t bytes_per_rank);
int comm_broadcast(dvp_ctx *ctx, void *buf, size_t bytes, int root);
int comm_group(dvp_ctx *ctx, bool start);
int comm_agree(dvp_ctx *ctx, int rc);
int comm_fold_points(dvp_ctx *ctx, const dvp::AffPt &mine, int rc_local, dvp::AffPt *total);
// the same for nb partial sums at once (one all-gather): total[b] = sum over the ranks of mine[b]
int comm_fold_points_n(dvp_ctx *ctx, const dvp::AffPt *mine, int rc_local, size_t nb, dvp::AffPt *total);
// nb MSMs over slot[offset, offset + n), pipelined: device work of MSM b+1 (and the upload of its scalars when they
// are host vectors) is enqueued before the host folds the partial sums of MSM b
int slot_msm_batch(dvp_ctx *ctx, int slot, size_t offset, const uint64_t *const *scalars, size_t n, size_t nb,
                   bool on_device, dvp::AffPt *out);

int ctx_decode_into(dvp_ctx *ctx, const uint8_t *pts30, size_t n, dvp::AffPt *d_out, int64_t *first_invalid);
