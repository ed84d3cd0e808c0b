// MSM engine interface (device-resident points, Montgomery Fr scalars) -- see msm.cu.
#pragma once
#include <cstddef>
#include <cstdint>
#include <vector>
#include <cuda_runtime.h>
#include "k233.cuh"

namespace dvp {

// Grow-only device scratch owned by a context.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes); // 0 on success
    void release();
    template <class T> T *as() { return reinterpret_cast<T *>(p); }
};

struct MsmStats {
    int window_bits = 0, windows = 0, rounds_main = 0, rounds_a = 0, rounds_b = 0, lanes = 0, tables = 0;
    unsigned long long launches = 0; // kernels launched by the last msm
    float ms_recode_sort = 0, ms_accumulate = 0, ms_reduce = 0, ms_tail = 0;
    // the dominant kernel: pass 2 of round 0 of the bucket accumulation (one launch)
    float ms_pass2_round0 = 0;
    float ms_tail_host = 0; // host fold of the per-bit sums (wall clock)
    float ms_device = 0; // recode .. partial sums on the host (CUDA events on the context stream), always measured
    unsigned long long adds_round0 = 0, adds_total = 0;
};

// One independent chain of tree rounds: its own stream and scratch.  The virtual windows of an MSM (bucket ranges)
// are split over the lanes so that the latency-bound inversion chains of one lane overlap the large kernels of another.
// launch categories of the development profiler
enum { PC_SORT = 0, PC_PLAN, PC_PASS1, PC_BINV_UP, PC_BINV_DIRECT, PC_BINV_DOWN, PC_PASS2, PC_MISC, PC_COUNT };

struct MsmLane {
    cudaStream_t stream = nullptr;    // high priority: the lane's timeline
    cudaStream_t stream_lo = nullptr; // low priority: large pass kernels detour through it
    cudaEvent_t ev_sw[2] = {nullptr, nullptr};
    cudaEvent_t done = nullptr;
    DevBuf seg_len[2], seg_start[2], c_len, c_start, blk, blk_flag, info, info_r0, pp[2], prefix, desc,
        thr_total, thr_inv, lvl_pre[2], lvl_tot[2], lvl_inv[2], buckets, rc, ents2, acc_ctl, acc_times;
    // The plan of all rounds of a k_accumulate launch, made ahead of it (k_pa_*): per-round segment tables, control
    // words and descriptors.  Two sets for the bucket accumulation (a pipelined batch plans MSM b+1 on the sort stream
    // while MSM b reads its own), one for reduction level A, whose plan depends only on the window layout and is kept
    // from one MSM to the next (key).
    struct PlanSet {
        DevBuf blk, start, len, ctl, desc;
        uint64_t key = 0;
        uint32_t stride = 0;
        void release() {
            blk.release(), start.release(), len.release(), ctl.release(), desc.release();
            key = 0;
        }
    };
    PlanSet plan_main[2], plan_a;
    unsigned long long launches = 0;
    uint32_t epoch = 0; // k_plan launch counter (the blocks' publish flag)
    // profiler: (category, start, stop) per bracket; events are pooled
    struct ProfRec {
        int cat;
        cudaEvent_t e0, e1;
    };
    std::vector<ProfRec> prof;
    size_t prof_used = 0;
    void prof_begin(int cat);
    void prof_end();
    // timing of the dominant kernel (lane 0 only)
    cudaEvent_t ev_k[2] = {nullptr, nullptr}, ev_s[3] = {nullptr, nullptr, nullptr};
    bool want_k = false;
    int init(int index);
    void destroy();
};

// Precomputed multiples of a resident point vector: T[j][i] = 2^(off_j) P_i, j < W (off_j = bit offset of window j),
// T[j] at points + j * stride.
// With them every window shares ONE bucket set (see MsmEngine::run), so wider windows pay off.
struct MsmTable {
    int W = 0;
    size_t stride = 0; // points per window table (the slot size)
    size_t offset = 0; // first point of the range this MSM runs over
};

struct MsmEngine {
    cudaStream_t stream = nullptr; // the context's stream: recode, final read-back
    std::vector<MsmLane> lanes;
    DevBuf keys, entries, len_all, start_all, cursor_all, scan_blk, lane_info, hb, msqr_tabs, mg_table;
    // second set of the sort's outputs (sorted entries, bucket lengths and starts): in a pipelined batch the sort of
    // the next MSM runs on `sort_stream` while the current MSM is still reading its own set
    DevBuf entries_b, len_all_b, start_all_b;
    cudaStream_t sort_stream = nullptr; // low priority: fills the latency-bound reduction phase of the MSM before
    int sort_ahead = 1;                 // 0: never sort ahead (every MSM starts after the previous one's read-back)
    int preplan = 1; // persistent path: tables and descriptors of all rounds before the launch (0: every round plans itself)
    // Host-side landing zones come in two sets (index k of MsmPending), so that the device work of the next MSM of a
    // batch can be enqueued before the host has folded the partial sums of the previous one.
    void *h_lane = nullptr; // pinned, 2 x 128 words: per-lane (entries, longest bucket) | control words of k_accumulate
    void *h_pts[2] = {nullptr, nullptr}; // pinned, receive the per-bit partial sums
    size_t h_pts_cap[2] = {0, 0};
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr}, ev_recode = nullptr, ev_t0[2] = {nullptr, nullptr},
                ev_t1[2] = {nullptr, nullptr};
    MsmStats last;
    int force_window_bits = 0; // 0 = choose from n
    int force_lanes = 0;       // 0 = choose from n
    bool timing = false;
    bool profile = false; // per-category CUDA-event timing of every launch (development; forces one lane)
    struct TimelineRow {
        float lane, cat, t0, t1;
    };
    std::vector<TimelineRow> timeline; // of the last profiled run: one row per bracket, ms since the start of the MSM
    float prof_ms[PC_COUNT] = {0};
    unsigned prof_n[PC_COUNT] = {0};
    size_t b16_min = (size_t)1 << 21; // rounds with at least this many additions chain 16 per thread (below: 4 or 1)
    size_t b64_min = (size_t)1 << 23; // rounds with at least this many additions chain 64 per thread (off by default)
    bool prio_split = true; // large pass kernels on low-priority streams
    int pass_b_max = 64;    // cap on the additions chained per thread
    int ld_tree_warp_a = 1; // level A's tree after the batched-affine rounds: one warp per addition (cooperative products)
    size_t ld_tree_max = 0; // 0 = automatic; a reduction level with more points starts with batched-affine rounds
    uint32_t binv_direct = 36864;    // batches up to this size are inverted by one cooperative launch (k_binv_coop)
    uint32_t binv_coop_warps = 2368; // ... in groups sized so that about this many warps run (one wave)
    size_t round_warp_max = 4096; // a round of at most this many additions runs one warp per addition (k_round_warp)
    int use_accumulate = 1; // all rounds of a lane in one persistent cooperative launch (k_accumulate): 0 never,
                            // 1 for the sizes where it wins, 2 always
    uint32_t acc_capacity = 0;  // blocks of k_accumulate that are co-resident on the device (0: no cooperative launch)
    int fused_rounds = 0; // separate-launch path: one kernel per round (pass 1, per-warp inversion, pass 2) instead of 3-7.
                          // Off: with chains of 16 the warp's own inversion (10 products + 6.7 us) is 16 % of its work,
                          // the hierarchical inversion 4 % (2^22: 24.9 against 22.3 ms, 2^24: 84.0 against 76.9)
    int pass2_minb = 1; // resident blocks per SM the pass-2 kernel is compiled for (register cap); 1 = 255 registers,
                        // no spills: with the staged operands 8 warps per SM are enough (2^22: 22.2 ms against 23.4 at 2)

    int init(cudaStream_t s);
    void destroy();
    // sum_i scalars[i] * points[i]; scalars are device pointers to n x 8 x u32 Montgomery limbs.
    // Result: affine E[r] point (or infinity) on the host.  Returns 0 or a DVP_ERR_* code.
    // With `tab`, d_points is the base of the tables and the MSM runs over points [tab->offset, tab->offset + n).
    int run(const AffPt *d_points, const uint32_t *d_scalars, size_t n, AffPt *h_result, const MsmTable *tab = nullptr);
    // The two halves of run(): enqueue() puts the whole device side of an MSM on the streams (no host wait on the
    // persistent path) and leaves the partial sums on their way to landing zone k; finish() waits for them and folds
    // them on the host.  Between the two the caller may enqueue the next MSM (with the other k): the engine's device
    // scratch is reused in stream order, only the landing zones are per-k.
    struct Pending {
        bool active = false, uniform = false, persistent_any = false, nosync = false;
        int k = 0, c = 0, cv = 0, base = 0, rem = 0, NL = 0;
        uint32_t V = 0, nbv = 0;
        unsigned long long launches = 0;
        MsmStats stt;
    };
    // `ahead`: the caller guarantees that the scalars are valid once `scalars_ready` has fired (nullptr: already) no
    // matter what else is queued on the engine's stream; the recode + counting sort may then start before the
    // previous MSM has finished (persistent path only).
    int enqueue(const AffPt *d_points, const uint32_t *d_scalars, size_t n, const MsmTable *tab, int k, Pending *P,
                bool ahead = false, cudaEvent_t scalars_ready = nullptr);
    int finish(Pending &P, AffPt *h_result);
    // d_tab[j * n + i] = 2^(off_j) d_points[i] for j < W (d_tab holds W n points; j = 0 is a copy)
    int build_table(const AffPt *d_points, size_t n, int W, AffPt *d_tab);
    // d_out[i] = scalars[i] * G (batched fixed-base multiplication of the generator), all on the device
    int mulgen(const uint32_t *d_scalars, size_t n, AffPt *d_out);
    int reserve_round(MsmLane &L, size_t task_ub);
};

int choose_window_bits(size_t n);
int choose_table_windows(size_t n);
int latency_probe(MsmEngine &E, int mode, int iters, float *us_per_op);
int selftest_warp(MsmEngine &E, int op, const void *d_a, const void *d_b, void *d_out, size_t n); // 0 mul, 1 inverse

} // namespace dvp
