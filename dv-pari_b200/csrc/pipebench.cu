// Integer-pipe microbenchmarks for sm_100a: establishes the issue-rate roofline that the MSM / field
// kernels are measured against (MEASURED_PEAKS.json only carries HBM and bf16 tensor peaks).
//   mode 0: IMAD.WIDE.U32 (32x32+64)     mode 1: LOP3           mode 2: IMAD (32-bit lo)
//   mode 3: 1 IMAD.WIDE : 2 LOP3 mix      mode 4: SHF funnel shift  mode 5: IADD3
// Each thread runs 8 independent dependency chains; result = thread-instructions per second.
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/dvpari.h"
#include "ctx.cuh"

template <int MODE>
__global__ void __launch_bounds__(256) k_pipe(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t x[8];
    uint64_t w[8];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = t * 2654435761u + i * 40503u + seed;
        w[i] = ((uint64_t)x[i] << 32) | (x[i] ^ 0x9e3779b9u);
    }
    const uint32_t m = seed | 0x11111111u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (MODE == 0) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(m));
                } else if (MODE == 1) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(m));
                } else if (MODE == 2) {
                    asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(m), "r"(seed));
                } else if (MODE == 3) {
                    asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"(x[i]), "r"(m));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(m));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xe8;" : "+r"(x[(i + 3) & 7]) : "r"(x[(i + 5) & 7]), "r"(m));
                } else if (MODE == 4) {
                    asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]), "r"(m));
                } else {
                    asm volatile("add.u32 %0, %0, %1;" : "+r"(x[i]) : "r"(x[(i + 1) & 7]));
                }
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= x[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
    if (s == 0x12345678u) sink[0] = s;
}

extern "C" int dvp_pipebench(dvp_ctx *ctx, int mode, int iters, int blocks_per_sm, double *instr_per_sec) {
    if (!ctx || !instr_per_sec || mode < 0 || mode > 5 || iters <= 0 || blocks_per_sm <= 0) return DVP_ERR_BAD_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return DVP_ERR_CUDA;
    if (ctx->small.reserve(128)) return DVP_ERR_OOM;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, ctx->device);
    const int blocks = prop.multiProcessorCount * blocks_per_sm;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    uint32_t *sink = (uint32_t *)ctx->small.p;
    for (int rep = 0; rep < 2; rep++) {
        const int it = rep == 0 ? 4 : iters;
        if (rep == 1) cudaEventRecord(e0, ctx->stream);
        switch (mode) {
        case 0: k_pipe<0><<<blocks, 256, 0, ctx->stream>>>(it, 7, sink); break;
        case 1: k_pipe<1><<<blocks, 256, 0, ctx->stream>>>(it, 7, sink); break;
        case 2: k_pipe<2><<<blocks, 256, 0, ctx->stream>>>(it, 7, sink); break;
        case 3: k_pipe<3><<<blocks, 256, 0, ctx->stream>>>(it, 7, sink); break;
        case 4: k_pipe<4><<<blocks, 256, 0, ctx->stream>>>(it, 7, sink); break;
        default: k_pipe<5><<<blocks, 256, 0, ctx->stream>>>(it, 7, sink); break;
        }
    }
    cudaEventRecord(e1, ctx->stream);
    if (cudaEventSynchronize(e1) != cudaSuccess) return DVP_ERR_CUDA;
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    const double per_iter = (mode == 3 ? 3.0 : 1.0) * 64.0;
    *instr_per_sec = (double)blocks * 256.0 * per_iter * iters / (ms * 1e-3);
    return DVP_OK;
}
