// Development probe (run under gpurun): which integer instructions overlap on sm_100a.
// Every mode interleaves two instruction streams on disjoint registers (8 independent chains each).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/pipebench2 scripts/pipebench2.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define A_NONE 0
#define A_WIDE 1   // mul.wide.u32 (64-bit result, no accumulate)
#define A_WIDEACC 2 // mad.wide.u32 accumulate in place
#define A_LO 3     // mad.lo.u32
#define A_HI 4     // mad.hi.u32
#define A_LOP 5    // lop3
#define A_SHF 6
#define A_ADD 7    // add.u32 (compiler's choice)
#define A_IADD3 8  // 3-input add
#define A_PRMT 9
#define A_LOPIMM 10 // lop3 with one immediate (2 register reads)
#define A_BREV 11
#define A_POPC 12

template <int OP> __device__ __forceinline__ void op(uint32_t &x, uint32_t &y, uint64_t &w, uint32_t m) {
    if (OP == A_WIDE) { // w <- lo(w) * hi(w): both halves feed the next product
        const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
        asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(w) : "r"(lo), "r"(hi));
    }
    if (OP == A_WIDEACC) { // w <- lo(w) * hi(w) + C, C loop-invariant
        const uint32_t lo = (uint32_t)w, hi = (uint32_t)(w >> 32);
        const uint64_t c = ((uint64_t)m << 32) | y;
        asm volatile("mad.wide.u32 %0, %1, %2, %3;" : "=l"(w) : "r"(lo), "r"(hi), "l"(c));
    }
    if (OP == A_LO) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(y));
    if (OP == A_HI) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x) : "r"(m), "r"(y));
    if (OP == A_LOP) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x) : "r"(y), "r"(m));
    if (OP == A_LOPIMM) asm volatile("lop3.b32 %0, %0, %1, 0x33333333, 0xe8;" : "+r"(x) : "r"(y));
    if (OP == A_SHF) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(m));
    if (OP == A_ADD) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y));
    if (OP == A_IADD3) asm volatile("{.reg .u32 t; add.u32 t, %0, %1; add.u32 %0, t, %2;}" : "+r"(x) : "r"(y), "r"(m));
    if (OP == A_BREV) asm volatile("brev.b32 %0, %0;" : "+r"(x));
    if (OP == A_POPC) asm volatile("popc.b32 %0, %0;" : "+r"(x));
    if (OP == A_PRMT) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(x) : "r"(y), "r"(m));
}

template <int OPA, int NA, int OPB, int NB>
__global__ void __launch_bounds__(256) k_mix(int iters, uint32_t seed, uint32_t *sink) {
    uint32_t xa[8], ya[8], xb[8], yb[8];
    uint64_t wa[8], wb[8];
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        xa[i] = t * 2654435761u + i * 40503u + seed;
        ya[i] = xa[i] ^ 0x1234567u;
        xb[i] = xa[i] * 3u + 1;
        yb[i] = xb[i] ^ 0x7654321u;
        wa[i] = xa[i];
        wb[i] = xb[i];
    }
    const uint32_t m = seed | 0x11111111u;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < 8; i++) {
#pragma unroll
                for (int k = 0; k < NA; k++) op<OPA>(xa[i], ya[(i + k) & 7], wa[i], m);
#pragma unroll
                for (int k = 0; k < NB; k++) op<OPB>(xb[(i + k) & 7], yb[(i + 3) & 7], wb[i], m);
            }
        }
    }
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s ^= xa[i] ^ xb[i] ^ (uint32_t)wa[i] ^ (uint32_t)(wa[i] >> 32) ^ (uint32_t)wb[i] ^ (uint32_t)(wb[i] >> 32);
    if (s == 0x12345678u) sink[0] = s;
}

template <int OPA, int NA, int OPB, int NB> void run(const char *name, uint32_t *sink, int nsm) {
    const int iters = 2000, blocks = nsm * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_mix<OPA, NA, OPB, NB><<<blocks, 256>>>(4, 7, sink);
    cudaEventRecord(e0);
    k_mix<OPA, NA, OPB, NB><<<blocks, 256>>>(iters, 7, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double n = (double)blocks * 256 * iters * 32.0 * (NA + NB);
    const double rate = n / (ms * 1e-3);
    // warp-instr per clk per SMSP at 1.965 GHz
    printf("%-40s %.3e thread-instr/s  IPC/SMSP(at 1.965GHz) %.3f\n", name, rate, rate / 32 / (nsm * 4) / 1.965e9);
}

int main(int argc, char **argv) {
    const bool only_ncu_cases = argc > 1; // `pipebench2 ncu`: the handful of mixes captured under ncu (profiles/)
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    uint32_t *sink;
    cudaMalloc(&sink, 64);
#define R(a, na, b, nb) run<a, na, b, nb>(#a "x" #na " + " #b "x" #nb, sink, nsm)
    if (only_ncu_cases) {
        R(A_WIDE, 1, A_NONE, 0);
        R(A_LOP, 1, A_NONE, 0);
        R(A_WIDE, 1, A_LOP, 1);
        R(A_WIDE, 1, A_LOP, 2);
        R(A_WIDE, 1, A_LOP, 3);
        R(A_LO, 1, A_LOP, 1);
        return 0;
    }
    R(A_WIDE, 1, A_NONE, 0);
    R(A_WIDEACC, 1, A_NONE, 0);
    R(A_LO, 1, A_NONE, 0);
    R(A_HI, 1, A_NONE, 0);
    R(A_LOP, 1, A_NONE, 0);
    R(A_LOPIMM, 1, A_NONE, 0);
    R(A_SHF, 1, A_NONE, 0);
    R(A_PRMT, 1, A_NONE, 0);
    R(A_ADD, 1, A_NONE, 0);
    R(A_IADD3, 1, A_NONE, 0);
    R(A_WIDE, 1, A_LOP, 1);
    R(A_WIDE, 1, A_LOP, 2);
    R(A_WIDE, 1, A_LOPIMM, 2);
    R(A_WIDEACC, 1, A_LOP, 1);
    R(A_WIDEACC, 1, A_LOP, 2);
    R(A_LO, 1, A_LOP, 1);
    R(A_LO, 1, A_LOP, 2);
    R(A_LO, 1, A_LOPIMM, 1);
    R(A_HI, 1, A_LOP, 1);
    R(A_HI, 1, A_LOP, 2);
    R(A_LO, 1, A_HI, 1);
    R(A_LO, 1, A_SHF, 1);
    R(A_WIDE, 1, A_SHF, 1);
    R(A_WIDE, 1, A_ADD, 1);
    R(A_LOP, 1, A_SHF, 1);
    R(A_LOP, 1, A_ADD, 1);
    R(A_LOP, 1, A_PRMT, 1);
    R(A_WIDE, 1, A_LO, 1);
    R(A_BREV, 1, A_NONE, 0);
    R(A_BREV, 1, A_LOP, 1);
    R(A_BREV, 1, A_LOP, 4);
    R(A_BREV, 1, A_LO, 4);
    R(A_LO, 1, A_LOP, 2);
    R(A_LO, 2, A_LOP, 1);
    return 0;
}
