// Development harness (run under gpurun): throughput of GF(2^233) multiplier variants vs register budget.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xptxas -v -o scripts/mulbench scripts/mulbench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../dv-pari_b200/csrc/gf233.cuh"

using namespace dvp;

template <int V> __device__ __forceinline__ gf mulv(const gf &a, const gf &b) {
    if (V == 0) return gf_mul_portable(a, b); // C form: compiler-chosen order
    if (V == 2) return gf_mul_dev2(a, b);     // 32-bit IMAD only, straight + bit-reversed streams
    return gf_mul_dev(a, b);                  // explicit mul.wide / mad.wide order (the device path of gf_mul)
}
__device__ __noinline__ gf mul_call0(const gf a, const gf b) { return gf_mul_portable(a, b); }
__device__ __noinline__ gf mul_call1(const gf a, const gf b) { return gf_mul_dev(a, b); }
__device__ __noinline__ gf mul_call2(const gf a, const gf b) { return gf_mul_dev2(a, b); }

template <int V, int THREADS, int MINB, int CALL>
__global__ void __launch_bounds__(THREADS, MINB) k_mul(const gf *__restrict__ in, gf *__restrict__ out, int iters) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    gf x = in[2 * i], y = in[2 * i + 1];
    for (int it = 0; it < iters; it++) {
        if (CALL) x = V == 2 ? mul_call2(x, y) : V ? mul_call1(x, y) : mul_call0(x, y);
        else x = mulv<V>(x, y);
        y.v[0] ^= x.v[3];
        y.v[5] ^= x.v[1];
    }
    out[i] = x;
}

template <int V, int THREADS, int MINB, int CALL> void run(const char *name, const gf *d_in, gf *d_out, gf *h_out, int nsm, int iters) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_mul<V, THREADS, MINB, CALL>, THREADS, 0);
    cudaFuncAttributes fa;
    cudaFuncGetAttributes(&fa, k_mul<V, THREADS, MINB, CALL>);
    const int blocks = nsm * occ * 4;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    k_mul<V, THREADS, MINB, CALL><<<blocks, THREADS>>>(d_in, d_out, 4);
    cudaEventRecord(e0);
    k_mul<V, THREADS, MINB, CALL><<<blocks, THREADS>>>(d_in, d_out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaMemcpy(h_out, d_out, 64 * sizeof(gf), cudaMemcpyDeviceToHost);
    uint32_t cs = 0;
    for (int k = 0; k < 64; k++)
        for (int w = 0; w < 8; w++) cs = cs * 31 + h_out[k].v[w];
    const double rate = (double)blocks * THREADS * iters / (ms * 1e-3);
    printf("%-28s regs=%3d lmem=%4zu occ=%d blk/SM (%2d warps/SM)  %.3e mul/s  checksum %08x  err=%s\n", name, fa.numRegs,
           (size_t)fa.localSizeBytes, occ, occ * THREADS / 32, rate, cs, cudaGetErrorString(cudaGetLastError()));
}

int main(int argc, char **argv) {
    const int iters = argc > 1 ? atoi(argv[1]) : 200;
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int nsm = prop.multiProcessorCount;
    const size_t nthr = (size_t)nsm * 2048 * 4;
    gf *h = (gf *)malloc(2 * nthr * sizeof(gf));
    uint64_t s = 0x9e3779b97f4a7c15ull;
    for (size_t i = 0; i < 2 * nthr; i++) {
        for (int w = 0; w < 8; w++) {
            s ^= s << 13; s ^= s >> 7; s ^= s << 17;
            h[i].v[w] = (uint32_t)(s >> 11);
        }
        h[i].v[7] &= 0x1ff;
    }
    gf *d_in, *d_out;
    cudaMalloc(&d_in, 2 * nthr * sizeof(gf));
    cudaMalloc(&d_out, nthr * sizeof(gf));
    cudaMemcpy(d_in, h, 2 * nthr * sizeof(gf), cudaMemcpyHostToDevice);
    gf *h_out = (gf *)malloc(64 * sizeof(gf));
#define R(V, T, M, C) run<V, T, M, C>("v" #V " thr" #T " minb" #M " call" #C, d_in, d_out, h_out, nsm, iters)
    R(0, 256, 1, 0); R(0, 256, 2, 0); R(0, 256, 3, 0); R(0, 256, 4, 0);
    R(0, 256, 2, 1); R(0, 256, 3, 1); R(0, 256, 4, 1);
    R(1, 256, 1, 0); R(1, 256, 2, 0); R(1, 256, 3, 0); R(1, 256, 4, 0);
    R(1, 256, 2, 1); R(1, 256, 3, 1); R(1, 256, 4, 1);
    R(2, 256, 1, 0); R(2, 256, 2, 0); R(2, 256, 3, 0); R(2, 256, 4, 0);
    R(2, 256, 2, 1); R(2, 256, 3, 1); R(2, 256, 4, 1); R(2, 128, 4, 0);
    R(1, 128, 5, 0); R(1, 128, 6, 0); R(1, 128, 8, 0);
    return 0;
}
