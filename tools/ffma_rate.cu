// microbenchmark: plain FFMA vs packed FFMA2 (fma.rn.f32x2) issue rate on sm_100a
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, float b, unsigned long long c) {
    unsigned long long r, bb; asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(bb), "l"(c)); return r; }
template <int MODE> __global__ void k(float* out, float s0, float s1, int iters) {
    float acc[32]; unsigned long long acc2[16];
    float x[4] = {s0 + threadIdx.x, s1, s0 * 2, s1 * 3};
    for (int i = 0; i < 32; ++i) acc[i] = i + threadIdx.x;
    for (int i = 0; i < 16; ++i) acc2[i] = ((unsigned long long)__float_as_uint(1.f + i) << 32) | __float_as_uint(2.f + threadIdx.x);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 32; ++i) acc[i] = fmaf(acc[i], x[i & 3], x[(i + 1) & 3]);
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) acc2[i] = fma2(acc2[i], x[i & 3], acc2[(i + 1) & 15]);
        }
    }
    float r = 0; for (int i = 0; i < 32; ++i) r += acc[i];
    for (int i = 0; i < 16; ++i) r += __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 1024 * 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int warps = 4; warps <= 32; warps *= 2) for (int mode = 0; mode < 2; ++mode) {
        const int iters = 20000; float ms;
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<0><<<148, warps * 32>>>(out, 1.0001f, 0.9999f, iters); else k<1><<<148, warps * 32>>>(out, 1.0001f, 0.9999f, iters);
            cudaEventRecord(e1); cudaEventSynchronize(e1); cudaEventElapsedTime(&ms, e0, e1);
        }
        double fmas = 148.0 * warps * 32 * 32.0 * iters;
        printf("warps/SM %2d %s: %.3f ms  %.1f TFLOP/s  (%.2f warp-instr/clk/SMSP at 1.965 GHz)\n", warps, mode ? "FFMA2" : "FFMA ", ms,
               2 * fmas / ms / 1e9, (148.0 * warps * (mode ? 16.0 : 32.0) * iters) / (ms * 1e-3 * 1.965e9) / (148 * 4));
    }
    return 0;
}
