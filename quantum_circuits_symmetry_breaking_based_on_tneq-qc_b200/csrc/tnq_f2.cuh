// tnq_f2.cuh -- packed fp32 pairs for sm_100 (shared by tnq_ladder_core.cuh and tnq_chain.cu).
#pragma once

#include <math.h>

#ifndef TNQ_HD
#ifdef __CUDACC__
#define TNQ_HD __device__ __forceinline__
#else
#define TNQ_HD inline
#endif
#endif

namespace tnq_ladder {

// ---- packed fp32 arithmetic ---------------------------------------------------------------------
// On sm_100 a plain FFMA issues every second cycle per scheduler (measured: 38.9 TFLOP/s over 148
// SMs, tools/ffma_rate.cu); only the packed form fma.rn.f32x2 (SASS FFMA2: two independent fp32
// FMAs on an aligned register pair, scalar operands broadcast for free) reaches the fp32 peak.  All
// hot loops below are therefore "vector += scalar * vector" over Vec<N> = N/2 register pairs (+ one
// plain float when N is odd).  Each half is an ordinary IEEE fma: the CPU emulation uses fmaf twice.
struct F2 {
    float lo, hi;
};
TNQ_HD F2 fma2(F2 a, F2 b, F2 c) {
#ifdef __CUDA_ARCH__
    unsigned long long A, B, C, R;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a.lo), "f"(a.hi));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b.lo), "f"(b.hi));
    asm("mov.b64 %0, {%1, %2};" : "=l"(C) : "f"(c.lo), "f"(c.hi));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(R) : "l"(A), "l"(B), "l"(C));
    F2 r;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(r.lo), "=f"(r.hi) : "l"(R));
    return r;
#else
    return F2{fmaf(a.lo, b.lo, c.lo), fmaf(a.hi, b.hi, c.hi)};
#endif
}
TNQ_HD F2 fma2(F2 a, float s, F2 c) { return fma2(a, F2{s, s}, c); }

}  // namespace tnq_ladder
