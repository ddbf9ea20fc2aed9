// tnq_chain.cu -- register-resident sweep for single-layer MPS networks (the reference's default
// graph: QCTNHelper.generate_example_graph(graph_type="mps"), examples/example_train_single_node.py,
// tests/test_probabilities.py), real float32, edge rank K in {2, 3, 4}.
//
// For this family every greedy group (tneq_qc/contractor/greedy_strategy.py:690-990) has the form
//     "cdef,aeg,higj,ahc,d,i->ajf":  env'[j,f] = sum Ls[c,e,f] M[e,g] Ls[h,g,j] env[h,c]
// with Ls[c,e,f] = sum_d G[c,d,e,f] s[d] (core folded with the next qubit's circuit state), the first
// group being the same formula with env = s0 s0^T, and the last one "acd,adc->a" a trace against M.
// The environment is K x K: it lives in REGISTERS.  One thread owns one sample and walks all n qubits:
//   T1[h,e,f] = env[h,c] Ls[c,e,f] ; T2[h,g,f] = T1[h,e,f] M[e,g] ; env'[j,f] = T2[h,g,f] Ls[h,g,j]
// (3 K^4 FMAs per qubit, everything unrolled at compile time, Ls read as shared-memory broadcasts,
// M streamed from HBM once: 4 K^2 bytes per qubit and sample -- the kernel's only real traffic).
// Training fuses the loss (engine_siamese.py:490-530) and the reverse sweep: the n left environments
// are kept per thread (local memory, L1 resident), the sweep runs back down the chain, and the
// per-sample core-gradient contributions are combined with warp shuffles, then per CTA, then by a
// small finalize kernel in a fixed order (deterministic).
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "tneq_b200.h"
#include "tnq_f2.cuh"

extern int tnq_internal_fail(const std::string& msg);
extern int tnq_internal_cuda_fail(cudaError_t e, const char* what);
extern void tnq_internal_count_launch();

namespace {

constexpr int MAXQ = TNQ_CHAIN_MAX_QUBITS;
constexpr int CHAIN_THREADS = 128;

struct ChainArgs {
    const float* core[MAXQ];     // n-1 cores, each [K][K][K][K] = G[c,d,e,f]
    const float* state[MAXQ];    // n circuit states, each [K]
    const float* mx[MAXQ];       // n measurement tensors, sample b at mx[q] + b * mx_stride[q], [K][K] row major
    long long mx_stride[MAXQ];
    float* grad[MAXQ];           // n-1 gradient outputs [K][K][K][K] (train / bwd)
    int n;
};

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one step of the chain, forward: env <- step(env, Ls, M)
template <int K>
__device__ __forceinline__ void chain_step(float (&env)[K][K], const float* __restrict__ Ls, const float (&M)[K][K]) {
    float T1[K][K][K];   // [h][e][f]
#pragma unroll
    for (int h = 0; h < K; ++h)
#pragma unroll
        for (int e = 0; e < K; ++e)
#pragma unroll
            for (int f = 0; f < K; ++f) {
                float s = 0.f;
#pragma unroll
                for (int c = 0; c < K; ++c) s = fmaf(env[h][c], Ls[(c * K + e) * K + f], s);
                T1[h][e][f] = s;
            }
    float T2[K][K][K];   // [h][g][f]
#pragma unroll
    for (int h = 0; h < K; ++h)
#pragma unroll
        for (int g = 0; g < K; ++g)
#pragma unroll
            for (int f = 0; f < K; ++f) {
                float s = 0.f;
#pragma unroll
                for (int e = 0; e < K; ++e) s = fmaf(T1[h][e][f], M[e][g], s);
                T2[h][g][f] = s;
            }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int f = 0; f < K; ++f) {
            float s = 0.f;
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int g = 0; g < K; ++g) s = fmaf(T2[h][g][f], Ls[(h * K + g) * K + j], s);
            env[j][f] = s;
        }
}

template <int K>
__device__ __forceinline__ void load_m(float (&M)[K][K], const float* __restrict__ p, bool valid) {
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) M[i][j] = valid ? __ldg(p + i * K + j) : 0.f;
}

// Coalesced fetch of one qubit's measurement matrices for the 32 samples of a warp: when the batch
// is contiguous (stride K*K) the 32*K*K floats are one contiguous run, read with K*K fully
// coalesced loads into registers (raw[i] = run[lane + 32 i]) ...
template <int K>
__device__ __forceinline__ void fetch_m(float (&raw)[K * K], const float* __restrict__ base, long long b0, long long B,
                                        int lane) {
    const long long total = (B - b0 < 32 ? B - b0 : 32) * (K * K);
    const float* p = base + b0 * (K * K);
#pragma unroll
    for (int i = 0; i < K * K; ++i) {
        const int idx = lane + 32 * i;
        raw[i] = idx < total ? __ldg(p + idx) : 0.f;
    }
}
// ... and transposed through a per-warp shared-memory slab into per-lane matrices (stride K*K is odd
// for K = 3 and the slab is padded for K = 2, 4, so both sides are conflict free).
template <int K>
__device__ __forceinline__ void unpack_m(float (&M)[K][K], const float (&raw)[K * K], float* slab, int lane) {
    constexpr int LD = (K * K) | 1;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < K * K; ++i) {
        const int idx = lane + 32 * i;
        slab[(idx / (K * K)) * LD + idx % (K * K)] = raw[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) M[i][j] = slab[lane * LD + i * K + j];
}

// MODE 0: values only.  MODE 1: fused loss + gradients.  MODE 2: gradients seeded by `seed` (autograd).
template <int K, int MODE>
__global__ void __launch_bounds__(CHAIN_THREADS)
tnq_chain_kernel(const __grid_constant__ ChainArgs a, long long B, const float* __restrict__ seed,
                 float* __restrict__ values, float* __restrict__ partials, float log_scale, float inv_count) {
    constexpr int K3 = K * K * K;
    extern __shared__ float sm[];
    const int n = a.n;
    float* Ls = sm;                                   // [n-1][K3]
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int SLAB = 32 * ((K * K) | 1);
    float* slab = sm + (n - 1) * K3 + warp * SLAB;    // per-warp staging of one qubit's Mx
    float* wacc = sm + (n - 1) * K3 + warps * SLAB;   // [warps][(n-1) * K3 + 1] gradient / loss accumulators
    const int acc_stride = (n - 1) * K3 + 1;
    bool packed = true;                               // every Mx contiguous over the batch?
    for (int q = 0; q < n; ++q) packed = packed && (a.mx_stride[q] == K * K);
    // Ls[q][c][e][f] = sum_d G_q[c,d,e,f] s_{q+1}[d]
    for (int i = threadIdx.x; i < (n - 1) * K3; i += blockDim.x) {
        const int q = i / K3, r = i % K3, c = r / (K * K), e = (r / K) % K, f = r % K;
        float s = 0.f;
        for (int d = 0; d < K; ++d) s = fmaf(__ldg(a.core[q] + ((c * K + d) * K + e) * K + f), __ldg(a.state[q + 1] + d), s);
        Ls[i] = s;
    }
    if (MODE != 0)
        for (int i = threadIdx.x; i < warps * acc_stride; i += blockDim.x) wacc[i] = 0.f;
    __syncthreads();
    float s0[K];
#pragma unroll
    for (int i = 0; i < K; ++i) s0[i] = __ldg(a.state[0] + i);
    float* my_acc = wacc + warp * acc_stride;

    const long long nwarp_iters = (B + 31) / 32;
    for (long long wi = (long long)blockIdx.x * warps + warp; wi < nwarp_iters; wi += (long long)gridDim.x * warps) {
        const long long b = wi * 32 + lane;
        const bool valid = b < B;
        float env[K][K];
#pragma unroll
        for (int h = 0; h < K; ++h)
#pragma unroll
            for (int c = 0; c < K; ++c) env[h][c] = s0[h] * s0[c];
        float tape[MODE != 0 ? MAXQ : 1][K][K];       // left environments (local memory when training)
        float M[K][K], Mn[K][K], raw[K * K];
        const long long b0 = wi * 32;
        if (packed) {
            fetch_m<K>(raw, a.mx[0], b0, B, lane);
            unpack_m<K>(M, raw, slab, lane);
        } else {
            load_m<K>(M, a.mx[0] + b * a.mx_stride[0], valid);
        }
        for (int q = 0; q < n - 1; ++q) {
            if (MODE != 0) {
#pragma unroll
                for (int h = 0; h < K; ++h)
#pragma unroll
                    for (int c = 0; c < K; ++c) tape[q][h][c] = env[h][c];
            }
            // next qubit's Mx is in flight while this one is consumed
            if (packed)
                fetch_m<K>(raw, a.mx[q + 1], b0, B, lane);
            else
                load_m<K>(Mn, a.mx[q + 1] + b * a.mx_stride[q + 1], valid);
            chain_step<K>(env, Ls + q * K3, M);
            if (packed) {
                unpack_m<K>(M, raw, slab, lane);
            } else {
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int j = 0; j < K; ++j) M[i][j] = Mn[i][j];
            }
        }
        float val = 0.f;                               // "acd,adc->a"
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int d = 0; d < K; ++d) val = fmaf(env[c][d], M[d][c], val);
        if (MODE != 2 && values != nullptr && valid) values[b] = val;
        if (MODE == 0) continue;

        // ---- seed: d loss / d value ----
        float dval;
        if (MODE == 1) {
            const float clamped = fmaxf(val, 1e-10f);
            float lpart = valid ? -(logf(clamped) + log_scale) * inv_count : 0.f;
            lpart = warp_sum(lpart);
            if (lane == 0) my_acc[acc_stride - 1] += lpart;
            dval = (valid && val >= 1e-10f) ? -inv_count / clamped : 0.f;
        } else {
            dval = valid ? __ldg(seed + b) : 0.f;
        }
        float denv[K][K];                              // adjoint of the environment after the last step
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int d = 0; d < K; ++d) denv[c][d] = dval * M[d][c];

        // ---- reverse sweep ----
        if (packed)
            fetch_m<K>(raw, a.mx[n - 2], b0, B, lane);
        else
            load_m<K>(Mn, a.mx[n - 2] + b * a.mx_stride[n - 2], valid);
        for (int q = n - 2; q >= 0; --q) {
            const float* L = Ls + q * K3;
            if (packed) {
                unpack_m<K>(M, raw, slab, lane);
                if (q > 0) fetch_m<K>(raw, a.mx[q - 1], b0, B, lane);
            } else {
#pragma unroll
                for (int i = 0; i < K; ++i)
#pragma unroll
                    for (int j = 0; j < K; ++j) M[i][j] = Mn[i][j];
                if (q > 0) load_m<K>(Mn, a.mx[q - 1] + b * a.mx_stride[q - 1], valid);
            }
            float e0[K][K];
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int c = 0; c < K; ++c) e0[h][c] = tape[q][h][c];
            // recompute T1[h][e][f], T2[h][g][f]
            float T1[K][K][K], T2[K][K][K];
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        float s = 0.f;
#pragma unroll
                        for (int c = 0; c < K; ++c) s = fmaf(e0[h][c], L[(c * K + e) * K + f], s);
                        T1[h][e][f] = s;
                    }
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int g = 0; g < K; ++g)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        float s = 0.f;
#pragma unroll
                        for (int e = 0; e < K; ++e) s = fmaf(T1[h][e][f], M[e][g], s);
                        T2[h][g][f] = s;
                    }
            float* gq = my_acc + q * K3;
            // R role: dLs[h][g][j] += sum_f T2[h][g][f] denv[j][f] ; dT2[h][g][f] = sum_j denv[j][f] Ls[h][g][j]
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int g = 0; g < K; ++g) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        float s = 0.f;
#pragma unroll
                        for (int f = 0; f < K; ++f) s = fmaf(T2[h][g][f], denv[j][f], s);
                        s = warp_sum(s);
                        if (lane == 0) gq[(h * K + g) * K + j] += s;
                    }
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        float s = 0.f;
#pragma unroll
                        for (int j = 0; j < K; ++j) s = fmaf(denv[j][f], L[(h * K + g) * K + j], s);
                        T2[h][g][f] = s;                       // T2 now holds dT2
                    }
                }
            // dT1[h][e][f] = sum_g dT2[h][g][f] M[e][g]   (T1 still needed for nothing else -> keep e0 for dLs_L)
            float dT1[K][K][K];
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        float s = 0.f;
#pragma unroll
                        for (int g = 0; g < K; ++g) s = fmaf(T2[h][g][f], M[e][g], s);
                        dT1[h][e][f] = s;
                    }
            // L role: dLs[c][e][f] += sum_h e0[h][c] dT1[h][e][f] ; denv[h][c] = sum_{e,f} dT1[h][e][f] Ls[c][e][f]
#pragma unroll
            for (int c = 0; c < K; ++c)
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        float s = 0.f;
#pragma unroll
                        for (int h = 0; h < K; ++h) s = fmaf(e0[h][c], dT1[h][e][f], s);
                        s = warp_sum(s);
                        if (lane == 0) gq[(c * K + e) * K + f] += s;
                    }
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int c = 0; c < K; ++c) {
                    float s = 0.f;
#pragma unroll
                    for (int e = 0; e < K; ++e)
#pragma unroll
                        for (int f = 0; f < K; ++f) s = fmaf(dT1[h][e][f], L[(c * K + e) * K + f], s);
                    denv[h][c] = s;
                }
        }
    }
    if (MODE != 0) {
        __syncthreads();
        for (int i = threadIdx.x; i < acc_stride; i += blockDim.x) {
            float s = 0.f;
            for (int w = 0; w < warps; ++w) s += wacc[w * acc_stride + i];
            partials[(size_t)blockIdx.x * acc_stride + i] = s;
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Forward values with TWO samples per thread (MODE 0, batch-contiguous measurements).  Every
// per-sample quantity is a packed pair (sample b0 + lane, sample b0 + 32 + lane) and the whole step
// runs on fma.rn.f32x2 (SASS FFMA2) with the folded cores as broadcast scalars: plain FFMA issues
// only every second cycle per scheduler on sm_100 (tools/ffma_rate.cu), which capped the one-sample
// kernel at 27 TFLOP/s; packed, the sweep is bound by streaming the measurement matrices from HBM.
// ------------------------------------------------------------------------------------------------
using tnq_ladder::F2;
using tnq_ladder::fma2;

template <int K>
__device__ __forceinline__ void chain_step2(F2 (&env)[K][K], const float* __restrict__ Ls, const F2 (&M)[K][K]) {
    F2 out[K][K];            // [j][f]
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int f = 0; f < K; ++f) out[j][f] = F2{0.f, 0.f};
#pragma unroll
    for (int h = 0; h < K; ++h) {
        F2 T1[K][K];         // [e][f]
#pragma unroll
        for (int e = 0; e < K; ++e)
#pragma unroll
            for (int f = 0; f < K; ++f) {
                F2 s{0.f, 0.f};
#pragma unroll
                for (int c = 0; c < K; ++c) s = fma2(env[h][c], Ls[(c * K + e) * K + f], s);
                T1[e][f] = s;
            }
#pragma unroll
        for (int g = 0; g < K; ++g)
#pragma unroll
            for (int f = 0; f < K; ++f) {
                F2 t2{0.f, 0.f};
#pragma unroll
                for (int e = 0; e < K; ++e) t2 = fma2(T1[e][f], M[e][g], t2);
#pragma unroll
                for (int j = 0; j < K; ++j) out[j][f] = fma2(t2, Ls[(h * K + g) * K + j], out[j][f]);
            }
    }
#pragma unroll
    for (int j = 0; j < K; ++j)
#pragma unroll
        for (int f = 0; f < K; ++f) env[j][f] = out[j][f];
}

// 64 samples x K*K floats, contiguous: raw[i] = run[lane + 32 i]
template <int K>
__device__ __forceinline__ void fetch_m2(float (&raw)[2 * K * K], const float* __restrict__ base, long long b0, long long B,
                                         int lane) {
    const long long total = (B - b0 < 64 ? B - b0 : 64) * (K * K);
    const float* p = base + b0 * (K * K);
#pragma unroll
    for (int i = 0; i < 2 * K * K; ++i) {
        const int idx = lane + 32 * i;
        raw[i] = idx < total ? __ldg(p + idx) : 0.f;
    }
}
template <int K>
__device__ __forceinline__ void unpack_m2(F2 (&M)[K][K], const float (&raw)[2 * K * K], float* slab, int lane) {
    constexpr int LD = (K * K) | 1;
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 2 * K * K; ++i) {
        const int idx = lane + 32 * i;
        slab[(idx / (K * K)) * LD + idx % (K * K)] = raw[i];
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < K; ++i)
#pragma unroll
        for (int j = 0; j < K; ++j) M[i][j] = F2{slab[lane * LD + i * K + j], slab[(lane + 32) * LD + i * K + j]};
}

template <int K>
__global__ void __launch_bounds__(CHAIN_THREADS, K <= 3 ? 4 : 1)
tnq_chain_fwd2_kernel(const __grid_constant__ ChainArgs a, long long B, float* __restrict__ values) {
    constexpr int K3 = K * K * K;
    extern __shared__ float sm[];
    const int n = a.n;
    float* Ls = sm;                                   // [n-1][K3]
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int SLAB = 64 * ((K * K) | 1);
    float* slab = sm + (n - 1) * K3 + warp * SLAB;
    for (int i = threadIdx.x; i < (n - 1) * K3; i += blockDim.x) {
        const int q = i / K3, r = i % K3, c = r / (K * K), e = (r / K) % K, f = r % K;
        float s = 0.f;
        for (int d = 0; d < K; ++d) s = fmaf(__ldg(a.core[q] + ((c * K + d) * K + e) * K + f), __ldg(a.state[q + 1] + d), s);
        Ls[i] = s;
    }
    __syncthreads();
    float s0[K];
#pragma unroll
    for (int i = 0; i < K; ++i) s0[i] = __ldg(a.state[0] + i);
    const long long ntiles = (B + 63) / 64;
    for (long long wi = (long long)blockIdx.x * warps + warp; wi < ntiles; wi += (long long)gridDim.x * warps) {
        const long long b0 = wi * 64;
        F2 env[K][K], M[K][K];
#pragma unroll
        for (int h = 0; h < K; ++h)
#pragma unroll
            for (int c = 0; c < K; ++c) env[h][c] = F2{s0[h] * s0[c], s0[h] * s0[c]};
        float raw[2 * K * K];
        fetch_m2<K>(raw, a.mx[0], b0, B, lane);
        unpack_m2<K>(M, raw, slab, lane);
        for (int q = 0; q < n - 1; ++q) {
            fetch_m2<K>(raw, a.mx[q + 1], b0, B, lane);   // next qubit's Mx in flight during this step
            chain_step2<K>(env, Ls + q * K3, M);
            unpack_m2<K>(M, raw, slab, lane);
        }
        F2 val{0.f, 0.f};                              // "acd,adc->a"
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int d = 0; d < K; ++d) val = fma2(env[c][d], M[d][c], val);
        if (b0 + lane < B) values[b0 + lane] = val.lo;
        if (b0 + 32 + lane < B) values[b0 + 32 + lane] = val.hi;
    }
}

// ---- opt-in: measurement matrices generated in registers from x (EngineSiamese.generate_data fused into the sweep) ----
// Reference: tneq_qc/core/engine_siamese.py:59-111 (weights, Hermite recurrence) and :133-254 (generate_data):
//   phi_k(x) = w_k * sqrt(exp(-x^2 / 2)) * He_k(x),  He_0 = 1, He_1 = x, He_i = x He_{i-1} - (i-1) He_{i-2},
//   Mx = phi phi^T, wrapped as TNTensor: divided by its global abs-max over the batch (tn_tensor.py:72-85).
// The chain kernel then reads n floats per sample instead of n K^2: the HBM roofline of the sweep itself moves.
struct HermiteW {
    float w[4];
};
template <int K>
__device__ __forceinline__ void hermite_phi(float x, const HermiteW& hw, float (&phi)[K]) {
    const float g = sqrtf(expf(-(x * x) / 2.0f));
    float hm2 = 1.0f, hm1 = x;
    phi[0] = hw.w[0] * g;                         // (w * g) * H, as the reference associates it
    if (K > 1) phi[1] = hw.w[1] * g * x;
#pragma unroll
    for (int i = 2; i < K; ++i) {
        const float h = x * hm1 - (float)(i - 1) * hm2;
        phi[i] = hw.w[i] * g * h;
        hm2 = hm1, hm1 = h;
    }
}

// scale[q] = max over the batch of max_{i,j} |phi_i phi_j| = max_b (max_i |phi_i|)^2  (positive floats order like ints)
template <int K>
__global__ void tnq_hermite_scale_kernel(const float* __restrict__ x, long long xs_b, long long xs_q, long long B, int n,
                                         HermiteW hw, float* __restrict__ scale) {
    const int q = blockIdx.y;
    float best = 0.f;
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < B; b += (long long)gridDim.x * blockDim.x) {
        float phi[K];
        hermite_phi<K>(__ldg(x + b * xs_b + q * xs_q), hw, phi);
        float a = 0.f;
#pragma unroll
        for (int i = 0; i < K; ++i) a = fmaxf(a, fabsf(phi[i]));
        best = fmaxf(best, a * a);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) best = fmaxf(best, __shfl_xor_sync(0xffffffffu, best, o));
    if ((threadIdx.x & 31) == 0) atomicMax(reinterpret_cast<int*>(scale) + q, __float_as_int(best));
}

template <int K>
__global__ void __launch_bounds__(CHAIN_THREADS, K <= 3 ? 4 : 1)
tnq_chain_fwd2x_kernel(const __grid_constant__ ChainArgs a, long long B, const float* __restrict__ x, long long xs_b,
                       long long xs_q, HermiteW hw, const float* __restrict__ scale, float* __restrict__ values) {
    constexpr int K3 = K * K * K;
    extern __shared__ float sm[];
    const int n = a.n;
    float* Ls = sm;                                   // [n-1][K3]
    float* inv_s = sm + (n - 1) * K3;                 // [n]
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int i = threadIdx.x; i < (n - 1) * K3; i += blockDim.x) {
        const int q = i / K3, r = i % K3, c = r / (K * K), e = (r / K) % K, f = r % K;
        float s = 0.f;
        for (int d = 0; d < K; ++d) s = fmaf(__ldg(a.core[q] + ((c * K + d) * K + e) * K + f), __ldg(a.state[q + 1] + d), s);
        Ls[i] = s;
    }
    for (int i = threadIdx.x; i < n; i += blockDim.x) inv_s[i] = 1.0f / __ldg(scale + i);
    __syncthreads();
    float s0[K];
#pragma unroll
    for (int i = 0; i < K; ++i) s0[i] = __ldg(a.state[0] + i);
    const long long ntiles = (B + 63) / 64;
    for (long long wi = (long long)blockIdx.x * warps + warp; wi < ntiles; wi += (long long)gridDim.x * warps) {
        const long long b0 = wi * 64, ba = b0 + lane, bb = b0 + 32 + lane;
        const float* xa = x + (ba < B ? ba : B - 1) * xs_b;
        const float* xb = x + (bb < B ? bb : B - 1) * xs_b;
        F2 env[K][K], M[K][K];
#pragma unroll
        for (int h = 0; h < K; ++h)
#pragma unroll
            for (int c = 0; c < K; ++c) env[h][c] = F2{s0[h] * s0[c], s0[h] * s0[c]};
        float xna = __ldg(xa), xnb = __ldg(xb);
        for (int q = 0; q < n; ++q) {
            float pa[K], pb[K];
            hermite_phi<K>(xna, hw, pa);
            hermite_phi<K>(xnb, hw, pb);
            if (q + 1 < n) xna = __ldg(xa + (q + 1) * xs_q), xnb = __ldg(xb + (q + 1) * xs_q);   // next qubit's x in flight
            const float is = inv_s[q];
#pragma unroll
            for (int i = 0; i < K; ++i)
#pragma unroll
                for (int j = 0; j < K; ++j) M[i][j] = F2{pa[i] * pa[j] * is, pb[i] * pb[j] * is};
            if (q < n - 1) chain_step2<K>(env, Ls + q * K3, M);
        }
        F2 val{0.f, 0.f};                              // "acd,adc->a"
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int d = 0; d < K; ++d) val = fma2(env[c][d], M[d][c], val);
        if (ba < B) values[ba] = val.lo;
        if (bb < B) values[bb] = val.hi;
    }
}

// Training / seeded reverse sweep with TWO samples per thread (MODE 1 / 2, batch-contiguous
// measurements, K <= 3): the same packed arithmetic as tnq_chain_fwd2_kernel for the forward sweep
// (left environments taped in local memory) and for the reverse sweep, whose per-h slices keep the
// working set in registers.  A thread's two gradient contributions are added before the warp
// reduction, so the shuffle tree runs once per 64 samples instead of once per 32.
template <int K, int MODE>
__global__ void __launch_bounds__(CHAIN_THREADS)
tnq_chain_train2_kernel(const __grid_constant__ ChainArgs a, long long B, const float* __restrict__ seed,
                        float* __restrict__ values, float* __restrict__ partials, float log_scale, float inv_count) {
    constexpr int K3 = K * K * K;
    extern __shared__ float sm[];
    const int n = a.n;
    float* Ls = sm;
    const int warps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int SLAB = 64 * ((K * K) | 1);
    float* slab = sm + (n - 1) * K3 + warp * SLAB;
    float* wacc = sm + (n - 1) * K3 + warps * SLAB;   // [warps][(n-1) * K3 + 1]
    const int acc_stride = (n - 1) * K3 + 1;
    float* rscr = wacc + warps * acc_stride + warp * (K3 * 33);   // per-warp lane-reduction scratch [K3][33]
    for (int i = threadIdx.x; i < (n - 1) * K3; i += blockDim.x) {
        const int q = i / K3, r = i % K3, c = r / (K * K), e = (r / K) % K, f = r % K;
        float s = 0.f;
        for (int d = 0; d < K; ++d) s = fmaf(__ldg(a.core[q] + ((c * K + d) * K + e) * K + f), __ldg(a.state[q + 1] + d), s);
        Ls[i] = s;
    }
    for (int i = threadIdx.x; i < warps * acc_stride; i += blockDim.x) wacc[i] = 0.f;
    __syncthreads();
    float s0[K];
#pragma unroll
    for (int i = 0; i < K; ++i) s0[i] = __ldg(a.state[0] + i);
    float* my_acc = wacc + warp * acc_stride;
    const long long ntiles = (B + 63) / 64;
    for (long long wi = (long long)blockIdx.x * warps + warp; wi < ntiles; wi += (long long)gridDim.x * warps) {
        const long long b0 = wi * 64;
        const bool vlo = b0 + lane < B, vhi = b0 + 32 + lane < B;
        F2 env[K][K], M[K][K];
#pragma unroll
        for (int h = 0; h < K; ++h)
#pragma unroll
            for (int c = 0; c < K; ++c) env[h][c] = F2{s0[h] * s0[c], s0[h] * s0[c]};
        F2 tape[MAXQ][K][K];
        float raw[2 * K * K];
        fetch_m2<K>(raw, a.mx[0], b0, B, lane);
        unpack_m2<K>(M, raw, slab, lane);
        for (int q = 0; q < n - 1; ++q) {
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int c = 0; c < K; ++c) tape[q][h][c] = env[h][c];
            fetch_m2<K>(raw, a.mx[q + 1], b0, B, lane);
            chain_step2<K>(env, Ls + q * K3, M);
            unpack_m2<K>(M, raw, slab, lane);
        }
        F2 val{0.f, 0.f};
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int d = 0; d < K; ++d) val = fma2(env[c][d], M[d][c], val);
        if (MODE != 2 && values != nullptr) {
            if (vlo) values[b0 + lane] = val.lo;
            if (vhi) values[b0 + 32 + lane] = val.hi;
        }
        // ---- seed ----
        F2 dval;
        if (MODE == 1) {
            const float clo = fmaxf(val.lo, 1e-10f), chi = fmaxf(val.hi, 1e-10f);
            float lpart = (vlo ? -(logf(clo) + log_scale) * inv_count : 0.f) + (vhi ? -(logf(chi) + log_scale) * inv_count : 0.f);
            lpart = warp_sum(lpart);
            if (lane == 0) my_acc[acc_stride - 1] += lpart;
            dval = F2{(vlo && val.lo >= 1e-10f) ? -inv_count / clo : 0.f, (vhi && val.hi >= 1e-10f) ? -inv_count / chi : 0.f};
        } else {
            dval = F2{vlo ? __ldg(seed + b0 + lane) : 0.f, vhi ? __ldg(seed + b0 + 32 + lane) : 0.f};
        }
        F2 denv[K][K];
#pragma unroll
        for (int c = 0; c < K; ++c)
#pragma unroll
            for (int d = 0; d < K; ++d) denv[c][d] = F2{dval.lo * M[d][c].lo, dval.hi * M[d][c].hi};
        // ---- reverse sweep ----
        fetch_m2<K>(raw, a.mx[n - 2], b0, B, lane);
        for (int q = n - 2; q >= 0; --q) {
            const float* L = Ls + q * K3;
            unpack_m2<K>(M, raw, slab, lane);
            if (q > 0) fetch_m2<K>(raw, a.mx[q - 1], b0, B, lane);
            float* gq = my_acc + q * K3;
            F2 accL[K][K][K];    // d Ls[c][e][f] of this thread's two samples: L role sum_h e0[h][c] dT1[h][e][f]
                                 // plus R role sum_f T2[h][g][f] denv[j][f] at [h][g][j]
            F2 dnew[K][K];       // [h][c]
#pragma unroll
            for (int c = 0; c < K; ++c)
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f) accL[c][e][f] = F2{0.f, 0.f};
#pragma unroll
            for (int h = 0; h < K; ++h) {
                F2 e0[K];
#pragma unroll
                for (int c = 0; c < K; ++c) e0[c] = tape[q][h][c];
                F2 T1[K][K], T2[K][K];   // [e][f], [g][f]
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        F2 t{0.f, 0.f};
#pragma unroll
                        for (int c = 0; c < K; ++c) t = fma2(e0[c], L[(c * K + e) * K + f], t);
                        T1[e][f] = t;
                    }
#pragma unroll
                for (int g = 0; g < K; ++g)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        F2 t{0.f, 0.f};
#pragma unroll
                        for (int e = 0; e < K; ++e) t = fma2(T1[e][f], M[e][g], t);
                        T2[g][f] = t;
                    }
                // R role: dLs[h][g][j] += sum_f T2[g][f] denv[j][f] ; dT2[g][f] = sum_j denv[j][f] Ls[h][g][j]
#pragma unroll
                for (int g = 0; g < K; ++g) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {   // both roles add into dLs[.][.][.]: one reduction for the two
#pragma unroll
                        for (int f = 0; f < K; ++f) accL[h][g][j] = fma2(T2[g][f], denv[j][f], accL[h][g][j]);
                    }
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        F2 t{0.f, 0.f};
#pragma unroll
                        for (int j = 0; j < K; ++j) t = fma2(denv[j][f], L[(h * K + g) * K + j], t);
                        T2[g][f] = t;                      // now dT2
                    }
                }
                // dT1[e][f] = sum_g dT2[g][f] M[e][g] ; L role: accL[c][e][f] += e0[c] dT1[e][f] ;
                // dnew[h][c] = sum_{e,f} dT1[e][f] Ls[c][e][f]
#pragma unroll
                for (int c = 0; c < K; ++c) dnew[h][c] = F2{0.f, 0.f};
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f) {
                        F2 d1{0.f, 0.f};
#pragma unroll
                        for (int g = 0; g < K; ++g) d1 = fma2(T2[g][f], M[e][g], d1);
#pragma unroll
                        for (int c = 0; c < K; ++c) {
                            accL[c][e][f] = fma2(e0[c], d1, accL[c][e][f]);
                            dnew[h][c] = fma2(d1, L[(c * K + e) * K + f], dnew[h][c]);
                        }
                    }
            }
            // lane reduction through shared memory: [K3][33] transposed, lane v sums row v in lane order
            // (3 shared-memory instructions per entry instead of a 5-step shuffle tree)
#pragma unroll
            for (int c = 0; c < K; ++c)
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int f = 0; f < K; ++f)
                        rscr[((c * K + e) * K + f) * 33 + lane] = accL[c][e][f].lo + accL[c][e][f].hi;
            __syncwarp();
            if (lane < K3) {
                float p0 = 0.f, p1 = 0.f, p2 = 0.f, p3 = 0.f;
#pragma unroll
                for (int l = 0; l < 32; l += 4) {
                    p0 += rscr[lane * 33 + l];
                    p1 += rscr[lane * 33 + l + 1];
                    p2 += rscr[lane * 33 + l + 2];
                    p3 += rscr[lane * 33 + l + 3];
                }
                gq[lane] += (p0 + p1) + (p2 + p3);
            }
            __syncwarp();
#pragma unroll
            for (int h = 0; h < K; ++h)
#pragma unroll
                for (int c = 0; c < K; ++c) denv[h][c] = dnew[h][c];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < acc_stride; i += blockDim.x) {
        float sacc = 0.f;
        for (int w = 0; w < warps; ++w) sacc += wacc[w * acc_stride + i];
        partials[(size_t)blockIdx.x * acc_stride + i] = sacc;
    }
}

// dG_q[c,d,e,f] = dLs_q[c,e,f] * s_{q+1}[d] ; loss
template <int K>
__global__ void tnq_chain_finalize_kernel(const __grid_constant__ ChainArgs a, const float* __restrict__ partials, int nparts,
                                          float* __restrict__ loss) {
    constexpr int K3 = K * K * K;
    const int acc_stride = (a.n - 1) * K3 + 1;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= acc_stride) return;
    float s = 0.f;
    for (int p = 0; p < nparts; ++p) s += partials[(size_t)p * acc_stride + i];
    if (i == acc_stride - 1) {
        if (loss) *loss = s;
        return;
    }
    const int q = i / K3, r = i % K3, c = r / (K * K), e = (r / K) % K, f = r % K;
    for (int d = 0; d < K; ++d) a.grad[q][((c * K + d) * K + e) * K + f] = s * __ldg(a.state[q + 1] + d);
}

template <int K>
int launch_chain(const ChainArgs& a, long long B, int mode, const float* seed, float* values, float* loss,
                 float log_scale, void* workspace, long long ws_bytes, cudaStream_t st) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int K3 = K * K * K, warps = CHAIN_THREADS / 32;
    const int acc_stride = (a.n - 1) * K3 + 1;
    const size_t smem = sizeof(float) * ((size_t)(a.n - 1) * K3 + (size_t)warps * 32 * ((K * K) | 1) +
                                         (mode ? (size_t)warps * acc_stride : 0));
    long long want = (B + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const long long cap = (long long)sms * (mode ? 4 : 8);
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    float* partials = reinterpret_cast<float*>(workspace);
    if (mode != 0 && (size_t)ws_bytes < sizeof(float) * (size_t)grid * acc_stride)
        return tnq_internal_fail("tnq_mps_chain: workspace too small");
    const float inv = 1.0f / (float)B;
    cudaError_t e = cudaSuccess;
    bool packed = true;
    for (int q = 0; q < a.n; ++q) packed = packed && (a.mx_stride[q] == K * K);
    if (mode == 0 && packed) {   // two samples per thread, packed FFMA2
        const size_t smem2 = sizeof(float) * ((size_t)(a.n - 1) * K3 + (size_t)warps * 64 * ((K * K) | 1));
        long long want2 = (B + 2 * CHAIN_THREADS - 1) / (2 * CHAIN_THREADS);
        const long long cap2 = (long long)sms * 4;
        const int grid2 = (int)(want2 < 1 ? 1 : (want2 > cap2 ? cap2 : want2));
        if (smem2 > 48 * 1024) e = cudaFuncSetAttribute(tnq_chain_fwd2_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
        if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(chain fwd2)");
        tnq_chain_fwd2_kernel<K><<<grid2, CHAIN_THREADS, smem2, st>>>(a, B, values);
        tnq_internal_count_launch();
        e = cudaGetLastError();
        if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_chain launch");
        return 0;
    }
    if constexpr (K <= 3) {
        if (mode != 0 && packed) {   // two samples per thread, packed FFMA2 (forward tape + reverse sweep)
            const size_t smem2 = sizeof(float) * ((size_t)(a.n - 1) * K3 + (size_t)warps * 64 * ((K * K) | 1) +
                                                  (size_t)warps * acc_stride + (size_t)warps * K3 * 33);
            long long want2 = (B + 2 * CHAIN_THREADS - 1) / (2 * CHAIN_THREADS);
            const int grid2 = (int)(want2 < 1 ? 1 : (want2 > cap ? cap : want2));   // same bound as the workspace
            if (mode == 1) {
                if (smem2 > 48 * 1024) e = cudaFuncSetAttribute(tnq_chain_train2_kernel<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                tnq_chain_train2_kernel<K, 1><<<grid2, CHAIN_THREADS, smem2, st>>>(a, B, seed, values, partials, log_scale, inv);
            } else {
                if (smem2 > 48 * 1024) e = cudaFuncSetAttribute(tnq_chain_train2_kernel<K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2);
                tnq_chain_train2_kernel<K, 2><<<grid2, CHAIN_THREADS, smem2, st>>>(a, B, seed, values, partials, log_scale, inv);
            }
            if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(chain train2)");
            tnq_internal_count_launch();
            tnq_chain_finalize_kernel<K><<<(acc_stride + 127) / 128, 128, 0, st>>>(a, partials, grid2, mode == 1 ? loss : nullptr);
            tnq_internal_count_launch();
            e = cudaGetLastError();
            if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_chain launch");
            return 0;
        }
    }
    if (mode == 0) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(tnq_chain_kernel<K, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tnq_chain_kernel<K, 0><<<grid, CHAIN_THREADS, smem, st>>>(a, B, seed, values, partials, log_scale, inv);
    } else if (mode == 1) {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(tnq_chain_kernel<K, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tnq_chain_kernel<K, 1><<<grid, CHAIN_THREADS, smem, st>>>(a, B, seed, values, partials, log_scale, inv);
    } else {
        if (smem > 48 * 1024) e = cudaFuncSetAttribute(tnq_chain_kernel<K, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        tnq_chain_kernel<K, 2><<<grid, CHAIN_THREADS, smem, st>>>(a, B, seed, values, partials, log_scale, inv);
    }
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(chain)");
    tnq_internal_count_launch();
    if (mode != 0) {
        tnq_chain_finalize_kernel<K><<<(acc_stride + 127) / 128, 128, 0, st>>>(a, partials, grid, mode == 1 ? loss : nullptr);
        tnq_internal_count_launch();
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_chain launch");
    return 0;
}


// ---- sample(): qubit-by-qubit inverse-CDF sampling with PREFIX ENVIRONMENTS ------------------------------
// Reference: EngineSiamese.sample (tneq_qc/core/engine_siamese.py:740-915) runs, for every qubit q, one full forward
// at batch num_samples x grid_size (grid expansion :802-822, forward :842-847, inverse CDF :855-905): the work on
// the qubits < q (already sampled) and > q (identity measurements) is redone for every grid point and every q.
// Here one thread owns one sample and walks the chain ONCE:
//   * left environment env (K x K, registers): the sweep over the sampled qubits, advanced by one chain_step per qubit;
//   * right environments R_q (K x K, shared by all samples: the later qubits are measured with the identity), built
//     once per CTA from the back of the chain and stored folded with the core, LR_q[h,g,f] = sum_j Ls_q[h,g,j] R_{q+1}[j,f];
//   * the value is linear in the measurement matrix of qubit q: value(M) = sum_{e,g} C[e,g] M[e,g] with
//     C[e,g] = sum_{h,c,f} env[h,c] Ls_q[c,e,f] LR_q[h,g,f] (2 K^4 multiply-adds), so a grid point costs K^2
//     multiply-adds instead of a contraction of the whole network;
//   * density = value clamped at 0 (abs_square is the identity for real dtypes, then clamp, :858-862), running sum =
//     cumsum, divided by (total + 1e-10),
//     index = #(cdf < u) clamped to G - 2, linear interpolation between the grid points idx and idx + 1 exactly as
//     :884-901 (including its extrapolation quirk: cdf[idx] is the first value NOT below u);
//   * the sampled value's matrix phi(y) phi(y)^T (generate_data, :133-254) is formed in registers.
// The uniform numbers are drawn by the caller in the reference's order (one (S,1) draw per qubit).
template <int K>
__global__ void __launch_bounds__(CHAIN_THREADS)
tnq_chain_sample_kernel(const __grid_constant__ ChainArgs a, long long S, int G, const float* __restrict__ grid_x,
                        const float* __restrict__ mx_grid, int grid_in_smem, const float* __restrict__ u, HermiteW hw,
                        float* __restrict__ samples) {
    constexpr int K2 = K * K, K3 = K * K * K;
    extern __shared__ float sm[];
    const int n = a.n;
    float* Ls = sm;                                   // [n-1][K3]
    float* LR = Ls + (n - 1) * K3;                    // [n-1][K3]: LR_q[h][g][f]
    float* R = LR + (n - 1) * K3;                     // [K2] scratch: R_{q+1}[j][f]
    float* mg_s = R + K2;                             // [G][K2] when it fits
    for (int i = threadIdx.x; i < (n - 1) * K3; i += blockDim.x) {
        const int q = i / K3, r = i % K3, c = r / K2, e = (r / K) % K, f = r % K;
        float s = 0.f;
        for (int d = 0; d < K; ++d) s = fmaf(__ldg(a.core[q] + ((c * K + d) * K + e) * K + f), __ldg(a.state[q + 1] + d), s);
        Ls[i] = s;
    }
    if (threadIdx.x < K2) R[threadIdx.x] = (threadIdx.x / K == threadIdx.x % K) ? 1.f : 0.f;     // R_{n-1} = identity
    if (grid_in_smem)
        for (int i = threadIdx.x; i < G * K2; i += blockDim.x) mg_s[i] = __ldg(mx_grid + i);
    __syncthreads();
    for (int q = n - 2; q >= 0; --q) {
        const float* L = Ls + q * K3;
        if (threadIdx.x < K3) {                       // LR_q[h][g][f] = sum_j Ls_q[h][g][j] R_{q+1}[j][f]
            const int hg = threadIdx.x / K, f = threadIdx.x % K;
            float s = 0.f;
            for (int j = 0; j < K; ++j) s = fmaf(L[hg * K + j], R[j * K + f], s);
            LR[q * K3 + threadIdx.x] = s;
        }
        __syncthreads();
        if (threadIdx.x < K2) {                       // R_q[h][c] = sum_{e,f} Ls_q[c][e][f] LR_q[h][e][f]
            const int h = threadIdx.x / K, c = threadIdx.x % K;
            float s = 0.f;
            for (int ef = 0; ef < K2; ++ef) s = fmaf(L[c * K2 + ef], LR[q * K3 + h * K2 + ef], s);
            R[threadIdx.x] = s;
        }
        __syncthreads();
    }
    const float* mg = grid_in_smem ? mg_s : mx_grid;
    float s0[K];
#pragma unroll
    for (int i = 0; i < K; ++i) s0[i] = __ldg(a.state[0] + i);
    for (long long b = (long long)blockIdx.x * blockDim.x + threadIdx.x; b < S; b += (long long)gridDim.x * blockDim.x) {
        float env[K][K];
#pragma unroll
        for (int h = 0; h < K; ++h)
#pragma unroll
            for (int c = 0; c < K; ++c) env[h][c] = s0[h] * s0[c];
        for (int q = 0; q < n; ++q) {
            float C[K][K];                            // value(M) = sum C[e][g] M[e][g]
            if (q < n - 1) {
                const float* L = Ls + q * K3;
                const float* W = LR + q * K3;
                float T1[K][K][K];                    // [h][e][f]
#pragma unroll
                for (int h = 0; h < K; ++h)
#pragma unroll
                    for (int e = 0; e < K; ++e)
#pragma unroll
                        for (int f = 0; f < K; ++f) {
                            float s = 0.f;
#pragma unroll
                            for (int c = 0; c < K; ++c) s = fmaf(env[h][c], L[(c * K + e) * K + f], s);
                            T1[h][e][f] = s;
                        }
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int g = 0; g < K; ++g) {
                        float s = 0.f;
#pragma unroll
                        for (int h = 0; h < K; ++h)
#pragma unroll
                            for (int f = 0; f < K; ++f) s = fmaf(T1[h][e][f], W[(h * K + g) * K + f], s);
                        C[e][g] = s;
                    }
            } else {                                  // "acd,adc->a": value = sum env[c][d] M[d][c]
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int g = 0; g < K; ++g) C[e][g] = env[g][e];
            }
            // pass 1: the total of the densities (the last entry of the cumulative sum)
            float total = 0.f;
#pragma unroll 4
            for (int i = 0; i < G; ++i) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < K2; ++k) v = fmaf(C[k / K][k % K], mg[i * K2 + k], v);
                total += fmaxf(v, 0.f);
            }
            const float denom = total + 1e-10f;
            const float uu = __ldg(u + b * n + q);
            // pass 2: idx = #(cdf < u); c0 = cdf[idx], c1 = cdf[idx + 1] after clamping idx to G - 2
            float run = 0.f, prev = 0.f, cur = 0.f, c0 = 0.f, c1 = 0.f;
            int idx = 0, state = 0;                   // state 0: still below u, 1: c0 taken, 2: c1 taken
            for (int i = 0; i < G; ++i) {
                float v = 0.f;
#pragma unroll
                for (int k = 0; k < K2; ++k) v = fmaf(C[k / K][k % K], mg[i * K2 + k], v);
                run += fmaxf(v, 0.f);
                prev = cur;
                cur = run / denom;
                if (state == 1) c1 = cur, state = 2;
                if (state == 0) {
                    if (cur < uu) ++idx;
                    else c0 = cur, state = 1;
                }
            }
            if (idx > G - 2) idx = G - 2, c0 = prev, c1 = cur;        // (also: first value not below u was the last one)
            const float x0 = __ldg(grid_x + idx), x1 = __ldg(grid_x + idx + 1);
            const float y = x0 + (uu - c0) / (c1 - c0 + 1e-10f) * (x1 - x0);
            samples[b * n + q] = y;
            if (q < n - 1) {
                float phi[K], M[K][K];
                hermite_phi<K>(y, hw, phi);
#pragma unroll
                for (int e = 0; e < K; ++e)
#pragma unroll
                    for (int g = 0; g < K; ++g) M[e][g] = phi[e] * phi[g];
                chain_step<K>(env, Ls + q * K3, M);
            }
        }
    }
}

}  // namespace

extern "C" {

int64_t tnq_mps_chain_workspace_bytes(int K, int n, int64_t B) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return (int64_t)sizeof(float) * (int64_t)sms * 8 * ((int64_t)(n - 1) * K * K * K + 1) + 256;
}

int tnq_mps_chain(int K, int n, const float* const* cores, const float* const* states, const float* const* mx,
                  const int64_t* mx_stride, int64_t B, int mode, const float* seed, float* values, float* loss,
                  float* const* grads, double log_scale, void* workspace, int64_t workspace_bytes, void* stream) {
    if (n < 2 || n > MAXQ) return tnq_internal_fail("tnq_mps_chain: between 2 and " + std::to_string(MAXQ) + " qubits");
    if (K < 2 || K > 4) return tnq_internal_fail("tnq_mps_chain: edge rank must be 2, 3 or 4");
    if (!cores || !states || !mx || !mx_stride || B <= 0 || mode < 0 || mode > 2) return tnq_internal_fail("tnq_mps_chain: bad arguments");
    if (mode != 0 && (!grads || !workspace)) return tnq_internal_fail("tnq_mps_chain: gradients need grads[] and a workspace");
    if (mode == 2 && !seed) return tnq_internal_fail("tnq_mps_chain: mode 2 needs a seed");
    ChainArgs a;
    a.n = n;
    for (int q = 0; q < n; ++q) {
        a.state[q] = states[q];
        a.mx[q] = mx[q];
        a.mx_stride[q] = mx_stride[q];
        a.core[q] = q < n - 1 ? cores[q] : nullptr;
        a.grad[q] = (mode != 0 && q < n - 1) ? grads[q] : nullptr;
        if (!states[q] || !mx[q] || (q < n - 1 && (!cores[q] || (mode != 0 && !grads[q]))))
            return tnq_internal_fail("tnq_mps_chain: null pointer at qubit " + std::to_string(q));
    }
    cudaStream_t st = (cudaStream_t)stream;
    switch (K) {
        case 2: return launch_chain<2>(a, B, mode, seed, values, loss, (float)log_scale, workspace, workspace_bytes, st);
        case 3: return launch_chain<3>(a, B, mode, seed, values, loss, (float)log_scale, workspace, workspace_bytes, st);
        default: return launch_chain<4>(a, B, mode, seed, values, loss, (float)log_scale, workspace, workspace_bytes, st);
    }
}

/* Forward values with the measurement matrices generated in registers from x (opt-in, no reference counterpart
 * as an API: it computes what contract(generate_data(x, ret_type='TNTensor')) computes, engine_siamese.py:133-254).
 *   x: sample b, qubit q at x[b * xs_b + q * xs_q];  weights[K]: the Hermite normalisation weights (:59-80)
 *   scale[n] (device, out): per-qubit TNTensor scale = global abs-max of phi phi^T over the batch
 *   values[B] (out): the contraction of the SCALED matrices; the true value is values * prod_q scale[q]. */
int tnq_mps_chain_x(int K, int n, const float* const* cores, const float* const* states, const float* x, int64_t xs_b,
                    int64_t xs_q, const float* weights, int64_t B, float* scale, float* values, void* stream) {
    if (n < 2 || n > MAXQ) return tnq_internal_fail("tnq_mps_chain_x: between 2 and " + std::to_string(MAXQ) + " qubits");
    if (K < 2 || K > 4) return tnq_internal_fail("tnq_mps_chain_x: edge rank must be 2, 3 or 4");
    if (!cores || !states || !x || !weights || !scale || !values || B <= 0) return tnq_internal_fail("tnq_mps_chain_x: bad arguments");
    ChainArgs a;
    a.n = n;
    for (int q = 0; q < n; ++q) {
        a.state[q] = states[q];
        a.mx[q] = nullptr;
        a.mx_stride[q] = 0;
        a.core[q] = q < n - 1 ? cores[q] : nullptr;
        a.grad[q] = nullptr;
        if (!states[q] || (q < n - 1 && !cores[q])) return tnq_internal_fail("tnq_mps_chain_x: null pointer at qubit " + std::to_string(q));
    }
    HermiteW hw;
    for (int i = 0; i < 4; ++i) hw.w[i] = i < K ? weights[i] : 0.f;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t e = cudaMemsetAsync(scale, 0, sizeof(float) * n, st);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaMemsetAsync(scale)");
    long long wantb = (B + 255) / 256;
    dim3 gs((unsigned)(wantb < 1 ? 1 : (wantb > 4 * sms ? 4 * sms : wantb)), (unsigned)n);
    const int K3 = K * K * K;
    const size_t smem2 = sizeof(float) * ((size_t)(n - 1) * K3 + (size_t)n);
    long long want2 = (B + 2 * CHAIN_THREADS - 1) / (2 * CHAIN_THREADS);
    const long long cap2 = (long long)sms * 4;
    const int grid2 = (int)(want2 < 1 ? 1 : (want2 > cap2 ? cap2 : want2));
#define TNQ_CHAIN_X(KK)                                                                                              \
    tnq_hermite_scale_kernel<KK><<<gs, 256, 0, st>>>(x, xs_b, xs_q, B, n, hw, scale);                                \
    tnq_internal_count_launch();                                                                                     \
    tnq_chain_fwd2x_kernel<KK><<<grid2, CHAIN_THREADS, smem2, st>>>(a, B, x, xs_b, xs_q, hw, scale, values);         \
    tnq_internal_count_launch();
    switch (K) {
        case 2: TNQ_CHAIN_X(2) break;
        case 3: TNQ_CHAIN_X(3) break;
        default: TNQ_CHAIN_X(4) break;
    }
#undef TNQ_CHAIN_X
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_chain_x launch");
    return 0;
}

/* sample() with prefix environments for single-layer MPS networks (see tnq_chain_sample_kernel):
 *   grid_x[G], mx_grid[G][K][K] (the grid's measurement matrices), u[S][n] uniform numbers (qubit q of sample s at
 *   u[s * n + q]), weights[K] (host: generate_data's Hermite weights) -> samples[S][n]. */
int tnq_mps_chain_sample(int K, int n, const float* const* cores, const float* const* states, int64_t S, int G,
                         const float* grid_x, const float* mx_grid, const float* u, const float* weights, float* samples,
                         void* stream) {
    if (n < 2 || n > MAXQ) return tnq_internal_fail("tnq_mps_chain_sample: between 2 and " + std::to_string(MAXQ) + " qubits");
    if (K < 2 || K > 4) return tnq_internal_fail("tnq_mps_chain_sample: edge rank must be 2, 3 or 4");
    if (!cores || !states || !grid_x || !mx_grid || !u || !weights || !samples || S <= 0 || G < 2)
        return tnq_internal_fail("tnq_mps_chain_sample: bad arguments");
    ChainArgs a;
    a.n = n;
    for (int q = 0; q < n; ++q) {
        a.state[q] = states[q];
        a.mx[q] = nullptr;
        a.mx_stride[q] = 0;
        a.core[q] = q < n - 1 ? cores[q] : nullptr;
        a.grad[q] = nullptr;
        if (!states[q] || (q < n - 1 && !cores[q])) return tnq_internal_fail("tnq_mps_chain_sample: null pointer at qubit " + std::to_string(q));
    }
    HermiteW hw;
    for (int i = 0; i < 4; ++i) hw.w[i] = i < K ? weights[i] : 0.f;
    cudaStream_t st = (cudaStream_t)stream;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int K2 = K * K, K3 = K * K * K;
    const size_t base = sizeof(float) * ((size_t)2 * (n - 1) * K3 + K2);
    const size_t gridb = sizeof(float) * (size_t)G * K2;
    const int in_smem = base + gridb <= 160 * 1024 ? 1 : 0;
    const size_t smem = base + (in_smem ? gridb : 0);
    long long want = (S + CHAIN_THREADS - 1) / CHAIN_THREADS;
    const long long cap = (long long)sms * 2;
    const int grid = (int)(want < 1 ? 1 : (want > cap ? cap : want));
    cudaError_t e = cudaSuccess;
#define TNQ_CHAIN_SAMPLE(KK)                                                                                          \
    e = cudaFuncSetAttribute(tnq_chain_sample_kernel<KK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(chain sample)");                     \
    tnq_chain_sample_kernel<KK><<<grid, CHAIN_THREADS, smem, st>>>(a, S, G, grid_x, mx_grid, in_smem, u, hw, samples);
    switch (K) {
        case 2: TNQ_CHAIN_SAMPLE(2) break;
        case 3: TNQ_CHAIN_SAMPLE(3) break;
        default: TNQ_CHAIN_SAMPLE(4) break;
    }
#undef TNQ_CHAIN_SAMPLE
    tnq_internal_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_chain_sample launch");
    return 0;
}

}  // extern "C"
