// tnq_ladder.cu -- the two-layer merged MPS network (QCTN.merge(mps_n, mps_n), BASELINE cfg3) as ONE
// persistent kernel of independent warps: forward sweep, fused loss and reverse sweep of
//   tneq_qc/contractor/greedy_strategy.py:461-598, tneq_qc/core/engine_siamese.py:490-530,
//   tneq_qc/backends/backend_pytorch.py:153-158
// for SPW = 32 / K^2 samples per warp.  The arithmetic (phases A, B, C and their adjoints) lives in
// tnq_ladder_core.cuh; this file holds the kernels, the launch geometry and the C ABI.
//
// Per launch: tnq_ladder_kernel (everything per sample) + tnq_ladder_finalize_kernel (sum of the
// per-warp gradient slices in warp order, circuit states folded back in, loss).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "tneq_b200.h"
#include "tnq_ladder_core.cuh"

extern int tnq_internal_fail(const std::string& msg);
extern int tnq_internal_cuda_fail(cudaError_t e, const char* what);
extern void tnq_internal_count_launch();

namespace {

using namespace tnq_ladder;

constexpr int WARPS_FWD = 16;     // warps per CTA, forward only (one CTA per SM, 128 registers per thread)
constexpr int WARPS_TRAIN = 8;    // with the reverse sweep: registers are allocated in units of 4 warps (8 x 32 x 255 <= 64 K)

template <int K>
__host__ __device__ constexpr int cst_floats(int n) {
    return ((n - 1) * Dims<K>::CSTEP + Dims<K>::K2 + 3) / 4 * 4;
}

template <int K, int MODE>
__global__ void __maxnreg__(MODE == 0 ? 128 : 255)
tnq_ladder_kernel(const __grid_constant__ Args a, long long B, long long ngroups, const float* __restrict__ seed,
                  float* __restrict__ values, float* __restrict__ gparts, float* __restrict__ lparts,
                  float* __restrict__ ckpt, float log_scale, float inv_count) {
    using D = Dims<K>;
    extern __shared__ __align__(16) float sm[];
    const int n = a.n;
    const int ncst = (n - 1) * D::CSTEP + D::K2;
    for (int i = threadIdx.x; i < ncst; i += blockDim.x) sm[i] = const_pool_element<K>(a, i);
    constexpr int WSZ = MODE == 0 ? D::WARP_FWD : D::WARP_TRAIN;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, wpc = blockDim.x >> 5;
    float* wb = sm + cst_floats<K>(n) + warp * WSZ;
    for (int i = lane; i < WSZ; i += 32) wb[i] = 0.f;
    __syncthreads();
    // warp numbering: consecutive groups go to different SMs first
    const long long wg = (long long)warp * gridDim.x + blockIdx.x;
    const long long wtotal = (long long)wpc * gridDim.x;
    if (wg >= ngroups) return;
    WarpCtx<K> c;
    c.cst = sm;
    c.E = wb;
    c.U = c.E + D::E_SZ;
    c.T2 = c.U + D::U_SZ;
    c.M = c.T2 + D::T2_SZ;
    c.V = c.M + D::M_SZ;
    c.D = c.V + D::V_SZ;
    c.dT2 = c.D + D::E_SZ;
    c.args = &a;
    c.B = B;
    c.seed = seed;
    c.values = values;
    c.log_scale = log_scale;
    c.inv_count = inv_count;
    c.ckE = nullptr, c.ckT2 = nullptr, c.gpart = nullptr;
    if (MODE != 0) {
        const int ng = D::grad_floats(n);
        c.gpart = gparts + (size_t)wg * ng;
        for (int i = lane; i < ng; i += 32) c.gpart[i] = 0.f;
        c.ckE = ckpt + (size_t)wg * D::ckpt_floats(n);
        c.ckT2 = c.ckE + (size_t)(n - 2) * D::E_SZ;
        __syncwarp();
    }
    LaneState<K> st;
    st.loss = 0.f;
    st.mnext = 0.f;
    for (long long g = wg; g < ngroups; g += wtotal) ladder_group<K, MODE>(c, st, lane, g * D::SPW);
    if (MODE == 1) lparts[wg * 32 + lane] = st.loss;
}

// One block = 32 consecutive elements of the compact gradient slice x 8 partial sums over the
// warps; the partial sums are combined in a fixed order, then expanded into the core gradients.
template <int K>
__global__ void __launch_bounds__(256)
tnq_ladder_finalize_kernel(const __grid_constant__ Args a, const float* __restrict__ gparts,
                           const float* __restrict__ lparts, int nwarps, float* __restrict__ loss) {
    using D = Dims<K>;
    __shared__ float part[8][33];
    const int n = a.n, ng = D::grad_floats(n);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nblk = (ng + 31) / 32;
    if ((int)blockIdx.x == nblk) {                 // loss: nwarps * 32 lane partials
        float t = 0.f;
        for (int i = threadIdx.x; i < nwarps * 32; i += 256) t += lparts[i];
        part[w][lane] = t;
        __syncthreads();
        if (threadIdx.x == 0 && loss != nullptr) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i)
                for (int l = 0; l < 32; ++l) s += part[i][l];
            *loss = s;
        }
        return;
    }
    const int e = blockIdx.x * 32 + lane;
    const int per = (nwarps + 7) / 8;
    const int w0 = w * per, w1 = min(nwarps, w0 + per);
    float t = 0.f;
    if (e < ng)
        for (int i = w0; i < w1; ++i) t += gparts[(size_t)i * ng + e];
    part[w][lane] = t;
    __syncthreads();
    if (w != 0 || e >= ng) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][lane];
    const int nX = (n - 1) * D::K4, nB = (n - 1) * D::K3;
    if (e < nX) {
        a.gradX[e / D::K4][e % D::K4] = s;
    } else if (e < nX + nB) {
        const int q = (e - nX) / D::K3, r = (e - nX) % D::K3;
        if (q == 0) return;                        // slot unused: A_0 is folded with both states (below)
        const int cc = r / D::K2, ee = (r / K) % K, f = r % K;
        for (int d = 0; d < K; ++d) a.gradA[q][((cc * K + d) * K + ee) * K + f] = s * __ldg(a.state[q + 1] + d);
    } else if (e < nX + nB + D::K2) {                // (the slice is padded to a multiple of 4 floats)
        const int ef = e - nX - nB;
        for (int cc = 0; cc < K; ++cc)
            for (int d = 0; d < K; ++d)
                a.gradA[0][(cc * K + d) * D::K2 + ef] = s * __ldg(a.state[0] + cc) * __ldg(a.state[1] + d);
    }
}

struct Geometry {
    int grid, wpc;
    long long ngroups, nwarps;
    size_t smem;
};

template <int K>
Geometry geometry(int n, long long B, int mode, int sms) {
    using D = Dims<K>;
    Geometry g;
    g.ngroups = (B + D::SPW - 1) / D::SPW;
    int wmax = mode == 0 ? WARPS_FWD : WARPS_TRAIN;
    {   // long chains have a bigger constant pool: fewer warps fit next to it
        int dev = 0, smem_max = 227 * 1024;
        if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        const long long per_warp = (long long)sizeof(float) * (mode == 0 ? D::WARP_FWD : D::WARP_TRAIN);
        const long long fit = ((long long)smem_max - (long long)sizeof(float) * cst_floats<K>(n)) / per_warp;
        if (fit < wmax) wmax = fit < 1 ? 1 : (int)fit;
    }
    // as many warps per CTA as fit (more resident warps hide more latency; measured: trimming the
    // warp count to balance the passes is slower), small batches spread over all SMs
    long long wpc = (g.ngroups + sms - 1) / sms;
    g.wpc = (int)(wpc < 1 ? 1 : (wpc > wmax ? wmax : wpc));
    long long grid = (g.ngroups + g.wpc - 1) / g.wpc;
    g.grid = (int)(grid > sms ? sms : grid);
    g.nwarps = (long long)g.grid * g.wpc;
    if (g.nwarps > g.ngroups) g.nwarps = g.ngroups;    // warps with index >= ngroups exit at once
    g.smem = sizeof(float) * ((size_t)cst_floats<K>(n) + (size_t)g.wpc * (mode == 0 ? D::WARP_FWD : D::WARP_TRAIN));
    return g;
}

template <int K>
size_t workspace_bytes(int n, long long B, int mode, int sms) {
    using D = Dims<K>;
    if (mode == 0) return 256;
    const Geometry g = geometry<K>(n, B, mode, sms);
    const size_t nw = (size_t)g.grid * g.wpc;
    return sizeof(float) * nw * ((size_t)D::grad_floats(n) + 32 + (size_t)D::ckpt_floats(n)) + 256;
}

template <int K, int MODE>
int launch_mode(const Args& a, const Geometry& g, long long B, const float* seed, float* values, float* gparts,
                float* lparts, float* ckpt, float log_scale, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(tnq_ladder_kernel<K, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(ladder)");
    tnq_ladder_kernel<K, MODE><<<g.grid, g.wpc * 32, g.smem, st>>>(a, B, g.ngroups, seed, values, gparts, lparts, ckpt,
                                                                    log_scale, 1.0f / (float)B);
    tnq_internal_count_launch();
    return 0;
}

template <int K>
int launch_ladder(const Args& a, long long B, int mode, const float* seed, float* values, float* loss, float log_scale,
                  void* workspace, long long ws_bytes, cudaStream_t st) {
    using D = Dims<K>;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const Geometry g = geometry<K>(a.n, B, mode, sms);
    float *gparts = nullptr, *lparts = nullptr, *ckpt = nullptr;
    if (mode != 0) {
        if ((size_t)ws_bytes < workspace_bytes<K>(a.n, B, mode, sms)) return tnq_internal_fail("tnq_mps_ladder: workspace too small");
        const size_t nw = (size_t)g.grid * g.wpc;
        gparts = reinterpret_cast<float*>(workspace);
        lparts = gparts + nw * D::grad_floats(a.n);
        ckpt = lparts + nw * 32;
    }
    int rc;
    if (mode == 0)
        rc = launch_mode<K, 0>(a, g, B, seed, values, gparts, lparts, ckpt, log_scale, st);
    else if (mode == 1)
        rc = launch_mode<K, 1>(a, g, B, seed, values, gparts, lparts, ckpt, log_scale, st);
    else
        rc = launch_mode<K, 2>(a, g, B, seed, values, gparts, lparts, ckpt, log_scale, st);
    if (rc) return rc;
    if (mode != 0) {
        const int nblk = (D::grad_floats(a.n) + 31) / 32;
        tnq_ladder_finalize_kernel<K><<<nblk + (mode == 1 ? 1 : 0), 256, 0, st>>>(a, gparts, lparts, (int)g.nwarps,
                                                                                  mode == 1 ? loss : nullptr);
        tnq_internal_count_launch();
    }
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_ladder launch");
    return 0;
}

}  // namespace

extern "C" {

// TNQ_LADDER_V2=1 routes edge rank 3 to the second-generation kernel (tnq_ladder2.cu): same results, 55x less HBM
// traffic, equal speed at 2048 samples per GPU, 15-20 % slower at 16384: opt-in until it is faster.
static bool use_second_generation(int K) {
    const char* v = getenv("TNQ_LADDER_V2");
    return K == 3 && v != nullptr && v[0] == '1';
}

int64_t tnq_mps_ladder_workspace_bytes(int K, int n, int64_t B, int mode) {
    if (use_second_generation(K)) return tnq_mps_ladder2_workspace_bytes(n, B, mode);
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (n < 3 || n > MAXQ || B <= 0) return 256;
    return (int64_t)(K == 2 ? workspace_bytes<2>(n, B, mode, sms) : workspace_bytes<3>(n, B, mode, sms));
}

int tnq_mps_ladder(int K, int n, const float* const* cores_a, const float* const* cores_x, const float* const* states,
                   const float* const* mx, const int64_t* mx_stride, int64_t B, int mode, const float* seed,
                   float* values, float* loss, float* const* grads_a, float* const* grads_x, double log_scale,
                   void* workspace, int64_t workspace_bytes, void* stream) {
    // edge rank 3 with TNQ_LADDER_V2=1: the second-generation kernel (tnq_ladder2.cu)
    if (use_second_generation(K))
        return tnq_mps_ladder2(n, cores_a, cores_x, states, mx, mx_stride, B, mode, seed, values, loss, grads_a, grads_x,
                               log_scale, workspace, workspace_bytes, stream);
    if (n < 3 || n > MAXQ) return tnq_internal_fail("tnq_mps_ladder: between 3 and " + std::to_string(MAXQ) + " qubits");
    if (K != 2 && K != 3) return tnq_internal_fail("tnq_mps_ladder: edge rank must be 2 or 3");
    if (!cores_a || !cores_x || !states || !mx || !mx_stride || B <= 0 || mode < 0 || mode > 2)
        return tnq_internal_fail("tnq_mps_ladder: bad arguments");
    if (mode != 0 && (!grads_a || !grads_x || !workspace)) return tnq_internal_fail("tnq_mps_ladder: gradients need grads[] and a workspace");
    if (mode == 0 && !values) return tnq_internal_fail("tnq_mps_ladder: mode 0 needs values");
    if (mode == 1 && !loss) return tnq_internal_fail("tnq_mps_ladder: mode 1 needs loss");
    if (mode == 2 && !seed) return tnq_internal_fail("tnq_mps_ladder: mode 2 needs a seed");
    Args a;
    a.n = n;
    for (int q = 0; q < MAXQ; ++q) {
        const bool hq = q < n, hc = q < n - 1;
        a.state[q] = hq ? states[q] : nullptr;
        a.mx[q] = hq ? mx[q] : nullptr;
        a.mx_stride[q] = hq ? mx_stride[q] : 0;
        a.coreA[q] = hc ? cores_a[q] : nullptr;
        a.coreX[q] = hc ? cores_x[q] : nullptr;
        a.gradA[q] = (hc && mode != 0) ? grads_a[q] : nullptr;
        a.gradX[q] = (hc && mode != 0) ? grads_x[q] : nullptr;
        if (hq && (!states[q] || !mx[q])) return tnq_internal_fail("tnq_mps_ladder: null pointer at qubit " + std::to_string(q));
        if (hc && (!cores_a[q] || !cores_x[q] || (mode != 0 && (!grads_a[q] || !grads_x[q]))))
            return tnq_internal_fail("tnq_mps_ladder: null core pointer at qubit " + std::to_string(q));
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (K == 2) return launch_ladder<2>(a, B, mode, seed, values, loss, (float)log_scale, workspace, workspace_bytes, st);
    return launch_ladder<3>(a, B, mode, seed, values, loss, (float)log_scale, workspace, workspace_bytes, st);
}

}  // extern "C"
