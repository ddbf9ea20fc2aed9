// tnq_sgdg.cu -- the Stiefel-manifold SGD step (Cayley transform) of the reference optimizer for ALL
// cores of a network in one launch: one CTA per core, every matrix in shared memory.
//
// Reference: tneq_qc/backends/backend_pytorch.py:349-468 (_sgdg_step), per core ~20 small torch
// kernels (norm, 4 mm, inverse, ...); for the 46 cores of the 24-qubit two-layer network that is
// ~1000 launches per optimizer step, far more than the fused contraction itself.
//
// Per core (p viewed as d x D, d = product of the first half of its dims; Stiefel case d <= D):
//   X  = p / (||p||_row + 1e-8)                              V  = momentum * V - g^T
//   MX = V X      XMX = X MX      W^ = MX - 1/2 X^T XMX      W = W^ - W^T
//   alpha = min(lr, 1 / (||W||_1 + 1e-8))                    (||.||_1 = max column abs sum)
//   Y  = (I - alpha/2 W)^-1 (I + alpha/2 W) X^T              p_new = Y^T        V_new = W X^T
// The 1 % random QR retraction of the reference (backend_pytorch.py:382) is decided on the host.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "tneq_b200.h"

extern int tnq_internal_fail(const std::string& msg);
extern int tnq_internal_cuda_fail(cudaError_t e, const char* what);
extern void tnq_internal_count_launch();

namespace {

constexpr int SG_THREADS = 128;

// C[m x n] = A[m x k] * B[k x n], all row major in shared memory with leading dimension ld
__device__ __forceinline__ void mm(float* C, const float* A, const float* B, int m, int k, int n, int ld, bool at = false,
                                   bool bt = false) {
    for (int i = threadIdx.x; i < m * n; i += blockDim.x) {
        const int r = i / n, c = i % n;
        float s = 0.f;
        for (int j = 0; j < k; ++j) s = fmaf(at ? A[j * ld + r] : A[r * ld + j], bt ? B[c * ld + j] : B[j * ld + c], s);
        C[r * ld + c] = s;
    }
    __syncthreads();
}

__global__ void __launch_bounds__(SG_THREADS)
tnq_sgdg_kernel(float* const* __restrict__ params, const float* const* __restrict__ grads, float* const* __restrict__ vel,
                const int* __restrict__ rows, const int* __restrict__ cols, int maxd, float lr, float momentum,
                float* pbase, const float* gbase, float* vbase, const long long* __restrict__ offs) {
    extern __shared__ float sm[];
    const int d = rows[blockIdx.x], D = cols[blockIdx.x];      // p is d x D with d <= D; W is D x D
    const int ld = maxd + 1, sz = maxd * ld;
    float *X = sm, *V = sm + sz, *MX = sm + 2 * sz, *T = sm + 3 * sz, *W = sm + 4 * sz, *L = sm + 5 * sz, *R = sm + 6 * sz;
    __shared__ float red[SG_THREADS];
    __shared__ float s_alpha;
    __shared__ int piv;
    // flat form: three base pointers + a static table of element offsets [3][ncores] (params | grads | velocity)
    float* p = offs ? pbase + offs[blockIdx.x] : params[blockIdx.x];
    const float* g = offs ? gbase + offs[gridDim.x + blockIdx.x] : grads[blockIdx.x];
    float* v = offs ? vbase + offs[2 * gridDim.x + blockIdx.x] : vel[blockIdx.x];
    const float eps = 1e-8f;
    // X (d x D) = row-normalised p ; V (D x d) = momentum * V - g^T
    for (int r = threadIdx.x; r < d; r += blockDim.x) {
        float nrm = 0.f;
        for (int c = 0; c < D; ++c) nrm = fmaf(p[r * D + c], p[r * D + c], nrm);
        nrm = sqrtf(nrm) + eps;
        for (int c = 0; c < D; ++c) X[r * ld + c] = p[r * D + c] / nrm;
    }
    for (int i = threadIdx.x; i < D * d; i += blockDim.x) {
        const int r = i / d, c = i % d;
        V[r * ld + c] = momentum * v[i] - g[c * D + r];
    }
    __syncthreads();
    mm(MX, V, X, D, d, D, ld);                 // MX   = V X        (D x D)
    mm(T, X, MX, d, D, D, ld);                 // XMX  = X MX       (d x D)
    mm(W, X, T, D, d, D, ld, /*at=*/true);     // XXMX = X^T XMX    (D x D, into W temporarily)
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
        const int r = i / D, c = i % D;
        L[r * ld + c] = MX[r * ld + c] - 0.5f * W[r * ld + c];      // W^
    }
    __syncthreads();
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
        const int r = i / D, c = i % D;
        W[r * ld + c] = L[r * ld + c] - L[c * ld + r];
    }
    __syncthreads();
    // ||W||_1 = max over columns of the absolute column sum
    float best = 0.f;
    for (int c = threadIdx.x; c < D; c += blockDim.x) {
        float s = 0.f;
        for (int r = 0; r < D; ++r) s += fabsf(W[r * ld + c]);
        best = fmaxf(best, s);
    }
    red[threadIdx.x] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int i = 0; i < blockDim.x; ++i) m = fmaxf(m, red[i]);
        s_alpha = fminf(0.5f * 2.f / (m + eps), lr);
    }
    __syncthreads();
    const float ha = 0.5f * s_alpha;
    // L = I - a/2 W  (inverted into R by Gauss-Jordan), MX <- I + a/2 W
    for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
        const int r = i / D, c = i % D;
        const float id = r == c ? 1.f : 0.f;
        L[r * ld + c] = id - ha * W[r * ld + c];
        MX[r * ld + c] = id + ha * W[r * ld + c];
        R[r * ld + c] = id;
    }
    __syncthreads();
    for (int k = 0; k < D; ++k) {              // Gauss-Jordan with partial pivoting
        if (threadIdx.x == 0) {
            int bi = k;
            float bv = fabsf(L[k * ld + k]);
            for (int r = k + 1; r < D; ++r)
                if (fabsf(L[r * ld + k]) > bv) bv = fabsf(L[r * ld + k]), bi = r;
            piv = bi;
        }
        __syncthreads();
        if (piv != k) {
            for (int c = threadIdx.x; c < D; c += blockDim.x) {
                float t0 = L[k * ld + c];
                L[k * ld + c] = L[piv * ld + c];
                L[piv * ld + c] = t0;
                t0 = R[k * ld + c];
                R[k * ld + c] = R[piv * ld + c];
                R[piv * ld + c] = t0;
            }
        }
        __syncthreads();
        const float inv = 1.f / L[k * ld + k];
        __syncthreads();
        for (int c = threadIdx.x; c < D; c += blockDim.x) {
            L[k * ld + c] *= inv;
            R[k * ld + c] *= inv;
        }
        __syncthreads();
        for (int i = threadIdx.x; i < D * D; i += blockDim.x) {
            const int r = i / D, c = i % D;
            if (r == k) continue;
            const float f = L[r * ld + k];
            // column k of L is read by every thread of this sweep: it is cleared afterwards
            if (c != k) L[r * ld + c] = fmaf(-f, L[k * ld + c], L[r * ld + c]);
            R[r * ld + c] = fmaf(-f, R[k * ld + c], R[r * ld + c]);
        }
        __syncthreads();
        for (int r = threadIdx.x; r < D; r += blockDim.x)
            if (r != k) L[r * ld + k] = 0.f;
        __syncthreads();
    }
    mm(T, R, MX, D, D, D, ld);                       // T = L^-1 (I + a/2 W)     (D x D)
    mm(L, T, X, D, D, d, ld, false, /*bt=*/true);    // Y = T X^T                (D x d)
    mm(R, W, X, D, D, d, ld, false, /*bt=*/true);    // V_new = W X^T            (D x d)
    for (int i = threadIdx.x; i < d * D; i += blockDim.x) {
        const int r = i / D, c = i % D;
        p[i] = L[c * ld + r];                        // p_new = Y^T              (d x D)
    }
    for (int i = threadIdx.x; i < D * d; i += blockDim.x) v[i] = R[(i / d) * ld + i % d];
}

}  // namespace

extern "C" int tnq_sgdg_step(float* const* params, const float* const* grads, float* const* velocity, const int* rows,
                             const int* cols, int ncores, int max_cols, float lr, float momentum, void* stream) {
    if (!params || !grads || !velocity || !rows || !cols || ncores <= 0)
        return tnq_internal_fail("tnq_sgdg_step: bad arguments");
    if (max_cols < 1 || max_cols > TNQ_SGDG_MAX_COLS)
        return tnq_internal_fail("tnq_sgdg_step: matrix width must be between 1 and " + std::to_string(TNQ_SGDG_MAX_COLS));
    const int d = max_cols;
    const size_t smem = sizeof(float) * 7 * (size_t)d * (d + 1);
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(tnq_sgdg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(sgdg)");
    tnq_sgdg_kernel<<<ncores, SG_THREADS, smem, (cudaStream_t)stream>>>(params, grads, velocity, rows, cols, d, lr, momentum,
                                                                         nullptr, nullptr, nullptr, nullptr);
    tnq_internal_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_sgdg_step launch");
    return 0;
}

extern "C" int tnq_sgdg_step_flat(float* params, const float* grads, float* velocity, const int64_t* offsets,
                                  const int* rows, const int* cols, int ncores, int max_cols, float lr, float momentum,
                                  void* stream) {
    if (!params || !grads || !velocity || !offsets || !rows || !cols || ncores <= 0)
        return tnq_internal_fail("tnq_sgdg_step_flat: bad arguments");
    if (max_cols < 1 || max_cols > TNQ_SGDG_MAX_COLS)
        return tnq_internal_fail("tnq_sgdg_step_flat: matrix width must be between 1 and " + std::to_string(TNQ_SGDG_MAX_COLS));
    const int d = max_cols;
    const size_t smem = sizeof(float) * 7 * (size_t)d * (d + 1);
    cudaError_t e = cudaSuccess;
    if (smem > 48 * 1024) e = cudaFuncSetAttribute(tnq_sgdg_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(sgdg)");
    tnq_sgdg_kernel<<<ncores, SG_THREADS, smem, (cudaStream_t)stream>>>(nullptr, nullptr, nullptr, rows, cols, d, lr, momentum,
                                                                         params, grads, velocity,
                                                                         reinterpret_cast<const long long*>(offsets));
    tnq_internal_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_sgdg_step_flat launch");
    return 0;
}
