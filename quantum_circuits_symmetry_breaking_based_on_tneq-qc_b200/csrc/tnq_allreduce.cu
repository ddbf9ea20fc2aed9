// tnq_allreduce.cu -- one-shot all-reduce of the packed gradient + loss buffer over NVLink peer
// memory (NVSwitch: every GPU reads every peer at full bandwidth).
//
// Replaces, for the ONE exchange step of the data-parallel hot path, the per-core blocking
// collectives of the reference (tneq_qc/distributed/comm/comm_torch.py:292-318, 510-522,
// DataParallelTrainer.sync_gradients data_parallel.py:194-204).  The message is 15 KB for the
// 24-qubit two-layer network: pure latency.  One kernel, one CTA per rank:
//   1. publish: copy the local values into this rank's slot of its SYMMETRIC buffer (memory that
//      every peer has mapped: torch.distributed._symmetric_memory), __threadfence_system();
//   2. signal: store the epoch into flags[parity][rank] of EVERY peer's buffer (st.release.sys);
//   3. wait until all peers' flags carry the epoch (ld.acquire.sys, bounded spin);
//   4. reduce: every rank reads all slots over NVLink in rank order (same order everywhere =>
//      bit-identical results on all ranks) and writes scale * sum to its output.
// Slots are double-buffered by epoch parity: a rank can only reach epoch e+2 after every peer
// has published e+1, i.e. after every peer finished reading epoch e.  The epoch counter lives in
// device memory, so the launch can be replayed from a CUDA graph.
//
// A peer that does not show up within the timeout (default 10 minutes, tnq_allreduce_set_timeout_ms;
// NCCL's default is of the same order) does NOT trap the context: the kernel records the epoch and
// the missing rank in the error words of its own buffer, fills the output with NaN (so the loss the
// host reads next is NaN and cannot be mistaken for a result) and returns; the host turns that into
// an exception (distributed/oneshot.py: check()) or falls back to NCCL.
//
// Symmetric buffer layout (32-bit words): [0,32) flags[2][16] | [32] epoch | [33] error epoch (0 = none) |
// [34] first missing rank | [64, 64+NMAX) slot 0 | [64+NMAX, 64+2 NMAX) slot 1.
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "tneq_b200.h"

extern int tnq_internal_fail(const std::string& msg);
extern int tnq_internal_cuda_fail(cudaError_t e, const char* what);
extern void tnq_internal_count_launch();

namespace {

constexpr int AR_THREADS = 1024;
constexpr int AR_MAX_WORLD = 16;
constexpr int AR_HEADER = 64;

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
    float v;
    asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

long long g_timeout_ns = 600LL * 1000 * 1000 * 1000;

__global__ void __launch_bounds__(AR_THREADS)
tnq_allreduce_oneshot_kernel(const uint64_t* __restrict__ peers, int rank, int world, long long nmax,
                             const float* __restrict__ src_a, long long na, const float* __restrict__ src_b,
                             long long nb, float* __restrict__ out, float scale, long long timeout_ns) {
    __shared__ uint32_t epoch_s;
    __shared__ int missing_s;
    uint32_t* own = reinterpret_cast<uint32_t*>(peers[rank]);
    if (threadIdx.x == 0) {
        epoch_s = own[32] + 1u;
        missing_s = -1;
    }
    __syncthreads();
    const uint32_t e = epoch_s, par = e & 1u;
    const long long n = na + nb;
    float* slot = reinterpret_cast<float*>(own) + AR_HEADER + (long long)par * nmax;
    for (long long i = threadIdx.x; i < n; i += AR_THREADS) slot[i] = i < na ? src_a[i] : src_b[i - na];
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < world) {
        uint32_t* peer = reinterpret_cast<uint32_t*>(peers[threadIdx.x]);
        st_release_sys(peer + par * AR_MAX_WORLD + rank, e);
        const unsigned long long t0 = global_ns();
        unsigned spins = 0;
        while (ld_acquire_sys(own + par * AR_MAX_WORLD + threadIdx.x) != e) {
            if ((++spins & 1023u) == 0 && global_ns() - t0 > (unsigned long long)timeout_ns) {
                atomicMax(&missing_s, (int)threadIdx.x);     // a missing peer surfaces as an error, not a hang
                break;
            }
        }
    }
    __syncthreads();
    if (missing_s >= 0) {                                    // no result is published
        for (long long i = threadIdx.x; i < n; i += AR_THREADS) out[i] = __int_as_float(0x7fc00000);
        if (threadIdx.x == 0) {
            own[33] = e;
            own[34] = (uint32_t)missing_s;
            own[32] = e;
        }
        return;
    }
    for (long long i = threadIdx.x; i < n; i += AR_THREADS) {
        float s = 0.f;
        for (int r = 0; r < world; ++r)
            s += ld_relaxed_sys(reinterpret_cast<const float*>(peers[r]) + AR_HEADER + (long long)par * nmax + i);
        out[i] = s * scale;
    }
    if (threadIdx.x == 0) own[32] = e;
}

}  // namespace

extern "C" {

int64_t tnq_allreduce_oneshot_words(int64_t nmax) { return AR_HEADER + 2 * nmax; }

int tnq_allreduce_set_timeout_ms(int64_t ms) {
    if (ms <= 0) return tnq_internal_fail("tnq_allreduce_set_timeout_ms: the timeout must be positive");
    g_timeout_ns = ms * 1000000LL;
    return 0;
}

int tnq_allreduce_oneshot(const uint64_t* peer_bufs_dev, int rank, int world, int64_t nmax, const float* src_a,
                          int64_t na, const float* src_b, int64_t nb, float* out, float scale, void* stream) {
    if (!peer_bufs_dev || !src_a || !out || world < 1 || world > AR_MAX_WORLD || rank < 0 || rank >= world)
        return tnq_internal_fail("tnq_allreduce_oneshot: bad arguments");
    if (na < 0 || nb < 0 || na + nb > nmax || (nb > 0 && !src_b))
        return tnq_internal_fail("tnq_allreduce_oneshot: message does not fit the symmetric buffer");
    tnq_allreduce_oneshot_kernel<<<1, AR_THREADS, 0, (cudaStream_t)stream>>>(peer_bufs_dev, rank, world, nmax, src_a, na,
                                                                              src_b, nb, out, scale, g_timeout_ns);
    tnq_internal_count_launch();
    const cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_allreduce_oneshot launch");
    return 0;
}

}  // extern "C"
