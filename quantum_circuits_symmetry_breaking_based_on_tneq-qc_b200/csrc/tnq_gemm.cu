// tnq_gemm.cu -- batched fp32 GEMM on the 5th-generation tensor cores (tcgen05) with
// fp32-faithful arithmetic: every operand is split on the fly into TF32 hi + lo parts and
// each product is issued as three tcgen05.mma.kind::tf32 (hi*hi + hi*lo + lo*hi, "3xTF32"),
// accumulated in fp32 in tensor memory (TMEM).
//
//     C[b] (M x N, row major, ldc)  (=|+=)  A[b] (M x K, row major, lda) * B[b]^T  (B[b] is N x K, ldb)
//
// Role in the contraction path: the large-bond-dimension regime (bond 64-128, complex64), where a
// pairwise contraction of the greedy sweep (tneq_qc/contractor/greedy_strategy.py:940,959: one
// torch.einsum per qubit group) is a real GEMM: rows = batch x kept indices, K = contracted
// indices.  Complex contractions arrive here as real GEMMs of doubled K and N (2x2-real form).
//
// Kernel anatomy (one 128 x 128 output tile per CTA):
//   warps 0-7  producers : 16-byte global loads of the fp32 A / B tiles (all loads of a stage are
//                          issued before any is consumed), split into TF32 hi and lo, stored to
//                          shared memory directly in the UMMA canonical K-major layout (8-row x
//                          16-byte core matrices, no swizzle), 3-stage ring, fence.proxy.async +
//                          mbarrier arrive;
//   warp 8     MMA issuer: one elected lane waits on the stage's "full" mbarrier and issues
//                          4 k-steps x 3 tcgen05.mma (M=128, N=128, K=8), then tcgen05.commit to
//                          the stage's "empty" mbarrier; owns the TMEM allocation;
//   drain (the producer warps again, while their global loads of the next stage are in flight):
//                          every pipeline stage owns one hi*hi accumulator in TMEM; before a stage
//                          is refilled, the partial sum its previous k-block left there is read out
//                          (tcgen05.ld 32 lanes x 32 columns) and added, round to nearest, to a
//                          running sum in registers; afterwards the epilogue: + the last partial
//                          sums + the correction accumulator -> 32x33 shared-memory transpose -> one
//                          coalesced 128-byte row segment per store instruction.
//
// Why the drain: the tensor core adds into its fp32 accumulator with TRUNCATION (round toward
// zero), i.e. every accumulating MMA costs ~1/2 ulp of the accumulator, always in the same
// direction.  Over the k-loop of one GEMM (1024 K=8 steps at bond 64) and the ~45 chained GEMMs
// of a 16-qubit sweep this is a systematic relative bias of ~2e-4 -- outside the 1e-5 parity
// bound.  With a FRESH accumulator per k-block (4 steps) the truncation acts on partial sums that
// are 1/nkb of the result, and the partial sums are added with IEEE round-to-nearest on the CUDA
// cores: the bias of a GEMM drops to ~1 ulp of its result, independent of K.  One hi*hi
// accumulator per pipeline stage (3 x 128 TMEM columns): the "full" barrier of a stage orders the
// drain before the next MMA into that accumulator, so no extra synchronisation is needed; the
// 2^-11 smaller correction terms (hi*lo + lo*hi) keep one accumulator for the whole k-loop
// (their truncation error is 2^-11 smaller too).
#include <cuda.h>            // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "tneq_b200.h"

extern int tnq_internal_fail(const std::string& msg);
extern int tnq_internal_cuda_fail(cudaError_t e, const char* what);
extern void tnq_internal_count_launch();

namespace {

constexpr int BM = 128, BN = 128, BK = 32;          // tile (BK floats = 128 bytes = 8 x 16-byte chunks)
constexpr int STAGES = 3;
constexpr int TILE_BYTES = BM * BK * 4;             // 16 KB, one of {A hi, A lo, B hi, B lo}
constexpr int STAGE_BYTES = 4 * TILE_BYTES;         // 64 KB
constexpr int PRODUCERS = 256;                      // 8 producer / drain / epilogue warps
constexpr int GEMM_THREADS = PRODUCERS + 32;
constexpr uint32_t TMEM_COLS = 512;                 // one hi*hi accumulator per stage (3) + the correction accumulator

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol error traps (surfacing as a CUDA error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try(bar, parity)) {
        if (clock64() - t0 > 4000000000LL) __trap();
    }
}

// UMMA shared-memory descriptor, K-major, no swizzle: 8 rows x 16 bytes core matrices stored
// contiguously (128 B); LBO = distance between the two 16-byte K chunks of one MMA (128 B here),
// SBO = distance between 8-row groups (1024 B here).  (cute/arch/mma_sm100_desc.hpp SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)(128 >> 4) << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;
}

// instruction descriptor: D=f32, A=B=tf32, both K-major, N=128, M=128 (InstrDescriptor bit layout)
constexpr uint32_t IDESC = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma_tf32(uint32_t tmem_c, uint64_t adesc, uint64_t bdesc, uint32_t accumulate,
                                          uint32_t idesc = IDESC) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t"
        "}\n" ::"r"(tmem_c),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate), "r"(0u)
        : "memory");
}

__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ float tf32_hi(float v) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
    return __uint_as_float(r);
}

// one 8-row x 4-chunk unit of a tile: global (row, chunk) -> register
template <bool ALIGNED>
__device__ __forceinline__ float4 load_unit(const float* __restrict__ src, long long ld, long long rows_left,
                                            long long k_left, int unit, int lane) {
    const int rg = unit >> 1, half = unit & 1;
    const int r8 = lane & 7, c4 = lane >> 3;
    const int row = rg * 8 + r8, chunk = half * 4 + c4;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (ALIGNED) {
        if (row < rows_left && chunk * 4 < k_left) v = __ldg(reinterpret_cast<const float4*>(src + row * ld + chunk * 4));
    } else if (row < rows_left) {   // odd K / leading dimension / base address: guarded scalar loads
        const float* r = src + row * ld + chunk * 4;
        const long long left = k_left - chunk * 4;
        if (left > 0) v.x = __ldg(r);
        if (left > 1) v.y = __ldg(r + 1);
        if (left > 2) v.z = __ldg(r + 2);
        if (left > 3) v.w = __ldg(r + 3);
    }
    return v;
}

// register -> TF32 hi / lo tiles in the canonical layout
__device__ __forceinline__ void store_unit(float4 v, int unit, int lane, uint32_t hi_base, uint32_t lo_base) {
    const int rg = unit >> 1, half = unit & 1;
    const int r8 = lane & 7, c4 = lane >> 3;
    const int chunk = half * 4 + c4;
    float4 h, l;
    h.x = tf32_hi(v.x), h.y = tf32_hi(v.y), h.z = tf32_hi(v.z), h.w = tf32_hi(v.w);
    // lo is rounded to nearest TF32 as well: the tensor core would otherwise truncate it (biased)
    l.x = tf32_hi(v.x - h.x), l.y = tf32_hi(v.y - h.y), l.z = tf32_hi(v.z - h.z), l.w = tf32_hi(v.w - h.w);
    const uint32_t off = (uint32_t)(rg * 64 + chunk * 8 + r8) * 16u;
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(hi_base + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo_base + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
}

// 32 lanes x 32 columns of an accumulator: r[j] = column c0 + j of this thread's TMEM lane
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// running sums of one thread: its TMEM lane (= output row) x 64 columns.  first == true overwrites.
__device__ __forceinline__ void drain_add(uint32_t taddr, float (&sum)[64], bool first) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)(cc * 32), r);
        if (first) {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[cc * 32 + j] = __uint_as_float(r[j]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) sum[cc * 32 + j] += __uint_as_float(r[j]);   // IEEE round to nearest
        }
    }
}

// epilogue of one warp: 32 rows x 64 columns of sums -> C (through a 32 x 33 transpose so that a
// store instruction covers one 128-byte row segment)
__device__ __forceinline__ void store_tile(const float (&sum)[64], float* xpose, int lane, float* __restrict__ C,
                                           long long ldc, long long row0, long long col0, long long M, long long N,
                                           int accumulate) {
#pragma unroll
    for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
        for (int j = 0; j < 32; ++j) xpose[lane * 33 + j] = sum[cc * 32 + j];
        __syncwarp();
        const long long col = col0 + cc * 32 + lane;
#pragma unroll 4
        for (int rr = 0; rr < 32; ++rr) {
            const long long row = row0 + rr;
            if (row < M && col < N) {
                float* d = C + row * ldc + col;
                const float v = xpose[rr * 33 + lane];
                if (accumulate == 2) atomicAdd(d, v);        // split-K: two CTAs add into a zeroed tile (two addends commute: deterministic)
                else *d = accumulate ? *d + v : v;
            }
        }
        __syncwarp();
    }
}

// the 12 MMAs of one k-block (issued by one thread), then the commits that release the stage / signal completion
__device__ __forceinline__ void issue_kblock(uint32_t st, uint32_t tmem_h, uint32_t tmem_corr, int kb, bool last,
                                             uint32_t empty_bar, uint32_t accum_bar) {
#pragma unroll
    for (int j = 0; j < BK / 8; ++j) {
        const uint32_t ko = (uint32_t)j * 2u * 128u;     // two 16-byte chunks per K=8 step
        const uint64_t a_hi = umma_desc(st + ko), a_lo = umma_desc(st + TILE_BYTES + ko);
        const uint64_t b_hi = umma_desc(st + 2 * TILE_BYTES + ko), b_lo = umma_desc(st + 3 * TILE_BYTES + ko);
        // correction terms: one accumulator for the whole k-loop (Ootomo & Yokota's split accumulators);
        // hi*hi: a fresh accumulator per k-block, drained and summed round-to-nearest on the CUDA cores
        umma_tf32(tmem_corr, a_lo, b_hi, (kb | j) != 0);
        umma_tf32(tmem_corr, a_hi, b_lo, 1u);
        umma_tf32(tmem_h, a_hi, b_hi, j != 0);
    }
    umma_commit(empty_bar);                              // frees the stage (and its accumulator) when the MMAs retire
    if (last) umma_commit(accum_bar);                    // everything complete
}

// SMALLK (K <= 256: a handful of k-blocks, the regime of the bond-64 sweep where K = 2 x bond):
// one pipeline stage, one hi*hi + one correction accumulator = 256 TMEM columns, <= 112 registers and EIGHT
// warps (lane 0 of warp 0 issues the MMAs after its own loads; with one stage the producers wait for the MMAs
// anyway), so that TWO CTAs share an SM -- the register file is partitioned per scheduler: 2 x 8 warps x 112
// registers fit, 2 x 9 do not -- and one tile's epilogue overlaps the other tile's loads and MMAs.
template <bool ALIGNED, bool SMALLK>
__global__ void __maxnreg__(SMALLK ? 128 : 168)
tnq_gemm_tf32x3_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ C, long long M,
                       long long N, long long K, long long lda, long long ldb, long long ldc, long long strideA,
                       long long strideB, long long strideC, int accumulate) {
    constexpr int NSTAGE = SMALLK ? 1 : STAGES;          // pipeline stages = hi*hi accumulators (the correction one follows)
    constexpr uint32_t NCOLS = SMALLK ? 256u : TMEM_COLS;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * STAGES + 1];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
    A += (long long)blockIdx.z * strideA + m0 * lda;
    B += (long long)blockIdx.z * strideB + n0 * ldb;
    C += (long long)blockIdx.z * strideC;
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t full0 = smem_u32(&bars[0]), empty0 = smem_u32(&bars[STAGES]), accum_bar = smem_u32(&bars[2 * STAGES]);
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; ++s) {
            mbar_init(full0 + 8 * s, PRODUCERS);
            mbar_init(empty0 + 8 * s, 1);
        }
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    constexpr int ALLOC_WARP = SMALLK ? 0 : PRODUCERS / 32;
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(NCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_corr = tmem_base + (uint32_t)(NSTAGE * BN);
    const int nkb = (int)((K + BK - 1) / BK);

    if (warp < PRODUCERS / 32) {
        // ---------------- producers (+ drain) ----------------
        // warp w owns TMEM lanes [32 (w % 4), +32) (hardware rule) x the column half w / 4
        const int quad = warp & 3, chalf = warp >> 2;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 64);
        float sum[64];
        bool have = false;                    // sum[] holds at least one partial result
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NSTAGE;
            const uint32_t phase = (uint32_t)(kb / NSTAGE) & 1u;
            mbar_wait(empty0 + 8 * s, phase ^ 1u);           // the MMAs of k-block kb - NSTAGE have retired
            const uint32_t st = smem_base + (uint32_t)s * STAGE_BYTES;
            const long long k0 = (long long)kb * BK;
            if (SMALLK) {
                float4 v[4];                  // A then B: half the registers in flight
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = load_unit<ALIGNED>(A + k0, lda, M - m0, K - k0, warp * 4 + i, lane);
                if (kb >= NSTAGE) {           // the drain of the previous k-block while the loads of A are in flight
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    drain_add(tlane + (uint32_t)(s * BN), sum, !have);
                    have = true;
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) store_unit(v[i], warp * 4 + i, lane, st, st + TILE_BYTES);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = load_unit<ALIGNED>(B + k0, ldb, N - n0, K - k0, warp * 4 + i, lane);
#pragma unroll
                for (int i = 0; i < 4; ++i) store_unit(v[i], warp * 4 + i, lane, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
            } else {
                float4 va[4], vb[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {        // all global loads of the stage first ...
                    va[i] = load_unit<ALIGNED>(A + k0, lda, M - m0, K - k0, warp * 4 + i, lane);
                    vb[i] = load_unit<ALIGNED>(B + k0, ldb, N - n0, K - k0, warp * 4 + i, lane);
                }
                if (kb >= NSTAGE) {                  // ... the drain of this stage's accumulator while they are in flight ...
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    drain_add(tlane + (uint32_t)(s * BN), sum, !have);
                    have = true;
                    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {        // ... then split and store
                    store_unit(va[i], warp * 4 + i, lane, st, st + TILE_BYTES);
                    store_unit(vb[i], warp * 4 + i, lane, st + 2 * TILE_BYTES, st + 3 * TILE_BYTES);
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> tensor-core reads
            mbar_arrive(full0 + 8 * s);          // (also: this stage's accumulator has been read)
            if (SMALLK && warp == 0) {           // no separate issuer warp in this variant
                mbar_wait(full0 + 8 * s, phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0)
                    issue_kblock(st, tmem_base + (uint32_t)(s * BN), tmem_corr, kb, kb == nkb - 1, empty0 + 8 * s, accum_bar);
                __syncwarp();
            }
        }
        // ---------------- epilogue ----------------
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int pending = nkb < NSTAGE ? nkb : NSTAGE;     // accumulators that still hold a partial sum
#pragma unroll 1
        for (int i = 0; i < pending; ++i) {
            drain_add(tlane + (uint32_t)(((nkb - pending + i) % NSTAGE) * BN), sum, !have);
            have = true;
        }
        drain_add(tmem_corr - tmem_base + tlane, sum, false);   // correction terms
        float* xpose = reinterpret_cast<float*>(smem) + warp * (32 * 33);   // pipeline smem is idle now
        store_tile(sum, xpose, lane, C, ldc, m0 + quad * 32, n0 + chalf * 64, M, N, accumulate);
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (!SMALLK) {
        // ---------------- MMA issuer ----------------
        for (int kb = 0; kb < nkb; ++kb) {
            const int s = kb % NSTAGE;
            const uint32_t phase = (uint32_t)(kb / NSTAGE) & 1u;
            mbar_wait(full0 + 8 * s, phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0)
                issue_kblock(smem_base + (uint32_t)s * STAGE_BYTES, tmem_base + (uint32_t)(s * BN), tmem_corr, kb, kb == nkb - 1,
                             empty0 + 8 * s, accum_bar);
            __syncwarp();
        }
    }
    __syncthreads();
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NCOLS) : "memory");
    }
}


// =====================================================================================================
// TMA-fed variant (the default whenever the operands qualify: 16-byte aligned base / leading dimension / batch
// stride).  Same arithmetic, accumulators, drain and epilogue as above; what changes is how the operands
// reach shared memory:
//   * ONE thread issues cp.async.bulk.tensor (SASS UTMALDG) per operand and k-block: the raw fp32 tile
//     (128 rows x 32 floats = 128-byte rows) lands in the 128-byte-swizzled K-major layout the tensor core
//     reads; out-of-range rows / columns (tile tails in M, N and K) arrive as zeros from the TMA unit;
//   * the tensor core reads the RAW tile as the TF32 "hi" operand: kind::tf32 ignores the 13 low mantissa
//     bits of a 32-bit operand, i.e. hi = trunc(v) without any instruction spent on it;
//   * the eight worker warps only produce lo = rna_tf32(v - trunc(v)): v - trunc(v) is exact in fp32, the
//     conversion is element-wise and therefore layout agnostic (read 16 bytes at an offset of the raw tile,
//     write 16 bytes at the same offset of the lo tile: the swizzle never has to be computed);
//   * raw ring (NRAW stages, filled by TMA) and lo ring (NLO stages, one hi*hi accumulator each) are
//     decoupled, so the loads run further ahead than the conversion.
// Barriers: full_raw[NRAW] (TMA transaction bytes), full_lo[NLO] (256 worker arrivals; also: the stage's
// accumulator has been drained), done[NRAW] (tcgen05.commit of a k-block: its raw stage, its lo stage and
// its accumulator are free), accum (everything retired).
// =====================================================================================================
constexpr int RAW_BYTES = 2 * TILE_BYTES;            // A | B of one k-block

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
        "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2, int c3, int c4) {
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];" ::"r"(dst),
        "l"((uint64_t)map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4)
        : "memory");
}
// UMMA shared-memory descriptor, K-major, 128-byte swizzle: rows of 128 bytes, 8-row groups 1024 bytes apart
// (SBO); LBO is not used by swizzled K-major layouts; a K=8 step advances the start address by 32 bytes inside
// the swizzle atom (the hardware applies the XOR to the absolute address bits: tiles are 1024-byte aligned).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;   // SWIZZLE_128B
    return d;
}
// MN-major operand (the rows are the contiguous direction in memory: [k][128 rows]), 128-byte swizzle: atoms of 32 floats
// (rows) x 8 k; the four row groups of a 128-row tile are LBO = 4096 bytes apart (one TMA box slice of 32 k-rows x 128 B),
// groups of 8 k-rows SBO = 1024 bytes apart; a K = 8 step advances the start address by 1024 bytes.
// (canonical layout ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units, cute/atom/mma_traits_sm100.hpp make_umma_desc<Major::MN>)
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr, uint32_t lbo, uint32_t sbo, uint32_t ltype) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(ltype & 7) << 61;
    return d;
}
struct MnDesc {
    uint32_t lbo, sbo, kstep, ltype;       // bytes, bytes, bytes per K = 8 step, UMMA layout type
};
__device__ __forceinline__ void issue_kblock_tma(uint32_t raw, uint32_t lo, uint32_t tmem_h, uint32_t tmem_corr, int kb, bool last,
                                                 uint32_t done_bar, uint32_t accum_bar, int a_mn = 0, int b_mn = 0,
                                                 MnDesc mn = MnDesc{4096, 512, 1024, 1}) {
    const uint32_t idesc = IDESC | (a_mn ? (1u << 15) : 0u) | (b_mn ? (1u << 16) : 0u);
#pragma unroll
    for (int j = 0; j < BK / 8; ++j) {
        const uint32_t ka = (uint32_t)j * (a_mn ? mn.kstep : 32u), kb_ = (uint32_t)j * (b_mn ? mn.kstep : 32u);
        const uint64_t a_hi = a_mn ? umma_desc_mn_sw128(raw + ka, mn.lbo, mn.sbo, mn.ltype) : umma_desc_sw128(raw + ka);
        const uint64_t a_lo = a_mn ? umma_desc_mn_sw128(lo + ka, mn.lbo, mn.sbo, mn.ltype) : umma_desc_sw128(lo + ka);
        const uint64_t b_hi = b_mn ? umma_desc_mn_sw128(raw + TILE_BYTES + kb_, mn.lbo, mn.sbo, mn.ltype) : umma_desc_sw128(raw + TILE_BYTES + kb_);
        const uint64_t b_lo = b_mn ? umma_desc_mn_sw128(lo + TILE_BYTES + kb_, mn.lbo, mn.sbo, mn.ltype) : umma_desc_sw128(lo + TILE_BYTES + kb_);
        umma_tf32(tmem_corr, a_lo, b_hi, (kb | j) != 0, idesc);
        umma_tf32(tmem_corr, a_hi, b_lo, 1u, idesc);
        umma_tf32(tmem_h, a_hi, b_hi, j != 0, idesc);
    }
    umma_commit(done_bar);
    if (last) umma_commit(accum_bar);
}

// SOLO: no dedicated TMA / MMA warps (256 threads, lane 0 of warp 0 issues both; NLO = 1) -- the small-K regime,
// two CTAs per SM so that one tile's epilogue overlaps the other's main loop.
template <int NRAW, int NLO, bool SOLO>
__global__ void __maxnreg__(SOLO ? 128 : 168)
tnq_gemm_tma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, int tma_store, float* __restrict__ C,
                    long long M, long long N, long long K, long long ldc, long long strideC, int zA, int zB, int accumulate,
                    int rewrite_hi, int splitk, int a_r0, int a_kdiv, int a_mn, int b_mn, int bk_kdiv, MnDesc mn) {
    // a_mn / b_mn: the operand is MN-major IN PLACE, memory [batch][row tile][k (bk_kdiv blocks of 32)][128 rows], read
    // through a 5-D tensor map (32-float chunk, k, chunk group, row tile, batch); the k-loop runs over (batch, k block):
    // the batch-into-K contraction of the core gradients without the transposition (tnq_gemm_tf32x3_bk)
    // a_r0 > 0: A is a strided VIEW described by a 4-D tensor map (k0, r0, k1, r1) -- rows m = r1 * a_r0 + r0,
    // contraction index k = k1 * (32 a_kdiv) + k0: the index permutation an explicit transposition kernel would
    // do is done by the TMA unit while it fills the tile (tnq_gemm_tf32x3_view)
    static_assert(!SOLO || NLO == 1, "the solo variant serialises conversion and MMA");
    static_assert(NRAW >= NLO, "done[] is indexed by the raw stage");
    constexpr uint32_t NCOLS = (NLO + 1) * BN <= 256 ? 256u : 512u;
    extern __shared__ __align__(1024) unsigned char smem[];
    __shared__ __align__(8) unsigned long long bars[2 * NRAW + NLO + 1];
    __shared__ uint32_t tmem_base_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const long long m0 = (long long)blockIdx.y * BM, n0 = (long long)blockIdx.x * BN;
    if (splitk <= 1) C += (long long)blockIdx.z * strideC;
    const uint32_t smem_base = (smem_u32(smem) + 1023u) & ~1023u;
    const uint32_t raw0 = smem_base, lo0 = smem_base + (uint32_t)NRAW * RAW_BYTES;
    const uint32_t full_raw0 = smem_u32(&bars[0]), done0 = smem_u32(&bars[NRAW]), full_lo0 = smem_u32(&bars[2 * NRAW]),
                   accum_bar = smem_u32(&bars[2 * NRAW + NLO]);
    if (tid == 0) {
        for (int s = 0; s < NRAW; ++s) {
            mbar_init(full_raw0 + 8 * s, 1);
            mbar_init(done0 + 8 * s, 1);
        }
        for (int s = 0; s < NLO; ++s) mbar_init(full_lo0 + 8 * s, PRODUCERS);
        mbar_init(accum_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmB) : "memory");
        if (tma_store) asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&tmC) : "memory");
    }
    constexpr int ALLOC_WARP = SOLO ? 1 : PRODUCERS / 32;
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_slot)),
                     "r"(NCOLS)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = tmem_base_slot;
    const uint32_t tmem_corr = tmem_base + (uint32_t)(NLO * BN);
    // split-K (splitk > 1, batch == 1): blockIdx.z owns the k-blocks [kb0, kb0 + nkb) and adds its partial tile atomically
    const int nkb_all = (int)((K + BK - 1) / BK);
    const int kb0 = splitk > 1 ? (int)((long long)nkb_all * blockIdx.z / splitk) : 0;
    const int nkb = splitk > 1 ? (int)((long long)nkb_all * (blockIdx.z + 1) / splitk) - kb0 : nkb_all;
    const int za = (zA && splitk <= 1) ? (int)blockIdx.z : 0, zb = (zB && splitk <= 1) ? (int)blockIdx.z : 0;
    if (splitk > 1) accumulate = 2;

    auto load_kblock = [&](int kb) {                     // one thread
        const int sr = kb % NRAW;
        const uint32_t dst = raw0 + (uint32_t)sr * RAW_BYTES, bar = full_raw0 + 8 * sr;
        mbar_expect_tx(bar, RAW_BYTES);
        const int kk = kb0 + kb;
        if (a_mn) {
            tma_load_5d(dst, &tmA, bar, 0, (kk % bk_kdiv) * BK, 0, (int)(m0 / BM), kk / bk_kdiv);
        } else if (a_r0 > 0) {
            tma_load_4d(dst, &tmA, bar, (kk % a_kdiv) * BK, a_r0 >= BM ? (int)(m0 % a_r0) : 0, kk / a_kdiv, (int)(m0 / a_r0));
        } else {
            tma_load_3d(dst, &tmA, bar, kk * BK, (int)m0, za);
        }
        if (b_mn) tma_load_5d(dst + TILE_BYTES, &tmB, bar, 0, (kk % bk_kdiv) * BK, 0, (int)(n0 / BN), kk / bk_kdiv);
        else tma_load_3d(dst + TILE_BYTES, &tmB, bar, kk * BK, (int)n0, zb);
    };

    if (warp < PRODUCERS / 32) {
        // ---------------- workers: drain, lo conversion, epilogue ----------------
        const int quad = warp & 3, chalf = warp >> 2;
        const uint32_t tlane = tmem_base + ((uint32_t)(quad * 32) << 16) + (uint32_t)(chalf * 64);
        float sum[64];
        bool have = false;
        if (SOLO && tid == 0)
            for (int kb = 0; kb < NRAW && kb < nkb; ++kb) load_kblock(kb);
        for (int kb = 0; kb < nkb; ++kb) {
            const int sr = kb % NRAW, sl = kb % NLO;
            if (kb >= NLO) {      // k-block kb - NLO has retired: its lo stage is free and its accumulator complete
                const int j = kb - NLO;
                mbar_wait(done0 + 8 * (j % NRAW), (uint32_t)(j / NRAW) & 1u);
                if (SOLO && tid == 0 && j + NRAW < nkb) load_kblock(j + NRAW);   // (NLO == 1: j = kb - 1, its raw stage is free too)
                if (SOLO) __syncwarp();
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                drain_add(tlane + (uint32_t)(sl * BN), sum, !have);
                have = true;
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            }
            mbar_wait(full_raw0 + 8 * sr, (uint32_t)(kb / NRAW) & 1u);      // the raw tiles have landed
            const uint32_t raw = raw0 + (uint32_t)sr * RAW_BYTES, lo = lo0 + (uint32_t)sl * RAW_BYTES;
            float4 v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t off = (uint32_t)(i * PRODUCERS + tid) * 16u;
                asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v[i].x), "=f"(v[i].y), "=f"(v[i].z), "=f"(v[i].w) : "r"(raw + off));
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint32_t off = (uint32_t)(i * PRODUCERS + tid) * 16u;
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v[i].x) & 0xFFFFE000u), h.y = __uint_as_float(__float_as_uint(v[i].y) & 0xFFFFE000u);
                h.z = __uint_as_float(__float_as_uint(v[i].z) & 0xFFFFE000u), h.w = __uint_as_float(__float_as_uint(v[i].w) & 0xFFFFE000u);
                // lo rounded to nearest TF32: the tensor core would truncate it (a biased error)
                l.x = tf32_hi(v[i].x - h.x), l.y = tf32_hi(v[i].y - h.y), l.z = tf32_hi(v[i].z - h.z), l.w = tf32_hi(v[i].w - h.w);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(lo + off), "f"(l.x), "f"(l.y), "f"(l.z), "f"(l.w) : "memory");
                if (rewrite_hi)   // diagnostics: do not rely on the tensor core ignoring the low mantissa bits
                    asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(raw + off), "f"(h.x), "f"(h.y), "f"(h.z), "f"(h.w) : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            mbar_arrive(full_lo0 + 8 * sl);
            if (SOLO && warp == 0) {
                mbar_wait(full_lo0 + 8 * sl, (uint32_t)(kb / NLO) & 1u);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (lane == 0)
                    issue_kblock_tma(raw, lo, tmem_base + (uint32_t)(sl * BN), tmem_corr, kb, kb == nkb - 1, done0 + 8 * sr, accum_bar, a_mn, b_mn, mn);
                __syncwarp();
            }
        }
        // ---------------- epilogue ----------------
        mbar_wait(accum_bar, 0);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int pending = nkb < NLO ? nkb : NLO;
#pragma unroll 1
        for (int i = 0; i < pending; ++i) {
            drain_add(tlane + (uint32_t)(((nkb - pending + i) % NLO) * BN), sum, !have);
            have = true;
        }
        drain_add(tmem_corr - tmem_base + tlane, sum, false);
        if (tma_store && accumulate == 0) {
            // the tile leaves through the TMA unit (SASS UTMASTG): every thread parks its 64 columns in four staging
            // boxes of 128 rows x 32 floats in the 128-byte swizzle (the rings are idle now; 4-way bank conflicts =
            // the minimum for 512 bytes per warp instruction), one thread per column half issues the two stores;
            // rows / columns beyond M / N are clipped by the tensor map
            const uint32_t stage = smem_base;
            const int row = quad * 32 + lane;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
                const int box = chalf * 2 + (c >> 3), ch = c & 7;
                const uint32_t at = stage + (uint32_t)box * TILE_BYTES + (uint32_t)row * 128u + (uint32_t)((ch ^ (row & 7)) * 16);
                asm volatile("st.shared.v4.f32 [%0], {%1,%2,%3,%4};" ::"r"(at), "f"(sum[4 * c]), "f"(sum[4 * c + 1]),
                             "f"(sum[4 * c + 2]), "f"(sum[4 * c + 3])
                             : "memory");
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            if (chalf == 0) asm volatile("bar.sync 1, 128;" ::: "memory");      // the four warps of this column half
            else asm volatile("bar.sync 2, 128;" ::: "memory");
            if ((warp & 3) == 0 && lane == 0) {
#pragma unroll
                for (int bx = 0; bx < 2; ++bx) {
                    const int box = chalf * 2 + bx;
                    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)&tmC),
                                 "r"(stage + (uint32_t)box * TILE_BYTES), "r"((int)n0 + box * 32), "r"((int)m0), "r"((int)blockIdx.z)
                                 : "memory");
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory is read before the CTA retires
            }
        } else {
            float* xpose = reinterpret_cast<float*>(smem + (smem_base - smem_u32(smem))) + warp * (32 * 33);   // the rings are idle now
            store_tile(sum, xpose, lane, C, ldc, m0 + quad * 32, n0 + chalf * 64, M, N, accumulate);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    } else if (!SOLO && warp == PRODUCERS / 32) {
        // ---------------- MMA issuer ----------------
        for (int kb = 0; kb < nkb; ++kb) {
            const int sr = kb % NRAW, sl = kb % NLO;
            mbar_wait(full_lo0 + 8 * sl, (uint32_t)(kb / NLO) & 1u);
            mbar_wait(full_raw0 + 8 * sr, (uint32_t)(kb / NRAW) & 1u);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (lane == 0)
                issue_kblock_tma(raw0 + (uint32_t)sr * RAW_BYTES, lo0 + (uint32_t)sl * RAW_BYTES, tmem_base + (uint32_t)(sl * BN), tmem_corr,
                                 kb, kb == nkb - 1, done0 + 8 * sr, accum_bar, a_mn, b_mn, mn);
            __syncwarp();
        }
    } else if (!SOLO) {
        // ---------------- TMA producer ----------------
        if (lane == 0) {
            for (int kb = 0; kb < nkb; ++kb) {
                if (kb >= NRAW) mbar_wait(done0 + 8 * (kb % NRAW), (uint32_t)((kb - NRAW) / NRAW) & 1u);
                load_kblock(kb);
            }
        }
        __syncwarp();
    }
    __syncthreads();
    if (warp == ALLOC_WARP) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(NCOLS) : "memory");
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn encode_tiled() {
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// (k, row, batch) view of a row-major operand with 128 x 32 boxes in the 128-byte swizzle.  false: not expressible
// (alignment, extents) -- the caller takes the register-staged kernel.
bool make_operand_map(CUtensorMap* map, const float* base, long long rows, long long K, long long ld, long long batch,
                      long long stride, int* uses_z) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (ld & 3) || ld < K || K > 0x7fffffffLL || rows > 0x7fffffffLL) return false;
    const bool z = batch > 1 && stride != 0;
    if (z && ((stride & 3) || stride < rows * ld || batch > 0x7fffffffLL)) return false;
    *uses_z = z ? 1 : 0;
    long long s2 = z ? stride : rows * ld;
    s2 = (s2 + 3) & ~3LL;
    const cuuint64_t dims[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)(z ? batch : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)ld * 4, (cuuint64_t)s2 * 4};
    const cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// MN-major descriptor parameters (diagnostic overrides: TNQ_MN_LBO / _SBO / _KSTEP / _LTYPE / _TMASW)
MnDesc mn_desc() {
    auto env = [](const char* n, int dflt) { const char* v = getenv(n); return (uint32_t)(v ? atoi(v) : dflt); };
    return MnDesc{env("TNQ_MN_LBO", 4096), env("TNQ_MN_SBO", 512), env("TNQ_MN_KSTEP", 1024), env("TNQ_MN_LTYPE", 1)};
}

// (k0, r0, k1, r1) view: element (m = r1 * R0 + r0, k = k1 * K0 + k0) at base + r1 sR1 + r0 sR0 + k1 sK1 + k0
bool make_view_map(CUtensorMap* map, const float* base, long long R1, long long R0, long long sR1, long long sR0,
                   long long K1, long long K0, long long sK1) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (sR1 & 3) || (sR0 & 3) || (sK1 & 3) || K0 % BK || R0 <= 0 || R1 <= 0 || K1 <= 0) return false;
    if (!(R0 % BM == 0 || BM % R0 == 0)) return false;
    if (sR0 <= 0 || (R1 > 1 && sR1 <= 0) || (K1 > 1 && sK1 <= 0)) return false;
    const cuuint64_t dims[4] = {(cuuint64_t)K0, (cuuint64_t)R0, (cuuint64_t)K1, (cuuint64_t)R1};
    const cuuint64_t strides[3] = {(cuuint64_t)sR0 * 4, (cuuint64_t)(K1 > 1 ? sK1 : K0) * 4, (cuuint64_t)(R1 > 1 ? sR1 : sR0 * R0) * 4};
    const cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(R0 >= BM ? BM : R0), 1, (cuuint32_t)(R0 >= BM ? 1 : BM / R0)};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// (n, m, batch) view of C for the TMA store: 128 x 32 boxes in the 128-byte swizzle
bool make_c_map(CUtensorMap* map, float* base, long long M, long long N, long long ldc, long long batch, long long strideC) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc) return false;
    if (((uintptr_t)base & 15) || (ldc & 3) || ldc < N || M > 0x7fffffffLL || N > 0x7fffffffLL) return false;
    if (batch > 1 && ((strideC & 3) || strideC < M * ldc || batch > 0x7fffffffLL)) return false;
    long long s2 = batch > 1 ? strideC : M * ldc;
    s2 = (s2 + 3) & ~3LL;
    const cuuint64_t dims[3] = {(cuuint64_t)N, (cuuint64_t)M, (cuuint64_t)(batch > 1 ? batch : 1)};
    const cuuint64_t strides[2] = {(cuuint64_t)ldc * 4, (cuuint64_t)s2 * 4};
    const cuuint32_t box[3] = {32, (cuuint32_t)BM, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int NRAW, int NLO, bool SOLO>
int launch_tma(const CUtensorMap& ta, const CUtensorMap& tb, float* C, long long M, long long N, long long K, long long ldc,
               long long strideC, long long batch, int zA, int zB, int accumulate, int rewrite_hi, cudaStream_t st,
               int a_r0 = 0, int a_kdiv = 1, int a_mn = 0, int b_mn = 0, int bk_kdiv = 1) {
    // split-K by two when a long k-loop would leave more than half of the SMs without a tile (the core-gradient GEMMs
    // of the bond-64 sweep: 64 tiles, K = 16 384): C is zeroed, both halves add atomically
    int splitk = 1;
    if (!SOLO && batch == 1 && !accumulate && ((N + BN - 1) / BN) * ((M + BM - 1) / BM) <= 74 && K >= 32 * BK) {
        static const bool no_split = getenv("TNQ_GEMM_NO_SPLITK") != nullptr;
        if (!no_split) splitk = 2;
    }
    if (splitk > 1) {
        cudaError_t e0 = ldc == N ? cudaMemsetAsync(C, 0, sizeof(float) * (size_t)M * (size_t)N, st)
                                  : cudaMemset2DAsync(C, sizeof(float) * (size_t)ldc, 0, sizeof(float) * (size_t)N, (size_t)M, st);
        if (e0 != cudaSuccess) return tnq_internal_cuda_fail(e0, "cudaMemsetAsync(gemm split-K)");
    }
    // the tile leaves through a TMA store when C can be described by a tensor map and is written (not added to)
    CUtensorMap tc;
    static const bool no_store = getenv("TNQ_GEMM_NO_TMA_STORE") != nullptr;
    const int tma_store = (!no_store && splitk == 1 && !accumulate && make_c_map(&tc, C, M, N, ldc, batch, strideC)) ? 1 : 0;
    if (!tma_store) tc = ta;                                       // (a valid descriptor; never used)
    const size_t smem = (size_t)(NRAW + NLO) * RAW_BYTES + 1024;
    auto kern = tnq_gemm_tma_kernel<NRAW, NLO, SOLO>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(gemm tma)");
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)(splitk > 1 ? splitk : batch));
    kern<<<grid, SOLO ? PRODUCERS : PRODUCERS + 64, smem, st>>>(ta, tb, tc, tma_store, C, M, N, K, ldc, strideC, zA, zB, accumulate, rewrite_hi, splitk, a_r0, a_kdiv, a_mn, b_mn, bk_kdiv, mn_desc());
    tnq_internal_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_gemm_tf32x3 (TMA) launch");
    return 0;
}

}  // namespace

// launch resources of one kernel variant (tests / diagnostics): out = {registers per thread, max threads per block,
// static shared bytes, threads the launch uses}
extern "C" int tnq_gemm_kernel_attrs(int aligned, int smallk, int* out) {
    auto kern = smallk ? (aligned ? tnq_gemm_tf32x3_kernel<true, true> : tnq_gemm_tf32x3_kernel<false, true>)
                       : (aligned ? tnq_gemm_tf32x3_kernel<true, false> : tnq_gemm_tf32x3_kernel<false, false>);
    cudaFuncAttributes at;
    const cudaError_t e = cudaFuncGetAttributes(&at, kern);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncGetAttributes(gemm)");
    out[0] = at.numRegs, out[1] = at.maxThreadsPerBlock, out[2] = (int)at.sharedSizeBytes;
    out[3] = smallk ? PRODUCERS : GEMM_THREADS;
    return 0;
}

extern "C" int tnq_gemm_tf32x3(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K, int64_t lda,
                               int64_t ldb, int64_t ldc, int64_t batch, int64_t strideA, int64_t strideB,
                               int64_t strideC, int accumulate, void* stream) {
    if (!A || !B || !C || M <= 0 || N <= 0 || K <= 0 || batch <= 0) return tnq_internal_fail("tnq_gemm_tf32x3: bad arguments");
    const bool aligned = !((K & 3) || (lda & 3) || (ldb & 3) || (strideA & 3) || (strideB & 3) || ((uintptr_t)A & 15) ||
                           ((uintptr_t)B & 15));
    if (batch > 65535) return tnq_internal_fail("tnq_gemm_tf32x3: batch too large for one launch (max 65535)");
    const bool smallk = K <= 256 && !getenv("TNQ_GEMM_NO_SMALLK");
    // TMA-fed kernels whenever both operands can be described by a tensor map (TNQ_GEMM_NO_TMA=1: the
    // register-staged kernels below; TNQ_GEMM_TMA_STAGES=33: raw ring of 3 + lo ring of 3 instead of 4 + 2, measured 6 % slower at K = 8192)
    static const bool no_tma = getenv("TNQ_GEMM_NO_TMA") != nullptr;
    if (!no_tma) {
        CUtensorMap ta, tb;
        int zA = 0, zB = 0;
        if (make_operand_map(&ta, A, M, K, lda, batch, strideA, &zA) && make_operand_map(&tb, B, N, K, ldb, batch, strideB, &zB)) {
            static const int rewrite_hi = getenv("TNQ_GEMM_TMA_REWRITE_HI") ? 1 : 0;
            static const int stages = getenv("TNQ_GEMM_TMA_STAGES") ? atoi(getenv("TNQ_GEMM_TMA_STAGES")) : 42;
            if (smallk)
                return launch_tma<2, 1, true>(ta, tb, C, M, N, K, ldc, strideC, batch, zA, zB, accumulate, rewrite_hi, (cudaStream_t)stream);
            if (stages == 42)
                return launch_tma<4, 2, false>(ta, tb, C, M, N, K, ldc, strideC, batch, zA, zB, accumulate, rewrite_hi, (cudaStream_t)stream);
            return launch_tma<3, 3, false>(ta, tb, C, M, N, K, ldc, strideC, batch, zA, zB, accumulate, rewrite_hi, (cudaStream_t)stream);
        }
    }
    const size_t smem = (size_t)(smallk ? 1 : STAGES) * STAGE_BYTES + 1024;
    auto kern = smallk ? (aligned ? tnq_gemm_tf32x3_kernel<true, true> : tnq_gemm_tf32x3_kernel<false, true>)
                       : (aligned ? tnq_gemm_tf32x3_kernel<true, false> : tnq_gemm_tf32x3_kernel<false, false>);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(gemm)");
    dim3 grid((unsigned)((N + BN - 1) / BN), (unsigned)((M + BM - 1) / BM), (unsigned)batch);
    kern<<<grid, smallk ? PRODUCERS : GEMM_THREADS, smem, (cudaStream_t)stream>>>(A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC,
                                                             accumulate);
    tnq_internal_count_launch();
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_gemm_tf32x3 launch");
    return 0;
}

/* C (M x N, row major, ldc) = A_view * B^T with A a strided 4-level VIEW of a tensor in HBM (see tneq_b200.h):
 * rows m = r1 * R0 + r0, contraction index k = k1 * K0 + k0, element at A + r1 sR1 + r0 sR0 + k1 sK1 + k0.
 * Returns TNQ_NOT_EXPRESSIBLE (-2) without launching when the view cannot be described to the TMA unit (alignment,
 * K0 not a multiple of 32, R0 neither a multiple nor a divisor of 128, K <= 256): the caller transposes explicitly. */
extern "C" int tnq_gemm_tf32x3_view(const float* A, int64_t R1, int64_t R0, int64_t sR1, int64_t sR0, int64_t K1, int64_t K0,
                                    int64_t sK1, const float* B, int64_t ldb, float* C, int64_t ldc, int64_t N, void* stream) {
    if (!A || !B || !C || N <= 0) return tnq_internal_fail("tnq_gemm_tf32x3_view: bad arguments");
    const long long M = R1 * R0, K = K1 * K0;
    static const bool no_tma = getenv("TNQ_GEMM_NO_TMA") != nullptr || getenv("TNQ_GEMM_NO_VIEW") != nullptr;
    if (no_tma || K <= 256 || M > 0x7fffffffLL) return -2;
    CUtensorMap ta, tb;
    int zB = 0;
    if (!make_view_map(&ta, A, R1, R0, sR1, sR0, K1, K0, sK1) || !make_operand_map(&tb, B, N, K, ldb, 1, 0, &zB)) return -2;
    return launch_tma<4, 2, false>(ta, tb, C, M, N, K, ldc, 0, 1, 0, 0, 0, 0, (cudaStream_t)stream, (int)R0, (int)(K0 / BK));
}


namespace {
// MN-major operand in place: memory [batch][tiles][Kin][128 rows]; (32-float chunk, k, chunk group, row tile, batch)
bool make_mn_map(CUtensorMap* map, const float* base, long long tiles, long long Kin, long long batch) {
    EncodeTiledFn enc = encode_tiled();
    if (!enc || ((uintptr_t)base & 15) || Kin % BK || tiles <= 0 || batch <= 0) return false;
    const cuuint64_t dims[5] = {32, (cuuint64_t)Kin, 4, (cuuint64_t)tiles, (cuuint64_t)batch};
    const cuuint64_t strides[4] = {128 * 4, 32 * 4, (cuuint64_t)Kin * 128 * 4, (cuuint64_t)tiles * Kin * 128 * 4};
    const cuuint32_t box[5] = {32, (cuuint32_t)BK, 4, 1, 1};
    const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    // 32-bit elements consumed MN-major need the 128-byte swizzle with 32-byte atoms (UMMA layout type SWIZZLE_128B_BASE32B)
    const char* sw = getenv("TNQ_MN_TMASW");
    const CUtensorMapSwizzle mode = (sw && atoi(sw) == 0) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B;
    return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 5, const_cast<float*>(base), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, mode, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
}  // namespace

/* Batch-into-K GEMM (the core gradients of the large-bond sweep): C[m, n] = sum_{b, k} A_b[m, k] * B_b[n, k], K = batch x Kin.
 * An operand is either K-major and contiguous, X[rows][batch * Kin] (x_mn = 0), or MN-major IN PLACE, X[batch][x_tiles][Kin][128]
 * with rows = x_tiles * 128 (x_mn = 1): the tensor the forward sweep produced, read through a 5-D tensor map and consumed
 * by the tensor core through MN-major descriptors -- no transposition.  Returns -2 without launching when not expressible. */
extern "C" int tnq_gemm_tf32x3_bk(const float* A, int a_mn, int64_t a_tiles, const float* B, int b_mn, int64_t b_tiles, float* C,
                                  int64_t M, int64_t N, int64_t batch, int64_t Kin, void* stream) {
    if (!A || !B || !C || M <= 0 || N <= 0 || batch <= 0 || Kin <= 0) return tnq_internal_fail("tnq_gemm_tf32x3_bk: bad arguments");
    static const bool off = getenv("TNQ_GEMM_NO_TMA") != nullptr || getenv("TNQ_GEMM_NO_MN") != nullptr;
    const long long K = batch * Kin;
    if (off || K <= 256 || Kin % BK || (a_mn && M != a_tiles * BM) || (b_mn && N != b_tiles * BN)) return -2;
    CUtensorMap ta, tb;
    int z = 0;
    if (!(a_mn ? make_mn_map(&ta, A, a_tiles, Kin, batch) : make_operand_map(&ta, A, M, K, K, 1, 0, &z))) return -2;
    if (!(b_mn ? make_mn_map(&tb, B, b_tiles, Kin, batch) : make_operand_map(&tb, B, N, K, K, 1, 0, &z))) return -2;
    return launch_tma<4, 2, false>(ta, tb, C, M, N, K, N, 0, 1, 0, 0, 0, 0, (cudaStream_t)stream, 0, 1, a_mn, b_mn, (int)(Kin / BK));
}
