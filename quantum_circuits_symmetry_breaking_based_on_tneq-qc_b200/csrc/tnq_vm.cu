// tnq_vm.cu -- the contraction "virtual machine": one persistent kernel walks the
// complete qubit sweep (and, for training, the reverse sweep) of a QCTN for a
// tile of samples whose working set stays in shared memory.
//
// Replaces, on the device, the hot loop of the reference:
//   tneq_qc/contractor/greedy_strategy.py:461-598  (one torch.einsum per qubit group)
//   tneq_qc/core/engine_siamese.py:490-530         (clamp / log / mean loss)
//   tneq_qc/backends/backend_pytorch.py:153-158    (torch.autograd.grad)
// The program format is documented in contractor/vm_program.py.
//
// Kernels (sm_100a):
//   tnq_prep_kernel    PREP section, one CTA   : batch-independent tensors -> CONST pool
//   tnq_body_kernel    BODY section, persistent: per-sample ops on SoA tiles in smem
//   tnq_reduce_kernel  cross-CTA sum of the per-CTA gradient accumulators (fixed order)
//   tnq_fin_kernel     FIN section, one CTA    : reverse of PREP, loss, outputs
//
// Data layout: a tile's FRAME is structure-of-arrays [element][S samples], so a
// warp's 32 lanes always touch 32 consecutive words: conflict-free shared memory
// and coalesced global traffic for any tensor shape (K = 2, 3, 4, ... alike).
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <atomic>
#include <string>
#include <vector>

#include "tneq_b200.h"

namespace {

constexpr int OP_WORDS = 24;
constexpr int SP_CONST = 0, SP_FRAME = 1, SP_GIN = 2, SP_GOUT = 3, SP_GACC = 4;
constexpr int OP_LIN = 1, OP_GEMM = 2, OP_RGEMM = 3, OP_SEED = 4;
constexpr int THREADS = 256;           // PREP / FIN / reduce kernels
constexpr int BODY_THREADS_MAX = 512;  // BODY kernel: 512 threads when one CTA owns the SM, else 256
constexpr long long MAGIC = 0x544E5142323030LL;

thread_local std::string g_error;
std::atomic<long long> g_launches{0};

int fail(const std::string& msg) {
    g_error = msg;
    return 1;
}
int cuda_fail(cudaError_t e, const char* what) {
    g_error = std::string(what) + ": " + cudaGetErrorString(e);
    return 2;
}

struct Op {
    int w[OP_WORDS];
};

struct RunArgs {
    const void* in_ptr[TNQ_MAX_INPUTS];
    long long in_hi[TNQ_MAX_INPUTS];
    long long in_lo[TNQ_MAX_INPUTS];
    void* out_ptr[TNQ_MAX_OUTPUTS];
    int out_elems[TNQ_MAX_OUTPUTS];
    unsigned char in_batched[TNQ_MAX_INPUTS];
    unsigned char out_batched[TNQ_MAX_OUTPUTS];
};

template <typename T>
struct Prog {
    const Op* ops;
    const int* itab;
    const T* ftab;
    int n_ops;
    int const_elems, gacc_elems, frame_elems;
    int nb;
};

// ------------------------------------------------------------------------------------
// BODY
// ------------------------------------------------------------------------------------
// All FRAME / CONST / GACC accesses of the hot loops go through 32-bit byte offsets that are
// pre-scaled (element offset x S x sizeof(T)) when an op's tables are staged in shared memory,
// so the inner loops are: one integer add + one ld.shared per A value, 16-byte broadcast loads
// for the packed shared operand, FMAs -- and the epilogue is one add + one st.shared per output.
constexpr int OTAB = 3072;  // ints of staged per-op tables

template <typename T>
__device__ __forceinline__ T lds(uint32_t a) {
    T v;
    if constexpr (sizeof(T) == 4)
        asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    else
        asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
template <typename T>
__device__ __forceinline__ void sts(uint32_t a, T v) {
    if constexpr (sizeof(T) == 4)
        asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
    else
        asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
// 16-byte load into consecutive elements of o[]
__device__ __forceinline__ void lds16(uint32_t a, float* o) {
    asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(o[0]), "=f"(o[1]), "=f"(o[2]), "=f"(o[3]) : "r"(a));
}
__device__ __forceinline__ void lds16(uint32_t a, double* o) {
    asm volatile("ld.shared.v2.f64 {%0,%1}, [%2];" : "=d"(o[0]), "=d"(o[1]) : "r"(a));
}

// FRAME accessor: shared memory (32-bit window address) or, for plans whose working set does
// not fit, a per-CTA slab of global memory.
template <typename T, bool SMEM>
struct Frame;
template <typename T>
struct Frame<T, true> {
    uint32_t base;
    __device__ __forceinline__ T ld(uint32_t off) const { return lds<T>(base + off); }
    __device__ __forceinline__ void st(uint32_t off, T v) const { sts<T>(base + off, v); }
};
template <typename T>
struct Frame<T, false> {
    char* base;
    __device__ __forceinline__ T ld(uint32_t off) const { return *reinterpret_cast<const T*>(base + off); }
    __device__ __forceinline__ void st(uint32_t off, T v) const { *reinterpret_cast<T*>(base + off) = v; }
};

template <typename T, bool SMEM>
struct Tile {
    Frame<T, SMEM> fr;
    uint32_t cpool;   // shared-window byte address of the CONST pool
    T* gacc;          // per-CTA accumulators (shared memory)
    int* tab;         // staged tables of the current op (shared memory)
    int S, logS;
    uint32_t sbytes;  // S * sizeof(T): byte stride between consecutive elements of a buffer
    int nvalid;       // valid samples in this tile
    long long s0;     // first global sample of the tile
    int nb;
    T log_scale, inv_count;
};

template <typename T>
__device__ __forceinline__ long long sample_offset(const RunArgs& a, int slot, long long sg, int nb) {
    return nb == 1 ? sg * a.in_hi[slot] : (sg / nb) * a.in_hi[slot] + (sg % nb) * a.in_lo[slot];
}

template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// dst[j] (=|+=) c0[j]*src[s0[j]] (+ c1[j]*src[s1[j]])
template <typename T, bool SMEM>
__device__ void body_lin(const Op& op, const Tile<T, SMEM>& t, const Prog<T>& p, const RunArgs& args) {
    const int acc = op.w[1], dsp = op.w[2], dbase = op.w[3], ssp = op.w[4], sbase = op.w[5];
    const int count = op.w[6], nt = op.w[7];
    const int* s0 = p.itab + op.w[9];
    const T* c0 = p.ftab + op.w[10];
    const int* s1 = p.itab + op.w[11];
    const T* c1 = p.ftab + op.w[12];
    const int sslot = op.w[13], dslot = op.w[14];
    const int S = t.S;
    const bool src_batched_gin = (ssp == SP_GIN) && args.in_batched[sslot];
    const T* gsrc = (ssp == SP_GIN) ? reinterpret_cast<const T*>(args.in_ptr[sslot]) + sbase : nullptr;
    T* gdst = (dsp == SP_GOUT) ? reinterpret_cast<T*>(args.out_ptr[dslot]) + dbase : nullptr;
    const int delems = (dsp == SP_GOUT) ? args.out_elems[dslot] : 0;
    const int total = count << t.logS;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int s = i & (S - 1), j = i >> t.logS;
        const bool valid = s < t.nvalid;
        T v = T(0);
        for (int term = 0; term < nt; ++term) {
            const int so = term == 0 ? __ldg(s0 + j) : __ldg(s1 + j);
            const T cf = term == 0 ? __ldg(c0 + j) : __ldg(c1 + j);
            T x;
            if (ssp == SP_FRAME) {
                x = t.fr.ld((uint32_t)((sbase + so) * S + s) * (uint32_t)sizeof(T));
            } else if (ssp == SP_CONST) {
                x = lds<T>(t.cpool + (uint32_t)(sbase + so) * (uint32_t)sizeof(T));
            } else if (src_batched_gin) {
                x = valid ? __ldg(gsrc + sample_offset<T>(args, sslot, t.s0 + s, t.nb) + so) : T(0);
            } else {
                x = __ldg(gsrc + so);
            }
            v += cf * x;
        }
        if (dsp == SP_FRAME) {
            const uint32_t d = (uint32_t)((dbase + j) * S + s) * (uint32_t)sizeof(T);
            t.fr.st(d, acc ? t.fr.ld(d) + v : v);
        } else if (valid) {  // batched GOUT, [nsamples][elems]
            T* d = gdst + (t.s0 + s) * (long long)delems + j;
            *d = acc ? *d + v : v;
        }
    }
}

// ---- GEMM ---------------------------------------------------------------------------
// C[cm[r]+cn[c]] (=|+=) sum_k A[am[r]+ak[k]] * B[bk[k]+bn[c]] for every sample of the tile.
//
// Thread mapping: one sample lane x RM rows x TN columns per thread (RM*TN accumulators in
// registers).  Consecutive lanes are consecutive samples, so every A / C access of a warp is
// 32 consecutive words (conflict free); the RM rows of a thread are strided by the number of
// row groups so that, for tiles narrower than a warp, neighbouring lanes touch neighbouring
// rows.  A shared B is read from its packed PREP copy with 16-byte vector loads that are a
// broadcast for the warp; per k-step a thread issues RM + TN/4 shared loads for RM*TN FMAs.
//
// Staged tables (ints, byte offsets): [0,nm) am  [nm,2nm) cm  then bn[nn], cn[nn], ak[nk], bk[nk].
template <typename T, bool SMEM, int RM, int TN, bool B_BATCHED>
__device__ __forceinline__ void gemm_tile(const Op& op, const Tile<T, SMEM>& t) {
    constexpr int V = 16 / (int)sizeof(T);
    constexpr int TNP = (TN + V - 1) / V * V;
    const int acc = op.w[1];
    const int nm = op.w[8], nn = op.w[9], nk = op.w[10];
    const uint32_t ldb = (uint32_t)op.w[21] * (uint32_t)sizeof(T);
    const int* am = t.tab;
    const int* cm = am + nm;
    const int* bn = cm + nm;
    const int* cn = bn + nn;
    const int* ak = cn + nn;
    const int* bk = ak + nk;
    const int nrg = (nm + RM - 1) / RM, ncg = nn / TN;
    const int total = (nrg * ncg) << t.logS;
    for (int it = threadIdx.x; it < total; it += blockDim.x) {
        const uint32_t sb = (uint32_t)(it & (t.S - 1)) * (uint32_t)sizeof(T);
        const int rc = it >> t.logS;
        const int cg = rc / nrg, rg = rc - cg * nrg;
        uint32_t aoff[RM];
        int crow[RM];
#pragma unroll
        for (int j = 0; j < RM; ++j) {
            const int r = rg + j * nrg;
            const int rr = r < nm ? r : rg;
            aoff[j] = (uint32_t)am[rr] + sb;
            crow[j] = r < nm ? cm[rr] + (int)sb : -1;
        }
        T sum[RM][TN];
#pragma unroll
        for (int j = 0; j < RM; ++j)
#pragma unroll
            for (int i = 0; i < TN; ++i) sum[j][i] = T(0);
        if (!B_BATCHED) {
            uint32_t brow = t.cpool + (uint32_t)op.w[7] * (uint32_t)sizeof(T) + (uint32_t)(cg * TNP) * (uint32_t)sizeof(T);
#pragma unroll 3
            for (int k = 0; k < nk; ++k) {
                const uint32_t ka = (uint32_t)ak[k];
                T b[TNP];
#pragma unroll
                for (int i = 0; i < TNP; i += V) lds16(brow + (uint32_t)i * (uint32_t)sizeof(T), b + i);
                brow += ldb;
#pragma unroll
                for (int j = 0; j < RM; ++j) {
                    const T a = t.fr.ld(aoff[j] + ka);
#pragma unroll
                    for (int i = 0; i < TN; ++i) sum[j][i] = fma(a, b[i], sum[j][i]);
                }
            }
        } else {
            uint32_t boff[TN];
#pragma unroll
            for (int i = 0; i < TN; ++i) boff[i] = (uint32_t)bn[cg * TN + i] + sb;
#pragma unroll 3
            for (int k = 0; k < nk; ++k) {
                const uint32_t ka = (uint32_t)ak[k], kb = (uint32_t)bk[k];
                T b[TN];
#pragma unroll
                for (int i = 0; i < TN; ++i) b[i] = t.fr.ld(boff[i] + kb);
#pragma unroll
                for (int j = 0; j < RM; ++j) {
                    const T a = t.fr.ld(aoff[j] + ka);
#pragma unroll
                    for (int i = 0; i < TN; ++i) sum[j][i] = fma(a, b[i], sum[j][i]);
                }
            }
        }
#pragma unroll
        for (int i = 0; i < TN; ++i) {
            const int co = cn[cg * TN + i];
#pragma unroll
            for (int j = 0; j < RM; ++j) {
                if (crow[j] >= 0) {
                    const uint32_t d = (uint32_t)(crow[j] + co);
                    t.fr.st(d, acc ? t.fr.ld(d) + sum[j][i] : sum[j][i]);
                }
            }
        }
    }
}

template <typename T, bool SMEM, int TN, bool B_BATCHED>
__device__ __forceinline__ void gemm_rows(const Op& op, const Tile<T, SMEM>& t) {
    // rows per thread: as many as the register budget allows while every thread still has work
    constexpr int RMAX = TN <= 4 ? 8 : 4;
    const int nm = op.w[8], ncg = op.w[9] / TN;
    int rm = RMAX;
    while (rm > 2 && ((((nm + rm - 1) / rm) * ncg) << t.logS) < (int)blockDim.x) rm >>= 1;
    if (RMAX == 8 && rm == 8)
        gemm_tile<T, SMEM, (RMAX == 8 ? 8 : 4), TN, B_BATCHED>(op, t);
    else if (rm >= 4)
        gemm_tile<T, SMEM, 4, TN, B_BATCHED>(op, t);
    else
        gemm_tile<T, SMEM, 2, TN, B_BATCHED>(op, t);
}

// generic fallback (any shape, nothing staged)
template <typename T, bool SMEM, bool B_BATCHED>
__device__ void gemm_generic(const Op& op, const Tile<T, SMEM>& t, const Prog<T>& p) {
    const int acc = op.w[1], cbase = op.w[3], abase = op.w[5], bbase = op.w[7];
    const int nm = op.w[8], nn = op.w[9], nk = op.w[10];
    const int* am = p.itab + op.w[11];
    const int* cm = p.itab + op.w[12];
    const int* ak = p.itab + op.w[13];
    const int* bk = p.itab + op.w[14];
    const int* bn = p.itab + op.w[15];
    const int* cn = p.itab + op.w[16];
    const int S = t.S;
    const uint32_t es = (uint32_t)sizeof(T);
    const int total = (nm * nn) << t.logS;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int s = i & (S - 1);
        const int rc = i >> t.logS;
        const int r = rc % nm, c = rc / nm;
        const int arow = abase + __ldg(am + r), bcol = bbase + __ldg(bn + c);
        T sum = T(0);
        for (int k = 0; k < nk; ++k) {
            const T a = t.fr.ld((uint32_t)((arow + __ldg(ak + k)) * S + s) * es);
            const T b = B_BATCHED ? t.fr.ld((uint32_t)((bcol + __ldg(bk + k)) * S + s) * es)
                                  : lds<T>(t.cpool + (uint32_t)(bcol + __ldg(bk + k)) * es);
            sum = fma(a, b, sum);
        }
        const uint32_t d = (uint32_t)((cbase + __ldg(cm + r) + __ldg(cn + c)) * S + s) * es;
        t.fr.st(d, acc ? t.fr.ld(d) + sum : sum);
    }
}

template <typename T, bool SMEM, bool B_BATCHED>
__device__ void body_gemm(const Op& op, const Tile<T, SMEM>& t, const Prog<T>& p) {
    const int nm = op.w[8], nn = op.w[9], nk = op.w[10], tn = op.w[20];
    if (2 * (nm + nn + nk) > OTAB || tn <= 0 || (!B_BATCHED && op.w[22] != 1)) {
        gemm_generic<T, SMEM, B_BATCHED>(op, t, p);
        return;
    }
    {   // stage this op's tables, pre-scaled to byte offsets
        const int sb = (int)t.sbytes;
        const int *am = p.itab + op.w[11], *cm = p.itab + op.w[12], *ak = p.itab + op.w[13];
        const int *bk = p.itab + op.w[14], *bn = p.itab + op.w[15], *cn = p.itab + op.w[16];
        const int cbase = op.w[3], abase = op.w[5], bbase = op.w[7];
        int* tab = t.tab;
        for (int i = threadIdx.x; i < nm; i += blockDim.x) {
            tab[i] = (abase + __ldg(am + i)) * sb;
            tab[nm + i] = (cbase + __ldg(cm + i)) * sb;
        }
        tab += 2 * nm;
        for (int i = threadIdx.x; i < nn; i += blockDim.x) {
            tab[i] = B_BATCHED ? (bbase + __ldg(bn + i)) * sb : 0;
            tab[nn + i] = __ldg(cn + i) * sb;
        }
        tab += 2 * nn;
        for (int i = threadIdx.x; i < nk; i += blockDim.x) {
            tab[i] = __ldg(ak + i) * sb;
            tab[nk + i] = B_BATCHED ? __ldg(bk + i) * sb : 0;
        }
    }
    __syncthreads();
    switch (tn) {
        case 1: gemm_rows<T, SMEM, 1, B_BATCHED>(op, t); break;
        case 2: gemm_rows<T, SMEM, 2, B_BATCHED>(op, t); break;
        case 3: gemm_rows<T, SMEM, 3, B_BATCHED>(op, t); break;
        case 4: gemm_rows<T, SMEM, 4, B_BATCHED>(op, t); break;
        case 5: gemm_rows<T, SMEM, 5, B_BATCHED>(op, t); break;
        case 8: gemm_rows<T, SMEM, 8, B_BATCHED>(op, t); break;
        case 9: gemm_rows<T, SMEM, 9, B_BATCHED>(op, t); break;
        default: gemm_generic<T, SMEM, B_BATCHED>(op, t, p); break;
    }
}

// ---- RGEMM --------------------------------------------------------------------------
// G[gk[i]+gn[c]] += sum over the tile's samples and rows r of A[am[r]+ak[i]] * D[dm[r]+dn[c]].
// One warp owns an RK x RN output tile; its lanes split the (row, sample) pairs and combine
// with a shuffle tree, so the accumulation order is fixed (deterministic).
// Staged tables (byte offsets): am[nm], dm[nm].
template <typename T, bool SMEM, int RK, int RN>
__device__ __forceinline__ void rgemm_tile(const Op& op, const Tile<T, SMEM>& t, const Prog<T>& p) {
    const int gbase = op.w[3], abase = op.w[5], dbase = op.w[7];
    const int nm = op.w[8], nn = op.w[9], nk = op.w[10];
    const int* ak = p.itab + op.w[13];
    const int* dn = p.itab + op.w[14];
    const int* gk = p.itab + op.w[15];
    const int* gn = p.itab + op.w[16];
    const int* am = t.tab;
    const int* dm = am + nm;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int ntn = nn / RN, tiles = (nk / RK) * ntn;
    const int rows = nm << t.logS;
    const int sb = (int)t.sbytes;
    for (int tile = warp; tile < tiles; tile += nwarp) {
        const int i0 = (tile / ntn) * RK, c0 = (tile % ntn) * RN;
        uint32_t aoff[RK], doff[RN];
#pragma unroll
        for (int i = 0; i < RK; ++i) aoff[i] = (uint32_t)((abase + __ldg(ak + i0 + i)) * sb);
#pragma unroll
        for (int j = 0; j < RN; ++j) doff[j] = (uint32_t)((dbase + __ldg(dn + c0 + j)) * sb);
        T sum[RK][RN];
#pragma unroll
        for (int i = 0; i < RK; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) sum[i][j] = T(0);
#pragma unroll 2
        for (int idx = lane; idx < rows; idx += 32) {
            const uint32_t s4 = (uint32_t)(idx & (t.S - 1)) * (uint32_t)sizeof(T);
            const int r = idx >> t.logS;
            const uint32_t ra = (uint32_t)am[r] + s4, rd = (uint32_t)dm[r] + s4;
            T a[RK], d[RN];
#pragma unroll
            for (int i = 0; i < RK; ++i) a[i] = t.fr.ld(aoff[i] + ra);
#pragma unroll
            for (int j = 0; j < RN; ++j) d[j] = t.fr.ld(doff[j] + rd);
#pragma unroll
            for (int i = 0; i < RK; ++i)
#pragma unroll
                for (int j = 0; j < RN; ++j) sum[i][j] = fma(a[i], d[j], sum[i][j]);
        }
#pragma unroll
        for (int i = 0; i < RK; ++i)
#pragma unroll
            for (int j = 0; j < RN; ++j) sum[i][j] = warp_sum(sum[i][j]);
        if (lane == 0) {
#pragma unroll
            for (int i = 0; i < RK; ++i) {
                const int grow = gbase + __ldg(gk + i0 + i);
#pragma unroll
                for (int j = 0; j < RN; ++j) t.gacc[grow + __ldg(gn + c0 + j)] += sum[i][j];
            }
        }
    }
}

template <typename T, bool SMEM, int RK>
__device__ __forceinline__ void rgemm_cols(int rn, const Op& op, const Tile<T, SMEM>& t, const Prog<T>& p) {
    switch (rn) {
        case 4: rgemm_tile<T, SMEM, RK, 4>(op, t, p); break;
        case 3: rgemm_tile<T, SMEM, RK, 3>(op, t, p); break;
        case 2: rgemm_tile<T, SMEM, RK, 2>(op, t, p); break;
        default: rgemm_tile<T, SMEM, RK, 1>(op, t, p); break;
    }
}

template <typename T, bool SMEM>
__device__ void body_rgemm(const Op& op, const Tile<T, SMEM>& t, const Prog<T>& p) {
    const int nm = op.w[8];
    int rk = op.w[20], rn = op.w[21];
    if (rk < 1 || rk > 4 || op.w[10] % rk) rk = 1;
    if (rn < 1 || rn > 4 || op.w[9] % rn) rn = 1;
    const int* am = p.itab + op.w[11];
    const int* dm = p.itab + op.w[12];
    const int sb = (int)t.sbytes;
    if (2 * nm <= OTAB) {
        for (int r = threadIdx.x; r < nm; r += blockDim.x) {
            t.tab[r] = __ldg(am + r) * sb;
            t.tab[nm + r] = __ldg(dm + r) * sb;
        }
        __syncthreads();
        switch (rk) {
            case 4: rgemm_cols<T, SMEM, 4>(rn, op, t, p); break;
            case 3: rgemm_cols<T, SMEM, 3>(rn, op, t, p); break;
            case 2: rgemm_cols<T, SMEM, 2>(rn, op, t, p); break;
            default: rgemm_cols<T, SMEM, 1>(rn, op, t, p); break;
        }
        return;
    }
    // huge row counts: unstaged, one output per warp pass
    const int gbase = op.w[3], abase = op.w[5], dbase = op.w[7], nn = op.w[9], nk = op.w[10];
    const int *ak = p.itab + op.w[13], *dn = p.itab + op.w[14], *gk = p.itab + op.w[15], *gn = p.itab + op.w[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int rows = nm << t.logS;
    for (int o = warp; o < nk * nn; o += nwarp) {
        const int i = o / nn, c = o % nn;
        const int ao = abase + __ldg(ak + i), dof = dbase + __ldg(dn + c);
        T sum = T(0);
        for (int idx = lane; idx < rows; idx += 32) {
            const int s = idx & (t.S - 1), r = idx >> t.logS;
            sum = fma(t.fr.ld((uint32_t)((ao + __ldg(am + r)) * t.S + s) * (uint32_t)sizeof(T)),
                      t.fr.ld((uint32_t)((dof + __ldg(dm + r)) * t.S + s) * (uint32_t)sizeof(T)), sum);
        }
        sum = warp_sum(sum);
        if (lane == 0) t.gacc[gbase + __ldg(gk + i) + __ldg(gn + c)] += sum;
    }
}

// Fused loss (engine_siamese.py:490-530): value = v or |v|^2 ; loss += -(log(max(value,1e-10)) +
// log_scale) / count ; d value = -1/(count*value) where value >= 1e-10 (torch.clamp passes the
// gradient on the boundary), chained through |v|^2 for complex amplitudes.
template <typename T, bool SMEM>
__device__ void body_seed(const Op& op, const Tile<T, SMEM>& t) {
    const int cplx = op.w[2], vb = op.w[3], dvb = op.w[4], lb = op.w[5];
    const int S = t.S;
    const uint32_t es = (uint32_t)sizeof(T);
    if ((threadIdx.x >> 5) == 0) {
        T part = T(0);
        for (int s = threadIdx.x; s < S; s += 32) {
            T vr = t.fr.ld((uint32_t)(vb * S + s) * es), vi = T(0), val;
            if (cplx) {
                vi = t.fr.ld((uint32_t)((vb + 1) * S + s) * es);
                val = vr * vr + vi * vi;
            } else {
                val = vr;
            }
            const bool valid = s < t.nvalid;
            const T clamped = val > T(1e-10) ? val : T(1e-10);
            if (valid) part -= (log(clamped) + t.log_scale) * t.inv_count;
            const T dval = (valid && val >= T(1e-10)) ? -t.inv_count / clamped : T(0);
            if (cplx) {
                t.fr.st((uint32_t)(dvb * S + s) * es, dval * T(2) * vr);
                t.fr.st((uint32_t)((dvb + 1) * S + s) * es, dval * T(2) * vi);
            } else {
                t.fr.st((uint32_t)(dvb * S + s) * es, dval);
            }
        }
        part = warp_sum(part);
        if (threadIdx.x == 0) t.gacc[lb] += part;
    }
}

template <typename T, bool FRAME_SMEM>
__global__ void __launch_bounds__(BODY_THREADS_MAX, 1)
tnq_body_kernel(Prog<T> p, const __grid_constant__ RunArgs args, const T* __restrict__ constg,
                T* __restrict__ partials, T* __restrict__ frame_g, long long nsamples, int S, int logS,
                long long ntiles, T log_scale, T inv_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int otab[OTAB];
    T* cpool = reinterpret_cast<T*>(smem_raw);
    T* gacc = cpool + ((p.const_elems + 3) & ~3);
    T* frame_s = gacc + ((p.gacc_elems + 3) & ~3);
    for (int i = threadIdx.x; i < p.const_elems; i += blockDim.x) cpool[i] = constg[i];
    for (int i = threadIdx.x; i < p.gacc_elems; i += blockDim.x) gacc[i] = T(0);
    Tile<T, FRAME_SMEM> t;
    if constexpr (FRAME_SMEM)
        t.fr.base = (uint32_t)__cvta_generic_to_shared(frame_s);
    else
        t.fr.base = reinterpret_cast<char*>(frame_g + (size_t)blockIdx.x * p.frame_elems * S);
    t.cpool = (uint32_t)__cvta_generic_to_shared(cpool);
    t.gacc = gacc;
    t.tab = otab;
    t.S = S;
    t.logS = logS;
    t.sbytes = (uint32_t)S * (uint32_t)sizeof(T);
    t.nb = p.nb;
    t.log_scale = log_scale;
    t.inv_count = inv_count;
    __syncthreads();
    for (long long tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        t.s0 = tile * S;
        const long long left = nsamples - t.s0;
        t.nvalid = left < S ? (int)left : S;
        for (int o = 0; o < p.n_ops; ++o) {
            const Op& op = p.ops[o];
            switch (op.w[0]) {
                case OP_LIN:
                    body_lin<T, FRAME_SMEM>(op, t, p, args);
                    break;
                case OP_GEMM:
                    if (op.w[6] == SP_FRAME)
                        body_gemm<T, FRAME_SMEM, true>(op, t, p);
                    else
                        body_gemm<T, FRAME_SMEM, false>(op, t, p);
                    break;
                case OP_RGEMM:
                    body_rgemm<T, FRAME_SMEM>(op, t, p);
                    break;
                case OP_SEED:
                    body_seed<T, FRAME_SMEM>(op, t);
                    break;
                default:
                    break;
            }
            __syncthreads();
        }
    }
    if (p.gacc_elems > 0 && partials != nullptr) {
        for (int i = threadIdx.x; i < p.gacc_elems; i += blockDim.x)
            partials[(size_t)blockIdx.x * p.gacc_elems + i] = gacc[i];
    }
}

// ------------------------------------------------------------------------------------
// PREP / FIN : batch independent, one CTA, generic addressing
// ------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ T* resolve(int space, int base, int slot, const RunArgs& a, T* constg, T* gaccg) {
    switch (space) {
        case SP_CONST:
            return constg + base;
        case SP_GACC:
            return gaccg + base;
        case SP_GIN:
            return const_cast<T*>(reinterpret_cast<const T*>(a.in_ptr[slot])) + base;
        case SP_GOUT:
            return reinterpret_cast<T*>(a.out_ptr[slot]) + base;
        default:
            return nullptr;
    }
}

template <typename T>
__device__ void shared_ops(const Prog<T>& p, const RunArgs& args, T* constg, T* gaccg) {
    for (int o = 0; o < p.n_ops; ++o) {
        const Op& op = p.ops[o];
        const int acc = op.w[1];
        if (op.w[0] == OP_LIN) {
            T* dst = resolve<T>(op.w[2], op.w[3], op.w[14], args, constg, gaccg);
            const T* src = resolve<T>(op.w[4], op.w[5], op.w[13], args, constg, gaccg);
            const int count = op.w[6], nt = op.w[7];
            const int *s0 = p.itab + op.w[9], *s1 = p.itab + op.w[11];
            const T *c0 = p.ftab + op.w[10], *c1 = p.ftab + op.w[12];
            for (int j = threadIdx.x; j < count; j += blockDim.x) {
                T v = c0[j] * src[s0[j]];
                if (nt == 2) v += c1[j] * src[s1[j]];
                dst[j] = acc ? dst[j] + v : v;
            }
        } else if (op.w[0] == OP_GEMM) {
            T* C = resolve<T>(op.w[2], op.w[3], op.w[19], args, constg, gaccg);
            const T* A = resolve<T>(op.w[4], op.w[5], op.w[17], args, constg, gaccg);
            const T* B = resolve<T>(op.w[6], op.w[7], op.w[18], args, constg, gaccg);
            const int nm = op.w[8], nn = op.w[9], nk = op.w[10];
            const int *am = p.itab + op.w[11], *cm = p.itab + op.w[12], *ak = p.itab + op.w[13];
            const int *bk = p.itab + op.w[14], *bn = p.itab + op.w[15], *cn = p.itab + op.w[16];
            for (int i = threadIdx.x; i < nm * nn; i += blockDim.x) {
                const int r = i / nn, c = i % nn;
                const T* a = A + am[r];
                const T* b = B + bn[c];
                T sum = T(0);
                for (int k = 0; k < nk; ++k) sum = fma(a[ak[k]], b[bk[k]], sum);
                T* d = C + cm[r] + cn[c];
                *d = acc ? *d + sum : sum;
            }
        }
        __syncthreads();
    }
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
tnq_prep_kernel(Prog<T> p, const __grid_constant__ RunArgs args, T* constg) {
    shared_ops<T>(p, args, constg, nullptr);
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
tnq_reduce_kernel(const T* __restrict__ partials, T* __restrict__ gaccg, int gacc_elems, int nparts) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= gacc_elems) return;
    T sum = T(0);
    for (int b = 0; b < nparts; ++b) sum += partials[(size_t)b * gacc_elems + j];
    gaccg[j] = sum;
}

template <typename T>
__global__ void __launch_bounds__(THREADS)
tnq_fin_kernel(Prog<T> p, const __grid_constant__ RunArgs args, T* constg, T* gaccg) {
    shared_ops<T>(p, args, constg, gaccg);
}

}  // namespace

// shared with the other translation units of the library (tnq_gemm.cu, tnq_permute.cu)
int tnq_internal_fail(const std::string& msg) { return fail(msg); }
int tnq_internal_cuda_fail(cudaError_t e, const char* what) { return cuda_fail(e, what); }
void tnq_internal_count_launch() { ++g_launches; }

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
struct tnq_plan {
    int dtype = 0;  // 0 f32, 1 f64
    int n_in = 0, n_out = 0;
    int const_elems = 0, frame_elems = 0, gacc_elems = 0;
    int n_prep = 0, n_body = 0, n_fin = 0;
    int nb = 1;
    std::vector<unsigned char> in_batched, out_batched;
    std::vector<int> out_elems;
    int device = 0;
    int sm_count = 0;
    int max_smem = 0;
    Op* d_ops = nullptr;
    int* d_itab = nullptr;
    void* d_ftab = nullptr;
    size_t elem_size() const { return dtype == 0 ? 4 : 8; }
};

namespace {

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct Geometry {
    int S = 0, logS = 0, grid = 0, threads = 256;
    bool frame_smem = true;
    size_t smem = 0;
    long long ntiles = 0;
    size_t off_const = 0, off_gacc = 0, off_part = 0, off_frame = 0, total = 0;
    int launches = 0;
};

int plan_geometry(const tnq_plan* pl, long long nsamples, Geometry* g) {
    if (nsamples <= 0) return fail("nsamples must be positive");
    const size_t es = pl->elem_size();
    const size_t fixed = (align_up(pl->const_elems, 4) + align_up(pl->gacc_elems, 4)) * es;
    const size_t budget = (size_t)pl->max_smem - OTAB * sizeof(int) - 1024;
    if (fixed > budget)
        return fail("plan too large for the shared-memory contraction kernel: prepared cores + gradient "
                    "accumulators need " + std::to_string(fixed) + " bytes of shared memory");
    // Samples per tile.  Every op of the program costs two barriers and a table staging per
    // tile, so tiles should be as large as possible while still giving every SM a tile; a large
    // tile whose working set does not fit shared memory keeps its FRAME in a per-CTA global slab
    // (L2 resident), which measured faster than small shared-memory tiles ("tiling").
    int S = 256;
    while (S > 8 && nsamples / S < pl->sm_count) S >>= 1;
    const size_t frame_per_sample = (size_t)(pl->frame_elems > 0 ? pl->frame_elems : 1) * es;
    g->frame_smem = fixed + frame_per_sample * S <= budget;
    if (!g->frame_smem && S > 64) S = 64;
    // experiment knobs (not part of the API): TNQ_TILE=<pow2>, TNQ_FORCE_GLOBAL=1, TNQ_CTAS_PER_SM=<n>
    if (const char* e = getenv("TNQ_FORCE_GLOBAL")) { if (atoi(e)) g->frame_smem = false; }
    if (const char* e = getenv("TNQ_TILE")) {
        int v = atoi(e);
        if (v >= 4 && (v & (v - 1)) == 0 && (!g->frame_smem || fixed + frame_per_sample * v <= budget)) S = v;
    }
    g->S = S;
    g->logS = 0;
    while ((1 << g->logS) < S) ++g->logS;
    g->ntiles = (nsamples + S - 1) / S;
    g->smem = fixed + (g->frame_smem ? frame_per_sample * S : 0);
    int per_sm = (int)(budget / (g->smem + OTAB * sizeof(int) + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    g->threads = per_sm == 1 ? 512 : 256;   // registers allow 512 x 128 or 2 x 256 x 128 per SM
    if (per_sm > 2) per_sm = 2;
    if (const char* e = getenv("TNQ_CTAS_PER_SM")) { int v = atoi(e); if (v >= 1 && v <= 16) per_sm = v; }
    if (const char* e = getenv("TNQ_THREADS")) { int v = atoi(e); if (v >= 32 && v <= 512 && v % 32 == 0) g->threads = v; }
    long long grid = (long long)pl->sm_count * per_sm;
    if (grid > g->ntiles) grid = g->ntiles;
    g->grid = (int)grid;
    size_t off = 0;
    g->off_const = off;
    off = align_up(off + (size_t)pl->const_elems * es, 256);
    g->off_gacc = off;
    off = align_up(off + (size_t)pl->gacc_elems * es, 256);
    g->off_part = off;
    off = align_up(off + (size_t)pl->gacc_elems * es * g->grid, 256);
    g->off_frame = off;
    if (!g->frame_smem) off = align_up(off + frame_per_sample * S * g->grid, 256);
    g->total = off + 256;
    g->launches = (pl->n_prep > 0 ? 1 : 0) + 1 + (pl->gacc_elems > 0 ? 1 : 0) + (pl->n_fin > 0 ? 1 : 0);
    return 0;
}

template <typename T>
int run_typed(tnq_plan* pl, long long nsamples, const RunArgs& args, const double* scalars, unsigned char* ws,
              const Geometry& g, cudaStream_t stream) {
    T* constg = reinterpret_cast<T*>(ws + g.off_const);
    T* gaccg = reinterpret_cast<T*>(ws + g.off_gacc);
    T* partials = reinterpret_cast<T*>(ws + g.off_part);
    T* frame_g = reinterpret_cast<T*>(ws + g.off_frame);
    Prog<T> p;
    p.itab = pl->d_itab;
    p.ftab = reinterpret_cast<const T*>(pl->d_ftab);
    p.const_elems = pl->const_elems;
    p.gacc_elems = pl->gacc_elems;
    p.frame_elems = pl->frame_elems;
    p.nb = pl->nb;
    if (pl->n_prep > 0) {
        p.ops = pl->d_ops;
        p.n_ops = pl->n_prep;
        tnq_prep_kernel<T><<<1, THREADS, 0, stream>>>(p, args, constg);
        ++g_launches;
    }
    p.ops = pl->d_ops + pl->n_prep;
    p.n_ops = pl->n_body;
    const T ls = (T)scalars[0], ic = (T)scalars[1];
    if (g.frame_smem) {
        auto k = tnq_body_kernel<T, true>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        k<<<g.grid, g.threads, g.smem, stream>>>(p, args, constg, pl->gacc_elems ? partials : nullptr, frame_g,
                                               nsamples, g.S, g.logS, g.ntiles, ls, ic);
    } else {
        auto k = tnq_body_kernel<T, false>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        k<<<g.grid, g.threads, g.smem, stream>>>(p, args, constg, pl->gacc_elems ? partials : nullptr, frame_g,
                                               nsamples, g.S, g.logS, g.ntiles, ls, ic);
    }
    ++g_launches;
    if (pl->gacc_elems > 0) {
        const int blocks = (pl->gacc_elems + THREADS - 1) / THREADS;
        tnq_reduce_kernel<T><<<blocks, THREADS, 0, stream>>>(partials, gaccg, pl->gacc_elems, g.grid);
        ++g_launches;
    }
    if (pl->n_fin > 0) {
        p.ops = pl->d_ops + pl->n_prep + pl->n_body;
        p.n_ops = pl->n_fin;
        tnq_fin_kernel<T><<<1, THREADS, 0, stream>>>(p, args, constg, gaccg);
        ++g_launches;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "kernel launch");
    return 0;
}

}  // namespace

extern "C" {

const char* tnq_last_error(void) { return g_error.c_str(); }

int64_t tnq_launch_count(void) { return g_launches.load(); }

int tnq_device_check(void) {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDevice");
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, dev);
    if (e != cudaSuccess) return cuda_fail(e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail("tneq_b200 is built for sm_100a (Blackwell B200) only; current device is sm_" +
                    std::to_string(prop.major) + std::to_string(prop.minor));
    return 0;
}

int tnq_plan_create(const int64_t* blob, int64_t nwords, tnq_plan_t** out) {
    if (!blob || !out || nwords < 16) return fail("tnq_plan_create: bad arguments");
    if (blob[0] != MAGIC || blob[1] != 1) return fail("tnq_plan_create: not a tneq_b200 program (magic/version)");
    tnq_plan* pl = new tnq_plan();
    pl->dtype = (int)blob[2];
    pl->n_in = (int)blob[3];
    pl->n_out = (int)blob[4];
    pl->const_elems = (int)blob[5];
    pl->frame_elems = (int)blob[6];
    pl->gacc_elems = (int)blob[7];
    pl->n_prep = (int)blob[8];
    pl->n_body = (int)blob[9];
    pl->n_fin = (int)blob[10];
    const int64_t n_itab = blob[11], n_ftab = blob[12];
    pl->nb = (int)blob[14];
    if (pl->n_in > TNQ_MAX_INPUTS || pl->n_out > TNQ_MAX_OUTPUTS) {
        delete pl;
        return fail("tnq_plan_create: too many operands (" + std::to_string(blob[3]) + " inputs, " +
                    std::to_string(blob[4]) + " outputs)");
    }
    int64_t at = 16;
    const int64_t nops = (int64_t)pl->n_prep + pl->n_body + pl->n_fin;
    const int64_t need = at + 2 * (pl->n_in + pl->n_out) + nops * OP_WORDS + n_itab + n_ftab;
    if (need != nwords) {
        delete pl;
        return fail("tnq_plan_create: blob length mismatch");
    }
    for (int i = 0; i < pl->n_in; ++i, at += 2) pl->in_batched.push_back((unsigned char)blob[at]);
    for (int i = 0; i < pl->n_out; ++i, at += 2) {
        pl->out_batched.push_back((unsigned char)blob[at]);
        pl->out_elems.push_back((int)blob[at + 1]);
    }
    std::vector<Op> ops((size_t)nops);
    for (int64_t o = 0; o < nops; ++o)
        for (int w = 0; w < OP_WORDS; ++w) ops[(size_t)o].w[w] = (int)blob[at + o * OP_WORDS + w];
    at += nops * OP_WORDS;
    std::vector<int> itab((size_t)n_itab);
    for (int64_t i = 0; i < n_itab; ++i) itab[(size_t)i] = (int)blob[at + i];
    at += n_itab;
    cudaError_t e = cudaGetDevice(&pl->device);
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, pl->device);
    if (e != cudaSuccess) {
        delete pl;
        return cuda_fail(e, "tnq_plan_create: no CUDA device");
    }
    pl->sm_count = prop.multiProcessorCount;
    pl->max_smem = (int)prop.sharedMemPerBlockOptin;
    const size_t es = pl->elem_size();
    e = cudaMalloc(&pl->d_ops, sizeof(Op) * (size_t)(nops > 0 ? nops : 1));
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_itab, sizeof(int) * (size_t)(n_itab > 0 ? n_itab : 1));
    if (e == cudaSuccess) e = cudaMalloc(&pl->d_ftab, es * (size_t)(n_ftab > 0 ? n_ftab : 1));
    if (e == cudaSuccess && nops) e = cudaMemcpy(pl->d_ops, ops.data(), sizeof(Op) * (size_t)nops, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_itab)
        e = cudaMemcpy(pl->d_itab, itab.data(), sizeof(int) * (size_t)n_itab, cudaMemcpyHostToDevice);
    if (e == cudaSuccess && n_ftab) {
        const double* f = reinterpret_cast<const double*>(blob + at);
        if (pl->dtype == 0) {
            std::vector<float> tmp((size_t)n_ftab);
            for (int64_t i = 0; i < n_ftab; ++i) tmp[(size_t)i] = (float)f[i];
            e = cudaMemcpy(pl->d_ftab, tmp.data(), 4 * (size_t)n_ftab, cudaMemcpyHostToDevice);
        } else {
            e = cudaMemcpy(pl->d_ftab, f, 8 * (size_t)n_ftab, cudaMemcpyHostToDevice);
        }
    }
    if (e != cudaSuccess) {
        tnq_plan_destroy(pl);
        return cuda_fail(e, "tnq_plan_create: upload");
    }
    *out = pl;
    return 0;
}

void tnq_plan_destroy(tnq_plan_t* pl) {
    if (!pl) return;
    if (pl->d_ops) cudaFree(pl->d_ops);
    if (pl->d_itab) cudaFree(pl->d_itab);
    if (pl->d_ftab) cudaFree(pl->d_ftab);
    delete pl;
}

int tnq_plan_num_inputs(const tnq_plan_t* pl) { return pl ? pl->n_in : -1; }
int tnq_plan_num_outputs(const tnq_plan_t* pl) { return pl ? pl->n_out : -1; }

int tnq_plan_query(const tnq_plan_t* pl, int64_t nsamples, tnq_run_info_t* info) {
    if (!pl || !info) return fail("tnq_plan_query: bad arguments");
    Geometry g;
    if (int rc = plan_geometry(pl, nsamples, &g)) return rc;
    info->tile_samples = g.S;
    info->grid = g.grid;
    info->frame_in_smem = g.frame_smem ? 1 : 0;
    info->launches = g.launches;
    info->smem_bytes = (int64_t)g.smem;
    info->workspace_bytes = (int64_t)g.total;
    return 0;
}

int tnq_plan_run(tnq_plan_t* pl, int64_t nsamples, const void* const* in_ptrs, const int64_t* in_stride_hi,
                 const int64_t* in_stride_lo, void* const* out_ptrs, const double* scalars, void* workspace,
                 int64_t workspace_bytes, void* stream) {
    if (!pl || !in_ptrs || !out_ptrs || !scalars) return fail("tnq_plan_run: bad arguments");
    Geometry g;
    if (int rc = plan_geometry(pl, nsamples, &g)) return rc;
    if (!workspace || (size_t)workspace_bytes < g.total)
        return fail("tnq_plan_run: workspace too small (" + std::to_string(workspace_bytes) + " < " +
                    std::to_string(g.total) + ")");
    RunArgs args;
    memset(&args, 0, sizeof(args));
    for (int i = 0; i < pl->n_in; ++i) {
        if (!in_ptrs[i]) return fail("tnq_plan_run: null input pointer at slot " + std::to_string(i));
        args.in_ptr[i] = in_ptrs[i];
        args.in_hi[i] = in_stride_hi ? in_stride_hi[i] : 0;
        args.in_lo[i] = in_stride_lo ? in_stride_lo[i] : 0;
        args.in_batched[i] = pl->in_batched[(size_t)i];
    }
    for (int i = 0; i < pl->n_out; ++i) {
        if (!out_ptrs[i]) return fail("tnq_plan_run: null output pointer at slot " + std::to_string(i));
        args.out_ptr[i] = out_ptrs[i];
        args.out_elems[i] = pl->out_elems[(size_t)i];
        args.out_batched[i] = pl->out_batched[(size_t)i];
    }
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    if (pl->dtype == 0) return run_typed<float>(pl, nsamples, args, scalars, ws, g, st);
    return run_typed<double>(pl, nsamples, args, scalars, ws, g, st);
}

}  // extern "C"
