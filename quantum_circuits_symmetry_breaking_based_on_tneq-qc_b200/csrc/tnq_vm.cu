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

#include <atomic>
#include <string>
#include <vector>

#include "tneq_b200.h"

namespace {

constexpr int OP_WORDS = 24;
constexpr int SP_CONST = 0, SP_FRAME = 1, SP_GIN = 2, SP_GOUT = 3, SP_GACC = 4;
constexpr int OP_LIN = 1, OP_GEMM = 2, OP_RGEMM = 3, OP_SEED = 4;
constexpr int THREADS = 256;
constexpr int KTAB = 1024;  // staged (ak, bk) pairs per GEMM
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
template <typename T>
struct Tile {
    T* cpool;        // CONST pool (shared memory copy)
    T* gacc;         // per-CTA accumulators (shared memory)
    T* frame;        // [frame_elems][S]
    int2* ktab;      // staged per-k offsets of the current GEMM
    int S, logS;
    int nvalid;      // valid samples in this tile
    long long s0;    // first global sample of the tile
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

template <typename T>
__device__ void body_lin(const Op& op, const Tile<T>& t, const Prog<T>& p, const RunArgs& args) {
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
#pragma unroll 2
        for (int term = 0; term < nt; ++term) {
            const int so = term == 0 ? __ldg(s0 + j) : __ldg(s1 + j);
            const T cf = term == 0 ? __ldg(c0 + j) : __ldg(c1 + j);
            T x;
            if (ssp == SP_FRAME) {
                x = t.frame[(size_t)(sbase + so) * S + s];
            } else if (ssp == SP_CONST) {
                x = t.cpool[sbase + so];
            } else if (src_batched_gin) {
                x = valid ? __ldg(gsrc + sample_offset<T>(args, sslot, t.s0 + s, t.nb) + so) : T(0);
            } else {
                x = __ldg(gsrc + so);
            }
            v += cf * x;
        }
        if (dsp == SP_FRAME) {
            T* d = t.frame + (size_t)(dbase + j) * S + s;
            *d = acc ? *d + v : v;
        } else if (valid) {  // batched GOUT, [nsamples][elems]
            T* d = gdst + (t.s0 + s) * (long long)delems + j;
            *d = acc ? *d + v : v;
        }
    }
}

// C[cm[r]+cn[c]] (=|+=) sum_k A[am[r]+ak[k]] * B[bk[k]+bn[c]] for every sample of the tile.
// Lanes run over samples, so A/C accesses are conflict free and a shared B is a broadcast.
template <typename T, bool B_BATCHED>
__device__ void body_gemm(const Op& op, const Tile<T>& t, const Prog<T>& p) {
    const int acc = op.w[1], cbase = op.w[3], abase = op.w[5], bbase = op.w[7];
    const int nm = op.w[8], nn = op.w[9], nk = op.w[10];
    const int* am = p.itab + op.w[11];
    const int* cm = p.itab + op.w[12];
    const int* ak = p.itab + op.w[13];
    const int* bk = p.itab + op.w[14];
    const int* bn = p.itab + op.w[15];
    const int* cn = p.itab + op.w[16];
    const int S = t.S;
    const bool staged = nk <= KTAB;
    if (staged) {
        for (int k = threadIdx.x; k < nk; k += blockDim.x)
            t.ktab[k] = make_int2(__ldg(ak + k) * S, B_BATCHED ? __ldg(bk + k) * S : __ldg(bk + k));
    }
    __syncthreads();
    const int nchunk = (nn + 3) >> 2;
    const int total = (nchunk * nm) << t.logS;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int s = i & (S - 1);
        const int rc = i >> t.logS;
        const int r = rc % nm, c0 = (rc / nm) << 2;
        const T* a_ptr = t.frame + (size_t)(abase + __ldg(am + r)) * S + s;
        int bo[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int c = min(c0 + j, nn - 1);
            bo[j] = B_BATCHED ? (bbase + __ldg(bn + c)) * S + s : bbase + __ldg(bn + c);
        }
        const T* b_ptr = B_BATCHED ? t.frame : t.cpool;
        T sum[4] = {T(0), T(0), T(0), T(0)};
        if (staged) {
#pragma unroll 3
            for (int k = 0; k < nk; ++k) {
                const int2 kb = t.ktab[k];
                const T a = a_ptr[kb.x];
#pragma unroll
                for (int j = 0; j < 4; ++j) sum[j] = fma(a, b_ptr[bo[j] + kb.y], sum[j]);
            }
        } else {
            for (int k = 0; k < nk; ++k) {
                const int ka = __ldg(ak + k) * S, kbv = B_BATCHED ? __ldg(bk + k) * S : __ldg(bk + k);
                const T a = a_ptr[ka];
#pragma unroll
                for (int j = 0; j < 4; ++j) sum[j] = fma(a, b_ptr[bo[j] + kbv], sum[j]);
            }
        }
        const int crow = cbase + __ldg(cm + r);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            if (c0 + j < nn) {
                T* d = t.frame + (size_t)(crow + __ldg(cn + c0 + j)) * S + s;
                *d = acc ? *d + sum[j] : sum[j];
            }
        }
    }
}

// G[gk[i]+gn[c]] += sum over the tile's samples and rows r of A[am[r]+ak[i]] * D[dm[r]+dn[c]].
// One warp owns an output strip; its lanes split (row, sample) pairs and combine
// with a shuffle tree, so the accumulation order is fixed (deterministic).
template <typename T>
__device__ void body_rgemm(const Op& op, const Tile<T>& t, const Prog<T>& p) {
    const int gbase = op.w[3], abase = op.w[5], dbase = op.w[7];
    const int nm = op.w[8], nn = op.w[9], nk = op.w[10];
    const int* am = p.itab + op.w[11];
    const int* dm = p.itab + op.w[12];
    const int* ak = p.itab + op.w[13];
    const int* dn = p.itab + op.w[14];
    const int* gk = p.itab + op.w[15];
    const int* gn = p.itab + op.w[16];
    const int S = t.S;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    const int nchunk = (nn + 3) >> 2;
    const int rows = nm << t.logS;
    for (int tile = warp; tile < nk * nchunk; tile += nwarp) {
        const int i = tile / nchunk, c0 = (tile % nchunk) << 2;
        const int aoff = abase + __ldg(ak + i);
        int doff[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) doff[j] = dbase + __ldg(dn + min(c0 + j, nn - 1));
        T sum[4] = {T(0), T(0), T(0), T(0)};
        for (int idx = lane; idx < rows; idx += 32) {
            const int s = idx & (S - 1), r = idx >> t.logS;
            const T a = t.frame[(size_t)(aoff + __ldg(am + r)) * S + s];
            const int dr = __ldg(dm + r);
#pragma unroll
            for (int j = 0; j < 4; ++j) sum[j] = fma(a, t.frame[(size_t)(doff[j] + dr) * S + s], sum[j]);
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) sum[j] = warp_sum(sum[j]);
        if (lane == 0) {
            const int grow = gbase + __ldg(gk + i);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (c0 + j < nn) t.gacc[grow + __ldg(gn + c0 + j)] += sum[j];
        }
    }
}

// Fused loss (engine_siamese.py:490-530): value = v or |v|^2 ; loss += -(log(max(value,1e-10)) +
// log_scale) / count ; d value = -1/(count*value) where value >= 1e-10 (torch.clamp passes the
// gradient on the boundary), chained through |v|^2 for complex amplitudes.
template <typename T>
__device__ void body_seed(const Op& op, const Tile<T>& t) {
    const int cplx = op.w[2], vb = op.w[3], dvb = op.w[4], lb = op.w[5];
    const int S = t.S;
    if ((threadIdx.x >> 5) == 0) {
        T part = T(0);
        for (int s = threadIdx.x; s < S; s += 32) {
            T vr = t.frame[(size_t)vb * S + s], vi = T(0), val;
            if (cplx) {
                vi = t.frame[(size_t)(vb + 1) * S + s];
                val = vr * vr + vi * vi;
            } else {
                val = vr;
            }
            const bool valid = s < t.nvalid;
            const T clamped = val > T(1e-10) ? val : T(1e-10);
            if (valid) part -= (log(clamped) + t.log_scale) * t.inv_count;
            const T dval = (valid && val >= T(1e-10)) ? -t.inv_count / clamped : T(0);
            if (cplx) {
                t.frame[(size_t)dvb * S + s] = dval * T(2) * vr;
                t.frame[(size_t)(dvb + 1) * S + s] = dval * T(2) * vi;
            } else {
                t.frame[(size_t)dvb * S + s] = dval;
            }
        }
        part = warp_sum(part);
        if (threadIdx.x == 0) t.gacc[lb] += part;
    }
}

template <typename T, bool FRAME_SMEM>
__global__ void __launch_bounds__(THREADS)
tnq_body_kernel(Prog<T> p, const __grid_constant__ RunArgs args, const T* __restrict__ constg,
                T* __restrict__ partials, T* __restrict__ frame_g, long long nsamples, int S, int logS,
                long long ntiles, T log_scale, T inv_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int2 ktab[KTAB];
    T* cpool = reinterpret_cast<T*>(smem_raw);
    T* gacc = cpool + ((p.const_elems + 3) & ~3);
    T* frame_s = gacc + ((p.gacc_elems + 3) & ~3);
    for (int i = threadIdx.x; i < p.const_elems; i += blockDim.x) cpool[i] = constg[i];
    for (int i = threadIdx.x; i < p.gacc_elems; i += blockDim.x) gacc[i] = T(0);
    Tile<T> t;
    t.cpool = cpool;
    t.gacc = gacc;
    t.frame = FRAME_SMEM ? frame_s : frame_g + (size_t)blockIdx.x * p.frame_elems * S;
    t.ktab = ktab;
    t.S = S;
    t.logS = logS;
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
                    body_lin<T>(op, t, p, args);
                    break;
                case OP_GEMM:
                    if (op.w[6] == SP_FRAME)
                        body_gemm<T, true>(op, t, p);
                    else
                        body_gemm<T, false>(op, t, p);
                    break;
                case OP_RGEMM:
                    body_rgemm<T>(op, t, p);
                    break;
                case OP_SEED:
                    body_seed<T>(op, t);
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
    int S = 0, logS = 0, grid = 0;
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
    const size_t budget = (size_t)pl->max_smem - KTAB * sizeof(int2) - 1024;
    if (fixed > budget)
        return fail("plan too large for the shared-memory contraction kernel: prepared cores + gradient "
                    "accumulators need " + std::to_string(fixed) + " bytes of shared memory");
    // samples per tile: as many as fit, but keep at least one tile per SM
    int want = 128;
    while (want > 8 && nsamples / want < pl->sm_count) want >>= 1;
    int S = want;
    const size_t frame_per_sample = (size_t)(pl->frame_elems > 0 ? pl->frame_elems : 1) * es;
    while (S >= 4 && fixed + frame_per_sample * S > budget) S >>= 1;
    g->frame_smem = S >= 4;
    if (!g->frame_smem) S = 32;
    g->S = S;
    g->logS = 0;
    while ((1 << g->logS) < S) ++g->logS;
    g->ntiles = (nsamples + S - 1) / S;
    g->smem = fixed + (g->frame_smem ? frame_per_sample * S : 0);
    int per_sm = (int)(budget / (g->smem + KTAB * sizeof(int2) + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
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
        k<<<g.grid, THREADS, g.smem, stream>>>(p, args, constg, pl->gacc_elems ? partials : nullptr, frame_g,
                                               nsamples, g.S, g.logS, g.ntiles, ls, ic);
    } else {
        auto k = tnq_body_kernel<T, false>;
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
        if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute");
        k<<<g.grid, THREADS, g.smem, stream>>>(p, args, constg, pl->gacc_elems ? partials : nullptr, frame_g,
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
