// tnq_ladder_core.cuh -- warp-level sweep of a TWO-LAYER merged MPS network ("ladder":
// QCTN.merge(mps_n, mps_n), tneq_qc/core/qctn.py:1296-1506; BASELINE cfg3), real float32, edge
// rank K in {2, 3}.  Shared by the CUDA kernels (tnq_ladder.cu) and by the lane-by-lane CPU
// emulation that tests/ build to check this exact code without a GPU (tests/emu/ladder_emu.cpp).
//
// Network (wires = qubits, time flows left to right):  A_0(0,1) A_1(1,2) ... A_{n-2}  then
// X_0(0,1) X_1(1,2) ... X_{n-2};  every core is G[in_q][in_q+1][out_q][out_q+1].
// The greedy sweep of the reference (tneq_qc/contractor/greedy_strategy.py:461-598) contracts, at
// qubit q, {A_q, X_q, Mx_q, their right-hand copies, state q+1} into the environment; the einsum
// strings are  first   "cdef,eghi,c,ahj,klmn,mojp,k,d,l->agnpfio"
//              middle  "cdef,ghij,aik,lmno,pqkr,aelpcgn,d,m->ahorfjq"
//              last    "cdef,gfhi,ahj,klmn,onjp,aekocgm,d,l->api"   and   "acd,adc->a".
// With Bs_q[c][e][f] = sum_d A_q[c][d][e][f] s_{q+1}[d] and the environment stored as
// E[c][l][n][p][e][g] (c/l: A-output on the L/R copy, e/n: pending X-input, g/p: X-output):
//   A:  T2[f][o][g][p]       = sum_{c,e,l,n} Bs[c][e][f] Bs[l][n][o] E[c][l][n][p][e][g]
//   B:  U[i][p][q][r]        = sum_k M[i][k] X[p][q][k][r]
//   C:  E'[f][o][q][r][h][j] = sum_{g,i,p} T2[f][o][g][p] U[i][p][q][r] X[g][h][i][j]
// (11907 multiply-adds per qubit and sample at K = 3; the K^8 term is C.)
//
// Mapping: one warp owns SPW = 32 / K^2 samples; each sample has K^2 "items" = lanes.  In A an
// item is a (p,g) pair and holds its K^4 slice of E in registers; in C an item is an (f,o) pair
// and holds V[g][i][q][r] = T2 o U in registers; all operands that are shared by the lanes of a
// sample are read from the warp's private shared-memory buffers, core tensors as warp-wide
// broadcasts.  A warp never synchronises with another warp.
//
// Training keeps, per warp, the environments E_q and the T2_q of the current group of samples in
// global memory (written by the forward sweep, read back by the reverse sweep of the same warp
// right afterwards), so the working set per sample in shared memory stays ~7 KB.
// Gradients: every lane accumulates its contributions in registers, the lanes of a warp are
// combined through shared memory in a fixed order, and each warp adds into its own global slice;
// a finalize kernel sums the slices in warp order (deterministic).
#pragma once

#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define TNQ_HOSTDEV __host__ __device__
#define TNQ_HD __device__ __forceinline__
#define TNQ_UNROLL _Pragma("unroll")
#define TNQ_NOUNROLL _Pragma("unroll 1")
#define TNQ_UNROLL3 _Pragma("unroll 3")
#else
#define TNQ_HOSTDEV
#define TNQ_HD inline
#define TNQ_UNROLL
#define TNQ_NOUNROLL
#define TNQ_UNROLL3
#endif

#include "tnq_f2.cuh"

namespace tnq_ladder {

constexpr int MAXQ = 64;

template <int K>
struct Dims {
    static constexpr int K2 = K * K, K3 = K2 * K, K4 = K2 * K2;
    static constexpr int IPS = K2;              // items (lanes) per sample
    static constexpr int SPW = 32 / IPS;        // samples per warp
    static constexpr int LANES = SPW * IPS;     // active lanes of a warp
    // pitch of one (f,o) block of the environment and sample stride, in floats: chosen so that
    // the strided gathers of phase A and the block-wise stores of phase C are bank-conflict free
    // (K = 3) or at most 2-way (K = 2)
    static constexpr int PITCH = K == 3 ? 85 : 16;
    static constexpr int SST = K == 3 ? 765 : 67;
    // U is kept in one of two layouts: [i][qr][p padded to UP] (reverse sweep) or
    // [i][p][qr padded to UQ] (forward sweep, where (q,r) is the packed-FMA vector index)
    static constexpr int UP = 4;
    static constexpr int UQ = (K2 + 3) / 4 * 4;
    static constexpr int USZ = K3 * UP > K2 * UQ ? K3 * UP : K2 * UQ;
    static constexpr int XTP = (K2 + 3) / 4 * 4;  // padded (g,i) vector of the transposed core
    static constexpr int BSP = 4;               // padded innermost extent of Bs
    static constexpr int LP = LANES | 1;        // row pitch of the lane-reduction scratch
    // per-step block of the constant pool: X natural | X transposed [h][j][(g,i)] | Bs [c][e][f] |
    // Bs transposed [f][(c,e)]
    // (each part starts on a 16-byte boundary)
    static constexpr int BTP = (K2 + 3) / 4 * 4;  // padded (l,n) vector of the transposed Bs
    static constexpr int OFF_XN = 0, OFF_XT = (K4 + 3) / 4 * 4, OFF_BS = OFF_XT + K2 * XTP;
    static constexpr int OFF_BT = (OFF_BS + K2 * BSP + 3) / 4 * 4;     // BsT[o][(l,n)]
    static constexpr int CSTEP = (OFF_BT + K * BTP + 3) / 4 * 4;
    // per-warp shared-memory buffers (floats, every offset a multiple of 4)
    static constexpr int E_SZ = (SPW * SST + 3) / 4 * 4;
    static constexpr int U_SZ = SPW * USZ;
    static constexpr int T2_SZ = (SPW * K4 + 3) / 4 * 4;
    static constexpr int M_SZ = (SPW * K2 + 3) / 4 * 4;
    static constexpr int V_SZ = 64;
    static constexpr int WARP_FWD = E_SZ + U_SZ + T2_SZ + M_SZ + V_SZ;
    static constexpr int WARP_TRAIN = WARP_FWD + E_SZ + T2_SZ;
    static_assert(SPW * SST >= K4 * LP, "environment buffer doubles as the lane-reduction scratch");
    static_assert(USZ >= K4, "U buffer doubles as the X-gradient side buffer");
    // per-warp gradient slice: [q][K4] dX_q | [q][K3] dBs_q (q >= 1) | [K2] dAs0
    // (padded to a multiple of 4 floats so that everything behind it stays 16-byte aligned)
    TNQ_HOSTDEV static constexpr int ckpt_floats(int n) { return (n - 2) * (E_SZ + T2_SZ); }
    TNQ_HOSTDEV static constexpr int grad_floats(int n) { return ((n - 1) * K4 + (n - 1) * K3 + K2 + 3) / 4 * 4; }
};

struct Args {
    const float* coreA[MAXQ];   // layer 1, A_q on wires (q, q+1), q < n-1
    const float* coreX[MAXQ];   // layer 2
    const float* state[MAXQ];   // n circuit states [K]
    const float* mx[MAXQ];      // sample b of qubit q at mx[q] + b * mx_stride[q], [K][K] row major
    long long mx_stride[MAXQ];
    float* gradA[MAXQ];
    float* gradX[MAXQ];
    int n;
};

template <int N>
struct Vec {
    static constexpr int NP = N / 2;
    F2 p[NP > 0 ? NP : 1];
    float s;                    // element N-1 when N is odd
    // i must be a compile-time constant after unrolling
    TNQ_HD float get(int i) const { return i < 2 * NP ? ((i & 1) ? p[i >> 1].hi : p[i >> 1].lo) : s; }
    TNQ_HD void set(int i, float v) {
        if (i < 2 * NP) {
            if (i & 1) p[i >> 1].hi = v; else p[i >> 1].lo = v;
        } else {
            s = v;
        }
    }
    TNQ_HD void zero() {
        TNQ_UNROLL
        for (int i = 0; i < NP; ++i) p[i] = F2{0.f, 0.f};
        s = 0.f;
    }
};
// acc += s * x
template <int N>
TNQ_HD void axpy(Vec<N>& acc, float s, const Vec<N>& x) {
    TNQ_UNROLL
    for (int i = 0; i < Vec<N>::NP; ++i) acc.p[i] = fma2(x.p[i], s, acc.p[i]);
    if (N & 1) acc.s = fmaf(x.s, s, acc.s);
}
// init + sum_i a[i] * b[i]
template <int N>
TNQ_HD float dot(const Vec<N>& a, const Vec<N>& b, float init) {
    F2 t{init, 0.f};
    TNQ_UNROLL
    for (int i = 0; i < Vec<N>::NP; ++i) t = fma2(a.p[i], b.p[i], t);
    if (N & 1) t.lo = fmaf(a.s, b.s, t.lo);
    return t.lo + t.hi;
}
// N consecutive floats (padded to a multiple of 4, 16-byte aligned) from shared memory
template <int N>
TNQ_HD void vload(Vec<N>& v, const float* p) {
    float t[(N + 3) / 4 * 4];
#ifdef __CUDA_ARCH__
    TNQ_UNROLL
    for (int i = 0; i < (N + 3) / 4; ++i) {
        const float4 q = *reinterpret_cast<const float4*>(p + 4 * i);
        t[4 * i] = q.x, t[4 * i + 1] = q.y, t[4 * i + 2] = q.z, t[4 * i + 3] = q.w;
    }
#else
    for (int i = 0; i < (N + 3) / 4 * 4; ++i) t[i] = p[i];
#endif
    TNQ_UNROLL
    for (int i = 0; i < N; ++i) v.set(i, t[i]);
}

template <int K>
struct LaneState {
    Vec<Dims<K>::K2> accX[Dims<K>::K2];  // this lane's partial d loss / d X_q: [g*K+i] over (h,j)
    Vec<K> accB[Dims<K>::K2];   // this lane's partial d loss / d Bs_q: [c*K+e] over f (or dAs0: [e] over f)
    float pf[Dims<K>::K2];      // last step: P_fo[i][p], later d P[i][p]
    float mnext;                // prefetched measurement-matrix element of the next step
    float loss;                 // running loss contribution (lanes with item 0)
};

template <int K>
struct WarpCtx {
    const float* cst;           // constant pool (shared memory)
    float *E, *U, *T2, *M, *V;  // per-warp buffers (shared memory)
    float *D, *dT2;             // training only
    float* ckE;                 // per-warp checkpoints (global): [q-1][E_SZ]
    float* ckT2;                //                                [q-1][T2_SZ]
    float* gpart;               // per-warp gradient slice (global)
    const Args* args;
    long long B;
    const float* seed;          // MODE 2
    float* values;              // MODE 0 / 1 (may be null in MODE 1)
    float log_scale, inv_count;
};

// N consecutive floats; 16-byte vector loads on the device when N is a multiple of 4
template <int N>
TNQ_HD void ldv(const float* p, float (&o)[N]) {
#ifdef __CUDA_ARCH__
    if constexpr (N % 4 == 0) {
        TNQ_UNROLL
        for (int i = 0; i < N; i += 4) {
            const float4 v = *reinterpret_cast<const float4*>(p + i);
            o[i] = v.x, o[i + 1] = v.y, o[i + 2] = v.z, o[i + 3] = v.w;
        }
    } else if constexpr (N % 2 == 0) {
        TNQ_UNROLL
        for (int i = 0; i < N; i += 2) {
            const float2 v = *reinterpret_cast<const float2*>(p + i);
            o[i] = v.x, o[i + 1] = v.y;
        }
    } else {
        TNQ_UNROLL
        for (int i = 0; i < N; ++i) o[i] = p[i];
    }
#else
    for (int i = 0; i < N; ++i) o[i] = p[i];
#endif
}
template <int N>
TNQ_HD void stv(float* p, const float (&o)[N]) {
#ifdef __CUDA_ARCH__
    if constexpr (N % 4 == 0) {
        TNQ_UNROLL
        for (int i = 0; i < N; i += 4) *reinterpret_cast<float4*>(p + i) = make_float4(o[i], o[i + 1], o[i + 2], o[i + 3]);
    } else if constexpr (N % 2 == 0) {
        TNQ_UNROLL
        for (int i = 0; i < N; i += 2) *reinterpret_cast<float2*>(p + i) = make_float2(o[i], o[i + 1]);
    } else {
        TNQ_UNROLL
        for (int i = 0; i < N; ++i) p[i] = o[i];
    }
#else
    for (int i = 0; i < N; ++i) p[i] = o[i];
#endif
}

TNQ_HD float ldg_f(const float* p) {
#ifdef __CUDA_ARCH__
    return __ldg(p);
#else
    return *p;
#endif
}

// element `it` of sample b's measurement matrix of qubit q (0 for samples past the batch)
template <int K>
TNQ_HD float load_m(const WarpCtx<K>& c, int q, long long b, int it) {
    return b < c.B ? ldg_f(c.args->mx[q] + b * c.args->mx_stride[q] + it) : 0.f;
}

// ---- phase A: E -> T2 (item = (p,g)) -------------------------------------------------------
// this item's K^4 slice of E as K^2 vectors over (l,n)
template <int K>
TNQ_HD void gather_epg(const float* e, Vec<Dims<K>::K2> (&Epg)[K][K]) {
    using D = Dims<K>;
    TNQ_UNROLL
    for (int cc = 0; cc < K; ++cc)
        TNQ_UNROLL
        for (int ee = 0; ee < K; ++ee)
            TNQ_UNROLL
            for (int l = 0; l < K; ++l)
                TNQ_UNROLL
                for (int n = 0; n < K; ++n) Epg[cc][ee].set(l * K + n, e[(cc * K + l) * D::PITCH + n * D::K3 + ee * K]);
}
// T1[f][(l,n)] = sum_{c,e} Bs[c][e][f] Epg[c][e][(l,n)]
template <int K>
TNQ_HD void contract_t1(const float* Bs, const Vec<Dims<K>::K2> (&Epg)[K][K], Vec<Dims<K>::K2> (&T1)[K]) {
    using D = Dims<K>;
    TNQ_UNROLL
    for (int f = 0; f < K; ++f) T1[f].zero();
    TNQ_UNROLL
    for (int cc = 0; cc < K; ++cc)
        TNQ_UNROLL
        for (int ee = 0; ee < K; ++ee) {
            float bs[D::BSP];
            ldv<D::BSP>(Bs + (cc * K + ee) * D::BSP, bs);
            TNQ_UNROLL
            for (int f = 0; f < K; ++f) axpy(T1[f], bs[f], Epg[cc][ee]);
        }
}

template <int K>
TNQ_HD void phase_a_fwd(const WarpCtx<K>& c, int lane, const float* Bs) {
    using D = Dims<K>;
    const int s = lane / D::IPS, it = lane % D::IPS, p = it / K, g = it % K;
    Vec<D::K2> Epg[K][K], T1[K];
    gather_epg<K>(c.E + s * D::SST + p * D::K2 + g, Epg);
    contract_t1<K>(Bs, Epg, T1);
    Vec<K> T2[K];            // [f] over o
    TNQ_UNROLL
    for (int f = 0; f < K; ++f) T2[f].zero();
    TNQ_UNROLL
    for (int ln = 0; ln < D::K2; ++ln) {
        Vec<K> bs;
        vload<K>(bs, Bs + ln * D::BSP);
        TNQ_UNROLL
        for (int f = 0; f < K; ++f) axpy(T2[f], T1[f].get(ln), bs);
    }
    float* t2 = c.T2 + s * D::K4 + g * K + p;
    TNQ_UNROLL
    for (int f = 0; f < K; ++f)
        TNQ_UNROLL
        for (int o = 0; o < K; ++o) t2[(f * K + o) * D::K2] = T2[f].get(o);
}

// ---- phase B: U[i][p][q][r] = sum_k M[i][k] X[p][q][k][r] (item = (q,r)) ---------------------
// LAYOUT 1: U[(i*K2 + qr)*UP + p] (reverse sweep); LAYOUT 2: U[(i*K + p)*UQ + qr] (forward sweep)
template <int K, int LAYOUT>
TNQ_HD void phase_b(const WarpCtx<K>& c, int lane, const float* Xn) {
    using D = Dims<K>;
    const int s = lane / D::IPS, qr = lane % D::IPS, q = qr / K, r = qr % K;
    float mm[K][K];
    TNQ_UNROLL
    for (int i = 0; i < K; ++i)
        TNQ_UNROLL
        for (int k = 0; k < K; ++k) mm[i][k] = c.M[s * D::K2 + i * K + k];
    float x[K][K];           // [p][k]
    TNQ_UNROLL
    for (int p = 0; p < K; ++p)
        TNQ_UNROLL
        for (int k = 0; k < K; ++k) x[p][k] = Xn[((p * K + q) * K + k) * K + r];
    TNQ_UNROLL
    for (int i = 0; i < K; ++i) {
        float u[D::UP];
        TNQ_UNROLL
        for (int p = 0; p < D::UP; ++p) {
            float a = 0.f;
            if (p < K) {
                TNQ_UNROLL
                for (int k = 0; k < K; ++k) a = fmaf(mm[i][k], x[p < K ? p : 0][k], a);
            }
            u[p] = a;
        }
        if (LAYOUT == 1) {
            stv<D::UP>(c.U + s * D::USZ + (i * D::K2 + qr) * D::UP, u);
        } else {
            TNQ_UNROLL
            for (int p = 0; p < K; ++p) c.U[s * D::USZ + (i * K + p) * D::UQ + qr] = u[p];
        }
    }
}

// T2 of the first step: the outer product As0[g][f] As0[p][o] (item = (f,o))
template <int K>
TNQ_HD void fill_t2_first(const WarpCtx<K>& c, int lane, const float* As0) {
    using D = Dims<K>;
    const int s = lane / D::IPS, it = lane % D::IPS, f = it / K, o = it % K;
    float* t2 = c.T2 + s * D::K4 + it * D::K2;
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) t2[g * K + p] = As0[g * K + f] * As0[p * K + o];
}

// ---- phase C: T2, U -> E' (item = (f,o)); vectors run over (q,r) ---------------------------------
template <int K>
TNQ_HD void phase_c_fwd(const WarpCtx<K>& c, int lane, const float* Xt) {
    using D = Dims<K>;
    const int s = lane / D::IPS, fo = lane % D::IPS;
    float t2[K][K];          // [g][p]
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) t2[g][p] = c.T2[s * D::K4 + fo * D::K2 + g * K + p];
    Vec<D::K2> V[D::K2];     // [g*K+i] over qr
    TNQ_UNROLL
    for (int gi = 0; gi < D::K2; ++gi) V[gi].zero();
    const float* U = c.U + s * D::USZ;
    TNQ_UNROLL
    for (int i = 0; i < K; ++i)
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) {
            Vec<D::K2> u;
            vload<D::K2>(u, U + (i * K + p) * D::UQ);
            TNQ_UNROLL
            for (int g = 0; g < K; ++g) axpy(V[g * K + i], t2[g][p], u);
        }
    float* e = c.E + s * D::SST + fo * D::PITCH;
    TNQ_UNROLL
    for (int hj = 0; hj < D::K2; ++hj) {
        float xt[D::XTP];
        ldv<D::XTP>(Xt + hj * D::XTP, xt);
        Vec<D::K2> out;
        out.zero();
        TNQ_UNROLL
        for (int gi = 0; gi < D::K2; ++gi) axpy(out, xt[gi], V[gi]);
        TNQ_UNROLL
        for (int qr = 0; qr < D::K2; ++qr) e[qr * D::K2 + hj] = out.get(qr);
    }
}

// ---- last composite step (qubit n-2), forward: P_fo[i][p] (item = (f,o)) -----------------------
//   P1[g][j][p'] = sum_p0 T2fo[g][p0] X[p0][o][j][p'] ; P2[g][h][p'] = sum_j M[h][j] P1[g][j][p']
//   P_fo[i][p']  = sum_{g,h} X[g][f][h][i] P2[g][h][p']
template <int K>
TNQ_HD void last_p2(const WarpCtx<K>& c, int lane, const float* Xn, float (&t2)[K][K], float (&P2)[K][K][K]) {
    using D = Dims<K>;
    const int s = lane / D::IPS, fo = lane % D::IPS, o = fo % K;
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) t2[g][p] = c.T2[s * D::K4 + fo * D::K2 + g * K + p];
    float P1[K][K][K];       // [g][j][p']
    TNQ_UNROLL
    for (int j = 0; j < K; ++j)
        TNQ_UNROLL
        for (int pp = 0; pp < K; ++pp) {
            float x[K];
            TNQ_UNROLL
            for (int p0 = 0; p0 < K; ++p0) x[p0] = Xn[((p0 * K + o) * K + j) * K + pp];
            TNQ_UNROLL
            for (int g = 0; g < K; ++g) {
                float a = 0.f;
                TNQ_UNROLL
                for (int p0 = 0; p0 < K; ++p0) a = fmaf(t2[g][p0], x[p0], a);
                P1[g][j][pp] = a;
            }
        }
    TNQ_UNROLL
    for (int h = 0; h < K; ++h) {
        float m[K];
        TNQ_UNROLL
        for (int j = 0; j < K; ++j) m[j] = c.M[s * D::K2 + h * K + j];
        TNQ_UNROLL
        for (int g = 0; g < K; ++g)
            TNQ_UNROLL
            for (int pp = 0; pp < K; ++pp) {
                float a = 0.f;
                TNQ_UNROLL
                for (int j = 0; j < K; ++j) a = fmaf(m[j], P1[g][j][pp], a);
                P2[g][h][pp] = a;
            }
    }
}

template <int K>
TNQ_HD void last_fwd(const WarpCtx<K>& c, LaneState<K>& st, int lane, const float* Xn) {
    using D = Dims<K>;
    const int f = (lane % D::IPS) / K;
    float t2[K][K], P2[K][K][K];
    last_p2<K>(c, lane, Xn, t2, P2);
    TNQ_UNROLL
    for (int i = 0; i < K; ++i)
        TNQ_UNROLL
        for (int pp = 0; pp < K; ++pp) {
            float a = 0.f;
            TNQ_UNROLL
            for (int g = 0; g < K; ++g)
                TNQ_UNROLL
                for (int h = 0; h < K; ++h) a = fmaf(Xn[((g * K + f) * K + h) * K + i], P2[g][h][pp], a);
            st.pf[i * K + pp] = a;
        }
}

// reverse of last_fwd: st.pf holds dP[i][p'] on entry; fills accX and the dT2 buffer
template <int K>
TNQ_HD void last_bwd(const WarpCtx<K>& c, LaneState<K>& st, int lane, const float* Xn) {
    using D = Dims<K>;
    const int s = lane / D::IPS, fo = lane % D::IPS, f = fo / K, o = fo % K;
    float t2[K][K], P2[K][K][K];
    last_p2<K>(c, lane, Xn, t2, P2);
    TNQ_UNROLL
    for (int v = 0; v < D::K2; ++v) st.accX[v].zero();
    float dP2[K][K][K];      // [g][h][p']
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int h = 0; h < K; ++h) {
            float x[K];
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) x[i] = Xn[((g * K + f) * K + h) * K + i];
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) {
                float cx = 0.f;   // d X[g][f][h][i]
                TNQ_UNROLL
                for (int pp = 0; pp < K; ++pp) cx = fmaf(st.pf[i * K + pp], P2[g][h][pp], cx);
                TNQ_UNROLL           // X[g][F][h][i]: accX[g*K + h] component (F, i)
                for (int F = 0; F < K; ++F)
                    st.accX[g * K + h].set(F * K + i, st.accX[g * K + h].get(F * K + i) + ((F == f) ? cx : 0.f));
            }
            TNQ_UNROLL
            for (int pp = 0; pp < K; ++pp) {
                float a = 0.f;
                TNQ_UNROLL
                for (int i = 0; i < K; ++i) a = fmaf(x[i], st.pf[i * K + pp], a);
                dP2[g][h][pp] = a;
            }
        }
    float dP1[K][K][K];      // [g][j][p']
    TNQ_UNROLL
    for (int j = 0; j < K; ++j) {
        float m[K];
        TNQ_UNROLL
        for (int h = 0; h < K; ++h) m[h] = c.M[s * D::K2 + h * K + j];
        TNQ_UNROLL
        for (int g = 0; g < K; ++g)
            TNQ_UNROLL
            for (int pp = 0; pp < K; ++pp) {
                float a = 0.f;
                TNQ_UNROLL
                for (int h = 0; h < K; ++h) a = fmaf(m[h], dP2[g][h][pp], a);
                dP1[g][j][pp] = a;
            }
    }
    float dt2[K][K];
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p0 = 0; p0 < K; ++p0) dt2[g][p0] = 0.f;
    TNQ_UNROLL
    for (int p0 = 0; p0 < K; ++p0)
        TNQ_UNROLL
        for (int j = 0; j < K; ++j)
            TNQ_UNROLL
            for (int pp = 0; pp < K; ++pp) {
                const float x = Xn[((p0 * K + o) * K + j) * K + pp];
                float cx = 0.f;   // d X[p0][o][j][p']
                TNQ_UNROLL
                for (int g = 0; g < K; ++g) {
                    dt2[g][p0] = fmaf(dP1[g][j][pp], x, dt2[g][p0]);
                    cx = fmaf(t2[g][p0], dP1[g][j][pp], cx);
                }
                TNQ_UNROLL           // X[p0][O][j][p']: accX[p0*K + j] component (O, p')
                for (int O = 0; O < K; ++O)
                    st.accX[p0 * K + j].set(O * K + pp, st.accX[p0 * K + j].get(O * K + pp) + ((O == o) ? cx : 0.f));
            }
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p0 = 0; p0 < K; ++p0) c.dT2[s * D::K4 + fo * D::K2 + g * K + p0] = dt2[g][p0];
}

// ---- reverse of phase C (item = (f,o)): D holds dE' on entry and Z = dV on exit ---------------
// per (q,r) slice:  v[g][i] = sum_p T2[g][p] U[i][p]      accX[g][i][(h,j)] += v[g][i] dE'[(h,j)]
//                   z[(g,i)] = sum_(h,j) dE'[(h,j)] Xt[(h,j)][(g,i)]     dT2[g][p] += z[g][i] U[i][p]
template <int K>
TNQ_HD void phase_c_bwd(const WarpCtx<K>& c, LaneState<K>& st, int lane, const float* Xt) {
    using D = Dims<K>;
    const int s = lane / D::IPS, fo = lane % D::IPS;
    Vec<K> t2T[K], dt2[K];   // t2T[p] over g ; dt2[g] over p
    TNQ_UNROLL
    for (int g = 0; g < K; ++g) {
        dt2[g].zero();
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) t2T[p].set(g, c.T2[s * D::K4 + fo * D::K2 + g * K + p]);
    }
    TNQ_UNROLL
    for (int v = 0; v < D::K2; ++v) st.accX[v].zero();
    const float* U = c.U + s * D::USZ;
    float* d = c.D + s * D::SST + fo * D::PITCH;
    TNQ_NOUNROLL             // (unrolling by 3 was measured: 3 % slower)
    for (int qr = 0; qr < D::K2; ++qr) {
        Vec<K> u[K];         // [i] over p
        TNQ_UNROLL
        for (int i = 0; i < K; ++i) vload<K>(u[i], U + (i * D::K2 + qr) * D::UP);
        Vec<D::K2> de;       // over (h,j)
        TNQ_UNROLL
        for (int hj = 0; hj < D::K2; ++hj) de.set(hj, d[qr * D::K2 + hj]);
        TNQ_UNROLL
        for (int i = 0; i < K; ++i) {
            Vec<K> v;        // over g
            v.zero();
            TNQ_UNROLL
            for (int p = 0; p < K; ++p) axpy(v, u[i].get(p), t2T[p]);
            TNQ_UNROLL
            for (int g = 0; g < K; ++g) axpy(st.accX[g * K + i], v.get(g), de);
        }
        Vec<D::K2> z;        // over (g,i)
        z.zero();
        TNQ_UNROLL
        for (int hj = 0; hj < D::K2; ++hj) {
            Vec<D::K2> xt;
            vload<D::K2>(xt, Xt + hj * D::XTP);
            axpy(z, de.get(hj), xt);
        }
        TNQ_UNROLL
        for (int g = 0; g < K; ++g)
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) {
                axpy(dt2[g], z.get(g * K + i), u[i]);
                d[qr * D::K2 + g * K + i] = z.get(g * K + i);
            }
    }
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) c.dT2[s * D::K4 + fo * D::K2 + g * K + p] = dt2[g].get(p);
}

// dU[i][p] of this item's (q,r) from Z and T2, then the right-hand-copy contribution to dX:
//   XR[p][q][k][r] = sum_i M[i][k] dU[i][p]     (item = (q,r); XR aliases the U buffer)
template <int K>
TNQ_HD void phase_du(const WarpCtx<K>& c, int lane) {
    using D = Dims<K>;
    const int s = lane / D::IPS, qr = lane % D::IPS, q = qr / K, r = qr % K;
    Vec<K> du[K];            // [i] over p
    TNQ_UNROLL
    for (int i = 0; i < K; ++i) du[i].zero();
    const float* z = c.D + s * D::SST + qr * D::K2;
    const float* t2 = c.T2 + s * D::K4;
    TNQ_UNROLL
    for (int fo = 0; fo < D::K2; ++fo)
        TNQ_UNROLL
        for (int g = 0; g < K; ++g) {
            Vec<K> tt;
            TNQ_UNROLL
            for (int p = 0; p < K; ++p) tt.set(p, t2[fo * D::K2 + g * K + p]);
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) axpy(du[i], z[fo * D::PITCH + g * K + i], tt);
        }
    float* xr = c.U + s * D::USZ;
    TNQ_UNROLL
    for (int k = 0; k < K; ++k) {
        Vec<K> w;            // over p
        w.zero();
        TNQ_UNROLL
        for (int i = 0; i < K; ++i) axpy(w, c.M[s * D::K2 + i * K + k], du[i]);
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) xr[((p * K + q) * K + k) * K + r] = w.get(p);
    }
}

// ---- reverse of phase A (item = (p,g)): E, dT2 -> dE (into D) and accB --------------------------
//   X2[f][(l,n)] = dT1 = sum_o Bs[l][n][o] dT2[f][o]
//   dBs[l][n][o] += T1[f][(l,n)] dT2[f][o]            dBs[c][e][f] += sum_(l,n) Epg[c][e][(l,n)] X2[f][(l,n)]
//   dE[c][e][(l,n)] = sum_f Bs[c][e][f] X2[f][(l,n)]
template <int K>
TNQ_HD void phase_a_bwd(const WarpCtx<K>& c, LaneState<K>& st, int lane, const float* Bs, const float* BsT) {
    using D = Dims<K>;
    const int s = lane / D::IPS, it = lane % D::IPS, p = it / K, g = it % K;
    Vec<D::K2> Epg[K][K], T1[K], X2[K];
    gather_epg<K>(c.E + s * D::SST + p * D::K2 + g, Epg);
    Vec<K> dt[K];            // [f] over o
    TNQ_UNROLL
    for (int f = 0; f < K; ++f)
        TNQ_UNROLL
        for (int o = 0; o < K; ++o) dt[f].set(o, c.dT2[s * D::K4 + (f * K + o) * D::K2 + g * K + p]);
    contract_t1<K>(Bs, Epg, T1);
    TNQ_UNROLL
    for (int ln = 0; ln < D::K2; ++ln) {
        st.accB[ln].zero();
        TNQ_UNROLL
        for (int f = 0; f < K; ++f) axpy(st.accB[ln], T1[f].get(ln), dt[f]);
    }
    TNQ_UNROLL
    for (int f = 0; f < K; ++f) X2[f].zero();
    TNQ_UNROLL
    for (int o = 0; o < K; ++o) {
        Vec<D::K2> bt;
        vload<D::K2>(bt, BsT + o * D::BTP);
        TNQ_UNROLL
        for (int f = 0; f < K; ++f) axpy(X2[f], dt[f].get(o), bt);
    }
    float* d = c.D + s * D::SST + p * D::K2 + g;
    TNQ_UNROLL
    for (int cc = 0; cc < K; ++cc)
        TNQ_UNROLL
        for (int ee = 0; ee < K; ++ee) {
            float bs[D::BSP];
            ldv<D::BSP>(Bs + (cc * K + ee) * D::BSP, bs);
            Vec<D::K2> de;
            de.zero();
            TNQ_UNROLL
            for (int f = 0; f < K; ++f) {
                axpy(de, bs[f], X2[f]);
                st.accB[cc * K + ee].set(f, dot(Epg[cc][ee], X2[f], st.accB[cc * K + ee].get(f)));
            }
            TNQ_UNROLL
            for (int l = 0; l < K; ++l)
                TNQ_UNROLL
                for (int n = 0; n < K; ++n) d[(cc * K + l) * D::PITCH + n * D::K3 + ee * K] = de.get(l * K + n);
        }
}

// first step, reverse of T2 = As0 (x) As0 (item = (f,o)): accB[e] over f = this lane's d As0[e][f]
template <int K>
TNQ_HD void first_a_bwd(const WarpCtx<K>& c, LaneState<K>& st, int lane, const float* As0) {
    using D = Dims<K>;
    const int s = lane / D::IPS, fo = lane % D::IPS, f = fo / K, o = fo % K;
    TNQ_UNROLL
    for (int v = 0; v < K; ++v) st.accB[v].zero();
    float dt[K][K];
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) dt[g][p] = c.dT2[s * D::K4 + fo * D::K2 + g * K + p];
    TNQ_UNROLL
    for (int g = 0; g < K; ++g) {
        float cl = 0.f;      // d As0[g][f]
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) cl = fmaf(dt[g][p], As0[p * K + o], cl);
        TNQ_UNROLL
        for (int F = 0; F < K; ++F) st.accB[g].set(F, st.accB[g].get(F) + ((F == f) ? cl : 0.f));
    }
    TNQ_UNROLL
    for (int p = 0; p < K; ++p) {
        float cr = 0.f;      // d As0[p][o]
        TNQ_UNROLL
        for (int g = 0; g < K; ++g) cr = fmaf(dt[g][p], As0[g * K + f], cr);
        TNQ_UNROLL
        for (int O = 0; O < K; ++O) st.accB[p].set(O, st.accB[p].get(O) + ((O == o) ? cr : 0.f));
    }
}

// += into the warp's own gradient slice: a reduction without return value on the device (no load
// latency on the critical path; the address has a single writer, so the order of adds is fixed)
TNQ_HD void gadd(float* p, float v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

// ---- lane reduction: scratch[v][lane] then row sums in lane order -> += gp[v] ---------------------
template <int K>
TNQ_HD void flush_put_x(float* scr, int lane, const Vec<Dims<K>::K2> (&acc)[Dims<K>::K2]) {
    using D = Dims<K>;
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int h = 0; h < K; ++h)
            TNQ_UNROLL
            for (int i = 0; i < K; ++i)
                TNQ_UNROLL
                for (int j = 0; j < K; ++j) scr[(((g * K + h) * K + i) * K + j) * D::LP + lane] = acc[g * K + i].get(h * K + j);
}
template <int K, int NROWS>
TNQ_HD void flush_put_b(float* scr, int lane, const Vec<K> (&acc)[Dims<K>::K2]) {
    using D = Dims<K>;
    TNQ_UNROLL
    for (int ab = 0; ab < NROWS; ++ab)
        TNQ_UNROLL
        for (int cc = 0; cc < K; ++cc) scr[(ab * K + cc) * D::LP + lane] = acc[ab].get(cc);
}
template <int K, int NV, bool WITH_XR>
TNQ_HD void flush_sum(const float* scr, const float* xr, int lane, float* gp) {
    using D = Dims<K>;
    TNQ_NOUNROLL
    for (int v = lane; v < NV; v += 32) {
        float t = 0.f;
        TNQ_UNROLL
        for (int l = 0; l < D::LANES; ++l) t += scr[v * D::LP + l];
        if (WITH_XR) {
            TNQ_UNROLL
            for (int s = 0; s < D::SPW; ++s) t += xr[s * D::USZ + v];
        }
        gadd(gp + v, t);
    }
}

// ---- checkpoints: cooperative 16-byte copies between a warp's buffers and global memory -----------
// store: shared -> global (fire and forget).  load: global -> shared with cp.async on the device,
// completed by ckpt_wait<PENDING>() (all but the PENDING most recently committed groups) followed
// by the phase's __syncwarp(); the CPU emulation copies at issue time.
template <int NFLOATS>
TNQ_HD void ckpt_store(float* dst, const float* src, int lane) {
    static_assert(NFLOATS % 4 == 0, "16-byte units");
#ifdef __CUDA_ARCH__
    TNQ_UNROLL
    for (int i = 0; i < (NFLOATS / 4 + 31) / 32; ++i) {
        const int j = lane + 32 * i;
        if (j < NFLOATS / 4) reinterpret_cast<float4*>(dst)[j] = reinterpret_cast<const float4*>(src)[j];
    }
#else
    for (int i = lane; i < NFLOATS; i += 32) dst[i] = src[i];
#endif
}
template <int NFLOATS>
TNQ_HD void ckpt_load_async(float* dst, const float* src, int lane) {
    static_assert(NFLOATS % 4 == 0, "16-byte units");
#ifdef __CUDA_ARCH__
    const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst);
    TNQ_UNROLL
    for (int i = 0; i < (NFLOATS / 4 + 31) / 32; ++i) {
        const int j = lane + 32 * i;
        if (j < NFLOATS / 4)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * j), "l"(src + 4 * j) : "memory");
    }
#else
    for (int i = lane; i < NFLOATS; i += 32) dst[i] = src[i];
#endif
}
TNQ_HD void ckpt_commit() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int PENDING>
TNQ_HD void ckpt_wait() {
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory");
#endif
}

}  // namespace tnq_ladder

// -------------------------------------------------------------------------------------------------
// The sweep of one group of SPW samples.  TNQ_PHASE(body) runs `body` for every lane and then
// makes the lanes' shared-memory writes visible to each other: on the device `lane` and `st` are
// the calling thread's, followed by __syncwarp(); the CPU emulation loops over 32 lanes.
// -------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define TNQ_PHASE(...) \
    { __VA_ARGS__ }    \
    __syncwarp();
#define TNQ_LANES_PARAM LaneState<K>&st, const int lane
#else
#define TNQ_PHASE(...)                         \
    for (int lane = 0; lane < 32; ++lane) {    \
        LaneState<K>& st = lanes[lane];        \
        (void)st;                              \
        __VA_ARGS__                            \
    }
#define TNQ_LANES_PARAM LaneState<K>* lanes
#endif

namespace tnq_ladder {

// MODE 0: values.  MODE 1: values + fused loss + gradients.  MODE 2: gradients seeded by c.seed.
template <int K, int MODE>
TNQ_HD void ladder_group(const WarpCtx<K>& c, TNQ_LANES_PARAM, long long b0) {
    using D = Dims<K>;
    const int n = c.args->n;
    const float* As0 = c.cst + (n - 1) * D::CSTEP;
    constexpr bool CK = MODE != 0;
#define TNQ_ACTIVE (lane < D::LANES)
#define TNQ_S (lane / D::IPS)
#define TNQ_IT (lane % D::IPS)
#define TNQ_XN(q) (c.cst + (q) * D::CSTEP + D::OFF_XN)
#define TNQ_XT(q) (c.cst + (q) * D::CSTEP + D::OFF_XT)
#define TNQ_BS(q) (c.cst + (q) * D::CSTEP + D::OFF_BS)
#define TNQ_BT(q) (c.cst + (q) * D::CSTEP + D::OFF_BT)
#define TNQ_GX(q) (c.gpart + (q) * D::K4)
#define TNQ_GB(q) (c.gpart + (n - 1) * D::K4 + (q) * D::K3)
#define TNQ_GA0 (c.gpart + (n - 1) * D::K4 + (n - 1) * D::K3)

    // ------------------------------- forward sweep -------------------------------
    TNQ_PHASE(if (TNQ_ACTIVE) {
        c.M[lane] = load_m<K>(c, 0, b0 + TNQ_S, TNQ_IT);
        st.mnext = load_m<K>(c, 1, b0 + TNQ_S, TNQ_IT);
        fill_t2_first<K>(c, lane, As0);
    })
    TNQ_PHASE(if (TNQ_ACTIVE) phase_b<K, 2>(c, lane, TNQ_XN(0));)
    TNQ_PHASE(if (TNQ_ACTIVE) phase_c_fwd<K>(c, lane, TNQ_XT(0));)
    for (int q = 1; q <= n - 2; ++q) {
        TNQ_PHASE(
            if (CK && q < n - 2) ckpt_store<D::E_SZ>(c.ckE + (size_t)(q - 1) * D::E_SZ, c.E, lane);
            if (TNQ_ACTIVE) {
                c.M[lane] = st.mnext;
                st.mnext = load_m<K>(c, q + 1, b0 + TNQ_S, TNQ_IT);
                phase_a_fwd<K>(c, lane, TNQ_BS(q));
            })
        if (q < n - 2) {
            TNQ_PHASE(
                if (CK) ckpt_store<D::T2_SZ>(c.ckT2 + (size_t)(q - 1) * D::T2_SZ, c.T2, lane);
                if (TNQ_ACTIVE) phase_b<K, 2>(c, lane, TNQ_XN(q));)
            TNQ_PHASE(if (TNQ_ACTIVE) phase_c_fwd<K>(c, lane, TNQ_XT(q));)
        }
    }
    TNQ_PHASE(if (TNQ_ACTIVE) last_fwd<K>(c, st, lane, TNQ_XN(n - 2));)
    TNQ_PHASE(if (TNQ_ACTIVE) c.M[lane] = st.mnext;)       // M_{n-1}
    TNQ_PHASE(if (TNQ_ACTIVE) {
        float v = 0.f;
        const float* m = c.M + TNQ_S * D::K2;
        TNQ_UNROLL
        for (int i = 0; i < D::K2; ++i) v = fmaf(st.pf[i], m[i], v);
        c.V[lane] = v;
    })
    TNQ_PHASE(if (TNQ_ACTIVE && TNQ_IT == 0) {
        float val = 0.f;
        TNQ_UNROLL
        for (int i = 0; i < D::IPS; ++i) val += c.V[lane + i];
        const long long b = b0 + TNQ_S;
        const bool valid = b < c.B;
        if (MODE != 2 && valid && c.values != nullptr) c.values[b] = val;
        float dv = 0.f;
        if (MODE == 1) {
            const float cl = val > 1e-10f ? val : 1e-10f;
            if (valid) {
                st.loss -= (logf(cl) + c.log_scale) * c.inv_count;
                dv = val >= 1e-10f ? -c.inv_count / cl : 0.f;
            }
        } else if (MODE == 2) {
            dv = valid ? ldg_f(c.seed + b) : 0.f;
        }
        c.V[32 + TNQ_S] = dv;
    })
    if (MODE == 0) return;

    // ------------------------------- reverse sweep -------------------------------
    // final trace and the last composite step (qubit n-2): T2 and E of that step are still in place
    TNQ_PHASE(if (TNQ_ACTIVE) {
        const float dv = c.V[32 + TNQ_S];
        const float* m = c.M + TNQ_S * D::K2;
        TNQ_UNROLL
        for (int i = 0; i < D::K2; ++i) st.pf[i] = dv * m[i];
        st.mnext = load_m<K>(c, n - 2, b0 + TNQ_S, TNQ_IT);
    })
    TNQ_PHASE(if (TNQ_ACTIVE) {
        c.M[lane] = st.mnext;
        if (n >= 3) st.mnext = load_m<K>(c, n - 3, b0 + TNQ_S, TNQ_IT);
    })
    TNQ_PHASE(if (TNQ_ACTIVE) {
        last_bwd<K>(c, st, lane, TNQ_XN(n - 2));
        flush_put_x<K>(c.D, lane, st.accX);
    })
    // cp.async groups are committed alternately: T2 of the next step (after this step's last use of
    // the T2 buffer), then E of that step (at its start); every wait leaves the newest group pending
    TNQ_PHASE(
        flush_sum<K, D::K4, false>(c.D, nullptr, lane, TNQ_GX(n - 2));
        if (n - 3 >= 1) ckpt_load_async<D::T2_SZ>(c.T2, c.ckT2 + (size_t)(n - 4) * D::T2_SZ, lane);
        ckpt_commit();)
    for (int q = n - 2; q >= 1; --q) {
        if (q < n - 2) {
            TNQ_PHASE(
                ckpt_load_async<D::E_SZ>(c.E, c.ckE + (size_t)(q - 1) * D::E_SZ, lane);
                ckpt_commit();
                if (TNQ_ACTIVE) {
                    c.M[lane] = st.mnext;
                    st.mnext = load_m<K>(c, q - 1, b0 + TNQ_S, TNQ_IT);
                })
            TNQ_PHASE(
                if (TNQ_ACTIVE) phase_b<K, 1>(c, lane, TNQ_XN(q));
                ckpt_wait<1>();)                                  // T2_q has landed
            TNQ_PHASE(if (TNQ_ACTIVE) phase_c_bwd<K>(c, st, lane, TNQ_XT(q));)
            TNQ_PHASE(if (TNQ_ACTIVE) phase_du<K>(c, lane);)
            TNQ_PHASE(
                if (TNQ_ACTIVE) flush_put_x<K>(c.D, lane, st.accX);
                if (q - 1 >= 1) ckpt_load_async<D::T2_SZ>(c.T2, c.ckT2 + (size_t)(q - 2) * D::T2_SZ, lane);
                ckpt_commit();)
            TNQ_PHASE(
                flush_sum<K, D::K4, true>(c.D, c.U, lane, TNQ_GX(q));
                ckpt_wait<1>();)                                  // E_q has landed
        }
        TNQ_PHASE(if (TNQ_ACTIVE) phase_a_bwd<K>(c, st, lane, TNQ_BS(q), TNQ_BT(q));)
        TNQ_PHASE(if (TNQ_ACTIVE) flush_put_b<K, D::K2>(c.E, lane, st.accB);)
        TNQ_PHASE(flush_sum<K, D::K3, false>(c.E, nullptr, lane, TNQ_GB(q));)
    }
    TNQ_PHASE(ckpt_wait<0>();)
    // first step
    TNQ_PHASE(if (TNQ_ACTIVE) {
        c.M[lane] = st.mnext;
        fill_t2_first<K>(c, lane, As0);
    })
    TNQ_PHASE(if (TNQ_ACTIVE) phase_b<K, 1>(c, lane, TNQ_XN(0));)
    TNQ_PHASE(if (TNQ_ACTIVE) phase_c_bwd<K>(c, st, lane, TNQ_XT(0));)
    TNQ_PHASE(if (TNQ_ACTIVE) phase_du<K>(c, lane);)
    TNQ_PHASE(if (TNQ_ACTIVE) flush_put_x<K>(c.D, lane, st.accX);)
    TNQ_PHASE(flush_sum<K, D::K4, true>(c.D, c.U, lane, TNQ_GX(0));)
    TNQ_PHASE(if (TNQ_ACTIVE) {
        first_a_bwd<K>(c, st, lane, As0);
        flush_put_b<K, K>(c.E, lane, st.accB);
    })
    TNQ_PHASE(flush_sum<K, D::K2, false>(c.E, nullptr, lane, TNQ_GA0);)
#undef TNQ_ACTIVE
#undef TNQ_S
#undef TNQ_IT
#undef TNQ_XN
#undef TNQ_XT
#undef TNQ_BS
#undef TNQ_BT
#undef TNQ_GX
#undef TNQ_GB
#undef TNQ_GA0
}

// Constant pool of one CTA, element `idx` of [0, (n-1)*CSTEP + K2): see Dims<K>::OFF_*.
template <int K>
TNQ_HD float const_pool_element(const Args& a, int idx) {
    using D = Dims<K>;
    const int n = a.n;
    if (idx >= (n - 1) * D::CSTEP) {           // As0[e][f] = sum_{c,d} A_0[c][d][e][f] s0[c] s1[d]
        const int ef = idx - (n - 1) * D::CSTEP;
        float v = 0.f;
        for (int cc = 0; cc < K; ++cc)
            for (int d = 0; d < K; ++d) v = fmaf(a.coreA[0][(cc * K + d) * D::K2 + ef], a.state[0][cc] * a.state[1][d], v);
        return v;
    }
    const int q = idx / D::CSTEP, r = idx % D::CSTEP;
    if (r < D::OFF_XT) return a.coreX[q][r];
    if (r < D::OFF_BS) {                        // Xt[h][j][g*K+i] = X[g][h][i][j]
        const int hj = (r - D::OFF_XT) / D::XTP, gi = (r - D::OFF_XT) % D::XTP;
        if (gi >= D::K2) return 0.f;
        const int h = hj / K, j = hj % K, g = gi / K, i = gi % K;
        return a.coreX[q][((g * K + h) * K + i) * K + j];
    }
    if (q == 0) return 0.f;
    int ce, f;                                  // Bs[c][e][f] = sum_d A_q[c][d][e][f] s_{q+1}[d]
    if (r < D::OFF_BT) {
        const int t = r - D::OFF_BS;
        if (t >= D::K2 * D::BSP) return 0.f;
        ce = t / D::BSP, f = t % D::BSP;
    } else {                                    // transposed copy BsT[f][(c,e)]
        const int t = r - D::OFF_BT;
        if (t >= K * D::BTP) return 0.f;
        f = t / D::BTP, ce = t % D::BTP;
    }
    if (f >= K || ce >= D::K2) return 0.f;
    const int cc = ce / K, e = ce % K;
    float v = 0.f;
    for (int d = 0; d < K; ++d) v = fmaf(a.coreA[q][((cc * K + d) * K + e) * K + f], a.state[q + 1][d], v);
    return v;
}

// Finalize: element v of the gradient of core A_q (layer == 0) or X_q (layer == 1), summed over the
// per-warp slices in warp order, with the circuit states folded back in for layer 1:
//   dA_q[c][d][e][f] = dBs_q[c][e][f] s_{q+1}[d]  (q >= 1),   dA_0[c][d][e][f] = dAs0[e][f] s_0[c] s_1[d]
template <int K>
TNQ_HD float grad_element(const Args& a, const float* gparts, int nwarps, int layer, int q, int v) {
    using D = Dims<K>;
    const int n = a.n, stride = D::grad_floats(n);
    int off;
    float w = 1.f;
    if (layer == 1) {
        off = q * D::K4 + v;
    } else {
        const int f = v % K, e = (v / K) % K, d = (v / D::K2) % K, cc = v / D::K3;
        if (q == 0) {
            off = (n - 1) * D::K4 + (n - 1) * D::K3 + e * K + f;
            w = a.state[0][cc] * a.state[1][d];
        } else {
            off = (n - 1) * D::K4 + q * D::K3 + (cc * K + e) * K + f;
            w = a.state[q + 1][d];
        }
    }
    float t = 0.f;
    for (int i = 0; i < nwarps; ++i) t += gparts[(size_t)i * stride + off];
    return t * w;
}

}  // namespace tnq_ladder
