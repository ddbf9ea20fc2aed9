// tnq_ladder2_core.cuh -- second-generation sweep of the TWO-LAYER merged MPS network (BASELINE cfg3,
// QCTN.merge(mps_n, mps_n), tneq_qc/core/qctn.py:1296-1506), real float32, edge rank K = 3.
// Same mathematics as tnq_ladder_core.cuh (the per-qubit group of the greedy sweep,
// tneq_qc/contractor/greedy_strategy.py:690-990: phases A, B, C below; the fused loss of
// tneq_qc/core/engine_siamese.py:490-530; its reverse sweep), different mapping:
//
//   * a LANE is a SAMPLE.  A CTA of 4 warps owns a tile of S = 32 / R samples; the 32 lanes of a warp
//     are R "slots" x S samples and the slots of a warp work on different ROW BLOCKS of the same step.
//     Everything that lives per sample is laid out [element][position][sample], so every shared-memory
//     access of a warp is R runs of S consecutive words: conflict free and broadcast free by
//     construction (the position maps below put the row blocks a warp works on at the same time into
//     different banks).
//   * everything that is shared by all samples (the core tensors, folded with the circuit states) sits in
//     CONSTANT memory: the operands arrive through the uniform datapath (LDCU -> uniform registers ->
//     FFMA2 with a uniform operand), never through the load/store unit.
//   * the environment E (K^6 floats per sample) never exists in memory: with row block (o,q',r)
//         V[f,g',i']  = sum_p' T2[f,o,g',p'] U[i',p',q',r]                    (phase C, first half)
//         E'[f,h,j]   = sum_{g',i'} V[f,g',i'] X[g',h,i',j]                   (phase C, second half)
//         T1'[f",j]   = sum_{f,h} Bs'[f,h,f"] E'[f,h,j]                       (phase A of the NEXT step)
//     are one pass over registers (27 row blocks of 405 multiply-adds); only T1' (K^5), T2 (K^4) and
//     U (K^4) go through shared memory.  The reverse sweep re-derives E' from the checkpointed T2 the
//     same way, so the only per-sample state that ever reaches HBM is T2 (81 floats per step).
//   * all hot loops are "vector += scalar * vector" on packed fp32 pairs (fma.rn.f32x2, tnq_f2.cuh).
//
// Per step q (0 <= q <= n-2), with Bs_q[c,e,f] = sum_d A_q[c,d,e,f] s_{q+1}[d]  (q >= 1),
// As0[e,f] = sum_{c,d} A_0[c,d,e,f] s_0[c] s_1[d], E_q[c,l,n,p,e,g] the environment entering step q:
//   A1: T1_q[f,l,n,p,g]      = sum_{c,e} Bs_q[c,e,f] E_q[c,l,n,p,e,g]
//   A2: T2_q[f,o,g,p]        = sum_{l,n} Bs_q[l,n,o] T1_q[f,l,n,p,g]          (T2_0 = As0[g,f] As0[p,o])
//   B : U_q[i,p,q',r]        = sum_k M_q[i,k] X_q[p,q',k,r]
//   C : E_{q+1}[f,o,q',r,h,j] = sum_{g,i,p} T2_q[f,o,g,p] U_q[i,p,q',r] X_q[g,h,i,j]
// and for the last step q = n-2 (einsum "cdef,gfhi,ahj,klmn,onjp,aekocgm,d,l->api" then "acd,adc->a"):
//   value = sum T2[f,o,g,p0] X[g,f,h,i] M_{n-2}[h,j] X[p0,o,j,p'] M_{n-1}[i,p'].
//
// Shared by the CUDA kernels (tnq_ladder2.cu) and by a thread-by-thread CPU emulation that the tests
// build to check this exact code without a GPU (tests/emu/ladder2_emu.cpp).
#pragma once

#include "tnq_ladder_core.cuh"   // Args, Vec<N>, axpy, dot, ldg_f, TNQ_* macros

namespace tnq_l2 {

using tnq_ladder::Args;
using tnq_ladder::Vec;
using tnq_ladder::axpy;
using tnq_ladder::dot;
using tnq_ladder::ldg_f;
using tnq_ladder::MAXQ;

constexpr int K = 3, K2 = 9, K3 = 27, K4 = 81;
// NW warps (4 or 8) per CTA: a template parameter next to R

// ---- constant pool (floats) ------------------------------------------------------------------------
// one block per step q, then a tail with the extra layouts of the last step and As0
constexpr int C_XA = 0;      // XA[(g,i)][(h,j)] = X_q[g,h,i,j], row pitch 10
constexpr int C_XB = 90;     // XB[(h,j)][(g,i)] = X_q[g,h,i,j], row pitch 10
constexpr int C_BA = 180;    // BA[(c,e)][f]     = Bs_q[c,e,f],  row pitch 4
constexpr int C_BB = 216;    // BB[o][(l,n)]     = Bs_q[l,n,o],  row pitch 10
constexpr int C_STEP = 248;
constexpr int T_XN = 0;      // XN[(a,b)][(c,d)] = X_{n-2}[a,b,c,d], row pitch 10
constexpr int T_XC = 90;     // XC[(c,d)][(a,b)] = X_{n-2}[a,b,c,d], row pitch 10
constexpr int T_AS0 = 180;   // As0[e][f]
constexpr int C_TAIL = 192;
constexpr int C_MAX = (MAXQ - 1) * C_STEP + C_TAIL;      // 15816 floats < 64 KB
TNQ_HOSTDEV constexpr int cst_floats(int n) { return (n - 1) * C_STEP + C_TAIL; }

// element idx of the constant pool, from the caller's cores and circuit states
TNQ_HD float cst_element(const Args& a, int idx) {
    const int n = a.n;
    if (idx >= (n - 1) * C_STEP) {
        const int t = idx - (n - 1) * C_STEP;
        const float* X = a.coreX[n - 2];
        if (t < T_XC) {
            const int r = t / 10, col = t % 10;
            return col < 9 ? X[r * 9 + col] : 0.f;
        }
        if (t < T_AS0) {
            const int r = (t - T_XC) / 10, col = (t - T_XC) % 10;
            return col < 9 ? X[col * 9 + r] : 0.f;
        }
        if (t < T_AS0 + K2) {                    // As0[e][f] = sum_{c,d} A_0[c,d,e,f] s0[c] s1[d]
            const int ef = t - T_AS0;
            float v = 0.f;
            for (int c = 0; c < K; ++c)
                for (int d = 0; d < K; ++d) v = fmaf(a.coreA[0][(c * K + d) * K2 + ef], a.state[0][c] * a.state[1][d], v);
            return v;
        }
        return 0.f;
    }
    const int q = idx / C_STEP, t = idx % C_STEP;
    const float* X = a.coreX[q];
    if (t < C_XB) {                              // XA[(g,i)][(h,j)]
        const int gi = t / 10, hj = t % 10;
        if (hj >= 9) return 0.f;
        return X[(((gi / 3) * 3 + hj / 3) * 3 + gi % 3) * 3 + hj % 3];
    }
    if (t < C_BA) {                              // XB[(h,j)][(g,i)]
        const int hj = (t - C_XB) / 10, gi = (t - C_XB) % 10;
        if (gi >= 9) return 0.f;
        return X[(((gi / 3) * 3 + hj / 3) * 3 + gi % 3) * 3 + hj % 3];
    }
    if (q == 0) return 0.f;
    int c, e, f;
    if (t < C_BB) {                              // BA[(c,e)][f]
        const int ce = (t - C_BA) / 4;
        f = (t - C_BA) % 4;
        c = ce / 3, e = ce % 3;
    } else if (t < C_BB + 30) {                  // BB[o][(l,n)]: Bs[l,n,o]
        f = (t - C_BB) / 10;
        const int ln = (t - C_BB) % 10;
        if (ln >= 9) return 0.f;
        c = ln / 3, e = ln % 3;
    } else {
        return 0.f;
    }
    if (f >= K) return 0.f;
    float v = 0.f;                               // Bs[c][e][f] = sum_d A_q[c][d][e][f] s_{q+1}[d]
    for (int d = 0; d < K; ++d) v = fmaf(a.coreA[q][((c * K + d) * K + e) * K + f], a.state[q + 1][d], v);
    return v;
}

// ---- geometry of one variant: R slots x S samples per warp ----------------------------------------------
template <int R, int NW>
struct Geo {
    static_assert(R == 1 || R == 2 || R == 4 || R == 8, "slots per warp");
    static_assert(NW == 4 || NW == 8, "warps per CTA");
    static constexpr int S = 32 / R;                         // samples per tile
    static constexpr int NU = (K3 + R - 1) / R;              // units (groups of R row blocks) per step
    static constexpr int PO = R == 1 ? 3 : (R == 8 ? 8 : 4);  // padded extent of the o / r position
    static constexpr int PQ = R == 1 ? 9 : (R == 2 ? 10 : (R == 4 ? 12 : 16));   // ... of the (q',r) position
    static constexpr int SP = S + 1;                         // pitch of the per-sample row buffers
    // shared-memory map (floats)
    static constexpr int T2_SZ = K3 * PO * S, U_SZ = K2 * PQ * S, T1_SZ = K2 * NU * 32, M_SZ = K2 * S;
    static constexpr int OFF_T2 = 0, OFF_U = OFF_T2 + T2_SZ, OFF_T1 = OFF_U + U_SZ, OFF_MA = OFF_T1 + T1_SZ,
                         OFF_MB = OFF_MA + M_SZ, OFF_VAL = OFF_MB + M_SZ;
    static constexpr int VAL_SZ = K2 * S + 2 * S + 32;        // last step: partial values | loss terms | d value | table
    static constexpr int OFF_POS = OFF_VAL + K2 * S + 2 * S;  // row block -> unit*32 + slot*S (27 ints)
    static constexpr int FWD_FLOATS = OFF_VAL + VAL_SZ;
    // training only
    static constexpr int OFF_DT2R = FWD_FLOATS;              // d T2 of the step above, [o''][f''][g] x position of p
    static constexpr int OFF_DT2P = OFF_DT2R + T2_SZ;        // per-thread partial sums of d T2, two flush slots
    static constexpr int DT2P_SZ = K3 * 2 * (NW * 32);
    static constexpr int OFF_DUP = OFF_DT2P + DT2P_SZ;       // per-row-block partial sums of d U
    static constexpr int OFF_P2P = OFF_DUP + T1_SZ;          // per-row-block part 2 of d Bs
    static constexpr int P2P_SZ = K * NU * 32;
    static constexpr int OFF_FL = OFF_P2P + P2P_SZ;          // per-warp flush scratch [27][33]
    static constexpr int FL_SZ = NW * K3 * 33;
    static constexpr int OFF_WS = OFF_FL + FL_SZ;            // per-warp row sums [NW][108]
    static constexpr int WS_SZ = NW * 108 + 108;            // + quarter sums of d Bs part 2
    static constexpr int OFF_ROWS = OFF_WS + WS_SZ;          // per-sample rows [81 + 9][SP]: right-copy d X, d As0
    static constexpr int ROWS_SZ = (K4 + K2) * SP;
    static constexpr int TRAIN_FLOATS = OFF_ROWS + ROWS_SZ;
    // global checkpoint of one tile: T2_q for q = 1 .. n-2, laid out like the shared copy
    TNQ_HOSTDEV static constexpr long long ckpt_floats(int n) { return (long long)(n - 2) * T2_SZ; }
};
// per-tile gradient slice: step q at q * GQ: [0,81) dX (left copy) | [81,162) dX (right copy) |
// [162,189) dBs part 1 | [189,216) dBs part 2;  step 0: [162,171) dAs0
constexpr int GQ = 216;
TNQ_HOSTDEV constexpr int grad_floats(int n) { return (n - 1) * GQ; }

// ---- row blocks: unit u, slot -> rb = o*9 + q'*3 + r (or -1: idle) -----------------------------------
// Built so that the row blocks of one unit have o / (q',r) / r positions that are pairwise equal or fall
// into different bank groups (tests/test_ladder2_emu.py checks this exhaustively).
template <int R, int NW>
TNQ_HOSTDEV constexpr int rb_of(int u, int slot) {
    if (R == 1) return u;
    if (R == 2) {
        if (u < 9) return (u / 3) * 9 + slot * 3 + (u % 3);            // (o, q' = slot, r)
        if (u < 12) return slot * 9 + 6 + (u - 9);                     // (o = slot, q' = 2, r)
        if (u == 12) return 24 + slot;                                 // (2, 2, r = slot)
        return slot == 0 ? 26 : -1;
    }
    if (R == 4) {
        if (u < 6) {
            const int o = u >> 1;
            const int qr = (u & 1) == 0 ? (slot >> 1) * 3 + (slot & 1)      // q' in {0,1} x r in {0,1}
                                        : (slot == 0 ? 2 : (slot == 1 ? 5 : (slot == 2 ? 6 : 7)));
            return o * 9 + qr;
        }
        return slot < 3 ? slot * 9 + 8 : -1;
    }
    if (u < 3) return u * 9 + slot;
    return slot < 3 ? slot * 9 + 8 : -1;
}
// inverse: rb -> (unit, slot)
template <int R, int NW>
TNQ_HD void uslot_of(int rb, int& u, int& slot) {
    const int o = rb / 9, qr = rb % 9, qq = qr / 3, r = qr % 3;
    if (R == 1) {
        u = rb, slot = 0;
    } else if (R == 2) {
        if (qq < 2) u = o * 3 + r, slot = qq;
        else if (o < 2) u = 9 + r, slot = o;
        else if (r < 2) u = 12, slot = r;
        else u = 13, slot = 0;
    } else if (R == 4) {
        if (qr == 8) {
            u = 6, slot = o;
        } else if (qq < 2 && r < 2) {
            u = o * 2, slot = qq * 2 + r;
        } else {
            u = o * 2 + 1, slot = qr == 2 ? 0 : (qr == 5 ? 1 : (qr == 6 ? 2 : 3));
        }
    } else {
        if (qr == 8) u = 3, slot = o;
        else u = o, slot = qr;
    }
}
template <int R, int NW>
TNQ_HD int pos_q(int qr) {                       // position of (q',r) in the U layout
    if (R == 4) return qr == 2 ? 4 : (qr == 3 ? 2 : (qr == 4 ? 3 : qr));
    return qr;
}
// units [ustart(w), ustart(w+1)) belong to warp w: NU units dealt as evenly as they go, the first warps one more
// (4 warps: 7,7,7,6 units at R = 1; 4,4,3,3 at R = 2; 2,2,2,1 at R = 4; 1,1,1,1 at R = 8)
template <int R, int NW>
TNQ_HOSTDEV constexpr int ustart(int w) {
    constexpr int NU = (K3 + R - 1) / R, base = NU / NW, rem = NU % NW;
    return w * base + (w < rem ? w : rem);
}

// which o the flush slot fs (0 / 1) of thread group g = warp*R + slot holds after phase_r (-1: unused):
// a thread parks its running d T2 sum whenever the o of its row blocks changes (at most once, by construction)
template <int R, int NW>
TNQ_HOSTDEV constexpr int flush_o(int g, int fs) {
    const int w = g / R, slot = g % R;
    int cur = -1, n = 0;
    for (int u = ustart<R, NW>(w); u < ustart<R, NW>(w + 1); ++u) {
        const int id = rb_of<R, NW>(u, slot);
        if (id < 0) continue;
        if (id / 9 != cur) {
            if (cur >= 0) ++n;
            cur = id / 9;
            if (n == fs) return cur;
        }
    }
    return -1;
}

// ---- context -----------------------------------------------------------------------------------------
struct Ctx {
    float* sm;                  // the CTA's shared memory
    const float* cst;           // constant pool (CPU emulation; the device reads the __constant__ copy)
    const Args* a;
    long long B, b0;            // batch size, first sample of the tile
    float* ck;                  // this CTA's T2 checkpoint slab (global)
    float* gpart;               // this TILE's gradient slice (global)
    float* lpart;               // this tile's loss partial
    const float* seed;          // MODE 2
    float* values;
    float log_scale, inv_count;
};

// per-thread state that lives across phases of the reverse sweep (registers on the device)
struct TS {
    Vec<K2> accX[K2];           // d X: [(g',i')] over (h,j)   (last step: [(p0,o)] over (j,p'))
    Vec<K> accB[K2];            // d Bs part 1: [(f,h)] over f''
};

#ifdef __CUDA_ARCH__
#define TNQ2_CST(off) tnq_l2_cst_dev[(off)]
#else
#define TNQ2_CST(off) c.cst[(off)]
#endif

}  // namespace tnq_l2

#ifdef __CUDACC__
// (this header is included by exactly one translation unit: tnq_ladder2.cu)
__constant__ float tnq_l2_cst_dev[tnq_l2::C_MAX + 8];
#endif

namespace tnq_l2 {

// N floats of the constant pool starting at an EVEN offset (warp-uniform on the device)
template <int N>
TNQ_HD void cvec(const Ctx& c, Vec<N>& v, int off) {
    (void)c;
    TNQ_UNROLL
    for (int i = 0; i < Vec<N>::NP; ++i) v.p[i] = tnq_ladder::F2{TNQ2_CST(off + 2 * i), TNQ2_CST(off + 2 * i + 1)};
    if (N & 1) v.s = TNQ2_CST(off + N - 1);
    else v.s = 0.f;
}

TNQ_HD float load_m(const Ctx& c, int q, long long b, int it) {
    return b < c.B ? ldg_f(c.a->mx[q] + b * c.a->mx_stride[q] + it) : 0.f;
}

// ================================= forward phases ======================================================

// T2_0 = As0 (x) As0 for every sample of the tile
template <int R, int NW>
TNQ_HD void fill_t2_first(const Ctx& c, int tid) {
    using G = Geo<R, NW>;
    float* T2s = c.sm + G::OFF_T2;
    const int base = (c.a->n - 1) * C_STEP + T_AS0;
    for (int e = tid; e < K4 * G::S; e += (NW * 32)) {
        const int s = e % G::S, x = e / G::S;          // x = ((p*3 + f)*3 + g)*3 + o
        const int o = x % 3, g = (x / 3) % 3, f = (x / 9) % 3, p = x / 27;
        T2s[((p * 9 + f * 3 + g) * G::PO + o) * G::S + s] = TNQ2_CST(base + g * 3 + f) * TNQ2_CST(base + p * 3 + o);
    }
}

// phase B: U_q from M_q (task = (i, p, sample); at most two tasks per thread).  The loads of M are split from
// the arithmetic so that callers can put independent work between them (the loads come from global memory).
template <int R, int NW>
struct UTask {
    static constexpr int N = (K2 * Geo<R, NW>::S + (NW * 32) - 1) / (NW * 32);   // tasks per thread
    float m[N][K];
};
template <int R, int NW>
TNQ_HD void phase_u_load(const Ctx& c, int q, int tid, UTask<R, NW>& ut) {
    using G = Geo<R, NW>;
    TNQ_UNROLL
    for (int k2 = 0; k2 < UTask<R, NW>::N; ++k2) {
        const int t = tid + k2 * (NW * 32);
        if (t < K2 * G::S) {
            const int i = t / G::S / 3, s = t % G::S;
            TNQ_UNROLL
            for (int k = 0; k < K; ++k) ut.m[k2][k] = load_m(c, q, c.b0 + s, i * K + k);
        }
    }
}
template <int R, int NW>
TNQ_HD void phase_u_compute(const Ctx& c, int q, int tid, const UTask<R, NW>& ut) {
    using G = Geo<R, NW>;
    float* Us = c.sm + G::OFF_U;
    const int cq = q * C_STEP;
    TNQ_UNROLL
    for (int k2 = 0; k2 < UTask<R, NW>::N; ++k2) {
        const int t = tid + k2 * (NW * 32);
        if (t < K2 * G::S) {
            const int ip = t / G::S, s = t % G::S, p = ip % 3;
            Vec<K2> u;
            u.zero();
            TNQ_UNROLL
            for (int k = 0; k < K; ++k) {
                // X[p,q',k,r] over (q',r): row (p,k) of XA; p differs between the tasks of a warp, so the row is
                // fetched per thread (an indexed constant load), not through the uniform datapath
                const int row = cq + C_XA + (p * 3 + k) * 10;
                TNQ_UNROLL
                for (int qr = 0; qr < K2; ++qr) u.set(qr, fmaf(ut.m[k2][k], TNQ2_CST(row + qr), u.get(qr)));
            }
            TNQ_UNROLL
            for (int qr = 0; qr < K2; ++qr) Us[(ip * G::PQ + pos_q<R, NW>(qr)) * G::S + s] = u.get(qr);
        }
    }
}
template <int R, int NW>
TNQ_HD void phase_u(const Ctx& c, int q, int tid) {
    UTask<R, NW> ut;
    phase_u_load<R, NW>(c, q, tid, ut);
    phase_u_compute<R, NW>(c, q, tid, ut);
}

// copy of M_q for the tasks that share it: Ms[(row*3+col)][sample]
template <int R, int NW>
TNQ_HD void load_ms(const Ctx& c, int q, float* Ms, int tid) {
    using G = Geo<R, NW>;
    for (int e = tid; e < K2 * G::S; e += (NW * 32)) Ms[e] = load_m(c, q, c.b0 + e % G::S, e / G::S);
}

// phase A2: T2_q from T1_q (task = (f, p, g, sample)); training keeps a copy in the checkpoint slab
template <int R, int NW, bool CK>
TNQ_HD void phase_a2(const Ctx& c, int q, int tid) {
    using G = Geo<R, NW>;
    const float* T1s = c.sm + G::OFF_T1;
    float* T2s = c.sm + G::OFF_T2;
    const int* pos = reinterpret_cast<const int*>(c.sm + G::OFF_POS);
    const int cq = q * C_STEP;
    TNQ_NOUNROLL
    for (int t = tid; t < K3 * G::S; t += (NW * 32)) {
        const int s = t % G::S, x = t / G::S;              // x = (f*3 + p)*3 + g
        const int g = x % 3, p = (x / 3) % 3, f = x / 9;
        Vec<K> acc;
        acc.zero();
        TNQ_UNROLL
        for (int ln = 0; ln < K2; ++ln) {
            const float v = T1s[(f * 3 + g) * G::NU * 32 + pos[ln * 3 + p] + s];
            Vec<K> b;
            cvec<K>(c, b, cq + C_BA + ln * 4);             // Bs[l,n,o] over o
            axpy(acc, v, b);
        }
        TNQ_UNROLL
        for (int o = 0; o < K; ++o) {
            const int at = ((p * 9 + f * 3 + g) * G::PO + o) * G::S + s;
            T2s[at] = acc.get(o);
            if (CK) c.ck[(long long)(q - 1) * G::T2_SZ + at] = acc.get(o);
        }
    }
}

// this thread's operands of one row block
template <int R, int NW>
struct RowBlock {
    int o, qr, u;
    const float* t2;            // + ((p'*9 + f*3+g') * PO) * S
    const float* uu;            // + ((i'*3+p') * PQ) * S
};
// (an idle slot -- there is at most one unit with idle slots per step -- computes on row block 0 and drops its
// results: no divergent control flow inside the hot loops)
template <int R, int NW>
TNQ_HD bool row_block(const Ctx& c, int u, int lane, RowBlock<R, NW>& rb) {
    using G = Geo<R, NW>;
    const int slot = lane / G::S, s = lane % G::S;
    int id = rb_of<R, NW>(u, slot);
    const bool active = id >= 0;
    id = active ? id : 0;
    rb.o = id / 9, rb.qr = id % 9, rb.u = u;
    rb.t2 = c.sm + G::OFF_T2 + rb.o * G::S + s;
    rb.uu = c.sm + G::OFF_U + pos_q<R, NW>(rb.qr) * G::S + s;
    return active;
}

// V[i'] over (f,g') = sum_p' T2[f,o,g',p'] U[i',p',q',r]
template <int R, int NW>
TNQ_HD void make_v(const RowBlock<R, NW>& rb, const float (&u9)[K2], Vec<K2> (&V)[K]) {
    using G = Geo<R, NW>;
    TNQ_UNROLL
    for (int i = 0; i < K; ++i) V[i].zero();
    TNQ_UNROLL
    for (int p = 0; p < K; ++p) {
        Vec<K2> t;
        TNQ_UNROLL
        for (int fg = 0; fg < K2; ++fg) t.set(fg, rb.t2[(p * 9 + fg) * G::PO * G::S]);
        TNQ_UNROLL
        for (int i = 0; i < K; ++i) axpy(V[i], u9[i * 3 + p], t);
    }
}
// E'[f] over (h,j) = sum_{g',i'} V[i'][(f,g')] X_q[g',h,i',j]
TNQ_HD void make_e(const Ctx& c, int cq, const Vec<K2> (&V)[K], Vec<K2> (&E)[K]) {
    TNQ_UNROLL
    for (int f = 0; f < K; ++f) E[f].zero();
    TNQ_UNROLL
    for (int g = 0; g < K; ++g)
        TNQ_UNROLL
        for (int i = 0; i < K; ++i) {
            Vec<K2> xa;
            cvec<K2>(c, xa, cq + C_XA + (g * 3 + i) * 10);
            TNQ_UNROLL
            for (int f = 0; f < K; ++f) axpy(E[f], V[i].get(f * 3 + g), xa);
        }
}

// fused phase C of step q and phase A1 of step q+1, one row block (o,q',r) at a time
template <int R, int NW>
TNQ_HD void phase_c_a1(const Ctx& c, int q, int warp, int lane) {
    using G = Geo<R, NW>;
    float* T1s = c.sm + G::OFF_T1;
    const int cq = q * C_STEP, cq1 = (q + 1) * C_STEP;
    TNQ_NOUNROLL
    for (int u = ustart<R, NW>(warp); u < ustart<R, NW>(warp + 1); ++u) {
        RowBlock<R, NW> rb;
        const bool active = row_block<R, NW>(c, u, lane, rb);
        float u9[K2];
        TNQ_UNROLL
        for (int x = 0; x < K2; ++x) u9[x] = rb.uu[x * G::PQ * G::S];
        Vec<K2> V[K], E[K];
        make_v<R, NW>(rb, u9, V);
        make_e(c, cq, V, E);
        Vec<K> T1[K];                      // [j] over f''
        TNQ_UNROLL
        for (int j = 0; j < K; ++j) T1[j].zero();
        TNQ_UNROLL
        for (int fh = 0; fh < K2; ++fh) {
            Vec<K> b;
            cvec<K>(c, b, cq1 + C_BA + fh * 4);            // Bs'[f,h,f''] over f''
            TNQ_UNROLL
            for (int j = 0; j < K; ++j) axpy(T1[j], E[fh / 3].get((fh % 3) * 3 + j), b);
        }
        if (active) {
            TNQ_UNROLL
            for (int f2 = 0; f2 < K; ++f2)
                TNQ_UNROLL
                for (int j = 0; j < K; ++j) T1s[((f2 * 3 + j) * G::NU + u) * 32 + lane] = T1[j].get(f2);
        }
    }
}

// last step, shared by the forward and the reverse pass: L[j][p'] and Z over (p0,o) of task (g,f)
template <int R, int NW>
TNQ_HD void last_lz(const Ctx& c, int gf, int s, float (&mh)[K][K], float (&mi)[K][K], float (&L)[K][K], Vec<K2>& Z) {
    using G = Geo<R, NW>;
    const int n = c.a->n;
    const float* MA = c.sm + G::OFF_MA;
    const float* MB = c.sm + G::OFF_MB;
    float x[K][K];
    TNQ_UNROLL
    for (int h = 0; h < K; ++h)
        TNQ_UNROLL
        for (int i = 0; i < K; ++i) {
            x[h][i] = ldg_f(c.a->coreX[n - 2] + gf * K2 + h * 3 + i);
            mh[h][i] = MA[(h * 3 + i) * G::S + s];
            mi[h][i] = MB[(h * 3 + i) * G::S + s];
        }
    float tmp[K][K];                       // [h][p'] = sum_i x[h][i] Mi[i][p']
    TNQ_UNROLL
    for (int h = 0; h < K; ++h)
        TNQ_UNROLL
        for (int pp = 0; pp < K; ++pp) {
            float v = 0.f;
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) v = fmaf(x[h][i], mi[i][pp], v);
            tmp[h][pp] = v;
        }
    TNQ_UNROLL
    for (int j = 0; j < K; ++j)
        TNQ_UNROLL
        for (int pp = 0; pp < K; ++pp) {
            float v = 0.f;
            TNQ_UNROLL
            for (int h = 0; h < K; ++h) v = fmaf(mh[h][j], tmp[h][pp], v);
            L[j][pp] = v;
        }
    Z.zero();
    const int tail = (n - 1) * C_STEP;
    TNQ_UNROLL
    for (int jp = 0; jp < K2; ++jp) {
        Vec<K2> xc;
        cvec<K2>(c, xc, tail + T_XC + jp * 10);            // X[p0,o,j,p'] over (p0,o)
        axpy(Z, L[jp / 3][jp % 3], xc);
    }
}

template <int R, int NW>
TNQ_HD void last_fwd(const Ctx& c, int tid) {
    using G = Geo<R, NW>;
    const float* T2s = c.sm + G::OFF_T2;
    float* valp = c.sm + G::OFF_VAL;
    TNQ_NOUNROLL
    for (int t = tid; t < K2 * G::S; t += (NW * 32)) {
        const int gf = t / G::S, s = t % G::S, g = gf / 3, f = gf % 3;
        float mh[K][K], mi[K][K], L[K][K];
        Vec<K2> Z;
        last_lz<R, NW>(c, gf, s, mh, mi, L, Z);
        float v = 0.f;
        TNQ_UNROLL
        for (int po = 0; po < K2; ++po)
            v = fmaf(T2s[(((po / 3) * 9 + f * 3 + g) * G::PO + po % 3) * G::S + s], Z.get(po), v);
        valp[gf * G::S + s] = v;
    }
}

// value, loss and d loss / d value of the tile's samples (MODE 0: values only)
template <int R, int NW, int MODE>
TNQ_HD void last_value(const Ctx& c, int tid) {
    using G = Geo<R, NW>;
    if (tid >= G::S) return;
    float* valp = c.sm + G::OFF_VAL;
    float val = 0.f;
    TNQ_UNROLL
    for (int gf = 0; gf < K2; ++gf) val += valp[gf * G::S + tid];
    const long long b = c.b0 + tid;
    const bool valid = b < c.B;
    if (MODE != 2 && valid && c.values != nullptr) c.values[b] = val;
    float dv = 0.f, lo = 0.f;
    if (MODE == 1) {
        const float cl = val > 1e-10f ? val : 1e-10f;
        if (valid) {
            lo = -(logf(cl) + c.log_scale) * c.inv_count;
            dv = val >= 1e-10f ? -c.inv_count / cl : 0.f;
        }
    } else if (MODE == 2) {
        dv = valid ? ldg_f(c.seed + b) : 0.f;
    }
    valp[K2 * G::S + tid] = lo;
    valp[K2 * G::S + G::S + tid] = dv;
}

// ================================= reverse phases ======================================================

// reverse of the last step (task = (g,f,sample)): d T2 (complete), d X of both copies
template <int R, int NW>
TNQ_HD void last_bwd(const Ctx& c, TS& ts, int tid) {
    using G = Geo<R, NW>;
    const int n = c.a->n;
    const float* T2s = c.sm + G::OFF_T2;
    float* dT2r = c.sm + G::OFF_DT2R;
    float* rows = c.sm + G::OFF_ROWS;      // XL rows [(g,f,h,i)][sample]
    const float* dval = c.sm + G::OFF_VAL + K2 * G::S + G::S;
    TNQ_UNROLL
    for (int x = 0; x < K2; ++x) ts.accX[x].zero();
    TNQ_UNROLL
    for (int x = 0; x < K2; ++x) ts.accB[x].zero();
    const int tail = (n - 1) * C_STEP;
    TNQ_NOUNROLL
    for (int t = tid; t < K2 * G::S; t += (NW * 32)) {
        const int gf = t / G::S, s = t % G::S, g = gf / 3, f = gf % 3;
        float mh[K][K], mi[K][K], L[K][K];
        Vec<K2> Z;
        last_lz<R, NW>(c, gf, s, mh, mi, L, Z);
        const float dv = dval[s];
        Vec<K2> Lv, dL;
        TNQ_UNROLL
        for (int jp = 0; jp < K2; ++jp) Lv.set(jp, L[jp / 3][jp % 3]);
        dL.zero();
        TNQ_UNROLL
        for (int po = 0; po < K2; ++po) {
            const int p0 = po / 3, o = po % 3;
            const float dz = dv * T2s[((p0 * 9 + f * 3 + g) * G::PO + o) * G::S + s];
            dT2r[((o * 9 + f * 3 + g) * G::PO + p0) * G::S + s] = dv * Z.get(po);
            axpy(ts.accX[po], dz, Lv);                          // d X[p0,o,j,p'] (right copy)
            Vec<K2> xn;
            cvec<K2>(c, xn, tail + T_XN + po * 10);             // X[p0,o,j,p'] over (j,p')
            axpy(dL, dz, xn);
        }
        float dtmp[K][K];                  // [h][p'] = sum_j Mh[h][j] dL[j][p']
        TNQ_UNROLL
        for (int h = 0; h < K; ++h)
            TNQ_UNROLL
            for (int pp = 0; pp < K; ++pp) {
                float v = 0.f;
                TNQ_UNROLL
                for (int j = 0; j < K; ++j) v = fmaf(mh[h][j], dL.get(j * 3 + pp), v);
                dtmp[h][pp] = v;
            }
        TNQ_UNROLL
        for (int h = 0; h < K; ++h)
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) {
                float v = 0.f;
                TNQ_UNROLL
                for (int pp = 0; pp < K; ++pp) v = fmaf(dtmp[h][pp], mi[i][pp], v);
                rows[(gf * K2 + h * 3 + i) * G::SP + s] = v;     // d X[g,f,h,i] (left copy), this sample
            }
    }
}

// flush scratch of warp `warp`: rows [27*ROUND, 27*ROUND+27) of this lane's accumulators
template <int R, int NW, int ROUND>
TNQ_HD void flush_put(const Ctx& c, const TS& ts, int warp, int lane) {
    using G = Geo<R, NW>;
    float* buf = c.sm + G::OFF_FL + warp * (K3 * 33);
    TNQ_UNROLL
    for (int r = 0; r < K3; ++r) {
        if (ROUND < 3) buf[r * 33 + lane] = ts.accX[(ROUND * K3 + r) / K2].get((ROUND * K3 + r) % K2);
        else buf[r * 33 + lane] = ts.accB[r / 3].get(r % 3);
    }
}
template <int R, int NW>
TNQ_HD void flush_sum(const Ctx& c, int round, int warp, int lane) {
    using G = Geo<R, NW>;
    if (lane >= K3) return;
    const float* buf = c.sm + G::OFF_FL + warp * (K3 * 33) + lane * 33;
    float t0 = 0.f, t1 = 0.f, t2 = 0.f, t3 = 0.f;         // (fixed association: deterministic)
    TNQ_UNROLL
    for (int l = 0; l < 32; l += 4) t0 += buf[l], t1 += buf[l + 1], t2 += buf[l + 2], t3 += buf[l + 3];
    c.sm[G::OFF_WS + warp * 108 + round * K3 + lane] = (t0 + t1) + (t2 + t3);
}

// row block -> position of its lane run in the row-block-indexed buffers (unit*32 + slot*S): once per CTA
template <int R, int NW>
TNQ_HD void build_pos(const Ctx& c, int tid) {
    using G = Geo<R, NW>;
    if (tid >= K3) return;
    int u, slot;
    uslot_of<R, NW>(tid, u, slot);
    reinterpret_cast<int*>(c.sm + G::OFF_POS)[tid] = u * 32 + slot * G::S;
}

// reverse of the fused phase (row block (o,q',r)): from d T1_{q+1}, d T2_{q+1}, T2_q, U_q
//   E' (recomputed), d Bs_{q+1} (both parts), d E', d X_q (left copy), d V, d T2_q and d U_q partial sums
template <int R, int NW>
TNQ_HD void phase_r(const Ctx& c, TS& ts, int q, int warp, int lane) {
    using G = Geo<R, NW>;
    const float* dT1s = c.sm + G::OFF_T1;
    const float* dT2r = c.sm + G::OFF_DT2R;
    float* dT2p = c.sm + G::OFF_DT2P;
    float* dUp = c.sm + G::OFF_DUP;
    float* p2p = c.sm + G::OFF_P2P;
    const int cq = q * C_STEP, cq1 = (q + 1) * C_STEP;
    const int s = lane % G::S, tid = warp * 32 + lane;
    TNQ_UNROLL
    for (int x = 0; x < K2; ++x) ts.accX[x].zero();
    TNQ_UNROLL
    for (int x = 0; x < K2; ++x) ts.accB[x].zero();
    Vec<K2> dT2acc[K];                     // [p'] over (f,g')
    TNQ_UNROLL
    for (int p = 0; p < K; ++p) dT2acc[p].zero();
    int cur_o = -1, fs = 0;
    TNQ_NOUNROLL
    for (int u = ustart<R, NW>(warp); u < ustart<R, NW>(warp + 1); ++u) {
        RowBlock<R, NW> rb;
        const bool active = row_block<R, NW>(c, u, lane, rb);
        if (active && rb.o != cur_o) {      // the running d T2 sum belongs to another o: park it
            if (cur_o >= 0) {
                TNQ_UNROLL
                for (int x = 0; x < K3; ++x) dT2p[(x * 2 + fs) * (NW * 32) + tid] = dT2acc[x / 9].get(x % 9);
                TNQ_UNROLL
                for (int p = 0; p < K; ++p) dT2acc[p].zero();
                ++fs;
            }
            cur_o = rb.o;
        }
        float u9[K2];
        TNQ_UNROLL
        for (int x = 0; x < K2; ++x) u9[x] = rb.uu[x * G::PQ * G::S];
        Vec<K> d1[K];                      // d T1_{q+1}: [j] over f''
        TNQ_UNROLL
        for (int j = 0; j < K; ++j)
            TNQ_UNROLL
            for (int f2 = 0; f2 < K; ++f2) d1[j].set(f2, active ? dT1s[((f2 * 3 + j) * G::NU + u) * 32 + lane] : 0.f);
        Vec<K2> V[K];
        make_v<R, NW>(rb, u9, V);
        {   // E' -> d Bs part 1, T1 (recomputed) -> d Bs part 2
            Vec<K2> E[K];
            make_e(c, cq, V, E);
            Vec<K> T1c[K];                 // [j] over f''
            TNQ_UNROLL
            for (int j = 0; j < K; ++j) T1c[j].zero();
            TNQ_UNROLL
            for (int fh = 0; fh < K2; ++fh) {
                Vec<K> b;
                cvec<K>(c, b, cq1 + C_BA + fh * 4);
                TNQ_UNROLL
                for (int j = 0; j < K; ++j) {
                    const float e = E[fh / 3].get((fh % 3) * 3 + j);
                    axpy(ts.accB[fh], e, d1[j]);
                    axpy(T1c[j], e, b);
                }
            }
            const int r = rb.qr % 3;
            Vec<K2> T1v;                   // over (f'',j): the order of the d T2 rows below
            TNQ_UNROLL
            for (int x = 0; x < K2; ++x) T1v.set(x, T1c[x % 3].get(x / 3));
            TNQ_UNROLL
            for (int o2 = 0; o2 < K; ++o2) {
                Vec<K2> t;                 // d T2_{q+1}[f'',o'',g=j,p=r] over (f'',j): all loads first, then one dot product
                TNQ_UNROLL
                for (int x = 0; x < K2; ++x) t.set(x, dT2r[((o2 * 9 + x) * G::PO + r) * G::S + s]);
                const float acc = dot(T1v, t, 0.f);
                if (active) p2p[(o2 * G::NU + u) * 32 + lane] = acc;
            }
        }
        Vec<K2> dE[K];                     // [f] over (h,j) = sum_f'' Bs'[f,h,f''] d T1[f'',j]
        TNQ_UNROLL
        for (int fh = 0; fh < K2; ++fh) {
            Vec<K> b;
            cvec<K>(c, b, cq1 + C_BA + fh * 4);
            TNQ_UNROLL
            for (int j = 0; j < K; ++j) dE[fh / 3].set((fh % 3) * 3 + j, dot(b, d1[j], 0.f));
        }
        TNQ_UNROLL
        for (int g = 0; g < K; ++g)
            TNQ_UNROLL
            for (int i = 0; i < K; ++i)
                TNQ_UNROLL
                for (int f = 0; f < K; ++f) axpy(ts.accX[g * 3 + i], V[i].get(f * 3 + g), dE[f]);
        Vec<K2> W[K];                      // d V: [i'] over (f,g')
        {
            Vec<K2> dV[K];                 // [f] over (g',i') = sum_(h,j) d E'[f][(h,j)] X[g',h,i',j]
            TNQ_UNROLL
            for (int f = 0; f < K; ++f) dV[f].zero();
            TNQ_UNROLL
            for (int hj = 0; hj < K2; ++hj) {
                Vec<K2> xb;
                cvec<K2>(c, xb, cq + C_XB + hj * 10);
                TNQ_UNROLL
                for (int f = 0; f < K; ++f) axpy(dV[f], dE[f].get(hj), xb);
            }
            TNQ_UNROLL
            for (int i = 0; i < K; ++i)
                TNQ_UNROLL
                for (int fg = 0; fg < K2; ++fg) W[i].set(fg, dV[fg / 3].get((fg % 3) * 3 + i));
        }
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) {
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) axpy(dT2acc[p], u9[i * 3 + p], W[i]);
            Vec<K2> t;
            TNQ_UNROLL
            for (int fg = 0; fg < K2; ++fg) t.set(fg, rb.t2[(p * 9 + fg) * G::PO * G::S]);
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) {
                const float du = dot(W[i], t, 0.f);
                if (active) dUp[((i * 3 + p) * G::NU + u) * 32 + lane] = du;
            }
        }
    }
    if (cur_o >= 0) {
        TNQ_UNROLL
        for (int x = 0; x < K3; ++x) dT2p[(x * 2 + fs) * (NW * 32) + tid] = dT2acc[x / 9].get(x % 9);
    }
}

// sum of per-sample rows [nrows][SP] (starting at row0 of the row buffer) over the samples -> gradient slice
template <int R, int NW>
TNQ_HD void rows_to_grad(const Ctx& c, float* dst, int row0, int nrows, int tid) {
    using G = Geo<R, NW>;
    if (tid >= nrows) return;
    const float* rows = c.sm + G::OFF_ROWS + (row0 + tid) * G::SP;
    float t0 = 0.f, t1 = 0.f;
    for (int s = 0; s < G::S; s += 2) t0 += rows[s], t1 += rows[s + 1];
    dst[tid] = t0 + t1;
}

// operands of phase_r(q): T2_q (checkpoint, or As0 (x) As0 for q = 0), U_q, and M_q in the buffer of q's parity.
// prepare_r_issue starts the global loads (the checkpoint travels with cp.async on the device); prepare_r_finish
// does the arithmetic and waits for the copy: callers put independent work between the two.
template <int R, int NW>
TNQ_HD void prepare_r_issue(const Ctx& c, int q, int tid, UTask<R, NW>& ut) {
    using G = Geo<R, NW>;
    if (q >= 1) {
        float* T2s = c.sm + G::OFF_T2;
        const float* ck = c.ck + (long long)(q - 1) * G::T2_SZ;
        static_assert(G::T2_SZ % 4 == 0, "16-byte units");
#ifdef __CUDA_ARCH__
        const uint32_t d = (uint32_t)__cvta_generic_to_shared(T2s);
        for (int e = tid; e < G::T2_SZ / 4; e += (NW * 32))
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d + 16u * e), "l"(ck + 4 * e) : "memory");
        asm volatile("cp.async.commit_group;" ::: "memory");
#else
        for (int e = tid; e < G::T2_SZ; e += (NW * 32)) T2s[e] = ck[e];
#endif
    }
    phase_u_load<R, NW>(c, q, tid, ut);
}
template <int R, int NW>
TNQ_HD void prepare_r_finish(const Ctx& c, int q, int tid, const UTask<R, NW>& ut) {
    using G = Geo<R, NW>;
    if (q == 0) fill_t2_first<R, NW>(c, tid);
    phase_u_compute<R, NW>(c, q, tid, ut);
    load_ms<R, NW>(c, q, c.sm + ((q & 1) ? G::OFF_MB : G::OFF_MA), tid);
#ifdef __CUDA_ARCH__
    asm volatile("cp.async.wait_group 0;" ::: "memory");
#endif
}

// Phase X(q): everything between two big phases of the reverse sweep, in ONE phase:
//   * flushed accumulators -> gradient slice (after phase_r(q): d X_q left copy and d Bs_{q+1} part 1; after the
//     reverse of the last step: d X right copy in natural order, and the left-copy rows)
//   * d Bs_{q+1} part 2: quarter sums (finished by finish_prev)
//   * d T2_q = sum of the parked partial sums (already complete after the last step), stored where phase_r(q-1)
//     reads it, and in the same task d T1_q = A2^T(d T2_q)                     (task = (p, f, g, sample))
//   * d U_q = sum over o of the partial sums, right-copy d X_q per sample -> row buffer (finished by finish_prev)
//   * the operands of phase_r(q-1)
template <int R, int NW>
TNQ_HD void phase_x(const Ctx& c, int q, bool after_r, int tid) {
    using G = Geo<R, NW>;
    const float* ws = c.sm + G::OFF_WS;
    const int* pos = reinterpret_cast<const int*>(c.sm + G::OFF_POS);
    float* gq = c.gpart + (long long)q * GQ;
    UTask<R, NW> ut;
    if (q >= 1) prepare_r_issue<R, NW>(c, q - 1, tid, ut);
    if (tid < 108) {
        float t = (ws[tid] + ws[108 + tid]) + (ws[216 + tid] + ws[324 + tid]);
        if (NW == 8) t += (ws[432 + tid] + ws[540 + tid]) + (ws[648 + tid] + ws[756 + tid]);
        if (!after_r) {
            if (tid < K4) gq[K4 + tid] = t;
        } else if (tid < K4) {
            const int gi = tid / 9, hj = tid % 9;           // rows are (g',i',h,j): d X[g',h,i',j] (left copy)
            gq[(((gi / 3) * 3 + hj / 3) * 3 + gi % 3) * 3 + hj % 3] = t;
        } else {
            gq[GQ + 162 + tid - K4] = t;                    // d Bs_{q+1} part 1, natural (c,e,f)
        }
    }
    if (!after_r) rows_to_grad<R, NW>(c, gq, 0, K4, tid);       // left-copy rows of the last step
    if (after_r && tid < 108) {
        // d Bs_{q+1} part 2: [(l,n)][o''] = sum over p and the samples; here: four quarter sums per entry
        const int row = tid >> 2, part = tid & 3, ln = row / 3, o2 = row % 3;
        const float* p2p = c.sm + G::OFF_P2P + o2 * G::NU * 32;
        float t = 0.f;
        TNQ_UNROLL
        for (int p = 0; p < K; ++p) {
            const float* src = p2p + pos[ln * 3 + p];
            for (int s = part; s < G::S; s += 4) t += src[s];
        }
        c.sm[G::OFF_WS + NW * 108 + tid] = t;
    }
    {
        const float* dT2p = c.sm + G::OFF_DT2P;
        float* dT2r = c.sm + G::OFF_DT2R;
        float* dT1s = c.sm + G::OFF_T1;
        const int cq = q * C_STEP;
        TNQ_NOUNROLL
        for (int t = tid; t < K3 * G::S; t += (NW * 32)) {
            const int s = t % G::S, x = t / G::S, p = x / 9, fg = x % 9;    // x = p*9 + f*3 + g
            float d[K] = {0.f, 0.f, 0.f};
            if (after_r) {
                TNQ_UNROLL
                for (int g = 0; g < NW * R; ++g)
                    TNQ_UNROLL
                    for (int fs = 0; fs < 2; ++fs) {
                        const int o = flush_o<R, NW>(g, fs);    // (compile time after unrolling)
                        if (o >= 0) d[o] += dT2p[(x * 2 + fs) * (NW * 32) + (g / R) * 32 + (g % R) * G::S + s];
                    }
                TNQ_UNROLL
                for (int o = 0; o < K; ++o) dT2r[((o * 9 + fg) * G::PO + p) * G::S + s] = d[o];
            } else {
                TNQ_UNROLL
                for (int o = 0; o < K; ++o) d[o] = dT2r[((o * 9 + fg) * G::PO + p) * G::S + s];
            }
            if (q >= 1) {
                Vec<K2> out;               // d T1_q[f,(l,n),p,g] over (l,n) = sum_o Bs_q[l,n,o] d T2_q[f,o,g,p]
                out.zero();
                TNQ_UNROLL
                for (int o = 0; o < K; ++o) {
                    Vec<K2> bb;
                    cvec<K2>(c, bb, cq + C_BB + o * 10);
                    axpy(out, d[o], bb);
                }
                float* dst = dT1s + fg * G::NU * 32 + s;
                TNQ_UNROLL
                for (int ln = 0; ln < K2; ++ln) dst[pos[ln * 3 + p]] = out.get(ln);
            }
        }
    }
    if (after_r) {
        const float* dUp = c.sm + G::OFF_DUP;
        const float* Mq = c.sm + ((q & 1) ? G::OFF_MB : G::OFF_MA);
        float* rows = c.sm + G::OFF_ROWS;
        TNQ_NOUNROLL
        for (int t = tid; t < K3 * G::S; t += (NW * 32)) {
            const int s = t % G::S, e = t / G::S, p = e / 9, qr = e % 9;    // e = p'*9 + qr
            const int a0 = pos[qr] + s, a1 = pos[9 + qr] + s, a2 = pos[18 + qr] + s;
            float du[K];
            TNQ_UNROLL
            for (int i = 0; i < K; ++i) {
                const float* dd = dUp + (i * 3 + p) * G::NU * 32;
                du[i] = (dd[a0] + dd[a1]) + dd[a2];
            }
            TNQ_UNROLL
            for (int k = 0; k < K; ++k) {
                float v = 0.f;
                TNQ_UNROLL
                for (int i = 0; i < K; ++i) v = fmaf(Mq[(i * 3 + k) * G::S + s], du[i], v);
                rows[(((p * 3 + qr / 3) * 3 + k) * 3 + qr % 3) * G::SP + s] = v;
            }
        }
    }
    if (q >= 1) prepare_r_finish<R, NW>(c, q - 1, tid, ut);
}

// what phase X(q) left unfinished (run one barrier later): right-copy d X_q rows and d Bs_{q+1} part 2 -> gradient slice
template <int R, int NW>
TNQ_HD void finish_prev(const Ctx& c, int q, int tid) {
    using G = Geo<R, NW>;
    // (given to the LAST warps: they own one row block less in the phase_r that follows)
    rows_to_grad<R, NW>(c, c.gpart + (long long)q * GQ + K4, 0, K4, (NW * 32) - 1 - tid);
    if (tid < K3) {
        const float* qs = c.sm + G::OFF_WS + NW * 108 + tid * 4;
        c.gpart[(long long)(q + 1) * GQ + 189 + tid] = (qs[0] + qs[1]) + (qs[2] + qs[3]);
    }
}

// first step: d As0[e][f'] = sum_{p,o} d T2_0[f',o,e,p] As0[p][o] + sum_{g,f} d T2_0[f,f',g,e] As0[g][f], per sample
template <int R, int NW>
TNQ_HD void das0_rows(const Ctx& c, int tid) {
    using G = Geo<R, NW>;
    const float* dT2r = c.sm + G::OFF_DT2R;
    float* rows = c.sm + G::OFF_ROWS + K4 * G::SP;
    const int base = (c.a->n - 1) * C_STEP + T_AS0;
    TNQ_NOUNROLL
    for (int t = tid; t < K2 * G::S; t += (NW * 32)) {
        const int s = t % G::S, ef = t / G::S, e = ef / 3, f1 = ef % 3;
        float v = 0.f;
        for (int a1 = 0; a1 < K; ++a1)
            for (int b1 = 0; b1 < K; ++b1) {
                // first term: (p,o) = (a1,b1);  second term: (g,f) = (a1,b1)
                v = fmaf(dT2r[((b1 * 9 + f1 * 3 + e) * G::PO + a1) * G::S + s], TNQ2_CST(base + a1 * 3 + b1), v);
                v = fmaf(dT2r[((f1 * 9 + b1 * 3 + a1) * G::PO + e) * G::S + s], TNQ2_CST(base + a1 * 3 + b1), v);
            }
        rows[ef * G::SP + s] = v;
    }
}

}  // namespace tnq_l2

// -------------------------------------------------------------------------------------------------------
// The sweep of one tile.  TNQ2_PH(body) runs `body` for every thread of the CTA and then makes the
// shared-memory writes visible to the whole CTA (device: __syncthreads(); CPU emulation: a loop over
// the 128 threads); TNQ2_PHW is the same with warp scope (device: __syncwarp()).
// -------------------------------------------------------------------------------------------------------
#ifdef __CUDACC__
#define TNQ2_PH(...)  \
    { __VA_ARGS__ }   \
    __syncthreads();
#define TNQ2_PHW(...) \
    { __VA_ARGS__ }   \
    __syncwarp();
#define TNQ2_THREAD_PARAM TS &ts, const int tid
#else
#define TNQ2_PH(...)                        \
    for (int tid = 0; tid < (NW * 32); ++tid) {    \
        TS& ts = tss[tid];                  \
        (void)ts;                           \
        __VA_ARGS__                         \
    }
#define TNQ2_PHW(...) TNQ2_PH(__VA_ARGS__)
#define TNQ2_THREAD_PARAM TS* tss
#endif

namespace tnq_l2 {

// MODE 0: values.  MODE 1: values + fused loss + gradients.  MODE 2: gradients seeded by c.seed.
template <int R, int NW, int MODE>
TNQ_HD void tile_sweep(const Ctx& c, TNQ2_THREAD_PARAM) {
    using G = Geo<R, NW>;
    const int n = c.a->n;
    constexpr bool CK = MODE != 0;
#define TNQ2_WARP (tid >> 5)
#define TNQ2_LANE (tid & 31)
    // ------------------------------- forward sweep -------------------------------
    TNQ2_PH(fill_t2_first<R, NW>(c, tid); phase_u<R, NW>(c, 0, tid);)
    for (int q = 0; q <= n - 3; ++q) {
        TNQ2_PH(phase_c_a1<R, NW>(c, q, TNQ2_WARP, TNQ2_LANE);)
        TNQ2_PH(
            phase_a2<R, NW, CK>(c, q + 1, tid);
            if (q + 1 <= n - 3) {
                phase_u<R, NW>(c, q + 1, tid);
            } else {
                load_ms<R, NW>(c, n - 2, c.sm + G::OFF_MA, tid);
                load_ms<R, NW>(c, n - 1, c.sm + G::OFF_MB, tid);
            })
    }
    TNQ2_PH(last_fwd<R, NW>(c, tid);)
    TNQ2_PH(last_value<R, NW, MODE>(c, tid);)
    if (MODE == 0) return;
    // ------------------------------- reverse sweep -------------------------------
    TNQ2_PHW(
        if (MODE == 1 && tid == 0) {
            float t = 0.f;
            for (int s = 0; s < G::S; ++s) t += c.sm[G::OFF_VAL + K2 * G::S + s];
            *c.lpart = t;
        }
        last_bwd<R, NW>(c, ts, tid);)
#define TNQ2_FLUSH(ROUND)                                                  \
    TNQ2_PHW(flush_put<R, NW, ROUND>(c, ts, TNQ2_WARP, TNQ2_LANE);)             \
    TNQ2_PHW(flush_sum<R, NW>(c, ROUND, TNQ2_WARP, TNQ2_LANE);)
#define TNQ2_FLUSH_LAST(ROUND)                                             \
    TNQ2_PHW(flush_put<R, NW, ROUND>(c, ts, TNQ2_WARP, TNQ2_LANE);)             \
    TNQ2_PH(flush_sum<R, NW>(c, ROUND, TNQ2_WARP, TNQ2_LANE);)
    TNQ2_FLUSH(0)
    TNQ2_FLUSH(1)
    TNQ2_FLUSH_LAST(2)
    TNQ2_PH(phase_x<R, NW>(c, n - 2, false, tid);)
    for (int q = n - 3; q >= 0; --q) {
        TNQ2_PHW(
            if (q < n - 3) finish_prev<R, NW>(c, q + 1, tid);
            phase_r<R, NW>(c, ts, q, TNQ2_WARP, TNQ2_LANE);)
        TNQ2_FLUSH(0)
        TNQ2_FLUSH(1)
        TNQ2_FLUSH(2)
        TNQ2_FLUSH_LAST(3)
        TNQ2_PH(phase_x<R, NW>(c, q, true, tid);)
    }
    TNQ2_PH(finish_prev<R, NW>(c, 0, tid); das0_rows<R, NW>(c, tid);)
    TNQ2_PH(rows_to_grad<R, NW>(c, c.gpart + 162, K4, K2, tid);)
#undef TNQ2_FLUSH_LAST
#undef TNQ2_FLUSH
#undef TNQ2_WARP
#undef TNQ2_LANE
}

// Finalize: element v of the gradient of core A_q (layer 0) or X_q (layer 1), summed over the tiles
// [t0, t1) in tile order, with the circuit states folded back in for layer 0 (the caller combines the
// chunks in a fixed order):  dA_q[c,d,e,f] = dBs_q[c,e,f] s_{q+1}[d] (q >= 1), dA_0 = dAs0[e,f] s_0[c] s_1[d]
TNQ_HD float grad_chunk(const Args& a, const float* gparts, long long t0, long long t1, int layer, int q, int v) {
    const int n = a.n;
    const long long stride = grad_floats(n);
    float t = 0.f;
    if (layer == 1) {
        const float* g = gparts + (long long)q * GQ + v;
        for (long long i = t0; i < t1; ++i) t += g[i * stride] + g[i * stride + K4];
        return t;
    }
    const int f = v % K, e = (v / K) % K, d = (v / K2) % K, cc = v / K3;
    if (q == 0) {
        const float* g = gparts + 162 + e * K + f;
        for (long long i = t0; i < t1; ++i) t += g[i * stride];
        return t * a.state[0][cc] * a.state[1][d];
    }
    const float* g = gparts + (long long)q * GQ + 162 + (cc * K + e) * K + f;
    for (long long i = t0; i < t1; ++i) t += g[i * stride] + g[i * stride + K3];
    return t * a.state[q + 1][d];
}

}  // namespace tnq_l2
