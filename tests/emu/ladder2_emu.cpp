// CPU emulation of the second-generation ladder kernel (csrc/tnq_ladder2_core.cuh): the SAME phase
// functions and the SAME tile driver as the CUDA kernel, with the 128 threads of a CTA executed one
// after the other inside every phase.  Test infrastructure only (built by tests/test_ladder2_emu.py
// with g++): the index bookkeeping, the row-block tables, the flush machinery and the reverse sweep are
// checked against the oracle on a machine without a GPU.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tnq_ladder2_core.cuh"

using namespace tnq_l2;

template <int R, int NW, int MODE>
static void run(const Args& a, long long B, const float* seed, float* values, float* loss, double log_scale) {
    using G = Geo<R, NW>;
    constexpr int NT = NW * 32;
    const int n = a.n;
    std::vector<float> cst(cst_floats(n));
    for (int i = 0; i < cst_floats(n); ++i) cst[i] = cst_element(a, i);
    const long long ntiles = (B + G::S - 1) / G::S;
    const int ng = grad_floats(n);
    std::vector<float> gparts((size_t)ntiles * ng, 0.f), lparts((size_t)ntiles, 0.f);
    std::vector<float> smem(G::TRAIN_FLOATS, 0.f), ck(MODE ? (size_t)G::ckpt_floats(n) : 1, 0.f);
    Ctx c;
    c.sm = smem.data();
    c.cst = cst.data();
    c.a = &a;
    c.B = B;
    c.seed = seed;
    c.values = values;
    c.log_scale = (float)log_scale;
    c.inv_count = 1.0f / (float)B;
    c.ck = ck.data();
    std::vector<TS> tss(NT);
    for (int tid = 0; tid < NT; ++tid) build_pos<R, NW>(c, tid);
    for (long long tile = 0; tile < ntiles; ++tile) {
        c.b0 = tile * G::S;
        c.gpart = gparts.data() + (size_t)tile * ng;
        c.lpart = lparts.data() + tile;
        tile_sweep<R, NW, MODE>(c, tss.data());
    }
    if (MODE == 0) return;
    for (int q = 0; q < n - 1; ++q)
        for (int v = 0; v < K4; ++v) {
            a.gradA[q][v] = grad_chunk(a, gparts.data(), 0, ntiles, 0, q, v);
            a.gradX[q][v] = grad_chunk(a, gparts.data(), 0, ntiles, 1, q, v);
        }
    if (MODE == 1 && loss != nullptr) {
        float t = 0.f;
        for (size_t i = 0; i < lparts.size(); ++i) t += lparts[i];
        *loss = t;
    }
}

template <int R, int NW>
static void run_r(const Args& a, long long B, int mode, const float* seed, float* values, float* loss, double ls) {
    if (mode == 0) run<R, NW, 0>(a, B, seed, values, loss, ls);
    else if (mode == 1) run<R, NW, 1>(a, B, seed, values, loss, ls);
    else run<R, NW, 2>(a, B, seed, values, loss, ls);
}

// R: slots per warp; warps: 4 or 8 warps per CTA
extern "C" int ladder2_emu(int R, int warps, int n, const float* const* coreA, const float* const* coreX, const float* const* states,
                           const float* const* mx, const long long* mx_stride, long long B, int mode, const float* seed,
                           float* values, float* loss, float* const* gradA, float* const* gradX, double log_scale) {
    if (n < 3 || n > MAXQ || mode < 0 || mode > 2) return 1;
    Args a;
    memset(&a, 0, sizeof(a));
    a.n = n;
    for (int q = 0; q < n; ++q) {
        a.state[q] = states[q];
        a.mx[q] = mx[q];
        a.mx_stride[q] = mx_stride[q];
    }
    for (int q = 0; q < n - 1; ++q) {
        a.coreA[q] = coreA[q];
        a.coreX[q] = coreX[q];
        a.gradA[q] = gradA ? gradA[q] : nullptr;
        a.gradX[q] = gradX ? gradX[q] : nullptr;
    }
    if (warps != 4 && warps != 8) return 3;
    switch (R * 16 + warps) {
        case 1 * 16 + 4: run_r<1, 4>(a, B, mode, seed, values, loss, log_scale); break;
        case 2 * 16 + 4: run_r<2, 4>(a, B, mode, seed, values, loss, log_scale); break;
        case 4 * 16 + 4: run_r<4, 4>(a, B, mode, seed, values, loss, log_scale); break;
        case 8 * 16 + 4: run_r<8, 4>(a, B, mode, seed, values, loss, log_scale); break;
        case 2 * 16 + 8: run_r<2, 8>(a, B, mode, seed, values, loss, log_scale); break;
        case 4 * 16 + 8: run_r<4, 8>(a, B, mode, seed, values, loss, log_scale); break;
        default: return 2;
    }
    return 0;
}

// row-block tables: rb_of / uslot_of round trip, every row block exactly once, and the bank rule (the positions
// of the row blocks of one unit are pairwise equal or distinct modulo R, for o, (q',r) and r)
template <int R>
static int check_tables() {
    constexpr int NW = 4;
    using G = Geo<R, NW>;
    int seen[27] = {0};
    for (int u = 0; u < G::NU; ++u) {
        int po[8], pq[8], pr[8], cnt = 0;
        for (int slot = 0; slot < R; ++slot) {
            const int rb = rb_of<R, NW>(u, slot);
            if (rb < 0) continue;
            if (rb >= 27) return 1;
            ++seen[rb];
            int uu, ss;
            uslot_of<R, NW>(rb, uu, ss);
            if (uu != u || ss != slot) return 2;
            po[cnt] = rb / 9, pq[cnt] = pos_q<R, NW>(rb % 9), pr[cnt] = rb % 3, ++cnt;
            if (pos_q<R, NW>(rb % 9) >= G::PQ || rb / 9 >= G::PO) return 3;
        }
        for (int i = 0; i < cnt; ++i)
            for (int j = i + 1; j < cnt; ++j) {
                if (po[i] != po[j] && po[i] % R == po[j] % R) return 4;
                if (pq[i] != pq[j] && pq[i] % R == pq[j] % R) return 5;
                if (pr[i] != pr[j] && pr[i] % R == pr[j] % R) return 6;
            }
    }
    for (int rb = 0; rb < 27; ++rb)
        if (seen[rb] != 1) return 7;
    // positions of (q',r) are a permutation of distinct values
    int used[16] = {0};
    for (int qr = 0; qr < 9; ++qr)
        if (used[pos_q<R, NW>(qr)]++) return 8;
    if (ustart<R, NW>(0) != 0 || ustart<R, NW>(NW) != G::NU) return 9;
    return 0;
}
extern "C" int ladder2_check_tables(int R) {
    switch (R) {
        case 1: return check_tables<1>();
        case 2: return check_tables<2>();
        case 4: return check_tables<4>();
        case 8: return check_tables<8>();
    }
    return -1;
}
