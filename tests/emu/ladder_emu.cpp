// CPU emulation of the ladder kernel (csrc/tnq_ladder_core.cuh): the SAME phase functions and the
// SAME sweep driver as the CUDA kernel, with the 32 lanes of a warp executed one after the other
// inside every phase.  Test infrastructure only (built by tests/test_ladder_emu.py with g++): it
// lets the kernel's index bookkeeping be checked against the oracle on a machine without a GPU.
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "tnq_ladder_core.cuh"

using namespace tnq_ladder;

template <int K, int MODE>
static void run(const Args& a, long long B, const float* seed, float* values, float* loss, double log_scale,
                int nwarps) {
    using D = Dims<K>;
    const int n = a.n;
    const int cst_n = (n - 1) * D::CSTEP + D::K2;
    std::vector<float> cst(cst_n);
    for (int i = 0; i < cst_n; ++i) cst[i] = const_pool_element<K>(a, i);
    const long long ngroups = (B + D::SPW - 1) / D::SPW;
    const int ng = D::grad_floats(n);
    std::vector<float> gparts((size_t)nwarps * ng, 0.f), lparts((size_t)nwarps * 32, 0.f);
    for (int w = 0; w < nwarps; ++w) {
        std::vector<float> smem(D::WARP_TRAIN, 0.f), ck(MODE ? D::ckpt_floats(n) : 1, 0.f);
        WarpCtx<K> c;
        c.cst = cst.data();
        c.E = smem.data();
        c.U = c.E + D::E_SZ;
        c.T2 = c.U + D::U_SZ;
        c.M = c.T2 + D::T2_SZ;
        c.V = c.M + D::M_SZ;
        c.D = c.V + D::V_SZ;
        c.dT2 = c.D + D::E_SZ;
        c.ckE = ck.data();
        c.ckT2 = ck.data() + (size_t)(n - 2) * D::E_SZ;
        c.gpart = gparts.data() + (size_t)w * ng;
        c.args = &a;
        c.B = B;
        c.seed = seed;
        c.values = values;
        c.log_scale = (float)log_scale;
        c.inv_count = 1.0f / (float)B;
        LaneState<K> lanes[32];
        memset(lanes, 0, sizeof(lanes));
        for (long long g = w; g < ngroups; g += nwarps) ladder_group<K, MODE>(c, lanes, g * D::SPW);
        for (int l = 0; l < 32; ++l) lparts[(size_t)w * 32 + l] = lanes[l].loss;
    }
    if (MODE == 0) return;
    for (int q = 0; q < n - 1; ++q)
        for (int v = 0; v < D::K4; ++v) {
            a.gradA[q][v] = grad_element<K>(a, gparts.data(), nwarps, 0, q, v);
            a.gradX[q][v] = grad_element<K>(a, gparts.data(), nwarps, 1, q, v);
        }
    if (MODE == 1 && loss != nullptr) {
        float t = 0.f;
        for (size_t i = 0; i < lparts.size(); ++i) t += lparts[i];
        *loss = t;
    }
}

extern "C" int ladder_emu(int K, int n, const float* const* coreA, const float* const* coreX,
                          const float* const* states, const float* const* mx, const long long* mx_stride, long long B,
                          int mode, const float* seed, float* values, float* loss, float* const* gradA,
                          float* const* gradX, double log_scale, int nwarps) {
    if (n < 3 || n > MAXQ || (K != 2 && K != 3) || mode < 0 || mode > 2) return 1;
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
#define RUN(KK, MM) run<KK, MM>(a, B, seed, values, loss, log_scale, nwarps)
    if (K == 3) {
        if (mode == 0) RUN(3, 0); else if (mode == 1) RUN(3, 1); else RUN(3, 2);
    } else {
        if (mode == 0) RUN(2, 0); else if (mode == 1) RUN(2, 1); else RUN(2, 2);
    }
    return 0;
}
