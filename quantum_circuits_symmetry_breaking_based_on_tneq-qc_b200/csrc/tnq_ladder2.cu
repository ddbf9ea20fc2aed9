// tnq_ladder2.cu -- the two-layer merged MPS network (QCTN.merge(mps_n, mps_n), BASELINE cfg3), edge rank 3,
// second generation: lanes are samples, the core tensors are uniform operands from constant memory, the
// environment never leaves the registers (tnq_ladder2_core.cuh has the mathematics and the mapping).
// Reference: tneq_qc/contractor/greedy_strategy.py:461-598 (the sweep), :690-990 (one qubit group),
// tneq_qc/core/engine_siamese.py:490-530 (loss), tneq_qc/backends/backend_pytorch.py:153-158 (gradients).
//
// Per call:  tnq_l2_prep_kernel      cores x circuit states -> constant-pool image (global), tile counter = 0
//            cudaMemcpyToSymbolAsync  image -> __constant__ (device to device, same stream)
//            tnq_ladder2_kernel       persistent CTAs of 4 warps fetch tiles of S samples from an atomic counter;
//                                     forward sweep, fused loss, reverse sweep; every TILE writes its own
//                                     gradient slice, so the result does not depend on which CTA ran it
//            tnq_l2_finalize_kernel   slices summed in tile order, circuit states folded back in, loss
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include <string>

#include "tneq_b200.h"
#include "tnq_ladder2_core.cuh"

extern int tnq_internal_fail(const std::string& msg);
extern int tnq_internal_cuda_fail(cudaError_t e, const char* what);
extern void tnq_internal_count_launch();

namespace {

using namespace tnq_l2;

__global__ void tnq_l2_prep_kernel(const __grid_constant__ Args a, float* __restrict__ image, int* __restrict__ counter) {
    const int nc = cst_floats(a.n);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < nc; i += gridDim.x * blockDim.x) image[i] = cst_element(a, i);
    if (blockIdx.x == 0 && threadIdx.x == 0) *counter = 0;
}

template <int R, int NW, int MODE>
__global__ void __maxnreg__(MODE == 0 ? 128 : 255)
tnq_ladder2_kernel(const __grid_constant__ Args a, long long B, long long ntiles, const float* __restrict__ seed,
                   float* __restrict__ values, float* __restrict__ gparts, float* __restrict__ lparts,
                   float* __restrict__ ckpt, int* __restrict__ counter, float log_scale, float inv_count) {
    using G = Geo<R, NW>;
    constexpr int NT = NW * 32;
    extern __shared__ __align__(16) float sm[];
    __shared__ long long tile_s;
    // (Measured dead end: reversing the warp order of every second CTA on an SM, so that the uneven 4,4,3,3 split of
    // the row blocks lands evenly on the four scheduler partitions, changed nothing: 1.438 vs 1.426 ms.)
    const int tid = threadIdx.x;
    constexpr int NSM = MODE == 0 ? G::FWD_FLOATS : G::TRAIN_FLOATS;
    for (int i = tid; i < NSM; i += NT) sm[i] = 0.f;
    Ctx c;
    c.sm = sm;
    c.cst = nullptr;
    c.a = &a;
    c.B = B;
    c.seed = seed;
    c.values = values;
    c.log_scale = log_scale;
    c.inv_count = inv_count;
    c.ck = MODE != 0 ? ckpt + (long long)blockIdx.x * G::ckpt_floats(a.n) : nullptr;
    __syncthreads();
    build_pos<R, NW>(c, tid);
    __syncthreads();
    TS ts;
    for (;;) {
        if (tid == 0) tile_s = atomicAdd(counter, 1);
        __syncthreads();
        const long long tile = tile_s;
        __syncthreads();
        if (tile >= ntiles) break;
        c.b0 = tile * G::S;
        c.gpart = MODE != 0 ? gparts + tile * grad_floats(a.n) : nullptr;
        c.lpart = MODE != 0 ? lparts + tile : nullptr;
        tile_sweep<R, NW, MODE>(c, ts, tid);
    }
}

// One block = 32 consecutive gradient elements x 8 chunks of tiles; the chunk sums are combined in a fixed
// order.  Elements: [0, (n-1)*81) = layer A, [(n-1)*81, 2(n-1)*81) = layer X; the last block sums the loss.
__global__ void __launch_bounds__(256)
tnq_l2_finalize_kernel(const __grid_constant__ Args a, const float* __restrict__ gparts, const float* __restrict__ lparts,
                       long long ntiles, float* __restrict__ loss) {
    __shared__ float part[8][33];
    const int n = a.n, ne = 2 * (n - 1) * K4;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int nblk = (ne + 31) / 32;
    const long long per = (ntiles + 7) / 8;
    const long long t0 = w * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
    if ((int)blockIdx.x == nblk) {
        float t = 0.f;
        for (long long i = t0 + lane; i < t1; i += 32) t += lparts[i];
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
    float t = 0.f;
    int layer = 0, q = 0, v = 0;
    if (e < ne) {
        layer = e / ((n - 1) * K4);
        q = (e % ((n - 1) * K4)) / K4;
        v = e % K4;
        if (t0 < t1) t = grad_chunk(a, gparts, t0, t1, layer, q, v);
    }
    part[w][lane] = t;
    __syncthreads();
    if (w != 0 || e >= ne) return;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += part[i][lane];
    (layer == 0 ? a.gradA[q] : a.gradX[q])[v] = s;
}

struct Plan {
    int R, NW, grid;
    long long ntiles;
    size_t smem;
};

int upw_of(int R, int NW) { return ((27 + R - 1) / R + NW - 1) / NW; }      // units the busiest warp owns per step
template <int R, int NW>
size_t smem_rw(int mode) { return sizeof(float) * (mode == 0 ? Geo<R, NW>::FWD_FLOATS : Geo<R, NW>::TRAIN_FLOATS); }
size_t smem_of(int R, int NW, int mode) {
    if (NW == 8) return R == 2 ? smem_rw<2, 8>(mode) : smem_rw<4, 8>(mode);
    switch (R) {
        case 1: return smem_rw<1, 4>(mode);
        case 2: return smem_rw<2, 4>(mode);
        case 4: return smem_rw<4, 4>(mode);
        default: return smem_rw<8, 4>(mode);
    }
}
long long ckpt_of(int R, int n) {
    switch (R) {
        case 1: return Geo<1, 4>::ckpt_floats(n);
        case 2: return Geo<2, 4>::ckpt_floats(n);
        case 4: return Geo<4, 4>::ckpt_floats(n);
        default: return Geo<8, 4>::ckpt_floats(n);
    }
}

// CTAs that fit on one SM: registers (16 K per scheduler partition; a CTA puts NW / 4 warps on each) and shared memory
int ctas_per_sm(int R, int NW, int mode, int smem_sm) {
    const int by_regs = (mode == 0 ? 4 : 2) * 4 / NW;
    const int by_smem = (int)((size_t)smem_sm / (smem_of(R, NW, mode) + 1024));
    const int c = by_smem < by_regs ? by_smem : by_regs;
    return c < 1 ? 1 : c;
}

// Geometry: tiles of 32 / R samples, NW warps per CTA: 4-warp CTAs, two per SM (training), the tile size that keeps the
// tail short.  8-warp CTAs (R = 2 or 4, half the row blocks per warp) exist for experiments (TNQ_LADDER_WARPS=8): measured
// on B200 at 2 048 / 4 096 samples they are SLOWER than 4-warp CTAs (0.33-0.47 ms vs 0.25 ms; 0.63 vs 0.41 ms of the
// first-generation kernel): the CTA barriers between the small phases cost more with twice the warps than the shorter
// row-block loop saves.
Plan make_plan(int n, long long B, int mode, int sms, int smem_sm) {
    (void)n;
    Plan best{};
    double best_cost = 0;
    const char* force = getenv("TNQ_LADDER_R");
    const char* force_w = getenv("TNQ_LADDER_WARPS");
    for (int NW = 4; NW <= 8; NW *= 2)
        for (int R = 1; R <= 8; R *= 2) {
            if (NW == 8 && R != 2 && R != 4) continue;
            if (force && atoi(force) != R) continue;
            if (force_w ? atoi(force_w) != NW : NW != 4) continue;     // 8-warp CTAs only on request (measured slower)
            if (smem_of(R, NW, mode) + 1024 > (size_t)smem_sm) continue;
            const int S = 32 / R;
            const long long ntiles = (B + S - 1) / S;
            const int per_sm = ctas_per_sm(R, NW, mode, smem_sm);
            const int slots = sms * per_sm;
            const double rounds = (double)ntiles / slots;
            const double eff = R == 1 ? 1.0 : (R == 8 ? 27.0 / 32 : 27.0 / 28);
            // time of a tile ~ units of its busiest warp (x CTAs sharing the SM's issue slots), + a fixed part per step
            const double tile = (upw_of(R, NW) / eff + 1.5) * per_sm;
            const double cost = (rounds < 1 ? 1 : rounds) * tile;
            if (best.R == 0 || cost < best_cost * 0.999) {
                best_cost = cost;
                best.R = R;
                best.NW = NW;
                best.ntiles = ntiles;
                best.grid = (int)(ntiles < slots ? ntiles : slots);
                best.smem = smem_of(R, NW, mode);
            }
        }
    return best;
}

void device_limits(int& sms, int& smem_sm) {
    int dev = 0;
    sms = 148, smem_sm = 227 * 1024;
    if (cudaGetDevice(&dev) == cudaSuccess) {
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    }
}

struct Workspace {
    float *gparts, *lparts, *image, *ckpt;
    int* counter;
    size_t bytes;
};
Workspace carve(void* base, const Plan& p, int n, int mode) {
    Workspace w{};
    size_t at = 0;
    auto take = [&](size_t nbytes) {
        void* ptr = base ? (char*)base + at : nullptr;
        at += (nbytes + 255) / 256 * 256;
        return ptr;
    };
    w.image = (float*)take(sizeof(float) * cst_floats(n));
    w.counter = (int*)take(256);
    if (mode != 0) {
        w.gparts = (float*)take(sizeof(float) * (size_t)p.ntiles * grad_floats(n));
        w.lparts = (float*)take(sizeof(float) * (size_t)p.ntiles);
        w.ckpt = (float*)take(sizeof(float) * (size_t)p.grid * ckpt_of(p.R, n));
    }
    w.bytes = at;
    return w;
}

template <int R, int NW, int MODE>
int launch_rm(const Args& a, const Plan& p, const Workspace& w, long long B, const float* seed, float* values,
              float log_scale, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(tnq_ladder2_kernel<R, NW, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaFuncSetAttribute(ladder2)");
    tnq_ladder2_kernel<R, NW, MODE><<<p.grid, NW * 32, p.smem, st>>>(a, B, p.ntiles, seed, values, w.gparts, w.lparts, w.ckpt, w.counter,
                                                             log_scale, 1.0f / (float)B);
    tnq_internal_count_launch();
    return 0;
}
template <int R, int NW>
int launch_r(const Args& a, const Plan& p, const Workspace& w, long long B, int mode, const float* seed, float* values,
             float log_scale, cudaStream_t st) {
    if (mode == 0) return launch_rm<R, NW, 0>(a, p, w, B, seed, values, log_scale, st);
    if (mode == 1) return launch_rm<R, NW, 1>(a, p, w, B, seed, values, log_scale, st);
    return launch_rm<R, NW, 2>(a, p, w, B, seed, values, log_scale, st);
}

}  // namespace

extern "C" {

int64_t tnq_mps_ladder2_workspace_bytes(int n, int64_t B, int mode) {
    if (n < 3 || n > MAXQ || B <= 0) return 256;
    int sms, smem_sm;
    device_limits(sms, smem_sm);
    const Plan p = make_plan(n, B, mode, sms, smem_sm);
    return (int64_t)carve(nullptr, p, n, mode).bytes + 256;
}

/* geometry the launch would use (tests / bench): out[6] = {R, samples per tile, tiles, grid, dynamic smem bytes, warps per CTA} */
int tnq_mps_ladder2_geometry(int n, int64_t B, int mode, int64_t* out) {
    int sms, smem_sm;
    device_limits(sms, smem_sm);
    const Plan p = make_plan(n, B, mode, sms, smem_sm);
    out[0] = p.R, out[1] = 32 / p.R, out[2] = p.ntiles, out[3] = p.grid, out[4] = (int64_t)p.smem, out[5] = p.NW;
    return 0;
}

int tnq_mps_ladder2(int n, const float* const* cores_a, const float* const* cores_x, const float* const* states,
                    const float* const* mx, const int64_t* mx_stride, int64_t B, int mode, const float* seed,
                    float* values, float* loss, float* const* grads_a, float* const* grads_x, double log_scale,
                    void* workspace, int64_t workspace_bytes, void* stream) {
    if (n < 3 || n > MAXQ) return tnq_internal_fail("tnq_mps_ladder2: between 3 and " + std::to_string(MAXQ) + " qubits");
    if (!cores_a || !cores_x || !states || !mx || !mx_stride || B <= 0 || mode < 0 || mode > 2 || !workspace)
        return tnq_internal_fail("tnq_mps_ladder2: bad arguments");
    if (mode != 0 && (!grads_a || !grads_x)) return tnq_internal_fail("tnq_mps_ladder2: gradients need grads[]");
    if (mode == 0 && !values) return tnq_internal_fail("tnq_mps_ladder2: mode 0 needs values");
    if (mode == 1 && !loss) return tnq_internal_fail("tnq_mps_ladder2: mode 1 needs loss");
    if (mode == 2 && !seed) return tnq_internal_fail("tnq_mps_ladder2: mode 2 needs a seed");
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
        if (hq && (!states[q] || !mx[q])) return tnq_internal_fail("tnq_mps_ladder2: null pointer at qubit " + std::to_string(q));
        if (hc && (!cores_a[q] || !cores_x[q] || (mode != 0 && (!grads_a[q] || !grads_x[q]))))
            return tnq_internal_fail("tnq_mps_ladder2: null core pointer at qubit " + std::to_string(q));
    }
    int sms, smem_sm;
    device_limits(sms, smem_sm);
    const Plan p = make_plan(n, B, mode, sms, smem_sm);
    if (p.R == 0) return tnq_internal_fail("tnq_mps_ladder2: no geometry fits the shared memory of this device");
    const Workspace w = carve(workspace, p, n, mode);
    if ((size_t)workspace_bytes < w.bytes) return tnq_internal_fail("tnq_mps_ladder2: workspace too small");
    cudaStream_t st = (cudaStream_t)stream;
    const int nc = cst_floats(n);
    tnq_l2_prep_kernel<<<(nc + 255) / 256, 256, 0, st>>>(a, w.image, w.counter);
    tnq_internal_count_launch();
    cudaError_t e = cudaMemcpyToSymbolAsync(tnq_l2_cst_dev, w.image, sizeof(float) * nc, 0, cudaMemcpyDeviceToDevice, st);
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "cudaMemcpyToSymbolAsync(ladder2 constants)");
    int rc;
    switch (p.R * 16 + p.NW) {
        case 1 * 16 + 4: rc = launch_r<1, 4>(a, p, w, B, mode, seed, values, (float)log_scale, st); break;
        case 2 * 16 + 4: rc = launch_r<2, 4>(a, p, w, B, mode, seed, values, (float)log_scale, st); break;
        case 4 * 16 + 4: rc = launch_r<4, 4>(a, p, w, B, mode, seed, values, (float)log_scale, st); break;
        case 8 * 16 + 4: rc = launch_r<8, 4>(a, p, w, B, mode, seed, values, (float)log_scale, st); break;
        case 2 * 16 + 8: rc = launch_r<2, 8>(a, p, w, B, mode, seed, values, (float)log_scale, st); break;
        default: rc = launch_r<4, 8>(a, p, w, B, mode, seed, values, (float)log_scale, st); break;
    }
    if (rc) return rc;
    if (mode != 0) {
        const int nblk = (2 * (n - 1) * K4 + 31) / 32;
        tnq_l2_finalize_kernel<<<nblk + (mode == 1 ? 1 : 0), 256, 0, st>>>(a, w.gparts, w.lparts, p.ntiles,
                                                                           mode == 1 ? loss : nullptr);
        tnq_internal_count_launch();
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) return tnq_internal_cuda_fail(e, "tnq_mps_ladder2 launch");
    return 0;
}

}  // extern "C"
