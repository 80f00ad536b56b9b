// sdes_rollout_simt.cu — persistent rollout kernel, control MLP on the fp32 FFMA pipe.
//
// One warp owns 32 trajectories for all T steps (one thread = one trajectory); warps pull
// 32-row tiles from a global counter so the 148 SMs stay evenly loaded whatever B is.  The
// weights, the GMM image and the prior/reference images sit in shared memory for the whole
// kernel; the state never leaves registers between steps.  This is the exact-fp32 engine
// (bit-for-bit fp32 FMA arithmetic like the reference's SGEMM up to summation order) and the
// GPU-side cross-check for the tcgen05 engine in sdes_rollout_mma.cu.
#include "sdes_step.cuh"

namespace sdes {

constexpr int SIMT_WARPS = 8;

// NN(s, x) = FourierMLP.forward (models/mlp.py:114-122) for this thread's row.
// act: this warp's [64][32] scratch (column = lane) — only ever read by the thread that wrote it.
template <int DPAD>
__device__ __forceinline__ void mlp_simt(const float (&x)[DPAD], float (&out)[DPAD], const float* __restrict__ wsm,
                                         const float* __restrict__ emb_row, float* act, int dim, int nh) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < DPAD; ++j) act[j * 32 + lane] = x[j];
    float acc[C];
    const float* w = wsm;
    // input layer + time embedding: h = W_in x + b_in + emb_t   (mlp.py:116-118)
    {
        const float* b = w + dim * C;
#pragma unroll
        for (int n = 0; n < C; ++n) acc[n] = b[n] + __ldg(emb_row + n);
        for (int k = 0; k < dim; ++k) {
            const float a = act[k * 32 + lane];
            const float4* w4 = reinterpret_cast<const float4*>(w + k * C);
#pragma unroll
            for (int q = 0; q < C / 4; ++q) {
                const float4 ww = w4[q];
                acc[4 * q + 0] = fmaf(ww.x, a, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(ww.y, a, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(ww.z, a, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(ww.w, a, acc[4 * q + 3]);
            }
        }
        w = b + C;
    }
    // hidden layers: h = W gelu(h) + b   (mlp.py:119-121)
    for (int l = 0; l < nh; ++l) {
#pragma unroll
        for (int n = 0; n < C; ++n) act[n * 32 + lane] = gelu_erf(acc[n]);
        const float* b = w + C * C;
#pragma unroll
        for (int n = 0; n < C; ++n) acc[n] = b[n];
#pragma unroll 2
        for (int k = 0; k < C; ++k) {
            const float a = act[k * 32 + lane];
            const float4* w4 = reinterpret_cast<const float4*>(w + k * C);
#pragma unroll
            for (int q = 0; q < C / 4; ++q) {
                const float4 ww = w4[q];
                acc[4 * q + 0] = fmaf(ww.x, a, acc[4 * q + 0]);
                acc[4 * q + 1] = fmaf(ww.y, a, acc[4 * q + 1]);
                acc[4 * q + 2] = fmaf(ww.z, a, acc[4 * q + 2]);
                acc[4 * q + 3] = fmaf(ww.w, a, acc[4 * q + 3]);
            }
        }
        w = b + C;
    }
    // output layer (mlp.py:122)
#pragma unroll
    for (int n = 0; n < C; ++n) act[n * 32 + lane] = gelu_erf(acc[n]);
    {
        const float* b = w + C * DPAD;
#pragma unroll
        for (int j = 0; j < DPAD; ++j) out[j] = b[j];
#pragma unroll 2
        for (int k = 0; k < C; ++k) {
            const float a = act[k * 32 + lane];
            const float4* w4 = reinterpret_cast<const float4*>(w + k * DPAD);
#pragma unroll
            for (int q = 0; q < DPAD / 4; ++q) {
                const float4 ww = w4[q];
                out[4 * q + 0] = fmaf(ww.x, a, out[4 * q + 0]);
                out[4 * q + 1] = fmaf(ww.y, a, out[4 * q + 1]);
                out[4 * q + 2] = fmaf(ww.z, a, out[4 * q + 2]);
                out[4 * q + 3] = fmaf(ww.w, a, out[4 * q + 3]);
            }
        }
    }
}

template <int DPAD>
__global__ void __launch_bounds__(SIMT_WARPS * 32, 1) rollout_simt_kernel(const __grid_constant__ KParams p) {
    extern __shared__ __align__(16) float smem[];
    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, T = d.n_steps, K = d.n_components;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    // ---- shared memory carve-up
    float* s_w = smem;
    const int K2 = (K + 1) & ~1;  // GMM images are padded to an even number of components
    float* s_mu = s_w + ((p.ws.w_simt_len + 3) & ~3ll);
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    float* s_prior = s_c + 64;
    float* s_ref = s_prior + 2 * DPAD + 4;
    float* s_act = s_ref + 2 * DPAD + 4;
    for (int64_t e = tid; e < p.ws.w_simt_len; e += blockDim.x) s_w[e] = ws[p.ws.w_simt + e];
    for (int e = tid; e < K2 * DPAD; e += blockDim.x) {
        s_mu[e] = ws[p.ws.gmm_mu + e];
        s_h[e] = ws[p.ws.gmm_h + e];
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c[e] = ws[p.ws.gmm_c + e];
    for (int e = tid; e < 2 * DPAD + 4; e += blockDim.x) {
        s_prior[e] = e <= 2 * DPAD ? ws[p.ws.prior + e] : 0.f;
        s_ref[e] = e <= 2 * DPAD ? ws[p.ws.ref + e] : 0.f;
    }
    __syncthreads();
    TargetSmem tsm{s_mu, s_h, s_c, reinterpret_cast<const uint32_t*>(ws + p.ws.counter)[1], s_prior, s_ref};
    float* act = s_act + warp * (C * 32);
    uint32_t* counter = reinterpret_cast<uint32_t*>(const_cast<float*>(ws) + p.ws.counter);
    const bool from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    const bool ret_traj = (d.flags & SDES_F_RETURN_TRAJ) != 0;
    const int64_t B = d.batch;

    for (;;) {
        uint32_t tile = 0;
        if (lane == 0) tile = atomicAdd(counter, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= (uint32_t)p.n_tiles) break;
        const int64_t row = (int64_t)tile * 32 + lane;
        const bool valid = row < B;
        const int64_t rrow = valid ? row : (B - 1);  // inactive lanes shadow the last row, never write

        float x[DPAD];
#pragma unroll
        for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(d.x0 + rrow * dim + j) : 0.f;
        if (ret_traj && valid) {
            const TrajRef o = traj_ref(d, d.xs, 0, rrow);
#pragma unroll
            for (int j = 0; j < DPAD; ++j)
                if (j < dim) o.p[j * o.stride] = x[j];
        }
        float rnd = initial_rnd<DPAD>(d, x, tsm);
        const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)rrow);

        for (int i = 0; i < T; ++i) {
            const float* tab = ws + p.ws.tab + (int64_t)i * TAB_STRIDE;
            float sc[DPAD], g[DPAD];
            score_part<DPAD>(d, x, sc, tsm, ws + p.ws.gate + (int64_t)i * DPAD, tab[TAB_SIGMA], tab[TAB_LERP_W]);
            mlp_simt<DPAD>(x, g, s_w, ws + p.ws.emb + (int64_t)i * C, act, dim, d.n_hidden);
            const StepCoef sc_ = make_step_coef(d, tab);
            const float* nrow = from_hbm ? d.noise + ((int64_t)i * B + rrow) * dim : nullptr;
            float cost = 0.f, ito = 0.f;
#pragma unroll
            for (int q = 0; q < DPAD / 4; ++q)
                update4(sc_, &x[4 * q], &g[4 * q], &sc[4 * q], s_prior + 4 * q, s_prior + DPAD + 4 * q, 4 * q, i, traj, nrow, cost, ito);
            finish_step(d, sc_, tab, cost, ito, rnd);
            if (ret_traj && valid) {
                const TrajRef o = traj_ref(d, d.xs, i + 1, rrow);
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) o.p[j * o.stride] = x[j];
            }
        }
        rnd += terminal_rnd<DPAD>(d, x, tsm);
        if (valid) {
#pragma unroll
            for (int j = 0; j < DPAD; ++j)
                if (j < dim) d.x_T[rrow * dim + j] = x[j];
            d.rnd[rrow] = rnd;
        }
    }
}

size_t simt_smem_bytes(const KParams& p) {
    const int dpad = p.ws.dpad, K = p.d.n_components;
    size_t fl = ((p.ws.w_simt_len + 3) & ~3ll) + 2 * (size_t)((K + 1) & ~1) * dpad + 64 + 2 * (2 * dpad + 4) + (size_t)SIMT_WARPS * C * 32;
    return fl * sizeof(float);
}

template <int DPAD>
static cudaError_t launch_simt_t(const KParams& p, int sm_count, cudaStream_t stream) {
    const size_t smem = simt_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(rollout_simt_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int warps_needed = p.n_tiles;
    int grid = (warps_needed + SIMT_WARPS - 1) / SIMT_WARPS;
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    rollout_simt_kernel<DPAD><<<grid, SIMT_WARPS * 32, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_rollout_simt(const KParams& p, int sm_count, cudaStream_t stream) {
    switch (p.ws.dpad) {
        case 4: return launch_simt_t<4>(p, sm_count, stream);
        case 8: return launch_simt_t<8>(p, sm_count, stream);
        case 12: return launch_simt_t<12>(p, sm_count, stream);
        case 16: return launch_simt_t<16>(p, sm_count, stream);
        case 32: return launch_simt_t<32>(p, sm_count, stream);
        case 52: return launch_simt_t<52>(p, sm_count, stream);
        case 64: return launch_simt_t<64>(p, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sdes
