// sdes_adjoint.cu — the reverse sweep of `loss.backward()` for the kl / kl_ito losses.
//
// With loss.method = kl the state is driven by the control WITH its graph (`sde_ctrl = generative_ctrl`,
// losses/oc.py:180, :305, :421), so d loss / d theta needs backpropagation through time.  Written as a discrete
// adjoint over the stored trajectory xs (T+1, B, d), with a_s = d loss / d x_s, w = d loss / d rnd_b and
// g = generative_ctrl(s, x_s):
//   Euler-Maruyama (TimeReversalLoss :204-219, ReferenceSDELoss :316-331), gm = g - r(x), r = sigma * prior score:
//       q       = w (gm dt + [ito] eps sqrt(dt))                       cotangent of the running cost on gm
//       delta_s = q + a_{s+1} sigma dt                                   cotangent of the control
//       a_s     = a_{s+1} (1 + mu dt) + J_x g^T delta_s + [r] q sigma / scale_prior^2
//   exponential integrator (ExponentialIntegratorSDELoss :429-443):
//       delta_s = w (beta_k^2 sigma^2 g + [ito] sigma beta_k eps) + a_{s+1} beta_k^2 sigma^2
//       a_s     = a_{s+1} alpha_k + J_x g^T delta_s
//   terminal:  a_T = w (grad log p_ref(x_T) - 1[|log rho| <= clip_target] grad log rho(x_T))     (:225, :337, :449-450)
// J_x g^T delta = (MLP input-gradient of delta * 1[|NN| <= clip_model]) + the score part's own x-dependence
// (models/reparam.py:78-83, :131-162): prior score exactly, target score through its analytic Hessian — unless the
// target's score is an autograd score evaluated WITHOUT create_graph (GMM: distr/base.py:130-137 as called from
// reparam.py:60, :135), which the reference's autograd treats as a constant (SDES_GRAD_TARGET_SCORE_CONST).
//
// This is the sweep of the fp32-FFMA engine (SDES_F_MLP_SIMT) and the cross-check of the tensor-core sweep in sdes_grad.cu
// (adj_step_kernel + per-step dgrad GEMMs), which is what the tcgen05 engine runs.
// This kernel only produces delta (T, B, d) — the cotangent of the control at every (trajectory, step).  The
// parameter gradient is then the SAME batched pass as for the lv loss (sdes_grad.cu: replayed forward, dgrad and
// wgrad GEMMs on tcgen05 over all B*T rows) with delta in place of the closed-form lv cotangent.
//
// Layout of the sweep: one thread owns one trajectory, a warp 32 of them; weights (transposed image of the fused
// fp32 engine), target and prior images live in shared memory; per warp a [64][32] activation scratch and
// (n_hidden+1) x [64][32] GELU' values of the replayed forward.  fp32 FFMA throughout (exact-fp32 arithmetic; the
// adjoint is a sequential T-step recursion per trajectory, the batched contraction is the GEMM pass that follows).
#include "sdes_step.cuh"

namespace sdes {

struct AdjArgs {
    KParams kp;          // SIMT-flagged descriptor: fp32 tables / weight image / target images of the fused prologue
    const float* xs;     // stored trajectory (layout per kp.d.flags & SDES_F_TRAJ_TILED)
    const float* w;      // (B) d loss / d rnd
    float* delta;        // (T, B, d) out, same layout family as xs
    uint32_t gflags;     // SDES_GRAD_*
    int warps;           // warps per CTA
};

__device__ __forceinline__ float dot64(const float* __restrict__ wrow, const float (&v)[C]) {
    const float4* w4 = reinterpret_cast<const float4*>(wrow);
    float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
    for (int q = 0; q < C / 4; ++q) {
        const float4 ww = w4[q];
        s0 = fmaf(ww.x, v[4 * q + 0], s0);
        s1 = fmaf(ww.y, v[4 * q + 1], s1);
        s2 = fmaf(ww.z, v[4 * q + 2], s2);
        s3 = fmaf(ww.w, v[4 * q + 3], s3);
    }
    return (s0 + s1) + (s2 + s3);
}

// FourierMLP.forward (models/mlp.py:114-122) for this thread's row, keeping GELU'(h_l) of every layer in gp.
template <int DPAD>
__device__ __forceinline__ void mlp_fwd_keep(const float (&x)[DPAD], float (&out)[DPAD], const float* __restrict__ wsm,
                                             const float* __restrict__ emb_row, float* act, float* gp, int dim, int nh) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int j = 0; j < DPAD; ++j) act[j * 32 + lane] = x[j];
    float acc[C];
    const float* w = wsm;
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
    for (int l = 0; l < nh; ++l) {
#pragma unroll
        for (int n = 0; n < C; ++n) {
            float y, dy;
            gelu_and_grad(acc[n], y, dy);  // one erfc / exp evaluation for both (|error| < 6e-7, sdes_common.cuh)
            act[n * 32 + lane] = y;
            gp[(l * C + n) * 32 + lane] = dy;
        }
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
#pragma unroll
    for (int n = 0; n < C; ++n) {
        float y, dy;
        gelu_and_grad(acc[n], y, dy);
        act[n * 32 + lane] = y;
        gp[(nh * C + n) * 32 + lane] = dy;
    }
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

// a += J_x NN(s, x)^T dnn: the input-gradient of the same network.  The transposed weight image serves both
// directions: row k of Wt is contiguous over the layer's outputs, so each input cotangent is one dot product.
template <int DPAD>
__device__ __forceinline__ void mlp_bwd_input(const float (&dnn)[DPAD], float (&a)[DPAD], const float* __restrict__ wsm,
                                              float* act, const float* gp, int dim, int nh) {
    const int lane = threadIdx.x & 31;
    const float* w_in = wsm;
    const float* w_h = wsm + dim * C + C;
    const float* w_out = w_h + nh * (C * C + C);
#pragma unroll 2
    for (int k = 0; k < C; ++k) {
        const float4* w4 = reinterpret_cast<const float4*>(w_out + k * DPAD);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int q = 0; q < DPAD / 4; ++q) {
            const float4 ww = w4[q];
            s0 = fmaf(ww.x, dnn[4 * q + 0], s0);
            s1 = fmaf(ww.y, dnn[4 * q + 1], s1);
            s0 = fmaf(ww.z, dnn[4 * q + 2], s0);
            s1 = fmaf(ww.w, dnn[4 * q + 3], s1);
        }
        act[k * 32 + lane] = (s0 + s1) * gp[(nh * C + k) * 32 + lane];
    }
    float v[C];
    for (int l = nh - 1; l >= 0; --l) {
#pragma unroll
        for (int n = 0; n < C; ++n) v[n] = act[n * 32 + lane];
        const float* w = w_h + l * (C * C + C);
#pragma unroll 2
        for (int k = 0; k < C; ++k) act[k * 32 + lane] = dot64(w + k * C, v) * gp[(l * C + k) * 32 + lane];
    }
#pragma unroll
    for (int n = 0; n < C; ++n) v[n] = act[n * 32 + lane];
#pragma unroll
    for (int k = 0; k < DPAD; ++k)
        if (k < dim) a[k] += dot64(w_in + k * C, v);
}

template <int DPAD>
__global__ void __launch_bounds__(256, 1) kl_adjoint_kernel(const __grid_constant__ AdjArgs a_) {
    extern __shared__ __align__(16) float smem[];
    const KParams& p = a_.kp;
    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, T = d.n_steps, K = d.n_components, nh = d.n_hidden;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

    float* s_w = smem;
    const int K2 = (K + 1) & ~1;
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
    float* act = s_act + warp * ((nh + 2) * C * 32);
    float* gp = act + C * 32;

    const int64_t B = d.batch;
    const bool ito = (d.flags & SDES_F_COMPUTE_ITO) != 0;
    const int ck = d.ctrl_kind;
    const bool score_detached = (a_.gflags & SDES_GRAD_SCORE_DETACHED) != 0 || ck == SDES_CTRL_CLIPPED;
    const bool target_in_ctrl = ck == SDES_CTRL_SCORE || ck == SDES_CTRL_LERP || ck == SDES_CTRL_LERP_TARGET;
    const bool target_hvp = target_in_ctrl && !score_detached && !(a_.gflags & SDES_GRAD_TARGET_SCORE_CONST);
    const bool prior_in_ctrl = ck == SDES_CTRL_LERP || ck == SDES_CTRL_LERP_PRIOR;

    for (int tile = blockIdx.x * a_.warps + warp; tile < p.n_tiles; tile += gridDim.x * a_.warps) {
        const int64_t row = (int64_t)tile * 32 + lane;
        const bool valid = row < B;
        const int64_t rrow = valid ? row : (B - 1);
        const float wb = a_.w[rrow];
        const bool dead = !(wb != 0.f) || !valid;  // filtered trajectories carry no gradient (rnd[mask], losses/oc.py:88-90)
        const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)rrow);

        float x[DPAD], a[DPAD];
        // ---- terminal cotangent
        {
            const TrajRef xr = traj_ref(d, const_cast<float*>(a_.xs), T, rrow);
#pragma unroll
            for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(xr.p + j * xr.stride) : 0.f;
            float tsc[DPAD];
            const float lp = target_eval<DPAD, true>(d, x, tsc, tsm);
            const float keep = fabsf(lp) <= d.clip_target ? 1.0f : 0.f;  // d clip(lp) / d lp (solver/oc.py:48-54)
#pragma unroll
            for (int j = 0; j < DPAD; ++j) {
                float v = -keep * tsc[j];
                if (d.loss_kind != SDES_LOSS_TIME_REVERSAL) v += (s_ref[j] - x[j]) * s_ref[DPAD + j];
                a[j] = (dead || j >= dim) ? 0.f : wb * v;
            }
        }
        // ---- reverse sweep
        for (int i = T - 1; i >= 0; --i) {
            const float* tab = ws + p.ws.tab + (int64_t)i * TAB_STRIDE;
            const StepCoef c = make_step_coef(d, tab);
            const float* gate_row = ws + p.ws.gate + (int64_t)i * DPAD;
            const float lerp_w = tab[TAB_LERP_W];
            {
                const TrajRef xr = traj_ref(d, const_cast<float*>(a_.xs), i, rrow);
#pragma unroll
                for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(xr.p + j * xr.stride) : 0.f;
            }
            // score part of the control and the factor that maps d g_j to the cotangent of its clipped inner value
            float sc[DPAD], fac[DPAD];
            if (ck == SDES_CTRL_CLIPPED) {
#pragma unroll
                for (int j = 0; j < DPAD; ++j) sc[j] = fac[j] = 0.f;
            } else {
                if (target_in_ctrl) {
                    target_eval<DPAD, true>(d, x, sc, tsm);
                } else {
#pragma unroll
                    for (int j = 0; j < DPAD; ++j) sc[j] = 0.f;
                }
                const float outer = (ck == SDES_CTRL_SCORE ? 1.0f : c.sigma) * d.scale_score;
#pragma unroll
                for (int j = 0; j < DPAD; ++j) {
                    const float ps = (s_prior[j] - x[j]) * s_prior[DPAD + j];
                    float inner;
                    if (ck == SDES_CTRL_LERP) inner = torch_lerp(ps, sc[j], lerp_w);
                    else if (ck == SDES_CTRL_LERP_PRIOR) inner = (1.0f - lerp_w) * ps;
                    else if (ck == SDES_CTRL_LERP_TARGET) inner = lerp_w * sc[j];
                    else inner = sc[j];
                    const float og = outer * gate_row[j];
                    sc[j] = og * clipf(inner, d.clip_score);
                    fac[j] = fabsf(inner) <= d.clip_score ? og : 0.f;
                }
            }
            float nn[DPAD];
            mlp_fwd_keep<DPAD>(x, nn, s_w, ws + p.ws.emb + (int64_t)i * C, act, gp, dim, nh);

            const float* nrow = (ito && c.from_hbm) ? d.noise + ((int64_t)i * B + rrow) * dim : nullptr;
            const TrajRef dr = traj_ref(d, a_.delta, i, rrow);
            const float a_mul = c.exp_int ? c.alpha_k : fmaf(c.mu, c.dt, 1.0f);
#pragma unroll
            for (int q = 0; q < DPAD / 4; ++q) {
                float e[4] = {0.f, 0.f, 0.f, 0.f};
                if (ito && 4 * q < dim) {
                    if (c.from_hbm) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? nrow[4 * q + r] : 0.f;
                    } else {
                        const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)i, (uint32_t)q);
                        e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
                    }
                }
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int j = 4 * q + r;
                    const float g = clipf(nn[j], c.cm) + sc[j];
                    float dg, an;
                    if (c.exp_int) {
                        dg = wb * (c.bb_ss * g + c.s_bk * e[r]) + a[j] * c.bb_ss;
                        an = a[j] * a_mul;
                    } else {
                        const float iv = s_prior[DPAD + j];
                        const float gm = c.ref_ctrl ? g - c.sigma * ((s_prior[j] - x[j]) * iv) : g;
                        const float qj = wb * (gm * c.dt + e[r] * c.sqrt_dt);
                        dg = fmaf(a[j], c.sigma * c.dt, qj);
                        an = a[j] * a_mul;
                        if (c.ref_ctrl) an = fmaf(qj, c.sigma * iv, an);
                    }
                    if (dead || j >= dim) dg = 0.f;
                    if (valid && j < dim) dr.p[j * dr.stride] = dg;
                    a[j] = an;
                    const float m = fabsf(nn[j]) <= c.cm ? dg : 0.f;  // d clip(NN) / d NN
                    nn[j] = m;
                    fac[j] *= dg;  // cotangent of inner_j
                }
            }
            if (!score_detached) {
                if (prior_in_ctrl) {
                    const float wp = 1.0f - lerp_w;
#pragma unroll
                    for (int j = 0; j < DPAD; ++j) a[j] = fmaf(-wp * s_prior[DPAD + j], fac[j], a[j]);
                }
                if (target_hvp) {
                    if (ck != SDES_CTRL_SCORE) {
#pragma unroll
                        for (int j = 0; j < DPAD; ++j) fac[j] *= lerp_w;
                    }
                    target_hvp_add<DPAD>(d, x, fac, a, tsm);
                }
            }
            mlp_bwd_input<DPAD>(nn, a, s_w, act, gp, dim, nh);
            if (dead) {
#pragma unroll
                for (int j = 0; j < DPAD; ++j) a[j] = 0.f;
            }
        }
    }
}

static size_t adjoint_smem_floats(const KParams& p, int warps) {
    const int dpad = p.ws.dpad, K = p.d.n_components;
    return ((p.ws.w_simt_len + 3) & ~3ll) + 2 * (size_t)((K + 1) & ~1) * dpad + 64 + 2 * (2 * dpad + 4) +
           (size_t)warps * (p.d.n_hidden + 2) * C * 32;
}

template <int DPAD>
static cudaError_t launch_adj_t(const AdjArgs& a, int sm_count, cudaStream_t stream) {
    const size_t smem = adjoint_smem_floats(a.kp, a.warps) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(kl_adjoint_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = (a.kp.n_tiles + a.warps - 1) / a.warps;
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    kl_adjoint_kernel<DPAD><<<grid, a.warps * 32, smem, stream>>>(a);
    return cudaGetLastError();
}

// kp: SIMT-flagged descriptor whose workspace already holds the fused prologue's outputs (launch_prepare ran on `stream`)
cudaError_t launch_kl_adjoint(const KParams& kp, const float* xs, const float* w, float* delta, uint32_t gflags, int sm_count,
                              cudaStream_t stream) {
    AdjArgs a;
    a.kp = kp;
    a.kp.n_tiles = (int)((kp.d.batch + 31) / 32);
    a.xs = xs; a.w = w; a.delta = delta; a.gflags = gflags;
    // as many warps per CTA as the per-warp scratch leaves room for (one CTA per SM, 227 KB)
    int warps = 8;
    while (warps > 1 && adjoint_smem_floats(a.kp, warps) * sizeof(float) > 227 * 1024) --warps;
    if (adjoint_smem_floats(a.kp, warps) * sizeof(float) > 227 * 1024) return cudaErrorInvalidConfiguration;
    a.warps = warps;
    switch (kp.ws.dpad) {
        case 4: return launch_adj_t<4>(a, sm_count, stream);
        case 8: return launch_adj_t<8>(a, sm_count, stream);
        case 12: return launch_adj_t<12>(a, sm_count, stream);
        case 16: return launch_adj_t<16>(a, sm_count, stream);
        case 32: return launch_adj_t<32>(a, sm_count, stream);
        case 52: return launch_adj_t<52>(a, sm_count, stream);
        case 64: return launch_adj_t<64>(a, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sdes
