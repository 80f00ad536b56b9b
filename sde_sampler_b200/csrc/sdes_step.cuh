// sdes_step.cuh — the per-trajectory arithmetic of one time step, thread-per-trajectory.
//
// Every rollout kernel (fp32-FFMA control MLP or tcgen05 control MLP) funnels into the same
// code here once the network output NN(s, x) is in registers: analytic target score, the
// control reparametrisation, the running-cost increments, the noise draw and the
// Euler-Maruyama / exponential-integrator update.  One thread owns one trajectory; the state
// x[DPAD], the network output and the score live in registers for all T steps.
#pragma once

#include "sdes_common.cuh"

namespace sdes {

// Shared-memory views of the parameter images (written by prepare_kernel, copied once per CTA).
struct TargetSmem {
    const float* gmm_mu;  // K * DPAD
    const float* gmm_h;   // K * DPAD
    const float* gmm_c;   // 64
    uint32_t gmm_mask;    // bit r: dimension pair (2r, 2r+1) differs between components
    const float* prior;   // loc[DPAD] | inv_var[DPAD] | lognorm
    const float* ref;     // same
};

// -------------------------------------------------------------------------------- targets
// GMM log-density and score (MixtureSameFamily log_prob distr/gauss.py:119-140; the reference
// differentiates it with autograd, distr/base.py:130-137 — analytic form SURVEY App. A.4).
// Direct (x-mu)^2 form: the expanded x^2 - 2 x mu + mu^2 form cancels catastrophically for
// modes at |mu| ~ 40 (SURVEY §7 "GMM log-density cancellation").
// One pass over the components with an online (running-max) softmax: no logits are stored.
// Per component: phase A accumulates the logit, phase B folds its responsibility-weighted
// score term into running sums that are rescaled whenever the running max moves.
//
// `ts.gmm_mask` has bit r set when the dimension pair (2r, 2r+1) differs between components.
// Pairs whose (mu, scale) are identical in every component factor out of the mixture exactly:
//     log rho(x) = logsumexp_k [ c_k - sum_{j in active} h_kj (x_j - mu_kj)^2 ] - sum_{j shared} h_j (x_j - mu_j)^2
//     score_j    = 2 h_j (mu_j - x_j)                                            for shared j
// so they cost O(d) instead of O(K d) (e.g. the zero-padded dims of GMM-40 in d=50).
// NP = number of leading pairs handled as "active" (a compile-time prefix, chosen by the caller
// to cover the highest set bit of the mask, so the inner loops carry no per-pair tests).
// Components are processed two at a time (independent accumulators) for instruction-level
// parallelism; the images are padded to an even K with h = 0, c = -inf.
template <int DPAD, int NP, bool NEED_SCORE, bool TWO = (NP <= 8), class X>
__device__ __forceinline__ float gmm_eval_na(const X& x, float (&score)[DPAD], const TargetSmem& ts, int K) {
    constexpr int NPAIR = DPAD / 2;
    float m = -INFINITY, ssum = 0.f;
    if (NEED_SCORE) {
#pragma unroll
        for (int j = 0; j < 2 * NP; ++j) score[j] = 0.f;
    }
#pragma unroll 1
    for (int k = 0; k < K; k += (TWO ? 2 : 1)) {
        const float2* mu2a = reinterpret_cast<const float2*>(ts.gmm_mu + k * DPAD);
        const float2* h2a = reinterpret_cast<const float2*>(ts.gmm_h + k * DPAD);
        const float2* mu2b = mu2a + NPAIR;
        const float2* h2b = h2a + NPAIR;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            const float2 mu = mu2a[r], h = h2a[r];
            const float d0 = x[2 * r + 0] - mu.x, d1 = x[2 * r + 1] - mu.y;
            a0 = fmaf(d0 * d0, h.x, a0);
            a1 = fmaf(d1 * d1, h.y, a1);
            if (TWO) {
                const float2 nu = mu2b[r], g = h2b[r];
                const float f0 = x[2 * r + 0] - nu.x, f1 = x[2 * r + 1] - nu.y;
                b0 = fmaf(f0 * f0, g.x, b0);
                b1 = fmaf(f1 * f1, g.y, b1);
            }
        }
        const float la = ts.gmm_c[k] - (a0 + a1), lb = TWO ? ts.gmm_c[k + 1] - (b0 + b1) : -INFINITY;
        const float m_new = fmaxf(m, fmaxf(la, lb));
        const float rescale = __expf(m - m_new);  // 1 when the max did not move, 0 on the first pair
        const float ea = __expf(la - m_new), eb = TWO ? __expf(lb - m_new) : 0.f;
        m = m_new;
        ssum = fmaf(ssum, rescale, ea + eb);
        if (NEED_SCORE) {
            // the rescale must be applied whenever it is != 1; the accumulation can be skipped when both
            // weights underflow to exactly 0 for the whole warp (bit-identical result)
            if (__any_sync(0xffffffffu, rescale != 1.0f || ea > 0.f || eb > 0.f)) {
                const float ea2 = 2.0f * ea, eb2 = 2.0f * eb;  // 1/var = 2h
#pragma unroll
                for (int r = 0; r < NP; ++r) {
                    const float2 mu = mu2a[r], h = h2a[r];
                    score[2 * r + 0] = fmaf(ea2 * h.x, mu.x - x[2 * r + 0], score[2 * r + 0] * rescale);
                    score[2 * r + 1] = fmaf(ea2 * h.y, mu.y - x[2 * r + 1], score[2 * r + 1] * rescale);
                    if (TWO) {
                        const float2 nu = mu2b[r], g = h2b[r];
                        score[2 * r + 0] = fmaf(eb2 * g.x, nu.x - x[2 * r + 0], score[2 * r + 0]);
                        score[2 * r + 1] = fmaf(eb2 * g.y, nu.y - x[2 * r + 1], score[2 * r + 1]);
                    }
                }
            }
        }
    }
    if (NEED_SCORE) {
        const float inv = 1.0f / ssum;
#pragma unroll
        for (int j = 0; j < 2 * NP; ++j) score[j] *= inv;
    }
    // pairs shared by all components (read from component 0)
    float shared = 0.f;
    const float2* mu2 = reinterpret_cast<const float2*>(ts.gmm_mu);
    const float2* h2 = reinterpret_cast<const float2*>(ts.gmm_h);
#pragma unroll
    for (int r = NP; r < NPAIR; ++r) {
        const float2 mu = mu2[r], h = h2[r];
        const float d0 = mu.x - x[2 * r + 0], d1 = mu.y - x[2 * r + 1];
        shared = fmaf(d0 * d0, h.x, shared);
        shared = fmaf(d1 * d1, h.y, shared);
        if (NEED_SCORE) {
            score[2 * r + 0] = 2.0f * h.x * d0;
            score[2 * r + 1] = 2.0f * h.y * d1;
        }
    }
    return m + logf(ssum) - shared;
}

template <int DPAD, bool NEED_SCORE, class X>
__device__ __forceinline__ float gmm_eval(const X& x, float (&score)[DPAD], const TargetSmem& ts, int K) {
    constexpr int NPAIR = DPAD / 2;
    const uint32_t mask = ts.gmm_mask;  // warp-uniform
    if (NPAIR > 1 && mask < 2u) return gmm_eval_na<DPAD, 1, NEED_SCORE>(x, score, ts, K);
    if (NPAIR > 2 && mask < 4u) return gmm_eval_na<DPAD, (NPAIR > 2 ? 2 : NPAIR), NEED_SCORE>(x, score, ts, K);
    if (NPAIR > 4 && mask < 16u) return gmm_eval_na<DPAD, (NPAIR > 4 ? 4 : NPAIR), NEED_SCORE>(x, score, ts, K);
    if (NPAIR > 8 && mask < 256u) return gmm_eval_na<DPAD, (NPAIR > 8 ? 8 : NPAIR), NEED_SCORE>(x, score, ts, K);
    if (NPAIR > 16 && mask < 65536u) return gmm_eval_na<DPAD, (NPAIR > 16 ? 16 : NPAIR), NEED_SCORE>(x, score, ts, K);
    return gmm_eval_na<DPAD, NPAIR, NEED_SCORE>(x, score, ts, K);
}

// MultiWell (distr/double_well.py:165-179; DoubleWell :39-45 is n_dw = d = 1).
template <int DPAD, bool NEED_SCORE, class X>
__device__ __forceinline__ float multiwell_eval(const X& x, float (&score)[DPAD], int dim, int n_dw,
                                                float sep, float shift) {
    float lp = 0.f;
#pragma unroll
    for (int j = 0; j < DPAD; ++j) {
        const float y = x[j] - shift;
        float s = 0.f;
        if (j < n_dw) {
            const float a = y * y - sep;
            lp -= a * a;
            s = -4.0f * a * y;
        } else if (j < dim) {
            lp -= 0.5f * y * y;
            s = -y;
        }
        if (NEED_SCORE) score[j] = s;
    }
    return lp;
}

// Funnel (distr/funnel.py:57-80): x_0 ~ N(0, var), x_{1:} | x_0 ~ N(0, exp(x_0) I).
template <int DPAD, bool NEED_SCORE, class X>
__device__ __forceinline__ float funnel_eval(const X& x, float (&score)[DPAD], int dim, float var) {
    float sq = 0.f;
#pragma unroll
    for (int j = 1; j < DPAD; ++j) sq = fmaf(x[j], x[j], sq);  // padded dims are 0
    const float x0 = x[0];
    const float inv = expf(-x0);
    const float dm1 = (float)(dim - 1);
    const float lp_first = -0.5f * logf(2.0f * 3.14159265358979323846f * var) - 0.5f * x0 * x0 / var;
    const float lp_other = -dm1 * (x0 + LOG_2PI) * 0.5f - 0.5f * sq * inv;
    if (NEED_SCORE) {
        score[0] = -x0 / var - 0.5f * dm1 + 0.5f * sq * inv;
#pragma unroll
        for (int j = 1; j < DPAD; ++j) score[j] = -x[j] * inv;
    }
    return lp_first + lp_other;
}

template <int DPAD, bool NEED_SCORE, class X>
__device__ __forceinline__ float target_eval(const SdesRolloutDesc& d, const X& x, float (&score)[DPAD],
                                             const TargetSmem& ts) {
    float lp;
    if (d.target_kind == SDES_TARGET_GMM)
        lp = gmm_eval<DPAD, NEED_SCORE>(x, score, ts, d.n_components);
    else if (d.target_kind == SDES_TARGET_MULTIWELL)
        lp = multiwell_eval<DPAD, NEED_SCORE>(x, score, d.dim, d.n_double_wells, d.separation, d.shift);
    else
        lp = funnel_eval<DPAD, NEED_SCORE>(x, score, d.dim, d.variance);
    return lp + d.log_norm_const;
}

// a += H v with H the Hessian of the target log-density at x (the Jacobian of its analytic score).
template <int DPAD>
__device__ __forceinline__ void target_hvp_add(const SdesRolloutDesc& d, const float (&x)[DPAD], const float (&v)[DPAD],
                                               float (&a)[DPAD], const TargetSmem& ts) {
    if (d.target_kind == SDES_TARGET_GMM) {
        // one component (Gauss / IsotropicGauss, distr/gauss.py:182-183, :222-223): score = (loc - x) / scale^2
#pragma unroll
        for (int j = 0; j < DPAD; ++j) a[j] = fmaf(-2.0f * ts.gmm_h[j], v[j], a[j]);
    } else if (d.target_kind == SDES_TARGET_MULTIWELL) {
        // distr/double_well.py:43-45, :174-179: -4 (y^2 - sep) y  ->  4 sep - 12 y^2;  Gaussian part: -1
#pragma unroll
        for (int j = 0; j < DPAD; ++j) {
            const float y = x[j] - d.shift;
            if (j < d.n_double_wells) a[j] = fmaf(4.0f * d.separation - 12.0f * y * y, v[j], a[j]);
            else if (j < d.dim) a[j] -= v[j];
        }
    } else {
        // distr/funnel.py:71-80
        float sq = 0.f, xv = 0.f;
#pragma unroll
        for (int j = 1; j < DPAD; ++j) {
            sq = fmaf(x[j], x[j], sq);
            xv = fmaf(x[j], v[j], xv);
        }
        const float inv = expf(-x[0]);
        a[0] += v[0] * (-1.0f / d.variance - 0.5f * sq * inv) + inv * xv;
#pragma unroll
        for (int j = 1; j < DPAD; ++j) a[j] += inv * (x[j] * v[0] - v[j]);
    }
}

// log N(x; loc, diag scale^2) from the image loc | inv_var | lognorm
template <int DPAD, class X>
__device__ __forceinline__ float diag_gauss_logp(const X& x, const float* img) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < DPAD; ++j) {
        const float y = x[j] - img[j];
        a = fmaf(y * y, img[DPAD + j], a);
    }
    return img[2 * DPAD] - 0.5f * a;
}

// ------------------------------------------------------------------------------- control
// generative_ctrl(s, x) = clip(NN(s, x), clip_model) + score_part(s, x)   (models/reparam.py):
//   ClippedCtrl :35-36     score_part = 0
//   ScoreCtrl   :78-83     scale_score * clip(grad log rho(x), clip_score) * gate(s)
//   LerpCtrl    :131-162   sigma(s) * scale_score * clip(lerp(grad log p_prior, grad log rho, s/T), clip_score) * gate(s)
//   LerpPriorCtrl :165-181 / LerpTargetCtrl :184-200: only the prior / target end of the lerp.
// score_part depends on x but not on the network, so it is evaluated BEFORE the MLP and kept in
// registers (sc[]); the network output is then streamed out of TMEM 8 columns at a time straight
// into the state update, and the full control vector never has to be materialised.
template <int DPAD, class X>
__device__ __forceinline__ void score_part(const SdesRolloutDesc& d, const X& x, float (&sc)[DPAD],
                                           const TargetSmem& ts, const float* __restrict__ gate_row, float sigma,
                                           float lerp_w) {
    if (d.ctrl_kind == SDES_CTRL_CLIPPED) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = 0.f;
        return;
    }
    if (d.ctrl_kind != SDES_CTRL_LERP_PRIOR) {
        target_eval<DPAD, true>(d, x, sc, ts);
    } else {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = 0.f;
    }
    const float* pl = ts.prior;
    const float cs = d.clip_score;
    const float outer = (d.ctrl_kind == SDES_CTRL_SCORE ? 1.0f : sigma) * d.scale_score;
    // one fused pass per control kind; the empty asm every 8 elements is a scheduling fence that keeps the
    // compiler from hoisting all DPAD operand loads at once (register pressure: 168 registers per thread)
#define SDES_SCORE_LOOP(EXPR)                                                        \
    _Pragma("unroll") for (int j = 0; j < DPAD; ++j) {                               \
        const float inner = (EXPR);                                                  \
        sc[j] = outer * (clipf(inner, cs) * gate_row[j]);                            \
        if ((j & 7) == 7) asm volatile("" ::: "memory");                             \
    }
    if (d.ctrl_kind == SDES_CTRL_LERP) {
        SDES_SCORE_LOOP(torch_lerp((pl[j] - x[j]) * pl[DPAD + j], sc[j], lerp_w))
    } else if (d.ctrl_kind == SDES_CTRL_LERP_PRIOR) {
        SDES_SCORE_LOOP((1.0f - lerp_w) * ((pl[j] - x[j]) * pl[DPAD + j]))
    } else if (d.ctrl_kind == SDES_CTRL_LERP_TARGET) {
        SDES_SCORE_LOOP(lerp_w * sc[j])
    } else {
        SDES_SCORE_LOOP(sc[j])
    }
#undef SDES_SCORE_LOOP
}

// ---------------------------------------------------------------------------------- step
// Per-step scalars shared by all dimensions of a trajectory.
struct StepCoef {
    float dt, sqrt_dt, mu, sigma, beta_k, alpha_k, bb_ss, s_bk, sg, cm;
    bool exp_int, ref_ctrl, from_hbm;
    uint32_t k0, k1;  // Philox key
    int dim;
};

__device__ __forceinline__ StepCoef make_step_coef(const SdesRolloutDesc& d, const float* __restrict__ tab) {
    StepCoef c;
    c.dt = tab[TAB_DT]; c.sqrt_dt = tab[TAB_SQRT_DT]; c.mu = tab[TAB_MU]; c.sigma = tab[TAB_SIGMA];
    c.beta_k = tab[TAB_BETA_K]; c.alpha_k = tab[TAB_ALPHA_K];
    c.sg = d.sigma;
    c.bb_ss = (c.beta_k * c.beta_k) * (c.sg * c.sg);
    c.s_bk = c.sg * c.beta_k;
    c.cm = d.clip_model;
    c.exp_int = d.loss_kind == SDES_LOSS_EXP_INTEGRATOR;
    c.ref_ctrl = (d.flags & SDES_F_REFERENCE_CTRL) != 0;
    c.from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    c.k0 = (uint32_t)d.seed; c.k1 = (uint32_t)(d.seed >> 32);
    c.dim = d.dim;
    return c;
}

// Four dimensions j0..j0+3 of one trajectory: assemble the control from the raw network output
// nn (bias included) and the precomputed score part, draw the noise, accumulate the cost / Ito
// sums and advance the state.
//   TimeReversalLoss  losses/oc.py:204-219   ReferenceSDELoss :316-331   ExponentialIntegrator :429-443
__device__ __forceinline__ void update4(const StepCoef& c, float* __restrict__ x4, const float* __restrict__ nn4,
                                        const float* __restrict__ sc4, const float* __restrict__ prior_loc4,
                                        const float* __restrict__ prior_iv4, int j0, int step, uint32_t traj,
                                        const float* __restrict__ noise_row, float& cost, float& ito) {
    if (j0 >= c.dim) return;  // whole chunk is padding: control 0, state stays 0 (warp-uniform)
    float e[4];
    if (c.from_hbm) {
#pragma unroll
        for (int r = 0; r < 4; ++r) e[r] = (j0 + r < c.dim) ? noise_row[j0 + r] : 0.f;
    } else {
        const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)step, (uint32_t)(j0 >> 2));
        e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
#pragma unroll
        for (int r = 0; r < 4; ++r) e[r] = (j0 + r < c.dim) ? e[r] : 0.f;
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        const float g = clipf(nn4[r], c.cm) + sc4[r];
        if (c.exp_int) {
            cost = fmaf(g, g, cost);
            ito = fmaf(c.sg * g * e[r], c.beta_k, ito);
            x4[r] = x4[r] * c.alpha_k + c.bb_ss * g + c.s_bk * e[r];
        } else {
            float gm = g;
            if (c.ref_ctrl) gm -= c.sigma * ((prior_loc4[r] - x4[r]) * prior_iv4[r]);  // solver/oc.py:305-306
            const float db = e[r] * c.sqrt_dt;
            cost = fmaf(gm, gm, cost);
            ito = fmaf(gm, db, ito);
            x4[r] = x4[r] + (c.mu * x4[r] + c.sigma * g) * c.dt + c.sigma * db;
        }
    }
}

__device__ __forceinline__ void finish_step(const SdesRolloutDesc& d, const StepCoef& c, const float* __restrict__ tab,
                                            float cost, float ito, float& rnd) {
    if (c.exp_int) {
        rnd += c.bb_ss * (0.5f * cost);
    } else {
        rnd += 0.5f * cost * c.dt;
        if (d.flags & SDES_F_SUB_DIV_INT) rnd -= tab[TAB_DIV_INT];
    }
    if (d.flags & SDES_F_COMPUTE_ITO) rnd += ito;
}

// initial cost (losses/oc.py:168-172, :296, :410)
template <int DPAD, class X>
__device__ __forceinline__ float initial_rnd(const SdesRolloutDesc& d, const X& x, const TargetSmem& ts) {
    if (d.loss_kind == SDES_LOSS_TIME_REVERSAL && !(d.flags & SDES_F_RND0_ZERO)) return diag_gauss_logp<DPAD>(x, ts.prior);
    return 0.f;
}

// terminal cost (losses/oc.py:225, :337, :449-450; clip: solver/oc.py:48-54)
template <int DPAD, class X>
__device__ __forceinline__ float terminal_rnd(const SdesRolloutDesc& d, const X& x, const TargetSmem& ts) {
    float dummy[DPAD];
    const float lp = clipf(target_eval<DPAD, false>(d, x, dummy, ts), d.clip_target);
    if (d.loss_kind == SDES_LOSS_TIME_REVERSAL) return -lp;
    return diag_gauss_logp<DPAD>(x, ts.ref) - lp;
}

}  // namespace sdes
