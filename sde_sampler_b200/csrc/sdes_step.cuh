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
    uint32_t gmm_mask;    // bit r: dimension chunk r (4 dims) differs between components
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
// `ts.gmm_mask` has bit r set when dimension chunk r (4 dims) differs between components.
// Chunks whose (mu, scale) are identical in every component factor out of the mixture exactly:
//     log rho(x) = logsumexp_k [ c_k - sum_{j in active} h_kj (x_j - mu_kj)^2 ] - sum_{j shared} h_j (x_j - mu_j)^2
//     score_j    = 2 h_j (mu_j - x_j)                                            for shared j
// so they cost O(d) instead of O(K d) (e.g. the zero-padded dims of GMM-40 in d=50).
template <int DPAD, bool NEED_SCORE>
__device__ __forceinline__ float gmm_eval(const float (&x)[DPAD], float (&score)[DPAD], const TargetSmem& ts, int K) {
    constexpr int NCH = DPAD / 4;
    const uint32_t mask = ts.gmm_mask;
    float m = -INFINITY, ssum = 0.f;
    if (NEED_SCORE) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) score[j] = 0.f;
    }
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
        const float4* mu4 = reinterpret_cast<const float4*>(ts.gmm_mu + k * DPAD);
        const float4* h4 = reinterpret_cast<const float4*>(ts.gmm_h + k * DPAD);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int r0 = 0; r0 < NCH; r0 += 4) {
            if ((mask >> r0) & 0xFu) {  // warp-uniform, two-level: 4 chunks, then each chunk
#pragma unroll
                for (int r = r0; r < r0 + 4 && r < NCH; ++r) {
                    if ((mask >> r) & 1u) {
                        const float4 mu = mu4[r], h = h4[r];
                        const float d0 = x[4 * r + 0] - mu.x, d1 = x[4 * r + 1] - mu.y;
                        const float d2 = x[4 * r + 2] - mu.z, d3 = x[4 * r + 3] - mu.w;
                        a0 = fmaf(d0 * d0, h.x, a0);
                        a1 = fmaf(d1 * d1, h.y, a1);
                        a2 = fmaf(d2 * d2, h.z, a2);
                        a3 = fmaf(d3 * d3, h.w, a3);
                    }
                }
            }
        }
        const float l = ts.gmm_c[k] - ((a0 + a1) + (a2 + a3));
        const float m_new = fmaxf(m, l);
        const float rescale = __expf(m - m_new);  // 1 when the max did not move, 0 on the first component
        const float e = __expf(l - m_new);
        m = m_new;
        ssum = fmaf(ssum, rescale, e);
        if (NEED_SCORE) {
            // the rescale must be applied whenever it is != 1; the accumulation can be skipped when this
            // component's weight underflows to exactly 0 for the whole warp (bit-identical result)
            const bool any_rescale = __any_sync(0xffffffffu, rescale != 1.0f);
            const bool any_weight = __any_sync(0xffffffffu, e > 0.f);
            if (any_rescale || any_weight) {
                const float e2 = 2.0f * e;  // 1/var = 2h
#pragma unroll
                for (int r0 = 0; r0 < NCH; r0 += 4) {
                    if ((mask >> r0) & 0xFu) {
#pragma unroll
                        for (int r = r0; r < r0 + 4 && r < NCH; ++r) {
                            if ((mask >> r) & 1u) {
                                const float4 mu = mu4[r], h = h4[r];
                                score[4 * r + 0] = fmaf(e2 * h.x, mu.x - x[4 * r + 0], score[4 * r + 0] * rescale);
                                score[4 * r + 1] = fmaf(e2 * h.y, mu.y - x[4 * r + 1], score[4 * r + 1] * rescale);
                                score[4 * r + 2] = fmaf(e2 * h.z, mu.z - x[4 * r + 2], score[4 * r + 2] * rescale);
                                score[4 * r + 3] = fmaf(e2 * h.w, mu.w - x[4 * r + 3], score[4 * r + 3] * rescale);
                            }
                        }
                    }
                }
            }
        }
    }
    // chunks shared by all components (read from component 0)
    const float inv = 1.0f / ssum;
    float shared = 0.f;
    const float4* mu4 = reinterpret_cast<const float4*>(ts.gmm_mu);
    const float4* h4 = reinterpret_cast<const float4*>(ts.gmm_h);
#pragma unroll
    for (int r = 0; r < NCH; ++r) {
        if ((mask >> r) & 1u) {
            if (NEED_SCORE) {
#pragma unroll
                for (int q = 0; q < 4; ++q) score[4 * r + q] *= inv;
            }
        } else {
            const float4 mu = mu4[r], h = h4[r];
            const float d0 = mu.x - x[4 * r + 0], d1 = mu.y - x[4 * r + 1];
            const float d2 = mu.z - x[4 * r + 2], d3 = mu.w - x[4 * r + 3];
            shared = fmaf(d0 * d0, h.x, shared);
            shared = fmaf(d1 * d1, h.y, shared);
            shared = fmaf(d2 * d2, h.z, shared);
            shared = fmaf(d3 * d3, h.w, shared);
            if (NEED_SCORE) {
                score[4 * r + 0] = 2.0f * h.x * d0;
                score[4 * r + 1] = 2.0f * h.y * d1;
                score[4 * r + 2] = 2.0f * h.z * d2;
                score[4 * r + 3] = 2.0f * h.w * d3;
            }
        }
    }
    return m + logf(ssum) - shared;
}

// MultiWell (distr/double_well.py:165-179; DoubleWell :39-45 is n_dw = d = 1).
template <int DPAD, bool NEED_SCORE>
__device__ __forceinline__ float multiwell_eval(const float (&x)[DPAD], float (&score)[DPAD], int dim, int n_dw,
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
template <int DPAD, bool NEED_SCORE>
__device__ __forceinline__ float funnel_eval(const float (&x)[DPAD], float (&score)[DPAD], int dim, float var) {
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

template <int DPAD, bool NEED_SCORE>
__device__ __forceinline__ float target_eval(const SdesRolloutDesc& d, const float (&x)[DPAD], float (&score)[DPAD],
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

// log N(x; loc, diag scale^2) from the image loc | inv_var | lognorm
template <int DPAD>
__device__ __forceinline__ float diag_gauss_logp(const float (&x)[DPAD], const float* img) {
    float a = 0.f;
#pragma unroll
    for (int j = 0; j < DPAD; ++j) {
        const float y = x[j] - img[j];
        a = fmaf(y * y, img[DPAD + j], a);
    }
    return img[2 * DPAD] - 0.5f * a;
}

// ------------------------------------------------------------------------------- control
// g = generative_ctrl(s, x) given nn = NN(s, x): ClippedCtrl reparam.py:35-36, ScoreCtrl
// :78-83, LerpCtrl :131-162, LerpPriorCtrl :165-181, LerpTargetCtrl :184-200.
// In: g[] holds the raw network output; out: g[] holds the control.
template <int DPAD>
__device__ __forceinline__ void control_assemble(const SdesRolloutDesc& d, const float (&x)[DPAD], float (&g)[DPAD],
                                                 const TargetSmem& ts, const float* __restrict__ gate_row,
                                                 float sigma, float lerp_w) {
    const float cm = d.clip_model, cs = d.clip_score;
    if (d.ctrl_kind == SDES_CTRL_CLIPPED) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) g[j] = clipf(g[j], cm);
        return;
    }
    float sc[DPAD];
    if (d.ctrl_kind != SDES_CTRL_LERP_PRIOR) {
        target_eval<DPAD, true>(d, x, sc, ts);
    } else {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = 0.f;
    }
    const float mult = d.ctrl_kind == SDES_CTRL_SCORE ? d.scale_score : d.scale_score;
    const float outer = d.ctrl_kind == SDES_CTRL_SCORE ? 1.0f : sigma;
    const float* pl = ts.prior;
    if (d.ctrl_kind == SDES_CTRL_LERP) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = torch_lerp((pl[j] - x[j]) * pl[DPAD + j], sc[j], lerp_w);
    } else if (d.ctrl_kind == SDES_CTRL_LERP_PRIOR) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = (1.0f - lerp_w) * ((pl[j] - x[j]) * pl[DPAD + j]);
    } else if (d.ctrl_kind == SDES_CTRL_LERP_TARGET) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = lerp_w * sc[j];
    }
#pragma unroll
    for (int j = 0; j < DPAD; ++j) {
        const float s = mult * clipf(sc[j], cs) * gate_row[j];
        g[j] = clipf(g[j], cm) + outer * s;
    }
}

// ---------------------------------------------------------------------------------- step
// Cost increments + noise + state update for one step, given the control g.
//   TimeReversalLoss  losses/oc.py:204-219   ReferenceSDELoss :316-331   ExponentialIntegrator :429-443
template <int DPAD>
__device__ __forceinline__ void step_update(const SdesRolloutDesc& d, float (&x)[DPAD], const float (&g)[DPAD],
                                            float& rnd, const TargetSmem& ts, const float* __restrict__ tab,
                                            int step, uint32_t traj, const float* __restrict__ noise_row) {
    const float dt = tab[TAB_DT], sqrt_dt = tab[TAB_SQRT_DT], mu = tab[TAB_MU], sigma = tab[TAB_SIGMA];
    const bool from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    const bool exp_int = d.loss_kind == SDES_LOSS_EXP_INTEGRATOR;
    const bool ref_ctrl = (d.flags & SDES_F_REFERENCE_CTRL) != 0;
    const float beta_k = tab[TAB_BETA_K], alpha_k = tab[TAB_ALPHA_K];
    const float sg = d.sigma;
    const float bb_ss = (beta_k * beta_k) * (sg * sg);
    const float s_bk = sg * beta_k;
    const float* pl = ts.prior;
    float cost = 0.f, ito = 0.f;
#pragma unroll
    for (int q = 0; q < DPAD / 4; ++q) {
        float e[4];
        if (from_hbm) {
#pragma unroll
            for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < d.dim) ? noise_row[4 * q + r] : 0.f;
        } else {
            const float4 n4 = normal4_call((uint32_t)d.seed, (uint32_t)(d.seed >> 32), traj, (uint32_t)step, (uint32_t)q);
            e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
#pragma unroll
            for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < d.dim) ? e[r] : 0.f;
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = 4 * q + r;
            if (exp_int) {
                cost = fmaf(g[j], g[j], cost);
                ito = fmaf(sg * g[j] * e[r], beta_k, ito);
                x[j] = x[j] * alpha_k + bb_ss * g[j] + s_bk * e[r];
            } else {
                float gm = g[j];
                if (ref_ctrl) gm -= sigma * ((pl[j] - x[j]) * pl[DPAD + j]);  // solver/oc.py:305-306
                const float db = e[r] * sqrt_dt;
                cost = fmaf(gm, gm, cost);
                ito = fmaf(gm, db, ito);
                x[j] = x[j] + (mu * x[j] + sigma * g[j]) * dt + sigma * db;
            }
        }
    }
    if (exp_int) {
        rnd += bb_ss * (0.5f * cost);
    } else {
        rnd += 0.5f * cost * dt;
        if (d.flags & SDES_F_SUB_DIV_INT) rnd -= tab[TAB_DIV_INT];
    }
    if (d.flags & SDES_F_COMPUTE_ITO) rnd += ito;
}

// initial cost (losses/oc.py:168-172, :296, :410)
template <int DPAD>
__device__ __forceinline__ float initial_rnd(const SdesRolloutDesc& d, const float (&x)[DPAD], const TargetSmem& ts) {
    if (d.loss_kind == SDES_LOSS_TIME_REVERSAL && !(d.flags & SDES_F_RND0_ZERO)) return diag_gauss_logp<DPAD>(x, ts.prior);
    return 0.f;
}

// terminal cost (losses/oc.py:225, :337, :449-450; clip: solver/oc.py:48-54)
template <int DPAD>
__device__ __forceinline__ float terminal_rnd(const SdesRolloutDesc& d, const float (&x)[DPAD], const TargetSmem& ts) {
    float dummy[DPAD];
    const float lp = clipf(target_eval<DPAD, false>(d, x, dummy, ts), d.clip_target);
    if (d.loss_kind == SDES_LOSS_TIME_REVERSAL) return -lp;
    return diag_gauss_logp<DPAD>(x, ts.ref) - lp;
}

}  // namespace sdes
