// sdes_timeembed.cuh — TimeEmbed forward for one time value by one thread block (models/mlp.py:43-82);
// shared by the prologue kernels of the fused (sdes_prepare.cu) and the wide (sdes_wide.cu) engines.
#pragma once

#include "sdes_common.cuh"

namespace sdes {

__device__ __forceinline__ float linspace_coeff(int c) {
    // torch.linspace(0.1, 100, 64) fp32 (models/mlp.py:58): step = (end-start)/(steps-1);
    // first half start + step*i, second half end - step*(steps-1-i)  (ATen RangeFactories).
    const float start = 0.1f, end = 100.0f;
    const float step = __fdiv_rn(__fsub_rn(end, start), (float)(C - 1));
    return (c < C / 2) ? __fadd_rn(start, __fmul_rn(step, (float)c))
                       : __fsub_rn(end, __fmul_rn(step, (float)(C - 1 - c)));
}

// One TimeEmbed forward (models/mlp.py:71-82) for a single time value, by one block.
// buf_a/buf_b: 2C floats each.  Result (n_out values) left in buf_out[0..n_out).
static __device__ void time_embed_row(const float* __restrict__ blob, int64_t o_phase, const int64_t* o_hw,
                               const int64_t* o_hb, int n_hidden, int64_t o_ow, int64_t o_ob, int n_out,
                               float s, float* buf_a, float* buf_b, float* buf_out) {
    const int tid = threadIdx.x;
    if (tid < 2 * C) {
        const int c = tid & (C - 1);
        // (coeff * t) + phase, separately rounded as the reference's two tensor ops
        const float arg = __fadd_rn(__fmul_rn(linspace_coeff(c), s), blob[o_phase + c]);
        buf_a[tid] = (tid < C) ? sinf(arg) : cosf(arg);
    }
    __syncthreads();
    float* in = buf_a;
    float* out = buf_b;
    int k_in = 2 * C;
    for (int l = 0; l < n_hidden; ++l) {
        if (tid < C) {
            const float* w = blob + o_hw[l] + (int64_t)tid * k_in;
            float acc = blob[o_hb[l] + tid];
            for (int k = 0; k < k_in; ++k) acc = fmaf(w[k], in[k], acc);
            out[tid] = gelu_erf(acc);
        }
        __syncthreads();
        float* tmp = in; in = out; out = tmp;
        k_in = C;
    }
    if (tid < n_out) {
        const float* w = blob + o_ow + (int64_t)tid * C;
        float acc = blob[o_ob + tid];
        for (int k = 0; k < C; ++k) acc = fmaf(w[k], in[k], acc);
        buf_out[tid] = acc;
    }
    __syncthreads();
}

// One row of the per-step scalar table (TAB_* columns of sdes_common.cuh) for step i, s = ts[i], t = ts[i+1]:
// SDE coefficients (eq/sdes.py:88-99, :141-145, :222-245), dt, sqrt(dt), the exponential integrator's alpha_k / beta_k
// (losses/oc.py:429-430).  Shared by the prologues of the fused and the wide engines.
__device__ __forceinline__ void write_step_table_row(const SdesRolloutDesc& d, int i, float* __restrict__ row) {
    const float s = d.ts[i], t = d.ts[i + 1];
    const float dt = __fsub_rn(t, s);
    float mu = 0.f, sigma = 0.f, div_int = 0.f, lerp_w = 0.f;
    if (d.sde_kind == SDES_SDE_VP) {
        // VP._diff_coeff_sq_t eq/sdes.py:222-229: generative (sign>0) runs beta_max -> beta_min
        const float ws_ = __fdiv_rn(s, d.terminal_t), wt_ = __fdiv_rn(t, d.terminal_t);
        const float b0 = d.sde_sign > 0.f ? d.beta_max : d.beta_min;
        const float b1 = d.sde_sign > 0.f ? d.beta_min : d.beta_max;
        const float beta_s = torch_lerp(b0, b1, ws_), beta_t = torch_lerp(b0, b1, wt_);
        mu = d.sde_sign * 0.5f * beta_s;                                     // :231-232
        sigma = d.scale_diff * sqrtf(beta_s);                                // :234-235
        div_int = d.sde_sign * 0.25f * (beta_t + beta_s) * dt * (float)d.dim;  // :237-245, :88-91
        lerp_w = ws_;
    } else if (d.sde_kind == SDES_SDE_CONST_OU) {
        mu = d.sde_sign * d.drift_coeff;                                     // eq/sdes.py:141-145
        sigma = d.diff_coeff;
        div_int = d.sde_sign * d.drift_coeff * dt * (float)d.dim;
        lerp_w = __fdiv_rn(s, d.terminal_t);
    }
    float beta_k = 0.f, alpha_k = 0.f;
    if (d.loss_kind == SDES_LOSS_EXP_INTEGRATOR) {
        beta_k = fminf(fmaxf(d.alpha * sqrtf(dt), 0.f), 1.f);                // losses/oc.py:429
        alpha_k = sqrtf(1.0f - beta_k * beta_k);                             // :430
    }
    row[TAB_DT] = dt;
    row[TAB_SQRT_DT] = sqrtf(dt);
    row[TAB_MU] = mu;
    row[TAB_SIGMA] = sigma;
    row[TAB_DIV_INT] = div_int;
    row[TAB_LERP_W] = lerp_w;
    row[TAB_BETA_K] = beta_k;
    row[TAB_ALPHA_K] = alpha_k;
}

}  // namespace sdes
