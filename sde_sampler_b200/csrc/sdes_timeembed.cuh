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

}  // namespace sdes
