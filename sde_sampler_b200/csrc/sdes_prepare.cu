// sdes_prepare.cu — the prologue kernel: everything on the rollout path that does not
// depend on the state x is hoisted out of the T-step loop into per-step tables.
//
//   * FourierMLP.timestep_embed(s_i)  (models/mlp.py:115-116 evaluates it on B identical rows
//     every step)                                               -> emb  (T, C)
//   * the scalar gate score_model(s_i), clipped (models/reparam.py:68-76) -> gate (T, dpad)
//   * SDE coefficients mu(s_i), sigma(s_i), int div mu (eq/sdes.py:88-99,:222-245), dt, sqrt(dt),
//     exponential-integrator alpha_k, beta_k (losses/oc.py:429-430)  -> tab  (T, 8)
//   * kernel-ready images of the weights and of the target / prior parameters.
//
// Blocks [0, T) do one time step each; the remaining blocks re-lay-out parameters.
#include <cuda_bf16.h>

#include "sdes_common.cuh"
#include "sdes_timeembed.cuh"

namespace sdes {

__global__ void __launch_bounds__(256) prepare_kernel(const KParams p) {
    const SdesRolloutDesc& d = p.d;
    float* ws = reinterpret_cast<float*>(d.workspace);
    const float* blob = d.params;
    const int T = d.n_steps, dim = d.dim, dpad = p.ws.dpad;
    const int tid = threadIdx.x;

    if ((int)blockIdx.x < T) {
        __shared__ float buf_a[2 * C], buf_b[2 * C], buf_out[C];
        const int i = blockIdx.x;
        const float s = d.ts[i];
        if (tid == 0) write_step_table_row(d, i, ws + p.ws.tab + (int64_t)i * TAB_STRIDE);
        // FourierMLP.timestep_embed(s)
        time_embed_row(blob, p.bl.te_phase, p.bl.te_h_w, p.bl.te_h_b, d.te_hidden, p.bl.te_out_w,
                       p.bl.te_out_b, C, s, buf_a, buf_b, buf_out);
        // the tcgen05 engine adds the input-layer bias here, once per step, instead of once per row
        if (tid < C) ws[p.ws.emb + (int64_t)i * C + tid] = buf_out[tid] + ((d.flags & SDES_F_MLP_SIMT) ? 0.f : blob[p.bl.in_b + tid]);
        __syncthreads();
        // gate
        float* grow = ws + p.ws.gate + (int64_t)i * dpad;
        if (d.flags & SDES_F_HAS_GATE) {
            time_embed_row(blob, p.bl.g_phase, p.bl.g_h_w, p.bl.g_h_b, d.gate_hidden, p.bl.g_out_w,
                           p.bl.g_out_b, d.gate_dim, s, buf_a, buf_b, buf_out);
            if (tid < dpad)
                grow[tid] = tid < dim ? clipf(buf_out[d.gate_dim == 1 ? 0 : tid], d.clip_model) : 0.f;
        } else if (tid < dpad) {
            grow[tid] = tid < dim ? 1.0f : 0.f;
        }
        return;
    }

    // ------------------------------------------------------------------ parameter images
    const int64_t nthreads = (int64_t)(gridDim.x - T) * blockDim.x;
    const int64_t gtid = (int64_t)(blockIdx.x - T) * blockDim.x + tid;
    const int nh = d.n_hidden;

    if ((int)blockIdx.x == T) {
        // first re-layout block: work counter reset and the GMM pair mask — bit r set when dims 2r, 2r+1
        // differ between components (sdes_step.cuh gmm_eval); one thread per dimension scans the components.
        __shared__ uint32_t s_mask;
        if (tid == 0) s_mask = 0u;
        __syncthreads();
        if (d.target_kind == SDES_TARGET_GMM && tid < dim) {
            bool differs = false;
            for (int k = 1; k < d.n_components; ++k)
                differs |= d.gmm_loc[(int64_t)k * dim + tid] != d.gmm_loc[tid] || d.gmm_scale[(int64_t)k * dim + tid] != d.gmm_scale[tid];
            if (differs) atomicOr(&s_mask, 1u << (tid >> 1));
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t* ctr = reinterpret_cast<uint32_t*>(ws + p.ws.counter);
            ctr[0] = ctr[2] = ctr[3] = 0u;  // [0] = dynamic work counter
            ctr[1] = s_mask;
        }
    }

    // per-tile progress words of the time-chunked scheduler
    if (!(d.flags & SDES_F_MLP_SIMT)) {
        uint32_t* prog = reinterpret_cast<uint32_t*>(ws + p.ws.progress);
        const int64_t tiles128 = (d.batch + 127) / 128;
        for (int64_t e = gtid; e < tiles128; e += nthreads) prog[e] = 0u;
    }

    // bf16 hi/lo operand images of the 4-group tcgen05 engine (sdes_rollout_mma.cu): per layer hi[N x K16] then
    // lo[N x K16] in the wimg16 layout, then the fp32 biases {b_h[64]} x nh, b_out[NOUT]
    if (!(d.flags & SDES_F_MLP_SIMT)) {
        __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(ws + p.ws.w_mma4);
        const int nout = (dpad + 15) / 16 * 16, k0b = (dpad + 15) & ~15;
        auto put = [&](int64_t base, int64_t half, int n, int k, int N, float v) {
            const __nv_bfloat16 hi = __float2bfloat16_rn(v);
            const int64_t off = (int64_t)(k / 8) * (N * 8) + (int64_t)n * 8 + (k % 8);
            w16[base + off] = hi;
            w16[base + half + off] = __float2bfloat16_rn(v - __bfloat162float(hi));
        };
        int64_t o = 0;
        for (int64_t e = gtid; e < (int64_t)C * k0b; e += nthreads) {
            const int n = (int)(e / k0b), k = (int)(e % k0b);
            put(o, (int64_t)C * k0b, n, k, C, k < dim ? blob[p.bl.in_w + (int64_t)n * dim + k] : 0.f);
        }
        o += 2ll * C * k0b;
        for (int l = 0; l < nh; ++l) {
            for (int64_t e = gtid; e < C * C; e += nthreads) put(o, C * C, (int)(e / C), (int)(e % C), C, blob[p.bl.h_w[l] + e]);
            o += 2ll * C * C;
        }
        for (int64_t e = gtid; e < (int64_t)nout * C; e += nthreads) {
            const int n = (int)(e / C), k = (int)(e % C);
            put(o, (int64_t)nout * C, n, k, nout, n < dim ? blob[p.bl.out_w + (int64_t)n * C + k] : 0.f);
        }
        o += 2ll * nout * C;
        float* wb = ws + p.ws.w_mma4 + o / 2;
        for (int l = 0; l < nh; ++l)
            for (int64_t e = gtid; e < C; e += nthreads) wb[l * C + e] = blob[p.bl.h_b[l] + e];
        for (int64_t e = gtid; e < nout; e += nthreads) wb[nh * C + e] = e < dim ? blob[p.bl.out_b + e] : 0.f;
    }

    // SIMT weight image: WtIn[d][C] bIn[C] {Wt[C][C] b[C]} x nh  WtOut[C][dpad] bOut[dpad]
    if (d.flags & SDES_F_MLP_SIMT) {
        float* w = ws + p.ws.w_simt;
        int64_t o = 0;
        for (int64_t e = gtid; e < (int64_t)dim * C; e += nthreads) {
            const int k = (int)(e / C), n = (int)(e % C);
            w[o + e] = blob[p.bl.in_w + (int64_t)n * dim + k];
        }
        o += (int64_t)dim * C;
        for (int64_t e = gtid; e < C; e += nthreads) w[o + e] = blob[p.bl.in_b + e];
        o += C;
        for (int l = 0; l < nh; ++l) {
            for (int64_t e = gtid; e < C * C; e += nthreads) {
                const int k = (int)(e / C), n = (int)(e % C);
                w[o + e] = blob[p.bl.h_w[l] + (int64_t)n * C + k];
            }
            o += C * C;
            for (int64_t e = gtid; e < C; e += nthreads) w[o + e] = blob[p.bl.h_b[l] + e];
            o += C;
        }
        for (int64_t e = gtid; e < (int64_t)C * dpad; e += nthreads) {
            const int k = (int)(e / dpad), n = (int)(e % dpad);
            w[o + e] = n < dim ? blob[p.bl.out_w + (int64_t)n * C + k] : 0.f;
        }
        o += (int64_t)C * dpad;
        for (int64_t e = gtid; e < dpad; e += nthreads) w[o + e] = e < dim ? blob[p.bl.out_b + e] : 0.f;
    }

    // GMM image (distr/gauss.py:119-140): mu, h = 0.5/scale^2, c_k = log w_k - sum log scale - d/2 log 2pi
    if (d.target_kind == SDES_TARGET_GMM) {
        const int K = d.n_components, K2 = (K + 1) & ~1;
        for (int64_t e = gtid; e < (int64_t)K2 * dpad; e += nthreads) {
            const int k = (int)(e / dpad), j = (int)(e % dpad);
            float mu = 0.f, h = 0.f;
            if (j < dim && k < K) {
                mu = d.gmm_loc[(int64_t)k * dim + j];
                const float sc = d.gmm_scale[(int64_t)k * dim + j];
                h = 0.5f / (sc * sc);
            }
            ws[p.ws.gmm_mu + e] = mu;
            ws[p.ws.gmm_h + e] = h;
        }
        for (int64_t k = gtid; k < 64; k += nthreads) {
            float c = -INFINITY;
            if (k < K) {
                float logw = 0.f;
                if (d.gmm_weights != nullptr) {
                    float tot = 0.f;
                    for (int q = 0; q < K; ++q) tot += d.gmm_weights[q];
                    logw = logf(d.gmm_weights[k] / tot);  // Categorical(probs=w) normalises
                }
                float sl = 0.f;
                for (int j = 0; j < dim; ++j) sl += logf(d.gmm_scale[k * dim + j]);
                c = logw - sl - 0.5f * (float)dim * LOG_2PI;
            }
            ws[p.ws.gmm_c + k] = c;
        }
    }

    // diagonal Gaussians: loc | 1/scale^2 | log-normaliser   (distr/gauss.py:131-140, :215-223)
    for (int which = 0; which < 2; ++which) {
        const float* loc = which == 0 ? d.prior_loc : d.ref_loc;
        const float* scale = which == 0 ? d.prior_scale : d.ref_scale;
        float* out = ws + (which == 0 ? p.ws.prior : p.ws.ref);
        for (int64_t j = gtid; j < dpad; j += nthreads) {
            float m = 0.f, iv = 0.f;
            if (loc != nullptr && j < dim) {
                m = loc[j];
                const float sc = scale[j];
                iv = 1.0f / (sc * sc);
            }
            out[j] = m;
            out[dpad + j] = iv;
        }
        if (gtid == 0) {
            float ln = 0.f;
            if (loc != nullptr) {
                for (int j = 0; j < dim; ++j) ln -= logf(scale[j]);
                ln -= 0.5f * (float)dim * LOG_2PI;
            }
            out[2 * dpad] = ln;
        }
    }
}

void launch_prepare(const KParams& p, cudaStream_t stream) {
    const int relayout_blocks = 16;
    prepare_kernel<<<p.d.n_steps + relayout_blocks, 256, 0, stream>>>(p);
}

}  // namespace sdes
