// sdes_trainer.cu — the caller's side of the rollout: what `Trainable.step` / `TrainableDiff.compute_loss`
// / `get_metrics` do around the loss call, as stream-ordered kernels with no host synchronisation.
//
//   sdes_sample_gauss_prior  x0 ~ prior (IsotropicGauss.sample, distr/gauss.py:228-242: torch.randn, or
//                            nn.init.trunc_normal_ at the truncate_quartile bounds) from the Philox stream
//   sdes_trainer_step        the tail of Trainable.step (solver/base.py:409-439) on the flat parameter blob:
//                            loss / gradient checks (max_loss, max_grad or all-finite), clip_grad_norm_
//                            (conf/utils/grad_clip.yaml), torch.optim.Adam (conf/solver/oc_base.yaml:26-29) and the
//                            EMA update (EMA, solver/base.py:620-684) — two launches, the skip decision stays on
//                            the device
//   sdes_eval_moments        the sample statistics of get_metrics (eval/metrics.py:120-131): effective sample size
//                            sums and per-dimension first / second moments in one pass
#include <cmath>

#include "sdes_common.cuh"

namespace sdes {

// ------------------------------------------------------------------------------- prior sampling
// (0, 1) uniform from 32 random bits, never 0 or 1 in fp32
__device__ __forceinline__ float u01_open(uint32_t r) { return fmaf((float)(r >> 8), 5.9604644775390625e-8f, 2.98023223876953125e-8f); }

// nn.init.trunc_normal_ (torch/nn/init.py `_no_grad_trunc_normal_`): u ~ U(2l-1, 2u-1), erfinv, * std sqrt 2, + mean, clamp
__device__ __forceinline__ float trunc_normal_from_u(float u, float lo2, float hi2, float mean, float std, float a, float b) {
    const float v = fmaf(u, hi2 - lo2, lo2);
    float x = erfinvf(v) * (std * 1.4142135623730951f) + mean;
    return fminf(fmaxf(x, a), b);
}

struct PriorArgs {
    float* out;
    const float* uniforms;  // parity mode: (B, d) uniforms in [0, 1) instead of Philox
    int64_t batch;
    int dim;
    float mean, std, a, b, lo2, hi2;
    int truncated;
    uint64_t seed, traj_offset;
};

__global__ void __launch_bounds__(256) sample_prior_kernel(const PriorArgs p) {
    const int chunks = (p.dim + 3) / 4;
    const int64_t total = p.batch * chunks;
    const uint32_t k0 = (uint32_t)p.seed, k1 = (uint32_t)(p.seed >> 32);
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t b = e / chunks;
        const int q = (int)(e % chunks);
        float v[4];
        if (p.truncated) {
            float u[4];
            if (p.uniforms != nullptr) {
#pragma unroll
                for (int r = 0; r < 4; ++r) u[r] = (4 * q + r < p.dim) ? p.uniforms[b * p.dim + 4 * q + r] : 0.5f;
            } else {
                // stream id 0xFFFFFFFF in the step slot keeps these draws apart from the rollout's (step < T)
                const uint4 r4 = philox4x32_10((uint32_t)(p.traj_offset + (uint64_t)b), 0xFFFFFFFFu, (uint32_t)q, PHILOX_STREAM, k0, k1);
                u[0] = u01_open(r4.x); u[1] = u01_open(r4.y); u[2] = u01_open(r4.z); u[3] = u01_open(r4.w);
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = trunc_normal_from_u(u[r], p.lo2, p.hi2, p.mean, p.std, p.a, p.b);
        } else {
            const float4 n4 = normal4_call(k0, k1, (uint32_t)(p.traj_offset + (uint64_t)b), 0xFFFFFFFFu, (uint32_t)q);
            v[0] = fmaf(p.std, n4.x, p.mean); v[1] = fmaf(p.std, n4.y, p.mean);
            v[2] = fmaf(p.std, n4.z, p.mean); v[3] = fmaf(p.std, n4.w, p.mean);
        }
#pragma unroll
        for (int r = 0; r < 4; ++r)
            if (4 * q + r < p.dim) p.out[b * p.dim + 4 * q + r] = v[r];
    }
}

// --------------------------------------------------------------------------------- trainer step
// state (8 doubles, persistent): [0] optimizer steps taken  [1] skipped steps  [2] EMA num_updates
//   [3] gradient L2 norm of this call  [4] gradient inf-norm  [5] 1 if this call stepped  [6] EMA decay used (or -1)
//   [7] clip coefficient applied
// scal (workspace, floats): [0] ok  [1] clip_coef  [2] step_size  [3] 1/sqrt(bias_correction2)  [4] ema mode (0 none, 1 copy, 2 lerp)
//   [5] 1 - decay
constexpr int TR_BLOCKS = 296, TR_THREADS = 256;

struct TrainerArgs {
    SdesTrainerStepDesc d;
    double* partial;    // TR_BLOCKS * 2 doubles (sum of squares, max |g|) + nonfinite flags folded into max as inf/NaN
    uint32_t* ticket;
    float* scal;
};

__global__ void __launch_bounds__(TR_THREADS) trainer_reduce_kernel(const TrainerArgs a) {
    const SdesTrainerStepDesc& d = a.d;
    __shared__ double s_sq[TR_THREADS / 32];
    __shared__ float s_mx[TR_THREADS / 32];
    __shared__ int s_bad[TR_THREADS / 32];
    __shared__ bool s_last;
    double sq = 0.0;
    float mx = 0.f;
    int bad = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
        const float g = d.grads[i];
        bad |= !isfinite(g);
        sq += (double)g * (double)g;
        mx = fmaxf(mx, fabsf(g));
    }
    for (int o = 16; o > 0; o >>= 1) {
        sq += __shfl_xor_sync(0xffffffffu, sq, o);
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        bad |= __shfl_xor_sync(0xffffffffu, bad, o);
    }
    const int warp = threadIdx.x >> 5;
    if ((threadIdx.x & 31) == 0) { s_sq[warp] = sq; s_mx[warp] = mx; s_bad[warp] = bad; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < TR_THREADS / 32; ++w) { sq += s_sq[w]; mx = fmaxf(mx, s_mx[w]); bad |= s_bad[w]; }
        a.partial[3 * blockIdx.x + 0] = sq;
        a.partial[3 * blockIdx.x + 1] = (double)mx;
        a.partial[3 * blockIdx.x + 2] = (double)bad;
        __threadfence();
        s_last = atomicAdd(a.ticket, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (!s_last || threadIdx.x != 0) return;
    // ---- last block: the scalar logic of Trainable.step (solver/base.py:409-439)
    __threadfence();
    double tsq = 0.0, tmx = 0.0, tbad = 0.0;
    for (unsigned b = 0; b < gridDim.x; ++b) {
        tsq += a.partial[3 * b + 0];
        tmx = fmax(tmx, a.partial[3 * b + 1]);
        tbad += a.partial[3 * b + 2];
    }
    *a.ticket = 0u;
    const double norm2 = sqrt(tsq);
    double* st = d.state;
    bool loss_ok = true;
    if (d.loss != nullptr) {
        const float lv = d.loss[0];
        loss_ok = isfinite(d.max_loss) ? (fabsf(lv) <= d.max_loss) : isfinite(lv);  // :410-412
    }
    bool grad_ok;
    if (isfinite(d.max_grad)) grad_ok = tbad == 0.0 && tmx <= (double)d.max_grad;    // :419-421 (a NaN norm fails the comparison)
    else grad_ok = tbad == 0.0;                                                     // :413-418
    const bool ok = loss_ok && grad_ok;
    st[3] = norm2; st[4] = tmx; st[5] = ok ? 1.0 : 0.0; st[6] = -1.0; st[7] = 1.0;
    a.scal[0] = ok ? 1.f : 0.f;
    a.scal[4] = 0.f;
    if (!ok) { st[1] += 1.0; return; }
    // clip_grad_norm_(max_norm, norm_type=2): coef = max_norm / (total_norm + 1e-6), clamped to 1 (torch/nn/utils/clip_grad.py)
    double coef = 1.0;
    if (isfinite(d.grad_clip_norm)) {
        coef = (double)d.grad_clip_norm / (norm2 + 1e-6);
        if (coef > 1.0) coef = 1.0;
    }
    st[7] = coef;
    a.scal[1] = (float)coef;
    // torch.optim.Adam (single-tensor path): step += 1, bias corrections in double like the Python scalars
    st[0] += 1.0;
    const double step = st[0];
    const double bc1 = 1.0 - pow((double)d.beta1, step), bc2 = 1.0 - pow((double)d.beta2, step);
    a.scal[2] = (float)((double)d.lr / bc1);
    a.scal[3] = (float)(1.0 / sqrt(bc2));
    // EMA.update (solver/base.py:652-684)
    if (d.ema_shadow != nullptr) {
        st[2] += 1.0;
        const long long nu = (long long)st[2];
        if (d.ema_update_every > 0 && nu % d.ema_update_every == 0) {
            if (nu <= d.ema_update_after_step) {
                a.scal[4] = 1.f;
                st[6] = 0.0;
            } else {
                // get_current_decay (:642-650)
                const double epoch = fmax((double)(nu - d.ema_update_after_step - 1), 0.0);
                double decay = 0.0;
                if (epoch > 0.0) {
                    const double value = 1.0 - pow(1.0 + epoch / d.ema_inv_gamma, -d.ema_power);
                    decay = fmin(fmax(value, d.ema_min_value), d.ema_decay);
                }
                a.scal[4] = 2.f;
                a.scal[5] = (float)(1.0 - decay);
                st[6] = decay;
            }
        }
    }
}

__global__ void __launch_bounds__(TR_THREADS) trainer_update_kernel(const TrainerArgs a) {
    const SdesTrainerStepDesc& d = a.d;
    if (a.scal[0] == 0.f) return;
    const float coef = a.scal[1], step_size = a.scal[2], inv_sqrt_bc2 = a.scal[3], one_minus_decay = a.scal[5];
    const int ema_mode = (int)a.scal[4];
    const float b1 = d.beta1, b2 = d.beta2;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < d.n; i += (int64_t)gridDim.x * blockDim.x) {
        float p = d.params[i];
        float g = d.grads[i] * coef;
        if (d.weight_decay != 0.f) g = fmaf(d.weight_decay, p, g);       // grad.add(param, alpha=weight_decay)
        float m = d.exp_avg[i], v = d.exp_avg_sq[i];
        m = fmaf(1.0f - b1, g - m, m);                                     // exp_avg.lerp_(grad, 1 - beta1)
        v = fmaf(1.0f - b2, g * g, v * b2);                                // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        const float denom = sqrtf(v) * inv_sqrt_bc2 + d.eps;
        p = p - step_size * (m / denom);                                   // param.addcdiv_(exp_avg, denom, value=-step_size)
        d.exp_avg[i] = m;
        d.exp_avg_sq[i] = v;
        d.params[i] = p;
        if (ema_mode == 1) d.ema_shadow[i] = p;
        else if (ema_mode == 2) {
            const float s = d.ema_shadow[i];
            d.ema_shadow[i] = s - (s - p) * one_minus_decay;               // tmp = s - p; tmp *= 1 - decay; s -= tmp
        }
    }
}

// ------------------------------------------------------------------------------- eval moments
// out (doubles): [0] sum w  [1] sum w^2  [2] n  [3] unused  [4 + j] sum_b x_bj  [4 + d + j] sum_b x_bj^2
__global__ void __launch_bounds__(256) eval_moments_kernel(const float* __restrict__ x, const float* __restrict__ w, int64_t B, int dim,
                                                           double* __restrict__ out) {
    // one warp per row group: lanes stride over dimensions so that rows are read coalesced
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int64_t gw = (int64_t)blockIdx.x * wpb + warp, nw = (int64_t)gridDim.x * wpb;
    for (int j0 = 0; j0 < dim; j0 += 32) {
        const int j = j0 + lane;
        double s1 = 0.0, s2 = 0.0;
        if (j < dim)
            for (int64_t b = gw; b < B; b += nw) {
                const double v = (double)x[b * dim + j];
                s1 += v;
                s2 += v * v;
            }
        if (j < dim) {
            atomicAdd(out + 4 + j, s1);
            atomicAdd(out + 4 + dim + j, s2);
        }
    }
    if (w != nullptr) {
        double a = 0.0, a2 = 0.0;
        for (int64_t b = gw * 32 + lane; b < B; b += nw * 32) {
            const double v = (double)w[b];
            a += v;
            a2 += v * v;
        }
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xffffffffu, a, o);
            a2 += __shfl_xor_sync(0xffffffffu, a2, o);
        }
        if (lane == 0) { atomicAdd(out + 0, a); atomicAdd(out + 1, a2); }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) out[2] = (double)B;
}

// ---------------------------------------------------------------------------------- host side
cudaError_t launch_sample_prior(float* out, const float* uniforms, int64_t batch, int dim, float mean, float std, int truncated, float a,
                                float b, uint64_t seed, uint64_t traj_offset, cudaStream_t stream) {
    PriorArgs p;
    p.out = out; p.uniforms = uniforms; p.batch = batch; p.dim = dim; p.mean = mean; p.std = std; p.a = a; p.b = b;
    p.truncated = truncated; p.seed = seed; p.traj_offset = traj_offset;
    // l = Phi((a - mean) / std), u = Phi((b - mean) / std) in double like the Python scalars of trunc_normal_
    const double l = 0.5 * (1.0 + erf(((double)a - mean) / std / sqrt(2.0))), u = 0.5 * (1.0 + erf(((double)b - mean) / std / sqrt(2.0)));
    p.lo2 = (float)(2.0 * l - 1.0);
    p.hi2 = (float)(2.0 * u - 1.0);
    const int64_t total = batch * ((dim + 3) / 4);
    int blocks = (int)((total + 255) / 256);
    if (blocks > 1184) blocks = 1184;
    if (blocks < 1) blocks = 1;
    sample_prior_kernel<<<blocks, 256, 0, stream>>>(p);
    return cudaGetLastError();
}

size_t trainer_workspace_bytes() { return 3 * TR_BLOCKS * sizeof(double) + 256; }

cudaError_t launch_trainer_step(const SdesTrainerStepDesc& d, cudaStream_t stream) {
    TrainerArgs a;
    a.d = d;
    uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
    a.partial = reinterpret_cast<double*>(ws);
    a.ticket = reinterpret_cast<uint32_t*>(ws + 3 * TR_BLOCKS * sizeof(double));
    a.scal = reinterpret_cast<float*>(ws + 3 * TR_BLOCKS * sizeof(double) + 64);
    int blocks = (int)((d.n + TR_THREADS - 1) / TR_THREADS);
    if (blocks > TR_BLOCKS) blocks = TR_BLOCKS;
    if (blocks < 1) blocks = 1;
    trainer_reduce_kernel<<<blocks, TR_THREADS, 0, stream>>>(a);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    trainer_update_kernel<<<blocks, TR_THREADS, 0, stream>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_eval_moments(const float* x, const float* w, int64_t B, int dim, double* out, cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(out, 0, (size_t)(4 + 2 * dim) * sizeof(double), stream);
    if (e != cudaSuccess) return e;
    int blocks = (int)((B + 255) / 256);
    if (blocks > 592) blocks = 592;
    if (blocks < 1) blocks = 1;
    eval_moments_kernel<<<blocks, 256, 0, stream>>>(x, w, B, dim, out);
    return cudaGetLastError();
}

}  // namespace sdes
