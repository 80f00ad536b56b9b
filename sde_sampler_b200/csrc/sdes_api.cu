// sdes_api.cu — the extern "C" surface declared in include/sdes_b200.h, descriptor validation,
// workspace layout, the rnd-statistics kernels and the noise-stream test hook.
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "sdes_common.cuh"

namespace sdes {
void launch_prepare(const KParams& p, cudaStream_t stream);
cudaError_t launch_rollout_simt(const KParams& p, int sm_count, cudaStream_t stream);
bool mma_supported(const KParams& p);
int64_t mma4_weight_image_floats(const SdesRolloutDesc& d);
int mma_groups_per_sm();
cudaError_t launch_rollout_tc(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);
cudaError_t launch_mma_selftest(const float* A, const float* W, float* D, int K, int N, cudaStream_t stream);
// wide engine (sdes_wide.cu): d > SDES_MAX_DIM or a NICE target
bool wide_engine_needed(const SdesRolloutDesc& d);
const char* wide_validate(const SdesRolloutDesc& d);
size_t wide_workspace_bytes(const SdesRolloutDesc& d);
int64_t launch_rollout_wide(const KParams& kp, cudaStream_t stream, cudaError_t* err);
// Langevin / Euler integrator (sdes_integrate.cu)
cudaError_t launch_langevin(const IntegrateParams& a, cudaStream_t stream);
cudaError_t launch_affine_integrate(const SdesAffineIntegrateDesc& g, cudaStream_t stream);
cudaError_t launch_expectations(const float* xs, int64_t n_rows, int dim, double* out4, cudaStream_t stream);
// lv gradient (sdes_grad.cu)
size_t lv_grad_workspace_bytes(const SdesRolloutDesc& d, int64_t fused_bytes, int64_t chunk_rows);
int64_t launch_lv_grad(const KParams& kp, const SdesLvGradDesc& g, int64_t fused_bytes, bool simt, cudaStream_t stream, cudaError_t* err,
                       bool bptt, int sm_count);
size_t kl_grad_workspace_bytes(const SdesRolloutDesc& d, int64_t fused_bytes, int64_t chunk_rows, bool simt);
int64_t launch_lv_grad_wide_desc(const KParams& kp, const SdesLvGradDesc& g, bool simt, cudaStream_t stream, cudaError_t* err, bool bptt);

// caller-side kernels (sdes_trainer.cu)
cudaError_t launch_sample_prior(float* out, const float* uniforms, int64_t batch, int dim, float mean, float std, int truncated, float a,
                                float b, uint64_t seed, uint64_t traj_offset, cudaStream_t stream);
size_t trainer_workspace_bytes();
cudaError_t launch_trainer_step(const SdesTrainerStepDesc& d, cudaStream_t stream);
cudaError_t launch_eval_moments(const float* x, const float* w, int64_t B, int dim, double* out, cudaStream_t stream);

static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

static int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

static int64_t align_up(int64_t v, int64_t a) { return (v + a - 1) / a * a; }

static bool blob_layout(const SdesRolloutDesc& d, BlobLayout& bl) {
    const int64_t dim = d.dim;
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o += n; return r; };
    bl.in_w = take(C * dim);
    bl.in_b = take(C);
    bl.te_phase = take(C);
    for (int l = 0; l < d.te_hidden; ++l) {
        bl.te_h_w[l] = take((int64_t)C * (l == 0 ? 2 * C : C));
        bl.te_h_b[l] = take(C);
    }
    bl.te_out_w = take(C * C);
    bl.te_out_b = take(C);
    for (int l = 0; l < d.n_hidden; ++l) {
        bl.h_w[l] = take(C * C);
        bl.h_b[l] = take(C);
    }
    bl.out_w = take(dim * C);
    bl.out_b = take(dim);
    if (d.flags & SDES_F_HAS_GATE) {
        bl.g_phase = take(C);
        for (int l = 0; l < d.gate_hidden; ++l) {
            bl.g_h_w[l] = take((int64_t)C * (l == 0 ? 2 * C : C));
            bl.g_h_b[l] = take(C);
        }
        bl.g_out_w = take((int64_t)d.gate_dim * C);
        bl.g_out_b = take(d.gate_dim);
    }
    bl.total = o;
    return true;
}

static void ws_layout(const SdesRolloutDesc& d, WsLayout& w) {
    const int dpad = (d.flags & SDES_F_MLP_SIMT) ? pad_dim(d.dim) : mma_pad_dim(d.dim);
    const int64_t T = d.n_steps, K = d.target_kind == SDES_TARGET_GMM ? d.n_components : 0;
    int64_t o = 0;
    auto take = [&](int64_t n) { int64_t r = o; o = align_up(o + n, 64); return r; };
    w.dpad = dpad;
    w.tab = take(T * TAB_STRIDE);
    w.emb = take(T * C);
    w.gate = take(T * dpad);
    w.gmm_mu = take(((K + 1) & ~1ll) * dpad);  // padded to an even number of components
    w.gmm_h = take(((K + 1) & ~1ll) * dpad);
    w.gmm_c = take(64);
    w.prior = take(2 * dpad + 4);
    w.ref = take(2 * dpad + 4);
    const bool simt = (d.flags & SDES_F_MLP_SIMT) != 0;
    w.w_simt_len = simt ? (int64_t)d.dim * C + C + (int64_t)d.n_hidden * (C * C + C) + (int64_t)C * dpad + dpad : 0;
    w.w_simt = take(w.w_simt_len);
    w.w_mma4_len = simt ? 0 : mma4_weight_image_floats(d);
    w.w_mma4 = take(w.w_mma4_len);
    w.counter = take(4);
    const int64_t tiles128 = simt ? 0 : (d.batch + 127) / 128;
    w.progress = take(tiles128);
    w.state = take(tiles128 * (dpad + 1) * 128);
    w.total = o;
}

static int validate(const SdesRolloutDesc* d, bool need_ptrs) {
    if (d == nullptr) return fail(-1, "desc is NULL");
    if (d->struct_bytes != sizeof(SdesRolloutDesc))
        return fail(-2, "desc.struct_bytes=%u but this library expects %zu", d->struct_bytes, sizeof(SdesRolloutDesc));
    if (d->abi_version != SDES_ABI_VERSION) return fail(-2, "desc.abi_version=%u, library is %d", d->abi_version, SDES_ABI_VERSION);
    if (d->dim < 1 || d->dim > SDES_MAX_WIDE_DIM) return fail(-3, "dim=%d not in [1,%d]", d->dim, SDES_MAX_WIDE_DIM);
    if (d->n_steps < 1) return fail(-3, "n_steps=%d < 1", d->n_steps);
    if (d->batch < 0) return fail(-3, "batch < 0");
    if (d->traj_offset + (uint64_t)d->batch > 0xFFFFFFFFull) return fail(-3, "traj_offset + batch exceeds the 32-bit Philox trajectory counter");
    if (d->n_hidden < 0 || d->n_hidden > SDES_MAX_HIDDEN) return fail(-3, "n_hidden=%d not in [0,%d]", d->n_hidden, SDES_MAX_HIDDEN);
    if (d->te_hidden < 1 || d->te_hidden > SDES_MAX_HIDDEN) return fail(-3, "te_hidden=%d not in [1,%d]", d->te_hidden, SDES_MAX_HIDDEN);
    if (d->loss_kind < 0 || d->loss_kind > SDES_LOSS_EXP_INTEGRATOR) return fail(-3, "bad loss_kind %d", d->loss_kind);
    if (d->ctrl_kind < 0 || d->ctrl_kind > SDES_CTRL_LERP_TARGET) return fail(-3, "bad ctrl_kind %d", d->ctrl_kind);
    if (d->sde_kind < 0 || d->sde_kind > SDES_SDE_CONST_OU) return fail(-3, "bad sde_kind %d", d->sde_kind);
    if (d->target_kind < 0 || d->target_kind > SDES_TARGET_NICE) return fail(-3, "bad target_kind %d", d->target_kind);
    if ((d->flags & SDES_F_TRAJ_TILED) && wide_engine_needed(*d)) return fail(-3, "SDES_F_TRAJ_TILED is a fused-engine layout (d <= %d)", SDES_MAX_DIM);
    if (wide_engine_needed(*d)) {
        const char* why = wide_validate(*d);
        if (why != nullptr) return fail(-3, "wide engine (dim=%d): %s", d->dim, why);
    }
    if (d->loss_kind != SDES_LOSS_EXP_INTEGRATOR && d->sde_kind == SDES_SDE_NONE) return fail(-3, "this loss needs an sde");
    if (d->ctrl_kind >= SDES_CTRL_LERP && d->sde_kind == SDES_SDE_NONE) return fail(-3, "Lerp controls need an sde");
    if (d->flags & SDES_F_HAS_GATE) {
        if (d->gate_hidden < 1 || d->gate_hidden > SDES_MAX_HIDDEN) return fail(-3, "gate_hidden=%d not in [1,%d]", d->gate_hidden, SDES_MAX_HIDDEN);
        if (d->gate_dim != 1 && d->gate_dim != d->dim) return fail(-3, "gate_dim=%d must be 1 or dim", d->gate_dim);
    }
    if (d->target_kind == SDES_TARGET_GMM && (d->n_components < 1 || d->n_components > SDES_MAX_COMPONENTS))
        return fail(-3, "n_components=%d not in [1,%d]", d->n_components, SDES_MAX_COMPONENTS);
    if (d->target_kind == SDES_TARGET_MULTIWELL && (d->n_double_wells < 0 || d->n_double_wells > d->dim))
        return fail(-3, "n_double_wells=%d not in [0,dim]", d->n_double_wells);
    if (d->target_kind == SDES_TARGET_FUNNEL && d->dim < 2) return fail(-3, "funnel needs dim >= 2");
    BlobLayout bl;
    blob_layout(*d, bl);
    if (d->n_params != bl.total) return fail(-4, "n_params=%lld but the layout for this descriptor has %lld floats", (long long)d->n_params, (long long)bl.total);
    if (!need_ptrs) return 0;
    if (!d->ts || !d->params || !d->x0 || !d->x_T || !d->rnd) return fail(-5, "ts/params/x0/x_T/rnd must be non-NULL");
    if (d->target_kind == SDES_TARGET_GMM && (!d->gmm_loc || !d->gmm_scale)) return fail(-5, "gmm_loc/gmm_scale are NULL");
    if (d->target_kind == SDES_TARGET_NICE && !d->nice_params) return fail(-5, "nice_params is NULL");
    const bool need_prior = (d->loss_kind == SDES_LOSS_TIME_REVERSAL && !(d->flags & SDES_F_RND0_ZERO)) ||
                            d->ctrl_kind == SDES_CTRL_LERP || d->ctrl_kind == SDES_CTRL_LERP_PRIOR ||
                            (d->flags & SDES_F_REFERENCE_CTRL);
    if (need_prior && (!d->prior_loc || !d->prior_scale)) return fail(-5, "prior_loc/prior_scale are NULL but this configuration reads the prior");
    if (d->loss_kind != SDES_LOSS_TIME_REVERSAL && (!d->ref_loc || !d->ref_scale)) return fail(-5, "ref_loc/ref_scale are NULL");
    if ((d->flags & SDES_F_NOISE_FROM_HBM) && !d->noise) return fail(-5, "SDES_F_NOISE_FROM_HBM set but noise is NULL");
    if ((d->flags & SDES_F_RETURN_TRAJ) && !d->xs) return fail(-5, "SDES_F_RETURN_TRAJ set but xs is NULL");
    if (!d->workspace) return fail(-5, "workspace is NULL");
    if (reinterpret_cast<uintptr_t>(d->workspace) % 256 != 0) return fail(-5, "workspace must be 256-byte aligned");
    if (d->gate_cot != nullptr && ((d->flags & SDES_F_MLP_SIMT) || wide_engine_needed(*d)))
        return fail(-3, "gate_cot is an output of the tensor-core fused engine only (d <= %d, no SDES_F_MLP_SIMT)", SDES_MAX_DIM);
    if (d->score_keep != nullptr && ((d->flags & SDES_F_MLP_SIMT) || wide_engine_needed(*d)))
        return fail(-3, "score_keep is an output of the tensor-core fused engine only (d <= %d, no SDES_F_MLP_SIMT)", SDES_MAX_DIM);
    if (d->score_keep != nullptr && d->gate_cot != nullptr) return fail(-3, "gate_cot (lv) and score_keep (kl) are alternatives");
    return 0;
}

// ------------------------------------------------------------------------- rnd statistics
// One CTA: B is at most a few 1e6 floats, the rollout that produced them took milliseconds.
__global__ void __launch_bounds__(1024) rnd_stats_kernel(const float* __restrict__ rnd, int64_t B, int mode,
                                                         float max_rnd, const uint8_t* __restrict__ smask,
                                                         double* __restrict__ out) {
    __shared__ double s_a[32], s_b[32], s_c[32];
    __shared__ float s_m[32];
    __shared__ float s_max;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double n = 0.0, s1 = 0.0, s2 = 0.0;
    float mx = -INFINITY;
    bool nan_seen = false;
    for (int64_t i = tid; i < B; i += blockDim.x) {
        const float r = rnd[i];
        // losses/oc.py:50-58; mode 2 = no mask (compute_results, :94-123)
        bool keep = mode == 2 ? true : (mode == 1 ? (r < max_rnd) : isfinite(r));
        if (smask != nullptr) keep = keep && smask[i] != 0;
        if (keep) {
            n += 1.0;
            s1 += (double)r;
            s2 += (double)r * (double)r;
            mx = fmaxf(mx, -r);
            nan_seen |= (r != r);
        }
    }
    if (nan_seen) mx = NAN;
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        mx = (mx != mx || om != om) ? NAN : fmaxf(mx, om);
    }
    if (lane == 0) { s_a[warp] = n; s_b[warp] = s1; s_c[warp] = s2; s_m[warp] = mx; }
    __syncthreads();
    if (warp == 0) {
        n = s_a[lane]; s1 = s_b[lane]; s2 = s_c[lane]; mx = s_m[lane];
        for (int o = 16; o > 0; o >>= 1) {
            n += __shfl_xor_sync(0xffffffffu, n, o);
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            mx = (mx != mx || om != om) ? NAN : fmaxf(mx, om);
        }
        if (lane == 0) {
            out[0] = n; out[1] = s1; out[2] = s2; out[3] = (double)mx; out[5] = (double)B;
            out[6] = (s2 - s1 * s1 / n) / (n - 1.0);  // unbiased variance of the kept rnd (lv loss of THIS shard)
            out[7] = s1 / n;                          // their mean (kl loss)
            s_max = mx;
        }
    }
    __syncthreads();
    const float gmx = s_max;
    double se = 0.0;
    for (int64_t i = tid; i < B; i += blockDim.x) {
        const float r = rnd[i];
        bool keep = mode == 2 ? true : (mode == 1 ? (r < max_rnd) : isfinite(r));
        if (smask != nullptr) keep = keep && smask[i] != 0;
        if (keep) se += (double)expf(-r - gmx);
    }
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    __syncthreads();
    if (lane == 0) s_a[warp] = se;
    __syncthreads();
    if (warp == 0) {
        se = s_a[lane];
        for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
        if (lane == 0) out[4] = se;
    }
}

// (world, 8) per-rank statistics -> the statistics of the global batch (include/sdes_b200.h: sdes_merge_stats); one warp
__global__ void merge_stats_kernel(const double* __restrict__ g, int world, double* __restrict__ out) {
    const int lane = threadIdx.x;
    double n = 0.0, s1 = 0.0, s2 = 0.0, nt = 0.0, mx = -INFINITY;
    bool nan_seen = false;
    for (int r = lane; r < world; r += 32) {
        const double* v = g + (int64_t)r * 8;
        n += v[0]; s1 += v[1]; s2 += v[2]; nt += v[5];
        nan_seen |= (v[3] != v[3]);
        if (v[0] > 0.0 && v[3] > mx) mx = v[3];
    }
    for (int o = 16; o > 0; o >>= 1) {
        n += __shfl_xor_sync(0xffffffffu, n, o);
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        nt += __shfl_xor_sync(0xffffffffu, nt, o);
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        nan_seen |= __shfl_xor_sync(0xffffffffu, (int)nan_seen, o) != 0;
    }
    double se = 0.0;
    for (int r = lane; r < world; r += 32) {
        const double* v = g + (int64_t)r * 8;
        if (v[0] > 0.0) se += v[4] * exp(v[3] - mx);
    }
    for (int o = 16; o > 0; o >>= 1) se += __shfl_xor_sync(0xffffffffu, se, o);
    if (lane == 0) {
        const double gmx = nan_seen ? (double)NAN : mx;
        out[0] = n; out[1] = s1; out[2] = s2; out[3] = gmx; out[4] = nan_seen ? (double)NAN : se; out[5] = nt;
        out[6] = (s2 - s1 * s1 / n) / (n - 1.0);
        out[7] = s1 / n;
    }
}

__global__ void weights_kernel(const float* __restrict__ rnd, int64_t B, const double* __restrict__ stats,
                               float* __restrict__ w) {
    const float mx = (float)stats[3];
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x)
        w[i] = expf(-rnd[i] - mx);  // losses/oc.py:104-105
}

// d (lv loss) / d rnd_b = 2 (rnd_b - mean) / (n - 1) for kept b, 0 otherwise, times the upstream scalar
__global__ void lv_weights_kernel(const float* __restrict__ rnd, int64_t B, int mode, float max_rnd, const uint8_t* __restrict__ smask,
                                  const double* __restrict__ stats, const float* __restrict__ upstream, float* __restrict__ w) {
    const double n = stats[0], mean = stats[1] / stats[0];
    const double up = upstream != nullptr ? (double)upstream[0] : 1.0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const float r = rnd[i];
        bool keep = mode == 2 ? true : (mode == 1 ? (r < max_rnd) : isfinite(r));
        if (smask != nullptr) keep = keep && smask[i] != 0;
        w[i] = keep ? (float)(2.0 * ((double)r - mean) / (n - 1.0) * up) : 0.f;
    }
}

// d (lv_traj loss) / d rnd[t, i] (losses/oc.py:78-84): the loss is the mean over kept samples of the unbiased variance across
// each sample's traj_per_sample trajectories
__global__ void __launch_bounds__(256) lv_traj_weights_kernel(const float* __restrict__ rnd, int64_t B0, int tps, int mode, float max_rnd,
                                                              const uint8_t* __restrict__ smask, const double* __restrict__ st3,
                                                              const float* __restrict__ upstream, float* __restrict__ w) {
    const double up = upstream != nullptr ? (double)upstream[0] : 1.0;
    const double scale = up * 2.0 / ((double)tps - 1.0) / st3[1];
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B0; b += (int64_t)gridDim.x * blockDim.x) {
        bool keep = true;
        double s1 = 0.0;
        for (int t = 0; t < tps; ++t) {
            const int64_t i = (int64_t)t * B0 + b;
            const float r = rnd[i];
            bool k = mode == 2 ? true : (mode == 1 ? (r < max_rnd) : isfinite(r));
            if (smask != nullptr) k = k && smask[i] != 0;
            keep = keep && k;
            s1 += (double)r;
        }
        const double mean = s1 / tps;
        for (int t = 0; t < tps; ++t) {
            const int64_t i = (int64_t)t * B0 + b;
            w[i] = keep ? (float)(((double)rnd[i] - mean) * scale) : 0.f;
        }
    }
}

// d (kl loss) / d rnd_b = 1 / n_kept for kept b (mean of the kept entries, losses/oc.py:90), times the upstream scalar
__global__ void kl_weights_kernel(const float* __restrict__ rnd, int64_t B, int mode, float max_rnd, const uint8_t* __restrict__ smask,
                                  const double* __restrict__ stats, const float* __restrict__ upstream, float* __restrict__ w) {
    const double up = upstream != nullptr ? (double)upstream[0] : 1.0;
    const float wk = (float)(up / stats[0]);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
        const float r = rnd[i];
        bool keep = mode == 2 ? true : (mode == 1 ? (r < max_rnd) : isfinite(r));
        if (smask != nullptr) keep = keep && smask[i] != 0;
        w[i] = keep ? wk : 0.f;
    }
}

// lv_traj (losses/oc.py:78-84): rnd is (tps, B0) — tps trajectories per initial sample; a sample is kept when all its
// copies pass the mask; the loss is the mean over kept samples of the unbiased variance across the copies.
// out: [0] sum of per-sample variances (kept), [1] kept samples, [2] all samples.
__global__ void __launch_bounds__(256) lv_traj_stats_kernel(const float* __restrict__ rnd, int64_t B0, int tps, int mode, float max_rnd,
                                                            const uint8_t* __restrict__ smask, double* __restrict__ out) {
    __shared__ double s_v[8], s_n[8];
    double vsum = 0.0, nk = 0.0;
    for (int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; b < B0; b += (int64_t)gridDim.x * blockDim.x) {
        bool keep = true;
        double s1 = 0.0, s2 = 0.0;
        for (int t = 0; t < tps; ++t) {
            const int64_t i = (int64_t)t * B0 + b;
            const float r = rnd[i];
            bool k = mode == 2 ? true : (mode == 1 ? (r < max_rnd) : isfinite(r));
            if (smask != nullptr) k = k && smask[i] != 0;
            keep = keep && k;
            s1 += (double)r;
            s2 += (double)r * (double)r;
        }
        if (keep) {
            vsum += (s2 - s1 * s1 / tps) / (tps - 1.0);
            nk += 1.0;
        }
    }
    for (int o = 16; o > 0; o >>= 1) {
        vsum += __shfl_xor_sync(0xffffffffu, vsum, o);
        nk += __shfl_xor_sync(0xffffffffu, nk, o);
    }
    if ((threadIdx.x & 31) == 0) { s_v[threadIdx.x >> 5] = vsum; s_n[threadIdx.x >> 5] = nk; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double v = 0.0, n = 0.0;
        for (int i = 0; i < 8; ++i) { v += s_v[i]; n += s_n[i]; }
        atomicAdd(out + 0, v);
        atomicAdd(out + 1, n);
        if (blockIdx.x == 0) out[2] = (double)B0;
    }
}

__global__ void philox_normal_kernel(uint64_t seed, uint64_t traj_offset, int64_t B, int T, int dim, float* __restrict__ out) {
    const int nchunk = (dim + 3) / 4;
    const int64_t total = (int64_t)T * B * nchunk;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(e % nchunk);
        const int64_t b = (e / nchunk) % B;
        const int i = (int)(e / ((int64_t)nchunk * B));
        float n[4];
        normal4(seed, (uint32_t)(traj_offset + (uint64_t)b), (uint32_t)i, (uint32_t)q, n[0], n[1], n[2], n[3]);
        for (int r = 0; r < 4; ++r)
            if (4 * q + r < dim) out[((int64_t)i * B + b) * dim + 4 * q + r] = n[r];
    }
}

__global__ void gelu_probe_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) y[i] = gelu_fast(x[i]);
}

__global__ void gelu_pair_probe_kernel(const float2* __restrict__ x, float2* __restrict__ y, int64_t n2) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x) y[i] = gelu_fast2(x[i]);
}

static int sm_count_cached() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace sdes

using namespace sdes;

extern "C" {

int sdes_version(void) { return SDES_ABI_VERSION; }

const char* sdes_last_error(void) { return g_err; }

int64_t sdes_launch_count(void) { return g_launches.load(); }

size_t sdes_workspace_bytes(const SdesRolloutDesc* desc) {
    if (validate(desc, false) != 0) return 0;
    if (wide_engine_needed(*desc)) return wide_workspace_bytes(*desc);
    WsLayout w;
    ws_layout(*desc, w);
    return (size_t)w.total * sizeof(float);
}

int sdes_tcgen05_supported(const SdesRolloutDesc* desc) {
    if (validate(desc, false) != 0) return 0;
    if (wide_engine_needed(*desc)) return 1;  // every Linear of the wide engine is a tcgen05 GEMM
    KParams p;
    memset(&p, 0, sizeof(p));
    p.d = *desc;
    blob_layout(p.d, p.bl);
    ws_layout(p.d, p.ws);
    return mma_supported(p) ? 1 : 0;
}

int sdes_rollout_fwd(const SdesRolloutDesc* desc, void* stream_) {
    g_err[0] = 0;
    int rc = validate(desc, true);
    if (rc != 0) return rc;
    KParams p;
    memset(&p, 0, sizeof(p));
    p.d = *desc;
    blob_layout(p.d, p.bl);
    if (wide_engine_needed(*desc)) {
        const size_t need = wide_workspace_bytes(*desc);
        if (need > desc->workspace_bytes) return fail(-6, "workspace_bytes=%zu < required %zu", desc->workspace_bytes, need);
        if (desc->batch == 0) return 0;
        cudaError_t we = cudaSuccess;
        g_launches += launch_rollout_wide(p, reinterpret_cast<cudaStream_t>(stream_), &we);
        if (we != cudaSuccess) return fail(-7, "wide engine launch failed: %s", cudaGetErrorString(we));
        return 0;
    }
    ws_layout(p.d, p.ws);
    if ((size_t)p.ws.total * sizeof(float) > desc->workspace_bytes)
        return fail(-6, "workspace_bytes=%zu < required %zu", desc->workspace_bytes, (size_t)p.ws.total * sizeof(float));
    if (desc->batch == 0) return 0;
    p.n_tiles = (int)((desc->batch + 31) / 32);
    const int sms = sm_count_cached();
    {
        // time-chunked scheduling of the tcgen05 engine: aim for >= 8 work items per resident group,
        // chunks of at least 8 steps (state parks in L2 between chunks: ~29 KB per item each way)
        const int64_t tiles128 = (desc->batch + 127) / 128, groups = (int64_t)mma_groups_per_sm() * sms;
        int64_t nc = (8 * groups + tiles128 - 1) / tiles128;
        const int64_t nc_max = desc->n_steps / 8 > 0 ? desc->n_steps / 8 : 1;
        if (nc > nc_max) nc = nc_max;
        if (nc < 1) nc = 1;
        p.chunk_steps = (int)((desc->n_steps + nc - 1) / nc);
        p.n_chunks = (desc->n_steps + p.chunk_steps - 1) / p.chunk_steps;
    }
    for (int r = 0; r < 10; ++r) {  // Philox4x32-10 key schedule of this call's seed (sdes_common.cuh philox4x32_10)
        p.philox_rk[2 * r] = (uint32_t)desc->seed + (uint32_t)r * PHILOX_W0;
        p.philox_rk[2 * r + 1] = (uint32_t)(desc->seed >> 32) + (uint32_t)r * PHILOX_W1;
    }
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    launch_prepare(p, stream);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "prepare kernel launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    if (desc->flags & SDES_F_MLP_SIMT) {
        e = launch_rollout_simt(p, sms, stream);
    } else {
        if (!mma_supported(p)) return fail(-8, "the tcgen05 engine does not support this descriptor (dim=%d, n_hidden=%d); set SDES_F_MLP_SIMT", desc->dim, desc->n_hidden);
        int n = 1;
        e = launch_rollout_tc(p, sms, stream, &n);
        g_launches += n - 1;
    }
    if (e != cudaSuccess) return fail(-7, "rollout kernel launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

// ---- Langevin / Euler integrator (sdes_integrate.cu)

// The integrator only reads the target part of the descriptor: complete the rest with a valid dummy configuration so
// that the common validation, the fused prologue (target images) and the workspace layout can be reused.
static int integrate_setup(const SdesRolloutDesc* desc, KParams& p) {
    if (desc == nullptr) return fail(-1, "desc is NULL");
    if (desc->struct_bytes != sizeof(SdesRolloutDesc)) return fail(-2, "desc.struct_bytes mismatch");
    memset(&p, 0, sizeof(p));
    p.d = *desc;
    p.d.loss_kind = SDES_LOSS_EXP_INTEGRATOR;
    p.d.ctrl_kind = SDES_CTRL_CLIPPED;
    p.d.sde_kind = SDES_SDE_NONE;
    p.d.flags = (desc->flags & SDES_F_NOISE_FROM_HBM) | SDES_F_MLP_SIMT;
    p.d.n_steps = 1;
    p.d.n_hidden = 0;
    p.d.te_hidden = 1;
    blob_layout(p.d, p.bl);
    p.d.n_params = p.bl.total;
    if (wide_engine_needed(p.d)) return fail(-8, "the Langevin integrator is implemented for d <= %d with analytic targets", SDES_MAX_DIM);
    int rc = validate(&p.d, false);
    if (rc != 0) return rc;
    ws_layout(p.d, p.ws);
    return 0;
}

static __global__ void langevin_images_kernel(const KParams p) {
    // the target images of the fused prologue (GMM mu / h / c and the pair mask) without the network tables
    const SdesRolloutDesc& d = p.d;
    float* ws = reinterpret_cast<float*>(d.workspace);
    const int dim = d.dim, dpad = p.ws.dpad, tid = threadIdx.x;
    __shared__ uint32_t s_mask;
    if (tid == 0) s_mask = 0u;
    __syncthreads();
    if (d.target_kind == SDES_TARGET_GMM) {
        const int K = d.n_components, K2 = (K + 1) & ~1;
        if (tid < dim) {
            bool differs = false;
            for (int k = 1; k < K; ++k)
                differs |= d.gmm_loc[(int64_t)k * dim + tid] != d.gmm_loc[tid] || d.gmm_scale[(int64_t)k * dim + tid] != d.gmm_scale[tid];
            if (differs) atomicOr(&s_mask, 1u << (tid >> 1));
        }
        for (int e = tid; e < K2 * dpad; e += blockDim.x) {
            const int k = e / dpad, j = e % dpad;
            float mu = 0.f, h = 0.f;
            if (j < dim && k < K) {
                mu = d.gmm_loc[(int64_t)k * dim + j];
                const float sc = d.gmm_scale[(int64_t)k * dim + j];
                h = 0.5f / (sc * sc);
            }
            ws[p.ws.gmm_mu + e] = mu;
            ws[p.ws.gmm_h + e] = h;
        }
        for (int k = tid; k < 64; k += blockDim.x) {
            float c = -INFINITY;
            if (k < K) {
                float logw = 0.f;
                if (d.gmm_weights != nullptr) {
                    float tot = 0.f;
                    for (int q = 0; q < K; ++q) tot += d.gmm_weights[q];
                    logw = logf(d.gmm_weights[k] / tot);
                }
                float sl = 0.f;
                for (int j = 0; j < dim; ++j) sl += logf(d.gmm_scale[k * dim + j]);
                c = logw - sl - 0.5f * (float)dim * LOG_2PI;
            }
            ws[p.ws.gmm_c + k] = c;
        }
    }
    __syncthreads();
    if (tid == 0) reinterpret_cast<uint32_t*>(ws + p.ws.counter)[1] = s_mask;
}

static int grad_setup(const SdesRolloutDesc* desc, const SdesLvGradDesc* g, KParams& p, bool& simt) {
    int rc = validate(desc, false);
    if (rc != 0) return rc;
    if (g == nullptr || g->struct_bytes != sizeof(SdesLvGradDesc)) return fail(-2, "SdesLvGradDesc is NULL or has the wrong struct_bytes");
    memset(&p, 0, sizeof(p));
    p.d = *desc;
    simt = (desc->flags & SDES_F_MLP_SIMT) != 0;
    if (wide_engine_needed(*desc)) {  // wide engine: the forward ran with SDES_F_KEEP_FOR_GRAD in the same workspace
        if (!(desc->flags & SDES_F_KEEP_FOR_GRAD)) return fail(-8, "wide-engine gradient needs the forward's SDES_F_KEEP_FOR_GRAD workspace");
        blob_layout(p.d, p.bl);
        return 0;
    }
    p.d.flags = (desc->flags | SDES_F_MLP_SIMT) & ~(uint32_t)SDES_F_RETURN_TRAJ;  // fp32 tables / target images of the fused prologue
    blob_layout(p.d, p.bl);
    ws_layout(p.d, p.ws);
    return 0;
}

size_t sdes_lv_grad_workspace_bytes(const SdesRolloutDesc* desc, const SdesLvGradDesc* g) {
    KParams p;
    bool simt;
    if (grad_setup(desc, g, p, simt) != 0) return 0;
    if (wide_engine_needed(*desc)) return wide_workspace_bytes(*desc);
    return lv_grad_workspace_bytes(p.d, p.ws.total * (int64_t)sizeof(float), g->chunk_rows);
}

int sdes_rollout_lv_grad(const SdesRolloutDesc* desc, const SdesLvGradDesc* g, void* stream_) {
    g_err[0] = 0;
    KParams p;
    bool simt;
    int rc = grad_setup(desc, g, p, simt);
    if (rc != 0) return rc;
    if (!desc->ts || !desc->params || !desc->workspace) return fail(-5, "ts/params/workspace must be non-NULL");
    if (wide_engine_needed(*desc)) {
        if (!g->w || !g->grad_params || !g->grad_emb) return fail(-5, "w/grad_params/grad_emb must be non-NULL");
        if ((desc->flags & SDES_F_NOISE_FROM_HBM) && !desc->noise) return fail(-5, "SDES_F_NOISE_FROM_HBM set but noise is NULL");
        if (wide_workspace_bytes(*desc) > desc->workspace_bytes) return fail(-6, "workspace_bytes too small for the keep-mode wide workspace");
        if (desc->batch == 0) return 0;
        cudaError_t we = cudaSuccess;
        g_launches += launch_lv_grad_wide_desc(p, *g, simt, reinterpret_cast<cudaStream_t>(stream_), &we, false);
        if (we != cudaSuccess) return fail(-7, "wide lv gradient launch failed: %s", cudaGetErrorString(we));
        return 0;
    }
    if (!g->xs || !g->w || !g->grad_params || !g->grad_emb) return fail(-5, "xs/w/grad_params/grad_emb must be non-NULL");
    if ((desc->flags & SDES_F_NOISE_FROM_HBM) && !desc->noise) return fail(-5, "SDES_F_NOISE_FROM_HBM set but noise is NULL");
    if (desc->target_kind == SDES_TARGET_GMM && (!desc->gmm_loc || !desc->gmm_scale)) return fail(-5, "gmm_loc/gmm_scale are NULL");
    if (reinterpret_cast<uintptr_t>(desc->workspace) % 256 != 0) return fail(-5, "workspace must be 256-byte aligned");
    const int64_t fused = p.ws.total * (int64_t)sizeof(float);
    const size_t need = lv_grad_workspace_bytes(p.d, fused, g->chunk_rows);
    if (need > desc->workspace_bytes) return fail(-6, "workspace_bytes=%zu < required %zu", desc->workspace_bytes, need);
    if (desc->batch == 0) return 0;
    cudaError_t e = cudaSuccess;
    g_launches += launch_lv_grad(p, *g, fused, simt, reinterpret_cast<cudaStream_t>(stream_), &e, false, sm_count_cached());
    if (e != cudaSuccess) return fail(-7, "lv gradient launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// ---- kl / kl_ito gradient (sdes_adjoint.cu + the GEMM passes of sdes_grad.cu)
static int kl_grad_setup(const SdesRolloutDesc* desc, const SdesLvGradDesc* g, KParams& p, bool& simt) {
    int rc = validate(desc, false);
    if (rc != 0) return rc;
    rc = grad_setup(desc, g, p, simt);
    if (rc != 0) return rc;
    if (wide_engine_needed(*desc)) {
        if (!(desc->flags & SDES_F_KEEP_SCORE)) return fail(-8, "wide-engine kl gradient needs the forward's SDES_F_KEEP_FOR_GRAD | SDES_F_KEEP_SCORE workspace");
        if (desc->target_kind == SDES_TARGET_MULTIWELL || desc->target_kind == SDES_TARGET_FUNNEL)
            return fail(-8, "kl gradient on the wide engine: NICE, GMM and Gaussian targets (a wide funnel / multi-well Hessian is not implemented)");
        if (desc->target_kind == SDES_TARGET_GMM && desc->n_components > 1 && desc->ctrl_kind != SDES_CTRL_CLIPPED &&
            desc->ctrl_kind != SDES_CTRL_LERP_PRIOR && !(g->flags & (SDES_GRAD_TARGET_SCORE_CONST | SDES_GRAD_SCORE_DETACHED)))
            return fail(-8, "kl gradient with a %d-component GMM score inside the control needs SDES_GRAD_TARGET_SCORE_CONST", desc->n_components);
        return 0;
    }
    if (desc->target_kind == SDES_TARGET_GMM && desc->n_components > 1 && desc->ctrl_kind != SDES_CTRL_CLIPPED &&
        desc->ctrl_kind != SDES_CTRL_LERP_PRIOR && !(g->flags & (SDES_GRAD_TARGET_SCORE_CONST | SDES_GRAD_SCORE_DETACHED)))
        return fail(-8, "kl gradient with a %d-component GMM score inside the control needs SDES_GRAD_TARGET_SCORE_CONST "
                        "(the reference differentiates an autograd score without create_graph)", desc->n_components);
    return 0;
}

size_t sdes_kl_grad_workspace_bytes(const SdesRolloutDesc* desc, const SdesLvGradDesc* g) {
    KParams p;
    bool simt;
    if (kl_grad_setup(desc, g, p, simt) != 0) return 0;
    if (wide_engine_needed(*desc)) return wide_workspace_bytes(*desc);
    return kl_grad_workspace_bytes(p.d, p.ws.total * (int64_t)sizeof(float), g->chunk_rows, simt);
}

int sdes_rollout_kl_grad(const SdesRolloutDesc* desc, const SdesLvGradDesc* g, void* stream_) {
    g_err[0] = 0;
    KParams p;
    bool simt;
    int rc = kl_grad_setup(desc, g, p, simt);
    if (rc != 0) return rc;
    if (!desc->ts || !desc->params || !desc->workspace) return fail(-5, "ts/params/workspace must be non-NULL");
    if (wide_engine_needed(*desc)) {
        if (!g->w || !g->grad_params || !g->grad_emb) return fail(-5, "w/grad_params/grad_emb must be non-NULL");
        if ((desc->flags & SDES_F_NOISE_FROM_HBM) && !desc->noise) return fail(-5, "SDES_F_NOISE_FROM_HBM set but noise is NULL");
        if (wide_workspace_bytes(*desc) > desc->workspace_bytes) return fail(-6, "workspace_bytes too small for the keep-mode wide workspace");
        if (desc->batch == 0) return 0;
        cudaError_t we = cudaSuccess;
        g_launches += launch_lv_grad_wide_desc(p, *g, simt, reinterpret_cast<cudaStream_t>(stream_), &we, true);
        if (we != cudaSuccess) return fail(-7, "wide kl gradient launch failed: %s", cudaGetErrorString(we));
        return 0;
    }
    if (!g->xs || !g->w || !g->grad_params || !g->grad_emb) return fail(-5, "xs/w/grad_params/grad_emb must be non-NULL");
    if ((desc->flags & SDES_F_NOISE_FROM_HBM) && !desc->noise) return fail(-5, "SDES_F_NOISE_FROM_HBM set but noise is NULL");
    if (desc->target_kind == SDES_TARGET_GMM && (!desc->gmm_loc || !desc->gmm_scale)) return fail(-5, "gmm_loc/gmm_scale are NULL");
    if (reinterpret_cast<uintptr_t>(desc->workspace) % 256 != 0) return fail(-5, "workspace must be 256-byte aligned");
    const int64_t fused = p.ws.total * (int64_t)sizeof(float);
    const size_t need = kl_grad_workspace_bytes(p.d, fused, g->chunk_rows, simt);
    if (need > desc->workspace_bytes) return fail(-6, "workspace_bytes=%zu < required %zu", desc->workspace_bytes, need);
    if (desc->batch == 0) return 0;
    cudaError_t e = cudaSuccess;
    g_launches += launch_lv_grad(p, *g, fused, simt, reinterpret_cast<cudaStream_t>(stream_), &e, true, sm_count_cached());
    if (e != cudaSuccess) return fail(-7, "kl gradient launch failed: %s", cudaGetErrorString(e));
    return 0;
}

size_t sdes_integrate_workspace_bytes(const SdesRolloutDesc* desc) {
    KParams p;
    if (integrate_setup(desc, p) != 0) return 0;
    return (size_t)p.ws.total * sizeof(float);
}

int sdes_langevin_integrate(const SdesRolloutDesc* desc, const SdesIntegrateDesc* g, void* stream_) {
    g_err[0] = 0;
    KParams p;
    int rc = integrate_setup(desc, p);
    if (rc != 0) return rc;
    if (g == nullptr || g->struct_bytes != sizeof(SdesIntegrateDesc)) return fail(-2, "SdesIntegrateDesc is NULL or has the wrong struct_bytes");
    if (g->n_steps < 1 || g->n_out < 0) return fail(-3, "n_steps must be >= 1 and n_out >= 0");
    if (!g->timesteps || !g->x_init || (g->n_out > 0 && (!g->out_ts || !g->xs_out))) return fail(-5, "timesteps/x_init/out_ts/xs_out must be non-NULL");
    if (desc->target_kind == SDES_TARGET_GMM && (!desc->gmm_loc || !desc->gmm_scale)) return fail(-5, "gmm_loc/gmm_scale are NULL");
    if ((desc->flags & SDES_F_NOISE_FROM_HBM) && !desc->noise) return fail(-5, "SDES_F_NOISE_FROM_HBM set but noise is NULL");
    if (!desc->workspace || reinterpret_cast<uintptr_t>(desc->workspace) % 256 != 0) return fail(-5, "workspace must be non-NULL and 256-byte aligned");
    if ((size_t)p.ws.total * sizeof(float) > desc->workspace_bytes) return fail(-6, "workspace_bytes too small");
    if (desc->batch == 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    langevin_images_kernel<<<1, 256, 0, stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "langevin prologue launch failed: %s", cudaGetErrorString(e));
    IntegrateParams a;
    a.d = p.d; a.ws = p.ws; a.timesteps = g->timesteps; a.out_ts = g->out_ts; a.n_steps = g->n_steps; a.n_out = g->n_out;
    a.diff_coeff = g->diff_coeff; a.clip_score = g->clip_score; a.eps = g->eps; a.x_init = g->x_init; a.xs_out = g->xs_out;
    a.noise_is_increment = g->noise_is_increment;
    e = launch_langevin(a, stream);
    if (e != cudaSuccess) return fail(-7, "langevin kernel launch failed: %s", cudaGetErrorString(e));
    g_launches += 2;
    return 0;
}

int sdes_affine_integrate(const SdesAffineIntegrateDesc* g, void* stream_) {
    g_err[0] = 0;
    if (g == nullptr || g->struct_bytes != sizeof(SdesAffineIntegrateDesc)) return fail(-2, "SdesAffineIntegrateDesc is NULL or has the wrong struct_bytes");
    if (g->dim < 1 || g->batch < 0 || g->n_steps < 1 || g->n_out < 0) return fail(-3, "dim >= 1, batch >= 0, n_steps >= 1, n_out >= 0 required");
    if (g->traj_offset + (uint64_t)g->batch > 0xFFFFFFFFull) return fail(-3, "traj_offset + batch exceeds the 32-bit Philox trajectory counter");
    if (!g->timesteps || !g->tab || !g->x_init || (g->n_out > 0 && (!g->out_ts || !g->xs_out))) return fail(-5, "timesteps/tab/x_init/out_ts/xs_out must be non-NULL");
    if ((g->dim & 3) == 0 && g->n_out > 0 && (reinterpret_cast<uintptr_t>(g->xs_out) & 15)) return fail(-5, "xs_out must be 16-byte aligned");
    if (g->batch == 0) return 0;
    cudaError_t e = launch_affine_integrate(*g, reinterpret_cast<cudaStream_t>(stream_));
    if (e != cudaSuccess) return fail(-7, "affine integrator launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_expectations(const float* xs, int64_t n_rows, int32_t dim, double* out4, void* stream_) {
    g_err[0] = 0;
    if (!xs || !out4) return fail(-5, "xs/out4 NULL");
    if (n_rows < 0 || dim < 1) return fail(-3, "n_rows >= 0 and dim >= 1 required");
    cudaError_t e = launch_expectations(xs, n_rows, dim, out4, reinterpret_cast<cudaStream_t>(stream_));
    if (e != cudaSuccess) return fail(-7, "expectations launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_rnd_stats(const float* rnd, int64_t batch, int mask_mode, float max_rnd, const uint8_t* sample_mask,
                   double* out_stats, void* stream_) {
    g_err[0] = 0;
    if (!rnd || !out_stats) return fail(-5, "rnd/out_stats NULL");
    if (mask_mode < 0 || mask_mode > 2) return fail(-3, "mask_mode must be 0 (isfinite), 1 (< max_rnd) or 2 (all)");
    rnd_stats_kernel<<<1, 1024, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(rnd, batch, mask_mode, max_rnd, sample_mask, out_stats);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "rnd_stats launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_merge_stats(const double* gathered, int32_t world, double* out_stats, void* stream_) {
    g_err[0] = 0;
    if (!gathered || !out_stats) return fail(-5, "gathered/out_stats NULL");
    if (world < 1) return fail(-3, "world must be >= 1");
    merge_stats_kernel<<<1, 32, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(gathered, world, out_stats);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "merge_stats launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_weights(const float* rnd, int64_t batch, const double* stats, float* weights, void* stream_) {
    g_err[0] = 0;
    if (!rnd || !stats || !weights) return fail(-5, "rnd/stats/weights NULL");
    if (batch == 0) return 0;
    const int blocks = (int)((batch + 255) / 256 < 1184 ? (batch + 255) / 256 : 1184);
    weights_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(rnd, batch, stats, weights);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "weights launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_lv_weights(const float* rnd, int64_t batch, int mask_mode, float max_rnd, const uint8_t* sample_mask, const double* stats,
                    const float* upstream, float* w, void* stream_) {
    g_err[0] = 0;
    if (!rnd || !stats || !w) return fail(-5, "rnd/stats/w NULL");
    if (mask_mode < 0 || mask_mode > 2) return fail(-3, "mask_mode must be 0, 1 or 2");
    if (batch == 0) return 0;
    const int blocks = (int)((batch + 255) / 256 < 1184 ? (batch + 255) / 256 : 1184);
    lv_weights_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(rnd, batch, mask_mode, max_rnd, sample_mask, stats, upstream, w);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "lv weights launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_lv_traj_weights(const float* rnd, int64_t n_samples, int32_t traj_per_sample, int mask_mode, float max_rnd,
                         const uint8_t* sample_mask, const double* out3, const float* upstream, float* w, void* stream_) {
    g_err[0] = 0;
    if (!rnd || !out3 || !w) return fail(-5, "rnd/out3/w NULL");
    if (traj_per_sample < 2) return fail(-3, "Cannot compute variance over a single trajectory.");
    if (mask_mode < 0 || mask_mode > 2) return fail(-3, "mask_mode must be 0, 1 or 2");
    if (n_samples == 0) return 0;
    const int blocks = (int)((n_samples + 255) / 256 < 592 ? (n_samples + 255) / 256 : 592);
    lv_traj_weights_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(rnd, n_samples, traj_per_sample, mask_mode, max_rnd,
                                                                                        sample_mask, out3, upstream, w);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "lv_traj weights launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_kl_weights(const float* rnd, int64_t batch, int mask_mode, float max_rnd, const uint8_t* sample_mask, const double* stats,
                    const float* upstream, float* w, void* stream_) {
    g_err[0] = 0;
    if (!rnd || !stats || !w) return fail(-5, "rnd/stats/w NULL");
    if (mask_mode < 0 || mask_mode > 2) return fail(-3, "mask_mode must be 0, 1 or 2");
    if (batch == 0) return 0;
    const int blocks = (int)((batch + 255) / 256 < 1184 ? (batch + 255) / 256 : 1184);
    kl_weights_kernel<<<blocks, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(rnd, batch, mask_mode, max_rnd, sample_mask, stats, upstream, w);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "kl weights launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_sample_gauss_prior(float* out, int64_t batch, int32_t dim, float mean, float std, int32_t truncated, float a, float b,
                            uint64_t seed, uint64_t traj_offset, const float* uniforms, void* stream_) {
    g_err[0] = 0;
    if (!out) return fail(-5, "out is NULL");
    if (batch < 0 || dim < 1) return fail(-3, "batch must be >= 0 and dim >= 1");
    if (!(std > 0.f)) return fail(-3, "std must be positive");
    if (truncated && !(a < b)) return fail(-3, "truncation bounds must satisfy a < b");
    if (traj_offset + (uint64_t)batch > 0xFFFFFFFFull) return fail(-3, "traj_offset + batch exceeds the 32-bit Philox trajectory counter");
    if (batch == 0) return 0;
    cudaError_t e = launch_sample_prior(out, uniforms, batch, dim, mean, std, truncated, a, b, seed, traj_offset, reinterpret_cast<cudaStream_t>(stream_));
    if (e != cudaSuccess) return fail(-7, "prior sampling launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

size_t sdes_trainer_workspace_bytes(void) { return trainer_workspace_bytes(); }

int sdes_trainer_step(const SdesTrainerStepDesc* d, void* stream_) {
    g_err[0] = 0;
    if (d == nullptr || d->struct_bytes != sizeof(SdesTrainerStepDesc)) return fail(-2, "SdesTrainerStepDesc is NULL or has the wrong struct_bytes");
    if (d->n < 0) return fail(-3, "n < 0");
    if (!d->params || !d->grads || !d->exp_avg || !d->exp_avg_sq || !d->state || !d->workspace) return fail(-5, "params/grads/exp_avg/exp_avg_sq/state/workspace must be non-NULL");
    if (d->workspace_bytes < trainer_workspace_bytes()) return fail(-6, "workspace_bytes=%zu < required %zu", d->workspace_bytes, trainer_workspace_bytes());
    if (!(d->beta1 >= 0.f && d->beta1 < 1.f) || !(d->beta2 >= 0.f && d->beta2 < 1.f)) return fail(-3, "Invalid beta parameter");
    if (!(d->lr >= 0.f) || !(d->eps >= 0.f) || !(d->weight_decay >= 0.f)) return fail(-3, "Invalid lr / eps / weight_decay");
    if (d->ema_shadow != nullptr && d->ema_update_every < 1) return fail(-3, "ema_update_every must be >= 1");
    if (d->n == 0) return 0;
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaMemsetAsync(reinterpret_cast<uint8_t*>(d->workspace) + trainer_workspace_bytes() - 256, 0, 256, stream);
    if (e == cudaSuccess) e = launch_trainer_step(*d, stream);
    if (e != cudaSuccess) return fail(-7, "trainer step launch failed: %s", cudaGetErrorString(e));
    g_launches += 2;
    return 0;
}

int sdes_eval_moments(const float* samples, const float* weights, int64_t batch, int32_t dim, double* out, void* stream_) {
    g_err[0] = 0;
    if (!samples || !out) return fail(-5, "samples/out NULL");
    if (batch < 0 || dim < 1) return fail(-3, "batch must be >= 0 and dim >= 1");
    cudaError_t e = launch_eval_moments(samples, weights, batch, dim, out, reinterpret_cast<cudaStream_t>(stream_));
    if (e != cudaSuccess) return fail(-7, "eval moments launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_lv_traj_stats(const float* rnd, int64_t n_samples, int32_t traj_per_sample, int mask_mode, float max_rnd,
                       const uint8_t* sample_mask, double* out3, void* stream_) {
    g_err[0] = 0;
    if (!rnd || !out3) return fail(-5, "rnd/out NULL");
    if (traj_per_sample < 2) return fail(-3, "Cannot compute variance over a single trajectory.");
    if (mask_mode < 0 || mask_mode > 2) return fail(-3, "mask_mode must be 0, 1 or 2");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    cudaError_t e = cudaMemsetAsync(out3, 0, 3 * sizeof(double), stream);
    if (e != cudaSuccess) return fail(-7, "memset failed: %s", cudaGetErrorString(e));
    if (n_samples == 0) return 0;
    const int blocks = (int)((n_samples + 255) / 256 < 592 ? (n_samples + 255) / 256 : 592);
    lv_traj_stats_kernel<<<blocks, 256, 0, stream>>>(rnd, n_samples, traj_per_sample, mask_mode, max_rnd, sample_mask, out3);
    e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "lv_traj stats launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_tcgen05_selftest(const float* a, const float* w, float* d, int32_t k, int32_t n, int32_t mode, void* stream_) {
    g_err[0] = 0;
    if (!a || !w || !d) return fail(-5, "a/w/d NULL");
    if (k < 8 || k > 64 || k % 8 || n < 16 || n > 64 || n % 16) return fail(-3, "k must be a multiple of 8 in [8,64], n a multiple of 16 in [16,64]");
    if (mode != 0) return fail(-3, "mode must be 0 (bf16 hi/lo split, three kind::f16 passes: the rollout's layer)");
    cudaError_t e = launch_mma_selftest(a, w, d, k, n, reinterpret_cast<cudaStream_t>(stream_));
    if (e != cudaSuccess) return fail(-7, "selftest launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_gelu_probe(const float* x, float* y, int64_t n, void* stream_) {
    g_err[0] = 0;
    if (!x || !y || n <= 0) return fail(-5, "x/y NULL or n <= 0");
    gelu_probe_kernel<<<592, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(x, y, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "gelu probe launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_gelu_pair_probe(const float* x, float* y, int64_t n, void* stream_) {
    g_err[0] = 0;
    if (!x || !y || n <= 0 || (n & 1)) return fail(-5, "x/y NULL or n not a positive even number");
    if ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 7) return fail(-5, "x/y must be 8-byte aligned");
    gelu_pair_probe_kernel<<<592, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(reinterpret_cast<const float2*>(x), reinterpret_cast<float2*>(y), n / 2);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "gelu pair probe launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

int sdes_philox_normal(uint64_t seed, uint64_t traj_offset, int64_t batch, int32_t n_steps, int32_t dim, float* out,
                       void* stream_) {
    g_err[0] = 0;
    if (!out) return fail(-5, "out NULL");
    if (batch <= 0 || n_steps <= 0 || dim <= 0) return fail(-3, "batch, n_steps, dim must be positive");
    philox_normal_kernel<<<592, 256, 0, reinterpret_cast<cudaStream_t>(stream_)>>>(seed, traj_offset, batch, n_steps, dim, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(-7, "philox launch failed: %s", cudaGetErrorString(e));
    g_launches++;
    return 0;
}

}  // extern "C"
