// sdes_wide.cu — the WIDE engine: rollouts whose state does not fit one thread's registers (d > 64) or whose
// target is a NICE flow (BASELINE cfg5: ExponentialIntegratorSDELoss + ScoreCtrl on nice/mnist, distr/nice.py).
//
// The fused single-kernel engines keep a trajectory's state in registers and the control MLP's weights in shared
// memory.  Here neither fits: at d = 784 a NICE score needs 48 Linear layers of up to 1000 x 1000 per step
// (forward + input-gradient backward of 4 couplings, 76 MB of weights), so the step is a sequence of launches
//     control MLP (4 GEMMs)  ->  NICE forward (couplings x layers GEMMs)  ->  latent kernel
//     ->  NICE backward (couplings x layers GEMMs)  ->  fused update kernel
// on the caller's stream, with the state, activations and gradients in HBM/L2.  Every Linear is ONE launch of
// `linear_mma_kernel`: a warp-specialised tcgen05 GEMM (TMA bulk-copy producer warp, single-thread MMA issuer,
// 4 epilogue warps reading the fp32 accumulator from TMEM) with the layer's whole epilogue fused: bias, ReLU /
// exact GELU / ReLU-backward mask, residual add (the additive coupling), fp32 store and — the point of the
// layout — the next layer's A operand written directly as a tensor-core-ready image.
//
// Operand images.  fp32 accuracy from bf16 tensor cores: every operand is split v = hi + lo with hi = bf16(v),
// lo = bf16(v - hi) (16 significant bits) and a product is three kind::f16 MMAs  A_lo*W_hi + A_hi*W_lo + A_hi*W_hi
// with fp32 accumulation in TMEM (relative error ~2^-16 of sum|a||w|; parity emulation in DESIGN.md).  Images are
// stored exactly as the UMMA shared-memory operand wants them (K-major, no swizzle, 8 x 16-byte core matrices) in
// blocks of 128 rows x 64 K-elements: [hi 16 KB | lo 16 KB], so a pipeline stage is filled by plain
// `cp.async.bulk` copies of contiguous memory and the epilogue's 16-byte stores are warp-contiguous.
//
// Planar state.  NICE's couplings act on the even / odd units of x (Coupling.forward, distr/nice.py:64-95).  The
// engine keeps every feature vector de-interleaved, [even units | odd units], each half padded to a multiple of 64
// (width P = 2 Hp): a coupling's "off" input is then a contiguous range of K-chunks of the state image and its
// "on" output a contiguous column range.  Weights of the control MLP are permuted to match when they are imaged.
#include <cuda_bf16.h>

#include <cstdlib>
#include <vector>

#include "sdes_linear.cuh"
#include "sdes_step.cuh"
#include "sdes_timeembed.cuh"

namespace sdes {
namespace wide {

constexpr int ROWS_PER_CTA = 8;  // elementwise kernels: one warp per trajectory



constexpr int MAX_COUP = 8, MAX_NLIN = 8;

struct Plan {
    int d, Hp, P, pc;      // pc = P / 64 chunks of the planar width
    int64_t B, Bp;
    int m_tiles, T, nh;
    bool nice;
    int n_coup, n_lin, mid, Mp, mc;
    // tables / vectors (byte offsets)
    int64_t tab, emb, gate, vec_prior, vec_ref, vec_es, scalars, gmm_mu, gmm_h, gmm_c;
    Lin mlp_in, mlp_h[SDES_MAX_HIDDEN], mlp_out;
    Lin nf[MAX_COUP][MAX_NLIN], nb[MAX_COUP][MAX_NLIN];   // NICE forward / transposed (backward) operands
    int64_t xst, nn, h, g, rnd, logp;                      // fp32 state arrays
    int64_t ximg, gimg, m_img[2], act_img, d_img[2];       // operand images
    int64_t act_stride_layer, act_stride_coup;
    // keep mode (SDES_F_KEEP_FOR_GRAD): one state image per time step, a separate image for the NICE running state,
    // per-(step, trajectory) gate cotangent sums, and the scratch of the gradient pass (sdes_grad.cu)
    bool keep, keep_score;
    int64_t ximg_slot;     // bytes between the state images of consecutive steps (0 = one shared slot)
    int64_t himg, qgate, grad_base;
    int64_t sc_keep;       // keep_score (kl gradient): target score of step i at sc_keep + i * plane bytes, terminal state at slot T
    int64_t total;
};

static bool make_plan(const SdesRolloutDesc& d, Plan& p) {
    p.d = d.dim;
    p.Hp = round_up((d.dim + 1) / 2, 64);
    p.P = 2 * p.Hp;
    p.pc = p.P / 64;
    p.B = d.batch;
    p.Bp = (d.batch + 127) / 128 * 128;
    if (p.Bp == 0) p.Bp = 128;
    p.m_tiles = (int)(p.Bp / 128);
    p.T = d.n_steps;
    p.nh = d.n_hidden;
    p.nice = d.target_kind == SDES_TARGET_NICE;
    p.n_coup = p.nice ? d.nice_couplings : 0;
    p.n_lin = p.nice ? d.nice_hidden + 1 : 0;
    p.mid = p.nice ? d.nice_mid : 0;
    p.Mp = round_up(p.mid, 64);
    p.mc = d.nice_mask_config;
    int64_t o = 0;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align256(o + bytes); return r; };
    const int64_t K = d.target_kind == SDES_TARGET_GMM ? d.n_components : 0;
    p.tab = take((int64_t)p.T * TAB_STRIDE * 4);
    p.emb = take((int64_t)p.T * C * 4);
    p.gate = take((int64_t)p.T * 4);
    p.vec_prior = take(2ll * p.P * 4);
    p.vec_ref = take(2ll * p.P * 4);
    p.vec_es = take((int64_t)p.P * 4);
    p.scalars = take(8 * 4);
    p.gmm_mu = take(K * p.P * 4);
    p.gmm_h = take(K * p.P * 4);
    p.gmm_c = take(64 * 4);
    auto lin = [&](Lin& l, int N, int Kin, bool bias) {
        set_tiling(l, N, Kin);
        // fewer tiles than half the SMs (e.g. the N = d/2 = 392 layers of cfg 5's 4 096-row shard: 2 x 32): halve the
        // column tile so the layer spreads over twice as many CTAs
        while ((int64_t)l.n_tiles * p.m_tiles <= 74 && l.tile_n >= 128 && l.tile_n % 32 == 0) {
            l.tile_n /= 2;
            l.n_tiles *= 2;
        }
        l.w_off = take(lin_image_bytes(l));
        l.b_off = bias ? take((int64_t)l.n_pad * 4) : -1;
    };
    lin(p.mlp_in, C, p.P, false);  // bias = time-embedding row (+ b_in), per step
    for (int l = 0; l < p.nh; ++l) lin(p.mlp_h[l], C, C, true);
    lin(p.mlp_out, p.P, C, true);
    for (int c = 0; c < p.n_coup; ++c)
        for (int l = 0; l < p.n_lin; ++l) {
            const int n_out = l == p.n_lin - 1 ? p.Hp : p.mid, n_in = l == 0 ? p.Hp : p.mid;
            lin(p.nf[c][l], n_out, n_in, true);
            lin(p.nb[c][l], n_in, n_out, false);  // W^T: contraction over the layer's outputs
        }
    const int64_t plane = p.Bp * (int64_t)p.P * 4;
    p.xst = take(plane);
    p.nn = take(plane);
    p.h = take(p.nice ? plane : 0);
    p.g = take(plane);
    p.rnd = take(p.Bp * 4);
    p.logp = take(p.Bp * 4);
    const int64_t img_p = (int64_t)p.m_tiles * p.pc * A_BLOCK;
    p.keep = (d.flags & SDES_F_KEEP_FOR_GRAD) != 0;
    p.ximg_slot = p.keep ? img_p : 0;
    p.ximg = take(p.keep ? img_p * (p.T + 1) : img_p);
    p.himg = (p.keep && p.nice) ? take(img_p) : p.ximg;  // without keep the couplings work in place on the state image
    p.qgate = take(p.keep ? (int64_t)p.T * p.Bp * 4 : 0);
    p.keep_score = p.keep && (d.flags & SDES_F_KEEP_SCORE) != 0;
    p.sc_keep = take(p.keep_score ? plane * (p.T + 1) : 0);
    p.gimg = take(p.nice ? img_p : 0);
    p.m_img[0] = take((int64_t)p.m_tiles * A_BLOCK);
    p.m_img[1] = take((int64_t)p.m_tiles * A_BLOCK);
    p.act_stride_layer = (int64_t)p.m_tiles * (p.Mp / 64) * A_BLOCK;
    p.act_stride_coup = p.act_stride_layer * (p.n_lin > 0 ? p.n_lin - 1 : 0);
    p.act_img = take(p.act_stride_coup * p.n_coup);
    p.d_img[0] = take(p.act_stride_layer);
    p.d_img[1] = take(p.act_stride_layer);
    p.grad_base = o;
    p.total = o;
    return true;
}



// ------------------------------------------------------------------------ prologue kernel
struct PrepArgs {
    SdesRolloutDesc d;
    BlobLayout bl;
    int Hp, P;
    float *tab, *emb, *gate, *vec_prior, *vec_ref, *vec_es, *scalars, *gmm_mu, *gmm_h, *gmm_c;
    float* mlp_out_bias;                 // P (planar)
    float* mlp_h_bias[SDES_MAX_HIDDEN];  // 64 each
};

__global__ void __launch_bounds__(256) prepare_kernel(const PrepArgs a) {
    const SdesRolloutDesc& d = a.d;
    const float* blob = d.params;
    const int T = d.n_steps, dim = d.dim, tid = threadIdx.x;
    if ((int)blockIdx.x < T) {
        __shared__ float buf_a[2 * C], buf_b[2 * C], buf_out[C];
        const int i = blockIdx.x;
        const float s = d.ts[i];
        if (tid == 0) write_step_table_row(d, i, a.tab + (int64_t)i * TAB_STRIDE);
        time_embed_row(blob, a.bl.te_phase, a.bl.te_h_w, a.bl.te_h_b, d.te_hidden, a.bl.te_out_w, a.bl.te_out_b, C, s,
                       buf_a, buf_b, buf_out);
        if (tid < C) a.emb[(int64_t)i * C + tid] = buf_out[tid] + blob[a.bl.in_b + tid];
        __syncthreads();
        float gate = 1.0f;
        if (d.flags & SDES_F_HAS_GATE) {
            time_embed_row(blob, a.bl.g_phase, a.bl.g_h_w, a.bl.g_h_b, d.gate_hidden, a.bl.g_out_w, a.bl.g_out_b, 1, s,
                           buf_a, buf_b, buf_out);
            gate = clipf(buf_out[0], d.clip_model);
        }
        if (tid == 0) a.gate[i] = gate;
        return;
    }
    const int64_t nthreads = (int64_t)(gridDim.x - T) * blockDim.x;
    const int64_t gtid = (int64_t)(blockIdx.x - T) * blockDim.x + tid;
    const int Hp = a.Hp, P = a.P;
    // per-dimension vectors in planar order; padding = 0
    for (int64_t pi = gtid; pi < P; pi += nthreads) {
        const int j = to_natural((int)pi, 1, Hp);
        const bool ok = j < dim;
        float m = 0.f, iv = 0.f;
        if (ok && d.prior_loc != nullptr) { m = d.prior_loc[j]; const float sc = d.prior_scale[j]; iv = 1.0f / (sc * sc); }
        a.vec_prior[pi] = m; a.vec_prior[P + pi] = iv;
        m = 0.f; iv = 0.f;
        if (ok && d.ref_loc != nullptr) { m = d.ref_loc[j]; const float sc = d.ref_scale[j]; iv = 1.0f / (sc * sc); }
        a.vec_ref[pi] = m; a.vec_ref[P + pi] = iv;
        a.mlp_out_bias[pi] = ok ? blob[a.bl.out_b + j] : 0.f;
        if (d.target_kind == SDES_TARGET_NICE) a.vec_es[pi] = ok ? expf(d.nice_params[d.n_nice_params - dim + j]) : 0.f;
    }
    for (int l = 0; l < d.n_hidden; ++l)
        for (int64_t e = gtid; e < C; e += nthreads) a.mlp_h_bias[l][e] = blob[a.bl.h_b[l] + e];
    if (gtid == 0) {
        float lp = 0.f, lr = 0.f, ss = 0.f;
        if (d.prior_loc != nullptr) { for (int j = 0; j < dim; ++j) lp -= logf(d.prior_scale[j]); lp -= 0.5f * (float)dim * LOG_2PI; }
        if (d.ref_loc != nullptr) { for (int j = 0; j < dim; ++j) lr -= logf(d.ref_scale[j]); lr -= 0.5f * (float)dim * LOG_2PI; }
        if (d.target_kind == SDES_TARGET_NICE) for (int j = 0; j < dim; ++j) ss += d.nice_params[d.n_nice_params - dim + j];
        a.scalars[0] = lp; a.scalars[1] = lr; a.scalars[2] = ss;
    }
    if (d.target_kind == SDES_TARGET_GMM) {
        const int K = d.n_components;
        for (int64_t e = gtid; e < (int64_t)K * P; e += nthreads) {
            const int k = (int)(e / P), j = to_natural((int)(e % P), 1, Hp);
            float mu = 0.f, h = 0.f;
            if (j < dim) { mu = d.gmm_loc[(int64_t)k * dim + j]; const float sc = d.gmm_scale[(int64_t)k * dim + j]; h = 0.5f / (sc * sc); }
            a.gmm_mu[e] = mu; a.gmm_h[e] = h;
        }
        for (int64_t k = gtid; k < 64; k += nthreads) {
            float c = -INFINITY;
            if (k < K) {
                float logw = 0.f;
                if (d.gmm_weights != nullptr) { float tot = 0.f; for (int q = 0; q < K; ++q) tot += d.gmm_weights[q]; logw = logf(d.gmm_weights[k] / tot); }
                float sl = 0.f;
                for (int j = 0; j < dim; ++j) sl += logf(d.gmm_scale[k * dim + j]);
                c = logw - sl - 0.5f * (float)dim * LOG_2PI;
            }
            a.gmm_c[k] = c;
        }
    }
}



// ---------------------------------------------------------------------- per-row kernels
struct RowArgs {
    SdesRolloutDesc d;
    int Hp, P, pc;
    int64_t Bp;
    const float *tab, *gate, *vec_prior, *vec_ref, *vec_es, *scalars, *gmm_mu, *gmm_h, *gmm_c;
    float *xst, *nn, *h, *g, *rnd, *logp;
    uint8_t *ximg, *gimg;   // ximg: where this kernel WRITES the state image (init: step 0; update: step + 1)
    float* qgate;           // keep mode: (T, Bp) gate cotangent sums, else NULL
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// write planar elements (k, k+1) of `plane` (0 = even units, 1 = odd units) of one row into a state image
__device__ __forceinline__ void img_store_pair(uint8_t* img, int pc, int Hp, int64_t row, int plane, int k, float v0, float v1) {
    const int mt = (int)(row >> 7), r = (int)(row & 127), kk = plane * Hp + k;
    uint32_t hi, lo;
    split_pair(v0, v1, hi, lo);
    uint8_t* o = img + (int64_t)mt * pc * A_BLOCK + img_group_offset(r, kk) + (kk & 7) * 2;
    *reinterpret_cast<uint32_t*>(o) = hi;
    *reinterpret_cast<uint32_t*>(o + A_HALF) = lo;
}

// x0 (B, d) -> planar fp32 state + state image (+ xs[0]); initial cost (losses/oc.py:168-172, :296, :410)
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) init_kernel(const RowArgs a) {
    const SdesRolloutDesc& d = a.d;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.Bp) return;
    const bool valid = row < d.batch;
    const int dim = d.dim, Hp = a.Hp;
    const bool want_prior = d.loss_kind == SDES_LOSS_TIME_REVERSAL && !(d.flags & SDES_F_RND0_ZERO);
    float acc = 0.f;
    float* xr = a.xst + row * a.P;
    for (int q = lane; q < Hp / 2; q += 32) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int j = 4 * q + e;
            v[e] = (valid && j < dim) ? d.x0[row * dim + j] : 0.f;
            if (valid && j < dim && (d.flags & SDES_F_RETURN_TRAJ)) d.xs[row * dim + j] = v[e];
        }
        // natural (4q, 4q+1, 4q+2, 4q+3) -> even plane (2q, 2q+1) = (v0, v2), odd plane (2q, 2q+1) = (v1, v3)
        *reinterpret_cast<float2*>(xr + 2 * q) = make_float2(v[0], v[2]);
        *reinterpret_cast<float2*>(xr + Hp + 2 * q) = make_float2(v[1], v[3]);
        img_store_pair(a.ximg, a.pc, Hp, row, 0, 2 * q, v[0], v[2]);
        img_store_pair(a.ximg, a.pc, Hp, row, 1, 2 * q, v[1], v[3]);
        if (want_prior) {
            const float2 me = *reinterpret_cast<const float2*>(a.vec_prior + 2 * q), mo = *reinterpret_cast<const float2*>(a.vec_prior + Hp + 2 * q);
            const float2 ie = *reinterpret_cast<const float2*>(a.vec_prior + a.P + 2 * q), io = *reinterpret_cast<const float2*>(a.vec_prior + a.P + Hp + 2 * q);
            acc = fmaf((v[0] - me.x) * (v[0] - me.x), ie.x, acc);
            acc = fmaf((v[2] - me.y) * (v[2] - me.y), ie.y, acc);
            acc = fmaf((v[1] - mo.x) * (v[1] - mo.x), io.x, acc);
            acc = fmaf((v[3] - mo.y) * (v[3] - mo.y), io.y, acc);
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) a.rnd[row] = want_prior ? a.scalars[0] - 0.5f * acc : 0.f;
}

// z = h * exp(scale); log-density of the standard-logistic latent and the seed of the backward pass
// (distr/nice.py:21-29, :109-124, :178-190): d/dz [-(softplus(z) + softplus(-z))] = -tanh(z/2).
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) latent_kernel(const RowArgs a, const int want_grad) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.Bp) return;
    const float* hr = a.h + row * a.P;
    float* gr = a.g + row * a.P;
    float lp = 0.f;
    for (int k = 2 * lane; k < a.P; k += 64) {
        const float2 hv = *reinterpret_cast<const float2*>(hr + k), es = *reinterpret_cast<const float2*>(a.vec_es + k);
        const float z0 = hv.x * es.x, z1 = hv.y * es.y;
        // softplus(z) + softplus(-z) = |z| + 2 log1p(exp(-|z|)); padded units have exp(scale) = 0 and are masked out
        if (es.x != 0.f) lp -= fabsf(z0) + 2.0f * log1pf(expf(-fabsf(z0)));
        if (es.y != 0.f) lp -= fabsf(z1) + 2.0f * log1pf(expf(-fabsf(z1)));
        if (want_grad) {
            const float g0 = -tanhf(0.5f * z0) * es.x, g1 = -tanhf(0.5f * z1) * es.y;
            *reinterpret_cast<float2*>(gr + k) = make_float2(g0, g1);
            img_store_pair(a.gimg, a.pc, a.Hp, row, k >= a.Hp ? 1 : 0, k >= a.Hp ? k - a.Hp : k, g0, g1);
        }
    }
    lp = warp_sum(lp);
    if (lane == 0) a.logp[row] = lp + a.scalars[2] + a.d.log_norm_const;
}

// GMM / diagonal-Gaussian target on a wide state: log-density and score in planar order (distr/gauss.py:119-140;
// analytic score).  One warp per trajectory, components looped.
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) gmm_kernel(const RowArgs a, const int want_score) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.Bp) return;
    const int K = a.d.n_components, P = a.P;
    const float* xr = a.xst + row * P;
    float l0 = -INFINITY, l1 = -INFINITY;  // logits of components lane and lane + 32
    for (int k = 0; k < K; ++k) {
        float s = 0.f;
        for (int p = lane; p < P; p += 32) {
            const float df = xr[p] - a.gmm_mu[(int64_t)k * P + p];
            s = fmaf(df * df, a.gmm_h[(int64_t)k * P + p], s);
        }
        s = warp_sum(s);
        const float l = a.gmm_c[k] - s;
        if ((k & 31) == lane) { if (k < 32) l0 = l; else l1 = l; }
    }
    const float m = warp_max(fmaxf(l0, l1));
    const float e0 = expf(l0 - m), e1 = expf(l1 - m);
    const float ssum = warp_sum(e0 + e1);
    if (lane == 0) a.logp[row] = m + logf(ssum) + a.d.log_norm_const;
    if (!want_score) return;
    const float w0 = e0 / ssum, w1 = e1 / ssum;
    float* gr = a.g + row * P;
    for (int p0 = 0; p0 < P; p0 += 32) {
        const int p = p0 + lane;
        const float x = xr[p];
        float sc = 0.f;
        for (int k = 0; k < K; ++k) {
            const float wk = __shfl_sync(0xffffffffu, k < 32 ? w0 : w1, k & 31);
            sc = fmaf(wk * 2.0f * a.gmm_h[(int64_t)k * P + p], a.gmm_mu[(int64_t)k * P + p] - x, sc);
        }
        gr[p] = sc;
    }
}

// MultiWell (distr/double_well.py:165-179) and Funnel (distr/funnel.py:57-80) on a wide state: log-density and score in
// planar order, one warp per trajectory (the d <= 64 forms are sdes_step.cuh::multiwell_eval / funnel_eval).
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) analytic_target_kernel(const RowArgs a, const int want_score) {
    const SdesRolloutDesc& d = a.d;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= a.Bp) return;
    const int P = a.P, Hp = a.Hp, dim = d.dim;
    const float* xr = a.xst + row * P;
    float* gr = a.g + row * P;
    float lp = 0.f;
    if (d.target_kind == SDES_TARGET_MULTIWELL) {
        for (int p = lane; p < P; p += 32) {
            const int j = to_natural(p, 1, Hp);
            const float y = xr[p] - d.shift;
            float sc = 0.f;
            if (j < d.n_double_wells) {
                const float w = y * y - d.separation;
                lp -= w * w;
                sc = -4.0f * w * y;
            } else if (j < dim) {
                lp -= 0.5f * y * y;
                sc = -y;
            }
            if (want_score) gr[p] = sc;
        }
        lp = warp_sum(lp);
    } else {  // funnel: x_0 ~ N(0, var), x_j | x_0 ~ N(0, exp(x_0))
        float sq = 0.f;
        for (int p = lane; p < P; p += 32) {
            const int j = to_natural(p, 1, Hp);
            if (j >= 1 && j < dim) sq = fmaf(xr[p], xr[p], sq);
        }
        sq = warp_sum(sq);
        const float x0 = xr[0], inv = expf(-x0), dm1 = (float)(dim - 1);
        lp = -0.5f * logf(2.0f * 3.14159265358979323846f * d.variance) - 0.5f * x0 * x0 / d.variance - dm1 * (x0 + LOG_2PI) * 0.5f - 0.5f * sq * inv;
        if (want_score) {
            for (int p = lane; p < P; p += 32) {
                const int j = to_natural(p, 1, Hp);
                float sc = 0.f;
                if (j == 0) sc = -x0 / d.variance - 0.5f * dm1 + 0.5f * sq * inv;
                else if (j < dim) sc = -xr[p] * inv;
                gr[p] = sc;
            }
        }
    }
    if (lane == 0) a.logp[row] = lp + d.log_norm_const;
}

// One time step for every trajectory: control assembly (models/reparam.py), noise, running cost / Ito sums and
// the Euler-Maruyama / exponential-integrator update (losses/oc.py:204-219, :316-331, :429-443) — the wide-state
// form of sdes_step.cuh::update4.  One warp per trajectory; a lane handles natural dims 4q..4q+3 (= one Philox
// call) which are planar elements (2q, 2q+1) of both halves.
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) update_kernel(const RowArgs a, const int step) {
    const SdesRolloutDesc& d = a.d;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= d.batch) return;
    const int dim = d.dim, Hp = a.Hp, P = a.P;
    const float* tab = a.tab + (int64_t)step * TAB_STRIDE;
    const StepCoef c = make_step_coef(d, tab);
    const float gate = a.gate[step], lerp_w = tab[TAB_LERP_W], cs = d.clip_score;
    const float outer = (d.ctrl_kind == SDES_CTRL_SCORE ? 1.0f : c.sigma) * d.scale_score;
    const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)row);
    float* xr = a.xst + row * P;
    const float* nr = a.nn + row * P;
    const float* gr = a.g + row * P;
    const float* noise = c.from_hbm ? d.noise + ((int64_t)step * d.batch + row) * dim : nullptr;
    float* xs_out = (d.flags & SDES_F_RETURN_TRAJ) ? d.xs + ((int64_t)(step + 1) * d.batch + row) * dim : nullptr;
    float cost = 0.f, ito = 0.f, qsum = 0.f;
    for (int q = lane; 4 * q < dim; q += 32) {
        float e[4];
        if (c.from_hbm) {
#pragma unroll
            for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? noise[4 * q + r] : 0.f;
        } else {
            const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)step, (uint32_t)q);
            e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
        }
        const float2 xe = *reinterpret_cast<const float2*>(xr + 2 * q), xo = *reinterpret_cast<const float2*>(xr + Hp + 2 * q);
        const float2 ne = *reinterpret_cast<const float2*>(nr + 2 * q), no = *reinterpret_cast<const float2*>(nr + Hp + 2 * q);
        float2 se = make_float2(0.f, 0.f), so = se;
        if (d.ctrl_kind != SDES_CTRL_CLIPPED && d.ctrl_kind != SDES_CTRL_LERP_PRIOR) {
            se = *reinterpret_cast<const float2*>(gr + 2 * q);
            so = *reinterpret_cast<const float2*>(gr + Hp + 2 * q);
        }
        const float2 ple = *reinterpret_cast<const float2*>(a.vec_prior + 2 * q), plo = *reinterpret_cast<const float2*>(a.vec_prior + Hp + 2 * q);
        const float2 pie = *reinterpret_cast<const float2*>(a.vec_prior + P + 2 * q), pio = *reinterpret_cast<const float2*>(a.vec_prior + P + Hp + 2 * q);
        // natural order within the quad: (even.x, odd.x, even.y, odd.y)
        float x4[4] = {xe.x, xo.x, xe.y, xo.y};
        const float nn4[4] = {ne.x, no.x, ne.y, no.y}, sc4[4] = {se.x, so.x, se.y, so.y};
        const float pl4[4] = {ple.x, plo.x, ple.y, plo.y}, pi4[4] = {pie.x, pio.x, pie.y, pio.y};
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            if (4 * q + r >= dim) { x4[r] = 0.f; continue; }
            const float pscore = (pl4[r] - x4[r]) * pi4[r];
            float part = 0.f;
            if (d.ctrl_kind == SDES_CTRL_SCORE) part = outer * (clipf(sc4[r], cs) * gate);
            else if (d.ctrl_kind == SDES_CTRL_LERP) part = outer * (clipf(torch_lerp(pscore, sc4[r], lerp_w), cs) * gate);
            else if (d.ctrl_kind == SDES_CTRL_LERP_PRIOR) part = outer * (clipf((1.0f - lerp_w) * pscore, cs) * gate);
            else if (d.ctrl_kind == SDES_CTRL_LERP_TARGET) part = outer * (clipf(lerp_w * sc4[r], cs) * gate);
            const float g = clipf(nn4[r], c.cm) + part;
            // d rnd / d gate: the Ito coefficient times the un-gated score part (sdes_grad.cu); gate != 0 assumed
            if (a.qgate != nullptr) qsum = fmaf((c.exp_int ? c.sg * c.beta_k : c.sqrt_dt) * e[r], part / gate, qsum);
            if (c.exp_int) {
                cost = fmaf(g, g, cost);
                ito = fmaf(c.sg * g * e[r], c.beta_k, ito);
                x4[r] = x4[r] * c.alpha_k + c.bb_ss * g + c.s_bk * e[r];
            } else {
                float gm = g;
                if (c.ref_ctrl) gm -= c.sigma * pscore;
                const float db = e[r] * c.sqrt_dt;
                cost = fmaf(gm, gm, cost);
                ito = fmaf(gm, db, ito);
                x4[r] = x4[r] + (c.mu * x4[r] + c.sigma * g) * c.dt + c.sigma * db;
            }
            if (xs_out != nullptr) xs_out[4 * q + r] = x4[r];
        }
        *reinterpret_cast<float2*>(xr + 2 * q) = make_float2(x4[0], x4[2]);
        *reinterpret_cast<float2*>(xr + Hp + 2 * q) = make_float2(x4[1], x4[3]);
        img_store_pair(a.ximg, a.pc, Hp, row, 0, 2 * q, x4[0], x4[2]);
        img_store_pair(a.ximg, a.pc, Hp, row, 1, 2 * q, x4[1], x4[3]);
    }
    cost = warp_sum(cost);
    ito = warp_sum(ito);
    if (a.qgate != nullptr) {
        qsum = warp_sum(qsum);
        if (lane == 0) a.qgate[(int64_t)step * a.Bp + row] = qsum;
    }
    if (lane == 0) {
        float rnd = a.rnd[row];
        finish_step(d, c, tab, cost, ito, rnd);
        a.rnd[row] = rnd;
    }
}

// terminal cost (losses/oc.py:225, :337, :449-450; clip: solver/oc.py:48-54) and the caller's outputs
__global__ void __launch_bounds__(32 * ROWS_PER_CTA) terminal_kernel(const RowArgs a) {
    const SdesRolloutDesc& d = a.d;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * ROWS_PER_CTA + (threadIdx.x >> 5);
    if (row >= d.batch) return;
    const int dim = d.dim, Hp = a.Hp, P = a.P;
    const float* xr = a.xst + row * P;
    float acc = 0.f;
    for (int j = lane; j < dim; j += 32) {
        const int p = (j & 1) * Hp + (j >> 1);
        const float x = xr[p];
        d.x_T[row * dim + j] = x;
        const float y = x - a.vec_ref[p];
        acc = fmaf(y * y, a.vec_ref[P + p], acc);
    }
    acc = warp_sum(acc);
    if (lane == 0) {
        const float lp = clipf(a.logp[row], d.clip_target);
        float term = -lp;
        if (d.loss_kind != SDES_LOSS_TIME_REVERSAL) term += a.scalars[1] - 0.5f * acc;
        d.rnd[row] = a.rnd[row] + term;
    }
}

}  // namespace wide

// ------------------------------------------------------------------------------ host side
using namespace wide;

bool wide_engine_needed(const SdesRolloutDesc& d) { return d.dim > SDES_MAX_DIM || d.target_kind == SDES_TARGET_NICE; }

int64_t wide_nice_param_count(const SdesRolloutDesc& d) {
    const int64_t half = d.dim / 2, mid = d.nice_mid;
    const int64_t per = mid * half + mid + (int64_t)(d.nice_hidden - 1) * (mid * mid + mid) + half * mid + half;
    return per * d.nice_couplings + d.dim;
}

const char* wide_validate(const SdesRolloutDesc& d) {
    if (d.dim > SDES_MAX_WIDE_DIM) return "dim exceeds SDES_MAX_WIDE_DIM";
    if ((d.flags & SDES_F_HAS_GATE) && d.gate_dim != 1) return "the wide engine supports a scalar gate only (gate_dim = 1)";
    if (d.target_kind == SDES_TARGET_NICE) {
        if (d.dim % 2) return "NICE needs an even dim";
        if (d.nice_couplings < 1 || d.nice_couplings > MAX_COUP) return "nice_couplings not in [1,8]";
        if (d.nice_hidden < 1 || d.nice_hidden + 1 > MAX_NLIN) return "nice_hidden not in [1,7]";
        if (d.nice_mid < 1 || d.nice_mid > 4096) return "nice_mid not in [1,4096]";
        if (d.n_nice_params != wide_nice_param_count(d)) return "n_nice_params does not match the NICE layout";
    }
    return nullptr;
}

void wide_grad_view(const SdesRolloutDesc& d, wide::WideGradView& v) {
    Plan p;
    make_plan(d, p);
    v.d = p.d; v.Hp = p.Hp; v.P = p.P; v.pc = p.pc; v.T = p.T; v.nh = p.nh; v.m_tiles = p.m_tiles; v.B = p.B; v.Bp = p.Bp;
    v.tab = p.tab; v.emb = p.emb; v.gate = p.gate; v.ximg = p.ximg; v.ximg_slot = p.ximg_slot; v.qgate = p.qgate; v.grad_base = p.grad_base;
    v.vec_prior = p.vec_prior; v.vec_ref = p.vec_ref; v.gmm_h = p.gmm_h; v.xst = p.xst; v.logp = p.logp; v.sc_keep = p.sc_keep;
    v.keep_score = p.keep_score;
    v.mlp_in = p.mlp_in; v.mlp_out = p.mlp_out;
    for (int l = 0; l < SDES_MAX_HIDDEN; ++l) v.mlp_h[l] = p.mlp_h[l];
}

int64_t lv_grad_wide_scratch_bytes(const wide::WideGradView& v, bool bptt);  // sdes_grad.cu

size_t wide_workspace_bytes(const SdesRolloutDesc& d) {
    Plan p;
    make_plan(d, p);
    if (p.keep) {
        wide::WideGradView v;
        wide_grad_view(d, v);
        return (size_t)(p.grad_base + lv_grad_wide_scratch_bytes(v, p.keep_score));
    }
    return (size_t)p.total;
}


int64_t launch_rollout_wide(const KParams& kp, cudaStream_t stream, cudaError_t* err) {
    const SdesRolloutDesc& d = kp.d;
    Plan p;
    make_plan(d, p);
    uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
    auto F = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
    const bool simt = (d.flags & SDES_F_MLP_SIMT) != 0;
    int64_t launches = 0;
    *err = cudaSuccess;
#define WIDE_CHECK(expr)                          \
    do {                                          \
        *err = (expr);                            \
        if (*err != cudaSuccess) return launches; \
    } while (0)

    // ---- prologue: tables, vectors, operand images of every weight matrix
    PrepArgs pa;
    pa.d = d; pa.bl = kp.bl; pa.Hp = p.Hp; pa.P = p.P;
    pa.tab = F(p.tab); pa.emb = F(p.emb); pa.gate = F(p.gate); pa.vec_prior = F(p.vec_prior); pa.vec_ref = F(p.vec_ref);
    pa.vec_es = F(p.vec_es); pa.scalars = F(p.scalars); pa.gmm_mu = F(p.gmm_mu); pa.gmm_h = F(p.gmm_h); pa.gmm_c = F(p.gmm_c);
    pa.mlp_out_bias = F(p.mlp_out.b_off);
    for (int l = 0; l < SDES_MAX_HIDDEN; ++l) pa.mlp_h_bias[l] = l < p.nh ? F(p.mlp_h[l].b_off) : nullptr;
    wide::prepare_kernel<<<p.T + 16, 256, 0, stream>>>(pa);
    ++launches;
    WIDE_CHECK(cudaGetLastError());
    auto image = [&](const Lin& l, const float* src, int src_ld, int N, int K, int transpose, int n_planar, int k_planar) {
        ImgArgs ia;
        ia.src = src; ia.src_ld = src_ld; ia.N = N; ia.K = K; ia.transpose = transpose; ia.n_planar = n_planar; ia.k_planar = k_planar;
        ia.Hp = p.Hp; ia.out = ws + l.w_off; ia.n_pad = l.n_pad; ia.tile_n = l.tile_n; ia.k_chunks = l.k_chunks;
        const int64_t groups = (int64_t)l.n_pad * l.k_chunks * 8;
        int blocks = (int)((groups + 255) / 256);
        if (blocks > 148 * 8) blocks = 148 * 8;
        weight_image_kernel<<<blocks, 256, 0, stream>>>(ia);
        ++launches;
        return cudaGetLastError();
    };
    const float* blob = d.params;
    WIDE_CHECK(image(p.mlp_in, blob + kp.bl.in_w, d.dim, C, d.dim, 0, 0, 1));
    for (int l = 0; l < p.nh; ++l) WIDE_CHECK(image(p.mlp_h[l], blob + kp.bl.h_w[l], C, C, C, 0, 0, 0));
    WIDE_CHECK(image(p.mlp_out, blob + kp.bl.out_w, C, d.dim, C, 0, 1, 0));
    if (p.nice) {
        const int half = d.dim / 2, mid = p.mid;
        const float* np_ = d.nice_params;
        int64_t o = 0;
        for (int c = 0; c < p.n_coup; ++c)
            for (int l = 0; l < p.n_lin; ++l) {
                const int n_out = l == p.n_lin - 1 ? half : mid, n_in = l == 0 ? half : mid;
                WIDE_CHECK(image(p.nf[c][l], np_ + o, n_in, n_out, n_in, 0, 0, 0));
                WIDE_CHECK(image(p.nb[c][l], np_ + o, n_in, n_in, n_out, 1, 0, 0));
                o += (int64_t)n_out * n_in;
                pad_bias_kernel<<<(p.nf[c][l].n_pad + 255) / 256, 256, 0, stream>>>(np_ + o, n_out, F(p.nf[c][l].b_off), p.nf[c][l].n_pad);
                ++launches;
                WIDE_CHECK(cudaGetLastError());
                o += n_out;
            }
    }

    RowArgs ra;
    ra.d = d; ra.Hp = p.Hp; ra.P = p.P; ra.pc = p.pc; ra.Bp = p.Bp;
    ra.tab = F(p.tab); ra.gate = F(p.gate); ra.vec_prior = F(p.vec_prior); ra.vec_ref = F(p.vec_ref); ra.vec_es = F(p.vec_es);
    ra.scalars = F(p.scalars); ra.gmm_mu = F(p.gmm_mu); ra.gmm_h = F(p.gmm_h); ra.gmm_c = F(p.gmm_c);
    ra.xst = F(p.xst); ra.nn = F(p.nn); ra.h = F(p.h); ra.g = F(p.g); ra.rnd = F(p.rnd); ra.logp = F(p.logp);
    ra.ximg = ws + p.ximg; ra.gimg = ws + p.gimg;
    ra.qgate = (p.keep && (d.flags & SDES_F_HAS_GATE) && d.ctrl_kind != SDES_CTRL_CLIPPED) ? F(p.qgate) : nullptr;
    const int row_blocks_pad = (int)((p.Bp + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
    const int row_blocks = (int)((p.B + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
    // keep mode: every step has its own state image; padding rows / columns of the later slots must be finite zeros
    // (the gradient's wgrad GEMMs contract over ALL rows of a tile)
    if (p.keep) WIDE_CHECK(cudaMemsetAsync(ws + p.ximg + p.ximg_slot, 0, (size_t)p.ximg_slot * p.T, stream));
    init_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra);
    ++launches;
    WIDE_CHECK(cudaGetLastError());

    const int64_t x_stride = (int64_t)p.pc * A_BLOCK, m_stride = A_BLOCK, act_stride = (int64_t)(p.Mp / 64) * A_BLOCK;
    auto base_args = [&](const Lin& l) {
        LinArgs a;
        a.a_img = nullptr; a.a_mt_stride = 0; a.w_img = ws + l.w_off; a.k_chunks = l.k_chunks; a.n_tiles = l.n_tiles; a.tile_n = l.tile_n;
        a.bias = l.b_off >= 0 ? F(l.b_off) : nullptr; a.bias_mt_div = 0; a.bias_mt_stride = 0; a.act = ACT_NONE;
        a.mask_img = nullptr; a.mask_mt_stride = 0; a.mul_img = nullptr; a.mul_mt_stride = 0; a.aux_img = nullptr; a.resid = nullptr;
        a.out_f32 = nullptr; a.ld_f32 = p.P; a.out_img = nullptr; a.out_mt_stride = 0;
        return a;
    };
    // NICE forward over the couplings, in place on the state image (x's image is rebuilt by update_kernel each step)
    auto nice_forward = [&](const uint8_t* x_img) -> cudaError_t {
        uint8_t* h_img = p.keep ? ws + p.himg : const_cast<uint8_t*>(x_img);
        for (int c = 0; c < p.n_coup; ++c) {
            const int on = ((p.mc + c) % 2) ? 0 : 1, off = 1 - on;  // plane index: 0 = even units (distr/nice.py:79-82)
            for (int l = 0; l < p.n_lin; ++l) {
                LinArgs a = base_args(p.nf[c][l]);
                uint8_t* act_out = ws + p.act_img + c * p.act_stride_coup + l * p.act_stride_layer;
                if (l == 0) { a.a_img = (c == 0 ? x_img : h_img) + (int64_t)off * (p.Hp / 64) * A_BLOCK; a.a_mt_stride = x_stride; }
                else { a.a_img = act_out - p.act_stride_layer; a.a_mt_stride = act_stride; }
                if (l < p.n_lin - 1) { a.act = ACT_RELU; a.out_img = act_out; a.out_mt_stride = act_stride; }
                else {
                    // on <- on + shift: the first two couplings still read x's planes, later ones the running state h
                    a.resid = (c < 2 ? F(p.xst) : F(p.h)) + on * p.Hp;
                    a.out_f32 = F(p.h) + on * p.Hp;
                    a.out_img = h_img + (int64_t)on * (p.Hp / 64) * A_BLOCK;
                    a.out_mt_stride = x_stride;
                }
                cudaError_t e = launch_linear(a, p.m_tiles, simt, stream, launches);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    };
    // NICE input-gradient backward (the autograd score of distr/base.py:130-137): g = d log rho / d x, seeded by latent_kernel
    auto nice_backward = [&]() -> cudaError_t {
        for (int c = p.n_coup - 1; c >= 0; --c) {
            const int on = ((p.mc + c) % 2) ? 0 : 1, off = 1 - on;
            int cur = 0;
            for (int l = p.n_lin - 1; l >= 0; --l) {
                LinArgs a = base_args(p.nb[c][l]);
                if (l == p.n_lin - 1) { a.a_img = ws + p.gimg + (int64_t)on * (p.Hp / 64) * A_BLOCK; a.a_mt_stride = x_stride; }
                else { a.a_img = ws + p.d_img[cur]; a.a_mt_stride = act_stride; cur = 1 - cur; }
                if (l > 0) {
                    a.mask_img = ws + p.act_img + c * p.act_stride_coup + (l - 1) * p.act_stride_layer;
                    a.mask_mt_stride = act_stride;
                    a.out_img = ws + p.d_img[cur]; a.out_mt_stride = act_stride;
                } else {
                    a.resid = F(p.g) + off * p.Hp; a.out_f32 = F(p.g) + off * p.Hp;
                    a.out_img = ws + p.gimg + (int64_t)off * (p.Hp / 64) * A_BLOCK; a.out_mt_stride = x_stride;
                }
                cudaError_t e = launch_linear(a, p.m_tiles, simt, stream, launches);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    };
    const bool need_score = d.ctrl_kind != SDES_CTRL_CLIPPED && d.ctrl_kind != SDES_CTRL_LERP_PRIOR;
    const int64_t plane_bytes = p.Bp * (int64_t)p.P * 4;

    for (int i = 0; i < p.T; ++i) {
        const uint8_t* x_img = ws + p.ximg + (int64_t)i * p.ximg_slot;   // image of the state at step i
        ra.ximg = ws + p.ximg + (int64_t)(i + 1) * p.ximg_slot;           // where update_kernel writes step i + 1
        // ---- control MLP (models/mlp.py:114-122): x image -> 64 -> ... -> 64 -> nn (fp32, planar)
        {
            LinArgs a = base_args(p.mlp_in);
            a.a_img = x_img; a.a_mt_stride = x_stride; a.bias = F(p.emb) + (int64_t)i * C; a.act = ACT_GELU;
            a.out_img = ws + p.m_img[0]; a.out_mt_stride = m_stride;
            WIDE_CHECK(launch_linear(a, p.m_tiles, simt, stream, launches));
            int cur = 0;
            for (int l = 0; l < p.nh; ++l) {
                a = base_args(p.mlp_h[l]);
                a.a_img = ws + p.m_img[cur]; a.a_mt_stride = m_stride; a.act = ACT_GELU;
                a.out_img = ws + p.m_img[1 - cur]; a.out_mt_stride = m_stride;
                WIDE_CHECK(launch_linear(a, p.m_tiles, simt, stream, launches));
                cur = 1 - cur;
            }
            a = base_args(p.mlp_out);
            a.a_img = ws + p.m_img[cur]; a.a_mt_stride = m_stride; a.out_f32 = F(p.nn);
            WIDE_CHECK(launch_linear(a, p.m_tiles, simt, stream, launches));
        }
        // ---- target score
        if (need_score) {
            if (p.nice) {
                if (p.n_coup == 1) WIDE_CHECK(cudaMemcpyAsync(F(p.h), F(p.xst), p.Bp * (int64_t)p.P * 4, cudaMemcpyDeviceToDevice, stream));
                WIDE_CHECK(nice_forward(x_img));
                latent_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 1);
                ++launches;
                WIDE_CHECK(cudaGetLastError());
                WIDE_CHECK(nice_backward());
            } else {
                if (d.target_kind == SDES_TARGET_GMM) gmm_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 1);
                else analytic_target_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 1);
                ++launches;
                WIDE_CHECK(cudaGetLastError());
            }
        }
        if (p.keep_score && need_score)  // kl gradient: the score enters the adjoint as a stored constant / through its mask
            WIDE_CHECK(cudaMemcpyAsync(ws + p.sc_keep + (int64_t)i * plane_bytes, F(p.g), (size_t)plane_bytes, cudaMemcpyDeviceToDevice, stream));
        update_kernel<<<row_blocks, 32 * ROWS_PER_CTA, 0, stream>>>(ra, i);
        ++launches;
        WIDE_CHECK(cudaGetLastError());
    }
    // ---- terminal cost
    if (p.nice) {
        if (p.n_coup == 1) WIDE_CHECK(cudaMemcpyAsync(F(p.h), F(p.xst), p.Bp * (int64_t)p.P * 4, cudaMemcpyDeviceToDevice, stream));
        WIDE_CHECK(nice_forward(ws + p.ximg + (int64_t)p.T * p.ximg_slot));
        latent_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, p.keep_score ? 1 : 0);
    } else if (d.target_kind == SDES_TARGET_GMM) {
        gmm_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, p.keep_score ? 1 : 0);
    } else {
        analytic_target_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, p.keep_score ? 1 : 0);
    }
    ++launches;
    WIDE_CHECK(cudaGetLastError());
    if (p.keep_score) {  // the terminal cost -log rho(x_T) IS differentiated by the reference: keep grad log rho(x_T)
        if (p.nice) WIDE_CHECK(nice_backward());
        WIDE_CHECK(cudaMemcpyAsync(ws + p.sc_keep + (int64_t)p.T * plane_bytes, F(p.g), (size_t)plane_bytes, cudaMemcpyDeviceToDevice, stream));
    }
    terminal_kernel<<<row_blocks, 32 * ROWS_PER_CTA, 0, stream>>>(ra);
    ++launches;
    WIDE_CHECK(cudaGetLastError());
#undef WIDE_CHECK
    return launches;
}

}  // namespace sdes
