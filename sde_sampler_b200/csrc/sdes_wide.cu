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

#include <vector>

#include "sdes_step.cuh"
#include "sdes_tc.cuh"
#include "sdes_timeembed.cuh"

namespace sdes {
namespace wide {

constexpr int KC = 64;                       // K elements per image block / pipeline stage
constexpr uint32_t A_HALF = 128u * KC * 2u;  // bytes of one bf16 half (hi or lo) of an activation block
constexpr uint32_t A_BLOCK = 2u * A_HALF;    // hi | lo
constexpr int LIN_THREADS = 192;             // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: epilogue
constexpr int ROWS_PER_CTA = 8;              // elementwise kernels: one warp per trajectory

static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }
static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

// ------------------------------------------------------------------------------ the plan
struct Lin {           // one Linear as a B operand image: [n_tile][k_chunk] blocks of tile_n x 64 (hi | lo)
    int N, K;          // logical out / in features
    int n_pad, n_tiles, tile_n, k_chunks;
    int64_t w_off;     // bytes from the workspace base
    int64_t b_off;     // bytes from the workspace base of the padded fp32 bias (n_pad), -1 = none
};

static void set_tiling(Lin& l, int N, int K) {
    l.N = N;
    l.K = K;
    l.n_pad = round_up(N, 64);
    int nt = (l.n_pad + 255) / 256;
    while ((l.n_pad / 16) % nt) ++nt;
    l.n_tiles = nt;
    l.tile_n = l.n_pad / nt;
    l.k_chunks = round_up(K, 64) / 64;
}
static int64_t lin_image_bytes(const Lin& l) { return (int64_t)l.n_pad * l.k_chunks * 64 * 4; }

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
    p.ximg = take(img_p);
    p.gimg = take(p.nice ? img_p : 0);
    p.m_img[0] = take((int64_t)p.m_tiles * A_BLOCK);
    p.m_img[1] = take((int64_t)p.m_tiles * A_BLOCK);
    p.act_stride_layer = (int64_t)p.m_tiles * (p.Mp / 64) * A_BLOCK;
    p.act_stride_coup = p.act_stride_layer * (p.n_lin > 0 ? p.n_lin - 1 : 0);
    p.act_img = take(p.act_stride_coup * p.n_coup);
    p.d_img[0] = take(p.act_stride_layer);
    p.d_img[1] = take(p.act_stride_layer);
    p.total = o;
    return true;
}

// ---------------------------------------------------------------------------- bf16 split
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = tc::pack_bf16x2(a, b);  // a in the low half
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
    lo = tc::pack_bf16x2(a - ha, b - hb);
}
__device__ __forceinline__ void unpack8(const uint4& w, float (&f)[8]) {
    f[0] = __uint_as_float(w.x << 16); f[1] = __uint_as_float(w.x & 0xFFFF0000u);
    f[2] = __uint_as_float(w.y << 16); f[3] = __uint_as_float(w.y & 0xFFFF0000u);
    f[4] = __uint_as_float(w.z << 16); f[5] = __uint_as_float(w.z & 0xFFFF0000u);
    f[6] = __uint_as_float(w.w << 16); f[7] = __uint_as_float(w.w & 0xFFFF0000u);
}
// byte offset of the 16-byte group holding elements (row r, k..k+7) inside an activation image row of blocks
__device__ __forceinline__ int64_t img_group_offset(int r, int k) {
    return (int64_t)(k >> 6) * A_BLOCK + (uint32_t)((((k & 63) >> 3) * 128 + r) * 16);
}

// ------------------------------------------------------------------------- weight images
// out image element (n, k) = src[rn][rk] (or src[rk][rn] when transposed) with rn / rk the logical indices behind
// the padded image indices: identity, or the planar permutation p -> natural j = 2 (p % Hp) + p / Hp.
struct ImgArgs {
    const float* src;
    int src_ld;           // columns of the row-major source
    int N, K;             // logical extents of the image's n and k axes (natural index space)
    int transpose;        // 0: src[n][k], 1: src[k][n]
    int n_planar, k_planar, Hp;
    uint8_t* out;
    int n_pad, tile_n, k_chunks;
};

__device__ __forceinline__ int to_natural(int p, int planar, int Hp) { return planar ? 2 * (p % Hp) + p / Hp : p; }

__global__ void __launch_bounds__(256) weight_image_kernel(const ImgArgs a) {
    const int64_t groups = (int64_t)a.n_pad * a.k_chunks * 8;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < groups; e += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(e % a.n_pad);
        const int kg = (int)(e / a.n_pad);  // global 8-wide k group
        const int rn = to_natural(n, a.n_planar, a.Hp);
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int rk = to_natural(kg * 8 + q, a.k_planar, a.Hp);
            float w = 0.f;
            if (rn < a.N && rk < a.K) w = a.transpose ? a.src[(int64_t)rk * a.src_ld + rn] : a.src[(int64_t)rn * a.src_ld + rk];
            v[q] = w;
        }
        uint4 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x);
        split_pair(v[2], v[3], hi.y, lo.y);
        split_pair(v[4], v[5], hi.z, lo.z);
        split_pair(v[6], v[7], hi.w, lo.w);
        const int nt = n / a.tile_n, nl = n % a.tile_n, kc = kg >> 3, kl = kg & 7;
        const int64_t half = (int64_t)a.tile_n * 128;  // bytes of one bf16 half of a block
        uint8_t* blk = a.out + ((int64_t)nt * a.k_chunks + kc) * 2 * half;
        const int64_t off = ((int64_t)kl * a.tile_n + nl) * 16;
        *reinterpret_cast<uint4*>(blk + off) = hi;
        *reinterpret_cast<uint4*>(blk + half + off) = lo;
    }
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
        const float s = d.ts[i], t = d.ts[i + 1];
        const float dt = __fsub_rn(t, s);
        if (tid == 0) {  // identical to the fused engines' table (sdes_prepare.cu)
            float mu = 0.f, sigma = 0.f, div_int = 0.f, lerp_w = 0.f;
            if (d.sde_kind == SDES_SDE_VP) {
                const float ws_ = __fdiv_rn(s, d.terminal_t), wt_ = __fdiv_rn(t, d.terminal_t);
                const float b0 = d.sde_sign > 0.f ? d.beta_max : d.beta_min;
                const float b1 = d.sde_sign > 0.f ? d.beta_min : d.beta_max;
                const float beta_s = torch_lerp(b0, b1, ws_), beta_t = torch_lerp(b0, b1, wt_);
                mu = d.sde_sign * 0.5f * beta_s;
                sigma = d.scale_diff * sqrtf(beta_s);
                div_int = d.sde_sign * 0.25f * (beta_t + beta_s) * dt * (float)dim;
                lerp_w = ws_;
            } else if (d.sde_kind == SDES_SDE_CONST_OU) {
                mu = d.sde_sign * d.drift_coeff;
                sigma = d.diff_coeff;
                div_int = d.sde_sign * d.drift_coeff * dt * (float)dim;
                lerp_w = __fdiv_rn(s, d.terminal_t);
            }
            float beta_k = 0.f, alpha_k = 0.f;
            if (d.loss_kind == SDES_LOSS_EXP_INTEGRATOR) {
                beta_k = fminf(fmaxf(d.alpha * sqrtf(dt), 0.f), 1.f);
                alpha_k = sqrtf(1.0f - beta_k * beta_k);
            }
            float* row = a.tab + (int64_t)i * TAB_STRIDE;
            row[TAB_DT] = dt; row[TAB_SQRT_DT] = sqrtf(dt); row[TAB_MU] = mu; row[TAB_SIGMA] = sigma;
            row[TAB_DIV_INT] = div_int; row[TAB_LERP_W] = lerp_w; row[TAB_BETA_K] = beta_k; row[TAB_ALPHA_K] = alpha_k;
        }
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

__global__ void pad_bias_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pad; e += gridDim.x * blockDim.x) dst[e] = e < n ? src[e] : 0.f;
}

// ------------------------------------------------------------------------ the GEMM layer
struct LinArgs {
    const uint8_t* a_img; int64_t a_mt_stride;    // A operand: [m_tile] rows of k_chunks blocks
    const uint8_t* w_img;                          // B operand: [n_tile][k_chunk] blocks
    int k_chunks, n_tiles, tile_n;
    const float* bias;                             // n_pad floats or NULL
    int act;                                       // 0 none, 1 ReLU, 2 exact GELU
    const uint8_t* mask_img; int64_t mask_mt_stride;  // ReLU backward: keep where the stored activation > 0
    const float* resid;                            // out = resid + acc (additive coupling / gradient accumulation)
    float* out_f32; int ld_f32;                    // row-major fp32 output (same leading dimension as resid)
    uint8_t* out_img; int64_t out_mt_stride;       // next layer's A operand
};
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2 };

// 8 consecutive output columns of one row: the fused epilogue of every layer
__device__ __forceinline__ void epilogue8(const LinArgs& a, int mt, int r, int col, float (&v)[8]) {
    const int64_t row = (int64_t)mt * 128 + r;
    if (a.bias != nullptr) {
        const float4 b0 = *reinterpret_cast<const float4*>(a.bias + col), b1 = *reinterpret_cast<const float4*>(a.bias + col + 4);
        v[0] += b0.x; v[1] += b0.y; v[2] += b0.z; v[3] += b0.w; v[4] += b1.x; v[5] += b1.y; v[6] += b1.z; v[7] += b1.w;
    }
    if (a.resid != nullptr) {
        const float4* rp = reinterpret_cast<const float4*>(a.resid + row * a.ld_f32 + col);
        const float4 r0 = rp[0], r1 = rp[1];
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w; v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    }
    if (a.act == ACT_RELU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
    } else if (a.act == ACT_GELU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = gelu_fast(v[q]);
    }
    const int64_t goff = img_group_offset(r, col);
    if (a.mask_img != nullptr) {
        const uint4 m = *reinterpret_cast<const uint4*>(a.mask_img + (int64_t)mt * a.mask_mt_stride + goff);
        const uint32_t w[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const uint32_t bits = (w[q >> 1] >> (16 * (q & 1))) & 0xFFFFu;  // bf16 hi half of the forward activation
            const bool pos = (bits & 0x8000u) == 0u && (bits & 0x7FFFu) != 0u;
            v[q] = pos ? v[q] : 0.f;
        }
    }
    if (a.out_f32 != nullptr) {
        float4* op = reinterpret_cast<float4*>(a.out_f32 + row * a.ld_f32 + col);
        op[0] = make_float4(v[0], v[1], v[2], v[3]);
        op[1] = make_float4(v[4], v[5], v[6], v[7]);
    }
    if (a.out_img != nullptr) {
        uint4 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x);
        split_pair(v[2], v[3], hi.y, lo.y);
        split_pair(v[4], v[5], hi.z, lo.z);
        split_pair(v[6], v[7], hi.w, lo.w);
        uint8_t* o = a.out_img + (int64_t)mt * a.out_mt_stride + goff;
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + A_HALF) = lo;
    }
}

__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}

// D[128, tile_n] = A[128, K] W[tile_n, K]^T for one (n_tile, m_tile); grid (n_tiles, m_tiles): the CTAs that share
// an A tile run together so A is read from HBM once and from L2 afterwards.
__global__ void __launch_bounds__(LIN_THREADS, 1) linear_mma_kernel(const __grid_constant__ LinArgs a, const int stages) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t s_full[4], s_empty[4], s_acc;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nt = blockIdx.x, mt = blockIdx.y;
    const uint32_t b_half = (uint32_t)a.tile_n * 128u;           // bytes of one bf16 half of a weight block
    const uint32_t stage_bytes = A_BLOCK + 2u * b_half;
    uint32_t ncols = 32;
    while ((int)ncols < a.tile_n) ncols <<= 1;

    if (warp == 1) {
        tc::tmem_alloc(&s_tmem, ncols);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            tc::mbar_init(&s_full[s], 1);
            tc::mbar_init(&s_empty[s], 1);
        }
        tc::mbar_init(&s_acc, 1);
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_d = s_tmem;

    if (warp == 0) {
        if (lane == 0) {  // ---- producer: one elected thread drives the TMA bulk copies
            const uint8_t* a_src = a.a_img + (int64_t)mt * a.a_mt_stride;
            const uint8_t* w_src = a.w_img + (int64_t)nt * a.k_chunks * (2ll * b_half);
            for (int kc = 0; kc < a.k_chunks; ++kc) {
                const int s = kc % stages, it = kc / stages;
                if (it > 0) tc::mbar_wait(&s_empty[s], (uint32_t)((it - 1) & 1));
                tc::mbar_arrive_expect_tx(&s_full[s], stage_bytes);
                uint8_t* dst = smem + (size_t)s * stage_bytes;
                const uint8_t* ap = a_src + (int64_t)kc * A_BLOCK;
                tc::bulk_g2s(dst, ap, A_HALF, &s_full[s]);
                tc::bulk_g2s(dst + A_HALF, ap + A_HALF, A_HALF, &s_full[s]);
                const uint8_t* wp = w_src + (int64_t)kc * (2ll * b_half);
                for (uint32_t off = 0; off < 2u * b_half; off += 16384u) {
                    const uint32_t n = 2u * b_half - off < 16384u ? 2u * b_half - off : 16384u;
                    tc::bulk_g2s(dst + A_BLOCK + off, wp + off, n, &s_full[s]);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = tc::idesc_bf16(128, a.tile_n);
            const uint32_t lbo_b = (uint32_t)a.tile_n * 16u;
            for (int kc = 0; kc < a.k_chunks; ++kc) {
                const int s = kc % stages, it = kc / stages;
                tc::mbar_wait(&s_full[s], (uint32_t)(it & 1));
                tc::fence_after();
                const uint32_t a_hi = tc::smem_u32(smem + (size_t)s * stage_bytes), a_lo = a_hi + A_HALF;
                const uint32_t b_hi = a_hi + A_BLOCK, b_lo = b_hi + b_half;
#pragma unroll
                for (int ks = 0; ks < KC / 16; ++ks) {
                    const uint64_t dah = tc::smem_desc_kmajor(a_hi + (uint32_t)ks * 4096u, 2048u, 128u);
                    const uint64_t dal = tc::smem_desc_kmajor(a_lo + (uint32_t)ks * 4096u, 2048u, 128u);
                    const uint64_t dbh = tc::smem_desc_kmajor(b_hi + (uint32_t)ks * 2u * lbo_b, lbo_b, 128u);
                    const uint64_t dbl = tc::smem_desc_kmajor(b_lo + (uint32_t)ks * 2u * lbo_b, lbo_b, 128u);
                    mma_f16_ss(tmem_d, dal, dbh, idesc, (kc > 0 || ks > 0) ? 1u : 0u);  // small terms first
                    mma_f16_ss(tmem_d, dah, dbl, idesc, 1u);
                    mma_f16_ss(tmem_d, dah, dbh, idesc, 1u);
                }
                tc::mma_commit(&s_empty[s]);  // the stage is free once these MMAs have read it
            }
            tc::mma_commit(&s_acc);
        }
    } else {  // ---- epilogue warps: TMEM lane quadrant = warp % 4
        const int q = warp & 3, r = q * 32 + lane;
        tc::mbar_wait(&s_acc, 0);
        tc::fence_after();
        const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
        float v[8], w[8];
        tc::tmem_ld8(taddr, v);
        for (int c0 = 0; c0 < a.tile_n; c0 += 16) {
            tc::wait_ld_tie<8>(v);
            tc::tmem_ld8(taddr + (uint32_t)c0 + 8u, w);
            epilogue8(a, mt, r, nt * a.tile_n + c0, v);
            tc::wait_ld_tie<8>(w);
            if (c0 + 16 < a.tile_n) tc::tmem_ld8(taddr + (uint32_t)c0 + 16u, v);
            epilogue8(a, mt, r, nt * a.tile_n + c0 + 8, w);
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_d, ncols);
}

// The same layer on the CUDA cores, reading the same operand images: the cross-check engine (SDES_F_MLP_SIMT).
__global__ void __launch_bounds__(128) linear_simt_kernel(const LinArgs a) {
    const int nt = blockIdx.x, mt = blockIdx.y, r = threadIdx.x;
    const uint8_t* A = a.a_img + (int64_t)mt * a.a_mt_stride;
    const int64_t b_half = (int64_t)a.tile_n * 128;
    const uint8_t* W = a.w_img + (int64_t)nt * a.k_chunks * 2 * b_half;
    for (int c0 = 0; c0 < a.tile_n; c0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int kc = 0; kc < a.k_chunks; ++kc) {
            for (int kg = 0; kg < 8; ++kg) {
                const uint8_t* ap = A + (int64_t)kc * A_BLOCK + (kg * 128 + r) * 16;
                float ah[8], al[8];
                unpack8(*reinterpret_cast<const uint4*>(ap), ah);
                unpack8(*reinterpret_cast<const uint4*>(ap + A_HALF), al);
#pragma unroll
                for (int q = 0; q < 8; ++q) {
                    const uint8_t* wp = W + (int64_t)kc * 2 * b_half + ((int64_t)kg * a.tile_n + c0 + q) * 16;
                    float wh[8], wl[8];
                    unpack8(*reinterpret_cast<const uint4*>(wp), wh);
                    unpack8(*reinterpret_cast<const uint4*>(wp + b_half), wl);
                    float s = 0.f;
#pragma unroll
                    for (int e = 0; e < 8; ++e) s = fmaf(al[e], wh[e], fmaf(ah[e], wl[e], fmaf(ah[e], wh[e], s)));
                    acc[q] += s;
                }
            }
        }
        epilogue8(a, mt, r, nt * a.tile_n + c0, acc);
    }
}

// ---------------------------------------------------------------------- per-row kernels
struct RowArgs {
    SdesRolloutDesc d;
    int Hp, P, pc;
    int64_t Bp;
    const float *tab, *gate, *vec_prior, *vec_ref, *vec_es, *scalars, *gmm_mu, *gmm_h, *gmm_c;
    float *xst, *nn, *h, *g, *rnd, *logp;
    uint8_t *ximg, *gimg;
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
    float cost = 0.f, ito = 0.f;
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
    if (d.target_kind == SDES_TARGET_MULTIWELL || d.target_kind == SDES_TARGET_FUNNEL)
        return "MultiWell / Funnel targets are implemented on the fused engines only (d <= 64)";
    if (d.target_kind == SDES_TARGET_NICE) {
        if (d.dim % 2) return "NICE needs an even dim";
        if (d.nice_couplings < 1 || d.nice_couplings > MAX_COUP) return "nice_couplings not in [1,8]";
        if (d.nice_hidden < 1 || d.nice_hidden + 1 > MAX_NLIN) return "nice_hidden not in [1,7]";
        if (d.nice_mid < 1 || d.nice_mid > 4096) return "nice_mid not in [1,4096]";
        if (d.n_nice_params != wide_nice_param_count(d)) return "n_nice_params does not match the NICE layout";
    }
    return nullptr;
}

size_t wide_workspace_bytes(const SdesRolloutDesc& d) {
    Plan p;
    make_plan(d, p);
    return (size_t)p.total;
}

static cudaError_t launch_linear(const LinArgs& a, int m_tiles, bool simt, cudaStream_t stream, int64_t& launches) {
    ++launches;
    if (simt) {
        linear_simt_kernel<<<dim3(a.n_tiles, m_tiles), 128, 0, stream>>>(a);
        return cudaGetLastError();
    }
    const uint32_t stage_bytes = A_BLOCK + 2u * (uint32_t)a.tile_n * 128u;
    int stages = (int)(200u * 1024u / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 1) stages = 1;
    linear_mma_kernel<<<dim3(a.n_tiles, m_tiles), LIN_THREADS, (size_t)stages * stage_bytes, stream>>>(a, stages);
    return cudaGetLastError();
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
    static bool attr_set = false;
    if (!attr_set) {
        *err = cudaFuncSetAttribute(linear_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (*err != cudaSuccess) return launches;
        attr_set = true;
    }
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
    const int row_blocks_pad = (int)((p.Bp + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
    const int row_blocks = (int)((p.B + ROWS_PER_CTA - 1) / ROWS_PER_CTA);
    init_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra);
    ++launches;
    WIDE_CHECK(cudaGetLastError());

    const int64_t x_stride = (int64_t)p.pc * A_BLOCK, m_stride = A_BLOCK, act_stride = (int64_t)(p.Mp / 64) * A_BLOCK;
    auto base_args = [&](const Lin& l) {
        LinArgs a;
        a.a_img = nullptr; a.a_mt_stride = 0; a.w_img = ws + l.w_off; a.k_chunks = l.k_chunks; a.n_tiles = l.n_tiles; a.tile_n = l.tile_n;
        a.bias = l.b_off >= 0 ? F(l.b_off) : nullptr; a.act = ACT_NONE; a.mask_img = nullptr; a.mask_mt_stride = 0; a.resid = nullptr;
        a.out_f32 = nullptr; a.ld_f32 = p.P; a.out_img = nullptr; a.out_mt_stride = 0;
        return a;
    };
    // NICE forward over the couplings, in place on the state image (x's image is rebuilt by update_kernel each step)
    auto nice_forward = [&]() -> cudaError_t {
        for (int c = 0; c < p.n_coup; ++c) {
            const int on = ((p.mc + c) % 2) ? 0 : 1, off = 1 - on;  // plane index: 0 = even units (distr/nice.py:79-82)
            for (int l = 0; l < p.n_lin; ++l) {
                LinArgs a = base_args(p.nf[c][l]);
                uint8_t* act_out = ws + p.act_img + c * p.act_stride_coup + l * p.act_stride_layer;
                if (l == 0) { a.a_img = ws + p.ximg + (int64_t)off * (p.Hp / 64) * A_BLOCK; a.a_mt_stride = x_stride; }
                else { a.a_img = act_out - p.act_stride_layer; a.a_mt_stride = act_stride; }
                if (l < p.n_lin - 1) { a.act = ACT_RELU; a.out_img = act_out; a.out_mt_stride = act_stride; }
                else {
                    // on <- on + shift: the first two couplings still read x's planes, later ones the running state h
                    a.resid = (c < 2 ? F(p.xst) : F(p.h)) + on * p.Hp;
                    a.out_f32 = F(p.h) + on * p.Hp;
                    a.out_img = ws + p.ximg + (int64_t)on * (p.Hp / 64) * A_BLOCK;
                    a.out_mt_stride = x_stride;
                }
                cudaError_t e = launch_linear(a, p.m_tiles, simt, stream, launches);
                if (e != cudaSuccess) return e;
            }
        }
        return cudaSuccess;
    };
    const bool need_score = d.ctrl_kind != SDES_CTRL_CLIPPED && d.ctrl_kind != SDES_CTRL_LERP_PRIOR;

    for (int i = 0; i < p.T; ++i) {
        // ---- control MLP (models/mlp.py:114-122): x image -> 64 -> ... -> 64 -> nn (fp32, planar)
        {
            LinArgs a = base_args(p.mlp_in);
            a.a_img = ws + p.ximg; a.a_mt_stride = x_stride; a.bias = F(p.emb) + (int64_t)i * C; a.act = ACT_GELU;
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
                WIDE_CHECK(nice_forward());
                if (p.n_coup == 1) {}  // (single coupling: the untouched plane of h was copied above)
                latent_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 1);
                ++launches;
                WIDE_CHECK(cudaGetLastError());
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
                        WIDE_CHECK(launch_linear(a, p.m_tiles, simt, stream, launches));
                    }
                }
            } else {
                gmm_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 1);
                ++launches;
                WIDE_CHECK(cudaGetLastError());
            }
        }
        update_kernel<<<row_blocks, 32 * ROWS_PER_CTA, 0, stream>>>(ra, i);
        ++launches;
        WIDE_CHECK(cudaGetLastError());
    }
    // ---- terminal cost
    if (p.nice) {
        if (p.n_coup == 1) WIDE_CHECK(cudaMemcpyAsync(F(p.h), F(p.xst), p.Bp * (int64_t)p.P * 4, cudaMemcpyDeviceToDevice, stream));
        WIDE_CHECK(nice_forward());
        latent_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 0);
    } else {
        gmm_kernel<<<row_blocks_pad, 32 * ROWS_PER_CTA, 0, stream>>>(ra, 0);
    }
    ++launches;
    WIDE_CHECK(cudaGetLastError());
    terminal_kernel<<<row_blocks, 32 * ROWS_PER_CTA, 0, stream>>>(ra);
    ++launches;
    WIDE_CHECK(cudaGetLastError());
#undef WIDE_CHECK
    return launches;
}

}  // namespace sdes
