// sdes_grad.cu — gradient of the log-variance loss with respect to the control network (SURVEY §8f-1:
// `loss.backward()` of `Trainable.step`, solver/base.py:404-407, for loss.method = lv).
//
// In the lv losses the state is driven by the DETACHED control (`sde_ctrl = generative_ctrl.detach()`,
// losses/oc.py:60-64), so x_s carries no gradient, and the running cost g.(u - g/2) dt has zero derivative at
// u = g.  What is left is the Ito term (oc.py:218-219, :330-331, :440-443):
//     d rnd_b / d theta = sum_s  J_theta g(s, x_{b,s})^T  c_{b,s},     c = eps sqrt(dt)   (sigma beta_k eps for DDS)
//     d loss  / d theta = sum_b  w_b  d rnd_b / d theta,               w_b = 2 (rnd_b - mean) / (n - 1)   (kept b)
// i.e. ONE backward pass of the control MLP over all B*T (trajectory, step) rows with the output cotangent
// w_b c_{b,s} 1[|NN| <= clip_model] — no backpropagation through time.  The rows come from the stored trajectory
// xs (T+1, B, d) of the forward rollout and the noise is re-drawn from the same Philox counters (or re-read from
// HBM in parity mode).  Rows are processed in time chunks; per chunk
//     pack x rows -> operand image | 4 forward GEMMs (GELU and GELU' images) | cotangent kernel |
//     3 dgrad GEMMs (x GELU') | 4 wgrad GEMMs + bias / time-embedding column sums
// all on tcgen05: forward and dgrad are `linear_mma_kernel` (sdes_linear.cuh), wgrad is `wgrad_mma_kernel` below,
// which contracts over the ROW dimension by reading the very same activation images as MN-major operands.
// Outputs: gradients of the x-dependent layers in the parameter-blob layout, d loss / d emb (T, 64) and
// d loss / d gate (T, gate_dim); the caller chains the last two through the two tiny time-embedding networks.
#include <cuda_bf16.h>

#include <cstdlib>
#include <cstdio>
#include <vector>

#include "sdes_linear.cuh"
#include "sdes_step.cuh"
#include "sdes_timeembed.cuh"
#include "sdes_grad_fused.cuh"

namespace sdes {
namespace grad {

using namespace wide;

struct GradPlan {
    int d, P, pc, nh, T;
    int64_t B, Bp;
    int chunk_steps, n_chunks;
    int64_t chunk_rows;
    int m_tiles;  // per full chunk
    Lin f_in, f_h[SDES_MAX_HIDDEN], f_out;   // forward operands
    Lin b_h[SDES_MAX_HIDDEN], b_out;         // transposed operands (dgrad)
    Lin b_in;                                // kl sweep: transposed input layer (dgrad down to x)
    int64_t adj;                             // kl sweep: the adjoint a_s = d loss / d x_s, fp32 (Bp, P)
    int64_t kl_flags;                        // one-kernel kl sweep: per-tile hand-off counters (uint32)
    int64_t dh_all[SDES_MAX_HIDDEN + 1];     // kl sweep: one delta_h image per layer for the whole chunk
    int64_t embb;                            // (T, 64): timestep_embed(s) + b_in
    int64_t ximg, a_img[SDES_MAX_HIDDEN + 1], gp_img[SDES_MAX_HIDDEN + 1], nn, dnn_img, dh_img[2], ones;
    int64_t total;
};

static void make_plan(const SdesRolloutDesc& d, int64_t chunk_rows_req, int64_t base, GradPlan& p, bool bptt_tc = false) {
    p.d = d.dim;
    p.P = round_up(d.dim, 64);
    p.pc = p.P / 64;
    p.nh = d.n_hidden;
    p.T = d.n_steps;
    p.B = d.batch;
    p.Bp = (d.batch + 127) / 128 * 128;
    int64_t want = chunk_rows_req > 0 ? chunk_rows_req : (1 << 20);
    p.chunk_steps = (int)(want / p.Bp);
    if (p.chunk_steps < 1) p.chunk_steps = 1;
    if (p.chunk_steps > p.T) p.chunk_steps = p.T;
    p.n_chunks = (p.T + p.chunk_steps - 1) / p.chunk_steps;
    p.chunk_rows = (int64_t)p.chunk_steps * p.Bp;
    p.m_tiles = (int)(p.chunk_rows / 128);
    int64_t o = base;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align256(o + bytes); return r; };
    auto lin = [&](Lin& l, int N, int Kin, bool bias) {
        set_tiling(l, N, Kin);
        l.w_off = take(lin_image_bytes(l));
        l.b_off = bias ? take((int64_t)l.n_pad * 4) : -1;
    };
    lin(p.f_in, C, p.P, false);
    for (int l = 0; l < p.nh; ++l) lin(p.f_h[l], C, C, true);
    lin(p.f_out, p.P, C, true);
    for (int l = 0; l < p.nh; ++l) lin(p.b_h[l], C, C, false);
    lin(p.b_out, C, p.P, false);
    p.embb = take((int64_t)p.T * C * 4);
    p.ones = take(64 * 4);
    const int64_t img1 = (int64_t)p.m_tiles * A_BLOCK, imgp = img1 * p.pc;
    p.ximg = take(imgp);
    for (int l = 0; l <= p.nh; ++l) {
        p.a_img[l] = take(img1);
        p.gp_img[l] = take(img1);
    }
    p.nn = take(p.chunk_rows * (int64_t)p.P * 4);
    p.dnn_img = take(imgp);
    p.dh_img[0] = take(img1);
    p.dh_img[1] = take(img1);
    p.adj = -1;
    p.kl_flags = -1;
    if (bptt_tc) {
        lin(p.b_in, p.P, C, false);
        p.adj = take(p.Bp * (int64_t)p.P * 4);
        p.kl_flags = take(p.Bp / 128 * 4 + 4096);  // + a small record area for the hand-off watchdog
        p.dh_all[0] = p.dh_img[0];
        p.dh_all[1] = p.dh_img[1];
        for (int l = 2; l <= p.nh; ++l) p.dh_all[l] = take(img1);
    }
    p.total = o;
}

// ---------------------------------------------------------------------------- small kernels
__global__ void embb_kernel(const float* __restrict__ emb, const float* __restrict__ in_b, float* __restrict__ out, int n, float* __restrict__ ones) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e < n) out[e] = emb[e] + in_b[e & (C - 1)];
    if (e < 64) ones[e] = 1.0f;
}

// rows (s, b) of the stored trajectory -> A operand image of the input layer (natural feature order, K = P)
__global__ void __launch_bounds__(128) pack_rows_kernel(const SdesRolloutDesc d, const float* __restrict__ xs, int64_t Bp, int pc, int s0,
                                                        uint8_t* __restrict__ ximg) {
    const int mt = blockIdx.x, r = threadIdx.x, dim = d.dim;
    const int64_t rr = (int64_t)mt * 128 + r, B = d.batch;
    const int64_t s = s0 + rr / Bp, b = rr % Bp;
    const TrajRef x = traj_ref(d, const_cast<float*>(xs), (int)s, b < B ? b : 0);
    for (int k0 = 0; k0 < pc * 64; k0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (b < B && k0 + q < dim) ? __ldg(x.p + (k0 + q) * x.stride) : 0.f;
        uint4 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x);
        split_pair(v[2], v[3], hi.y, lo.y);
        split_pair(v[4], v[5], hi.z, lo.z);
        split_pair(v[6], v[7], hi.w, lo.w);
        uint8_t* o = ximg + (int64_t)mt * pc * A_BLOCK + img_group_offset(r, k0);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + A_HALF) = lo;
    }
}

// Output cotangent of the control network for every row of the chunk, and the gate's gradient.
//   delta_j = w_b c_j 1[|NN_j| <= clip_model],   c_j = eps_j sqrt(dt)  (DDS: sigma beta_k eps_j)
//   d gate(s) += w_b sum_j c_j * outer * clip(inner_j)            (the score part of models/reparam.py without its gate)
struct CotArgs {
    KParams kp;               // descriptor + fused-engine workspace layout (tables and target images of sdes_prepare.cu)
    const float* xs;
    const float* w;
    const float* nn;          // (rows, P)
    const float* ones;
    const float* delta;       // kl / kl_ito: control cotangent (T, B, d) of the adjoint sweep (sdes_adjoint.cu); NULL = lv
    float* adj;               // kl sweep on the tensor cores: adjoint (Bp, P) fp32, updated in place per step
    uint32_t gflags;          // SDES_GRAD_*
    int step;                 // kl sweep: the time step this launch handles
    uint8_t* dnn_img;
    float* grad_gate;         // (T, gate_dim) or NULL
    int P, pc, s0;
    int64_t Bp;
    int skip_gate;            // lv with the forward's gate_cot: the gate gradient is reduced from it, no target score here
};

template <int DPAD>
__global__ void __launch_bounds__(128) cotangent_kernel(const __grid_constant__ CotArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    float* smem = smem_f;
    __shared__ float s_red[4];
    const KParams& p = a.kp;
    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, K = d.n_components, tid = threadIdx.x;
    const int K2 = (K + 1) & ~1;
    float* s_mu = smem;
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    float* s_prior = s_c + 64;
    float* s_gsum = s_prior + 2 * DPAD + 4;  // DPAD per-dimension gate sums
    for (int e = tid; e < K2 * DPAD; e += blockDim.x) {
        s_mu[e] = ws[p.ws.gmm_mu + e];
        s_h[e] = ws[p.ws.gmm_h + e];
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c[e] = ws[p.ws.gmm_c + e];
    for (int e = tid; e < 2 * DPAD + 4; e += blockDim.x) s_prior[e] = e <= 2 * DPAD ? ws[p.ws.prior + e] : 0.f;
    for (int e = tid; e < DPAD; e += blockDim.x) s_gsum[e] = 0.f;
    __syncthreads();
    TargetSmem tsm{s_mu, s_h, s_c, reinterpret_cast<const uint32_t*>(ws + p.ws.counter)[1], s_prior, s_prior};

    const int mt = blockIdx.x;
    const int64_t rr = (int64_t)mt * 128 + tid, B = d.batch;
    const int s = a.s0 + (int)(rr / a.Bp);
    const int64_t b = rr % a.Bp;
    const bool valid = b < B;
    const int64_t bb = valid ? b : 0;
    float x[DPAD];
    const TrajRef xr = traj_ref(d, const_cast<float*>(a.xs), s, bb);
#pragma unroll
    for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(xr.p + j * xr.stride) : 0.f;
    const float* tab = ws + p.ws.tab + (int64_t)s * TAB_STRIDE;
    const StepCoef c = make_step_coef(d, tab);
    const float wb = valid ? a.w[bb] : 0.f;
    const float cscale = wb * (c.exp_int ? c.sg * c.beta_k : c.sqrt_dt);
    const bool bptt = a.delta != nullptr;
    const TrajRef dref = traj_ref(d, const_cast<float*>(bptt ? a.delta : a.xs), s, bb);
    const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)bb);
    const float* nrow = c.from_hbm ? d.noise + ((int64_t)s * B + bb) * dim : nullptr;
    const float* nnrow = a.nn + rr * a.P;  // (staging this tile through shared memory was measured slower here: 637 vs 457 us)
    const bool want_gate = !a.skip_gate && a.grad_gate != nullptr && (d.flags & SDES_F_HAS_GATE) && d.ctrl_kind != SDES_CTRL_CLIPPED;
    float sc[DPAD];
    if (want_gate) {
        score_part<DPAD>(d, x, sc, tsm, a.ones, tab[TAB_SIGMA], tab[TAB_LERP_W]);  // gate = 1: outer * clip(inner)
    } else {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = 0.f;
    }
    float gsum = 0.f;
    float cot[DPAD];
#pragma unroll
    for (int q = 0; q < DPAD / 4; ++q) {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (bptt) {
            // the cotangent of the control was produced by the reverse sweep; cscale * e below reproduces it
#pragma unroll
            for (int r = 0; r < 4; ++r) e[r] = (valid && 4 * q + r < dim) ? __ldg(dref.p + (4 * q + r) * dref.stride) : 0.f;
        } else if (4 * q < dim) {
            if (c.from_hbm) {
#pragma unroll
                for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? nrow[4 * q + r] : 0.f;
            } else {
                const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)s, (uint32_t)q);
                e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = 4 * q + r;
            const float cj = (j < dim) ? (bptt ? e[r] : cscale * e[r]) : 0.f;
            const float nnj = (j < dim) ? nnrow[j] : 0.f;
            cot[j] = (fabsf(nnj) <= c.cm) ? cj : 0.f;  // d clip(NN) / d NN (torch.clip passes the gradient on [-c, c])
            if (d.gate_dim == 1) gsum = fmaf(cj, sc[j], gsum);
            else if (want_gate && cj != 0.f) atomicAdd(&s_gsum[j], cj * sc[j]);
        }
    }
    // delta image (K = P = 64, natural feature order; this path serves d <= 64)
#pragma unroll
    for (int k0 = 0; k0 < 64; k0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (k0 + q < DPAD) ? cot[(k0 + q < DPAD) ? k0 + q : 0] : 0.f;
        uint4 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x);
        split_pair(v[2], v[3], hi.y, lo.y);
        split_pair(v[4], v[5], hi.z, lo.z);
        split_pair(v[6], v[7], hi.w, lo.w);
        uint8_t* o = a.dnn_img + (int64_t)mt * a.pc * A_BLOCK + img_group_offset(tid, k0);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + A_HALF) = lo;
    }
    if (!want_gate) return;
    const float gmask = fabsf(ws[p.ws.gate + (int64_t)s * p.ws.dpad]) < c.cm ? 1.0f : 0.f;  // d clip(gate) / d gate (scalar gate)
    if (d.gate_dim == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = gsum;
        __syncthreads();
        if (tid == 0) atomicAdd(a.grad_gate + s, gmask * (s_red[0] + s_red[1] + s_red[2] + s_red[3]));
    } else {
        __syncthreads();
        for (int j = tid; j < dim; j += blockDim.x) {
            const float gm = fabsf(ws[p.ws.gate + (int64_t)s * p.ws.dpad + j]) < c.cm ? 1.0f : 0.f;
            atomicAdd(a.grad_gate + (int64_t)s * dim + j, gm * s_gsum[j]);
        }
    }
}

// ---- kl / kl_ito reverse sweep with the control MLP on the tensor cores.  The math is that of sdes_adjoint.cu; here the
// replayed forward of the chunk (a_img / gp_img / nn) already exists, so one step of the sweep is
//     adj_step_kernel (thread per trajectory): control cotangent delta_s from a_{s+1}, w and the stored NN output -> the
//         masked delta image of this step's row tiles, the gate gradient, and a <- a * (1 + mu dt | alpha_k) + the score
//         term's own x-derivatives (+ reference-control term)
//     nh + 2 dgrad GEMMs on this step's row tiles (x GELU'), the last one accumulating J_x NN^T delta into a (resid)
// and the weight gradients are taken once per chunk from the per-layer delta images the sweep leaves behind.
template <int DPAD>
__global__ void __launch_bounds__(128) adj_init_kernel(const __grid_constant__ CotArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    float* smem = smem_f;
    const KParams& p = a.kp;
    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, K = d.n_components, tid = threadIdx.x;
    const int K2 = (K + 1) & ~1;
    float* s_mu = smem;
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    float* s_prior = s_c + 64;
    float* s_ref = s_prior + 2 * DPAD + 4;
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
    const int64_t b = (int64_t)blockIdx.x * 128 + tid, B = d.batch;
    const bool valid = b < B;
    const int64_t bb = valid ? b : 0;
    const float wb = valid ? a.w[bb] : 0.f;
    const bool dead = !(wb != 0.f);
    float x[DPAD], tsc[DPAD];
    const TrajRef xr = traj_ref(d, const_cast<float*>(a.xs), d.n_steps, bb);
#pragma unroll
    for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(xr.p + j * xr.stride) : 0.f;
    const float lp = target_eval<DPAD, true>(d, x, tsc, tsm);
    const float keep = fabsf(lp) <= d.clip_target ? 1.0f : 0.f;
    float* arow = a.adj + b * a.P;
#pragma unroll
    for (int j = 0; j < DPAD; ++j) {
        float v = -keep * tsc[j];
        if (d.loss_kind != SDES_LOSS_TIME_REVERSAL) v += (s_ref[j] - x[j]) * s_ref[DPAD + j];
        arow[j] = (j < dim && !dead) ? wb * v : 0.f;
    }
    for (int j = DPAD; j < a.P; ++j) arow[j] = 0.f;
}

template <int DPAD>
__global__ void __launch_bounds__(128) adj_step_kernel(const __grid_constant__ CotArgs a) {
    extern __shared__ __align__(16) float smem_f[];
    float* smem = smem_f;
    __shared__ float s_red[4];
    const KParams& p = a.kp;
    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, K = d.n_components, tid = threadIdx.x;
    const int K2 = (K + 1) & ~1;
    float* s_mu = smem;
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    float* s_prior = s_c + 64;
    float* s_gsum = s_prior + 2 * DPAD + 4;
    for (int e = tid; e < K2 * DPAD; e += blockDim.x) {
        s_mu[e] = ws[p.ws.gmm_mu + e];
        s_h[e] = ws[p.ws.gmm_h + e];
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c[e] = ws[p.ws.gmm_c + e];
    for (int e = tid; e < 2 * DPAD + 4; e += blockDim.x) s_prior[e] = e <= 2 * DPAD ? ws[p.ws.prior + e] : 0.f;
    for (int e = tid; e < DPAD; e += blockDim.x) s_gsum[e] = 0.f;
    const int s = a.step, mt_step = blockIdx.x;
    const int tiles_per_step = (int)(a.Bp / 128);
    const int mt = (s - a.s0) * tiles_per_step + mt_step;  // row tile inside the chunk's images
    // this tile's rows of the network output and of the adjoint, transposed through shared memory (coalesced segments in,
    // conflict-free rows out: stride DPAD + 1); the updated adjoint goes back the same way
    float* s_nn = s_gsum + DPAD;
    float* s_adj = s_nn + 128 * (DPAD + 1);
    {
        const float* src = a.nn + (int64_t)mt * 128 * a.P;
        const float* asrc = a.adj + (int64_t)mt_step * 128 * a.P;
        for (int e = tid; e < 128 * dim; e += blockDim.x) {
            const int r = e / dim, j = e - r * dim;
            s_nn[r * (DPAD + 1) + j] = src[(int64_t)r * a.P + j];
            s_adj[r * (DPAD + 1) + j] = asrc[(int64_t)r * a.P + j];
        }
    }
    __syncthreads();
    TargetSmem tsm{s_mu, s_h, s_c, reinterpret_cast<const uint32_t*>(ws + p.ws.counter)[1], s_prior, s_prior};

    const int64_t b = (int64_t)mt_step * 128 + tid, B = d.batch;
    const bool valid = b < B;
    const int64_t bb = valid ? b : 0;
    const float wb = valid ? a.w[bb] : 0.f;
    const bool dead = !(wb != 0.f);
    const float* tab = ws + p.ws.tab + (int64_t)s * TAB_STRIDE;
    const StepCoef c = make_step_coef(d, tab);
    const float* gate_row = ws + p.ws.gate + (int64_t)s * p.ws.dpad;
    const float lerp_w = tab[TAB_LERP_W];
    const int ck = d.ctrl_kind;
    const bool ito = (d.flags & SDES_F_COMPUTE_ITO) != 0;
    const bool score_detached = (a.gflags & SDES_GRAD_SCORE_DETACHED) != 0 || ck == SDES_CTRL_CLIPPED;
    const bool target_in_ctrl = ck == SDES_CTRL_SCORE || ck == SDES_CTRL_LERP || ck == SDES_CTRL_LERP_TARGET;
    const bool target_hvp = target_in_ctrl && !score_detached && !(a.gflags & SDES_GRAD_TARGET_SCORE_CONST);
    const bool prior_in_ctrl = ck == SDES_CTRL_LERP || ck == SDES_CTRL_LERP_PRIOR;
    const bool want_gate = a.grad_gate != nullptr && (d.flags & SDES_F_HAS_GATE) && ck != SDES_CTRL_CLIPPED;

    float x[DPAD];
    const TrajRef xr = traj_ref(d, const_cast<float*>(a.xs), s, bb);
#pragma unroll
    for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(xr.p + j * xr.stride) : 0.f;
    // score part: base = outer * clip(inner) (the gate's cotangent multiplies it), sc = base * gate, fac maps d g to d inner
    float sc[DPAD], fac[DPAD], base[DPAD];
    if (ck == SDES_CTRL_CLIPPED) {
#pragma unroll
        for (int j = 0; j < DPAD; ++j) sc[j] = fac[j] = base[j] = 0.f;
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
            base[j] = outer * clipf(inner, d.clip_score);
            sc[j] = base[j] * gate_row[j];
            fac[j] = fabsf(inner) <= d.clip_score ? outer * gate_row[j] : 0.f;
        }
    }
    const float* nnrow = s_nn + tid * (DPAD + 1);
    float* arow = s_adj + tid * (DPAD + 1);
    const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)bb);
    const float* nrow = (ito && c.from_hbm) ? d.noise + ((int64_t)s * B + bb) * dim : nullptr;
    const float a_mul = c.exp_int ? c.alpha_k : fmaf(c.mu, c.dt, 1.0f);
    float cot[DPAD], an[DPAD];
    float gsum = 0.f;
#pragma unroll
    for (int q = 0; q < DPAD / 4; ++q) {
        float e[4] = {0.f, 0.f, 0.f, 0.f};
        if (ito && 4 * q < dim) {
            if (c.from_hbm) {
#pragma unroll
                for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? nrow[4 * q + r] : 0.f;
            } else {
                const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)s, (uint32_t)q);
                e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
            }
        }
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int j = 4 * q + r;
            const float nnj = (j < dim) ? nnrow[j] : 0.f;
            const float ap = (j < dim) ? arow[j] : 0.f;
            const float g = clipf(nnj, c.cm) + sc[j];
            float dg, nx;
            if (c.exp_int) {
                dg = wb * (c.bb_ss * g + c.s_bk * e[r]) + ap * c.bb_ss;
                nx = ap * a_mul;
            } else {
                const float iv = s_prior[DPAD + j];
                const float gm = c.ref_ctrl ? g - c.sigma * ((s_prior[j] - x[j]) * iv) : g;
                const float qj = wb * (gm * c.dt + e[r] * c.sqrt_dt);
                dg = fmaf(ap, c.sigma * c.dt, qj);
                nx = ap * a_mul;
                if (c.ref_ctrl) nx = fmaf(qj, c.sigma * iv, nx);
            }
            if (dead || j >= dim) dg = 0.f;
            an[j] = nx;
            cot[j] = fabsf(nnj) <= c.cm ? dg : 0.f;  // d clip(NN) / d NN
            if (d.gate_dim == 1) gsum = fmaf(dg, base[j], gsum);
            else if (want_gate && dg != 0.f) atomicAdd(&s_gsum[j], dg * base[j]);
            fac[j] *= dg;  // cotangent of inner_j
        }
    }
    if (!score_detached) {
        if (prior_in_ctrl) {
            const float wp = 1.0f - lerp_w;
#pragma unroll
            for (int j = 0; j < DPAD; ++j) an[j] = fmaf(-wp * s_prior[DPAD + j], fac[j], an[j]);
        }
        if (target_hvp) {
            if (ck != SDES_CTRL_SCORE) {
#pragma unroll
                for (int j = 0; j < DPAD; ++j) fac[j] *= lerp_w;
            }
            target_hvp_add<DPAD>(d, x, fac, an, tsm);
        }
    }
#pragma unroll
    for (int j = 0; j < DPAD; ++j)
        if (j < dim) arow[j] = (dead || !valid) ? 0.f : an[j];
    __syncthreads();
    {
        float* adst = a.adj + (int64_t)mt_step * 128 * a.P;
        for (int e = tid; e < 128 * dim; e += blockDim.x) {
            const int r = e / dim, j = e - r * dim;
            adst[(int64_t)r * a.P + j] = s_adj[r * (DPAD + 1) + j];
        }
    }
    // masked delta image of this step's rows (K = P = 64, natural feature order)
#pragma unroll
    for (int k0 = 0; k0 < 64; k0 += 8) {
        float v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = (k0 + q < DPAD) ? cot[(k0 + q < DPAD) ? k0 + q : 0] : 0.f;
        uint4 hi, lo;
        split_pair(v[0], v[1], hi.x, lo.x);
        split_pair(v[2], v[3], hi.y, lo.y);
        split_pair(v[4], v[5], hi.z, lo.z);
        split_pair(v[6], v[7], hi.w, lo.w);
        uint8_t* o = a.dnn_img + (int64_t)mt * a.pc * A_BLOCK + img_group_offset(tid, k0);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + A_HALF) = lo;
    }
    if (!want_gate) return;
    const float gmask = fabsf(ws[p.ws.gate + (int64_t)s * p.ws.dpad]) < c.cm ? 1.0f : 0.f;
    if (d.gate_dim == 1) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
        if ((tid & 31) == 0) s_red[tid >> 5] = gsum;
        __syncthreads();
        if (tid == 0) atomicAdd(a.grad_gate + s, gmask * (s_red[0] + s_red[1] + s_red[2] + s_red[3]));
    } else {
        __syncthreads();
        for (int j = tid; j < dim; j += blockDim.x) {
            const float gm = fabsf(ws[p.ws.gate + (int64_t)s * p.ws.dpad + j]) < c.cm ? 1.0f : 0.f;
            atomicAdd(a.grad_gate + (int64_t)s * dim + j, gm * s_gsum[j]);
        }
    }
}

// ---- fused dgrad chain of the kl sweep: all n_hidden + 2 transposed layers of one step's row tile in ONE kernel.
// Per 128-row tile: TMA brings the masked delta image in as the first A operand; every layer is 12 bf16 tcgen05 MMAs
// (hi/lo split, K = 64) against its weight image, all of which stay resident in shared memory; the epilogue warps
// multiply the accumulator by GELU'(h) (stored image), write the delta_h image the weight gradients need AND the same
// 16-byte groups straight into the shared-memory A buffer of the next layer (the image layout is the canonical
// no-swizzle K-major operand layout), so no delta_h is ever read back; the last layer (W_in^T) accumulates into the
// fp32 adjoint.  One launch per time step instead of n_hidden + 2, no operand re-reads.  Layers are a dependency chain
// per tile, so two CTAs share an SM and overlap each other's MMA / epilogue phases, and each row's 64 accumulator columns
// are split over CHAIN_EPW epilogue warps to shorten every hop of the chain.
struct ChainArgs {
    const uint8_t* dnn_img;                       // [m_tile] blocks of A_BLOCK (K = P = 64)
    const uint8_t* w_img[SDES_MAX_HIDDEN + 2];    // W_out^T, W_h[nh-1]^T, ..., W_h[0]^T, W_in^T: 16 KB each (hi | lo)
    const uint8_t* gp_img[SDES_MAX_HIDDEN + 1];   // GELU'(h) image multiplying the output of layer i < L - 1
    uint8_t* dh_img[SDES_MAX_HIDDEN + 1];         // delta_h image written by layer i < L - 1
    float* adj;                                   // (rows, 64) fp32: += output of the last layer
    int n_layers, m_tiles;
};
// epilogue warps per TMEM lane quadrant: a single warp runs an 8-column group's ~100 dependent instructions at ~6-10 cycles
// each, so a row's 64 columns are shared by CHAIN_EPW warps (16 columns each) to shorten the per-layer hop of the chain
constexpr int CHAIN_EPW = 4;
constexpr int CHAIN_THREADS = 64 + 128 * CHAIN_EPW;
constexpr int CHAIN_COLS = 64 / CHAIN_EPW, CHAIN_GROUPS = CHAIN_COLS / 8;
constexpr uint32_t CHAIN_W_BYTES = 16384u;        // one 64 x 64 weight image

static __global__ void __launch_bounds__(CHAIN_THREADS, 2) dgrad_chain_kernel(const __grid_constant__ ChainArgs a) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t s_wfull, s_afull, s_acc, s_aready;
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int L = a.n_layers;
    uint8_t* s_a = smem;                  // A operand: hi | lo (32 KB)
    uint8_t* s_w = smem + A_BLOCK;        // L weight images
    if (warp == 1) {
        tc::tmem_alloc(&s_tmem, 64u);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(&s_wfull, 1);
        tc::mbar_init(&s_afull, 1);
        tc::mbar_init(&s_acc, 1);
        tc::mbar_init(&s_aready, 128 * CHAIN_EPW);
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        if (lane == 0) {  // ---- control thread: TMA loads and MMA issue (the chain is sequential per tile anyway)
            tc::mbar_arrive_expect_tx(&s_wfull, (uint32_t)L * CHAIN_W_BYTES);
            for (int l = 0; l < L; ++l) tc::bulk_g2s(s_w + (size_t)l * CHAIN_W_BYTES, a.w_img[l], CHAIN_W_BYTES, &s_wfull);
            tc::mbar_wait(&s_wfull, 0u);
            const uint32_t idesc = tc::idesc_bf16(128, 64);
            const uint32_t a_hi = tc::smem_u32(s_a), a_lo = a_hi + A_HALF;
            uint32_t ph_a = 0u, ph_ready = 0u;
            for (int tile = (int)blockIdx.x; tile < a.m_tiles; tile += (int)gridDim.x) {
                tc::mbar_arrive_expect_tx(&s_afull, A_BLOCK);
                const uint8_t* src = a.dnn_img + (int64_t)tile * A_BLOCK;
                tc::bulk_g2s(s_a, src, A_HALF, &s_afull);
                tc::bulk_g2s(s_a + A_HALF, src + A_HALF, A_HALF, &s_afull);
                tc::mbar_wait(&s_afull, ph_a);
                ph_a ^= 1u;
                for (int l = 0; l < L; ++l) {
                    if (l > 0) {  // the epilogue has drained the accumulator and written this layer's A operand
                        tc::mbar_wait(&s_aready, ph_ready);
                        ph_ready ^= 1u;
                    }
                    tc::fence_after();
                    const uint32_t b_hi = tc::smem_u32(s_w + (size_t)l * CHAIN_W_BYTES), b_lo = b_hi + 8192u;
#pragma unroll
                    for (int ks = 0; ks < KC / 16; ++ks) {
                        const uint64_t dah = tc::smem_desc_kmajor(a_hi + (uint32_t)ks * 4096u, 2048u, 128u);
                        const uint64_t dal = tc::smem_desc_kmajor(a_lo + (uint32_t)ks * 4096u, 2048u, 128u);
                        const uint64_t dbh = tc::smem_desc_kmajor(b_hi + (uint32_t)ks * 2048u, 1024u, 128u);
                        const uint64_t dbl = tc::smem_desc_kmajor(b_lo + (uint32_t)ks * 2048u, 1024u, 128u);
                        mma_f16_ss(tmem_base, dal, dbh, idesc, ks > 0 ? 1u : 0u);  // small terms first
                        mma_f16_ss(tmem_base, dah, dbl, idesc, 1u);
                        mma_f16_ss(tmem_base, dah, dbh, idesc, 1u);
                    }
                    tc::mma_commit(&s_acc);
                }
                // the last epilogue has read the accumulator (and the last MMAs have read A): the next tile may start
                tc::mbar_wait(&s_aready, ph_ready);
                ph_ready ^= 1u;
            }
        }
    } else if (warp >= 2) {  // ---- epilogue warps: TMEM lane quadrant = warp % 4, thread = row
        const int q = warp & 3, r = q * 32 + lane;
        const int c_lo = ((warp - 2) >> 2) * CHAIN_COLS;  // this warp's share of the row's columns
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c_lo;
        uint32_t ph_acc = 0u;
        for (int tile = (int)blockIdx.x; tile < a.m_tiles; tile += (int)gridDim.x) {
            for (int l = 0; l < L; ++l) {
                if (l < L - 1) {
                    // this layer's GELU' row does not depend on the accumulator: all 16 loads are issued BEFORE waiting for
                    // the MMAs, so their latency hides behind the tensor-core phase of the chain
                    const uint8_t* gpp = a.gp_img[l] + (int64_t)tile * A_BLOCK;
                    uint8_t* dhp = a.dh_img[l] + (int64_t)tile * A_BLOCK;
                    uint4 mh[CHAIN_GROUPS], ml[CHAIN_GROUPS];
#pragma unroll
                    for (int g = 0; g < CHAIN_GROUPS; ++g) {
                        const int64_t go = img_group_offset(r, c_lo + 8 * g);
                        mh[g] = *reinterpret_cast<const uint4*>(gpp + go);
                        ml[g] = *reinterpret_cast<const uint4*>(gpp + go + A_HALF);
                    }
                    tc::mbar_wait(&s_acc, ph_acc);
                    ph_acc ^= 1u;
                    tc::fence_after();
                    float v[8];
                    tc::tmem_ld8(taddr, v);
#pragma unroll
                    for (int g = 0; g < CHAIN_GROUPS; ++g) {
                        tc::wait_ld_tie<8>(v);
                        float w[8], fh[8], fl[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] = v[e];
                        if (g < CHAIN_GROUPS - 1) tc::tmem_ld8(taddr + (uint32_t)(8 * g + 8), v);
                        unpack8(mh[g], fh);
                        unpack8(ml[g], fl);
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] *= fh[e] + fl[e];
                        uint4 hi, lo;
                        split_pair(w[0], w[1], hi.x, lo.x);
                        split_pair(w[2], w[3], hi.y, lo.y);
                        split_pair(w[4], w[5], hi.z, lo.z);
                        split_pair(w[6], w[7], hi.w, lo.w);
                        const int64_t goff = img_group_offset(r, c_lo + 8 * g);
                        *reinterpret_cast<uint4*>(s_a + goff) = hi;          // next layer's A operand, same layout
                        *reinterpret_cast<uint4*>(s_a + goff + A_HALF) = lo;
                        *reinterpret_cast<uint4*>(dhp + goff) = hi;
                        *reinterpret_cast<uint4*>(dhp + goff + A_HALF) = lo;
                    }
                    tc::fence_proxy_async();  // generic-proxy writes of A -> visible to the tensor-core (async) proxy
                } else {
                    float* arow = a.adj + ((int64_t)tile * 128 + r) * 64 + c_lo;
                    float4 acc[2 * CHAIN_GROUPS];  // this warp's part of the adjoint row, fetched while the last layer's MMAs run
#pragma unroll
                    for (int g = 0; g < 2 * CHAIN_GROUPS; ++g) acc[g] = reinterpret_cast<const float4*>(arow)[g];
                    tc::mbar_wait(&s_acc, ph_acc);
                    ph_acc ^= 1u;
                    tc::fence_after();
                    float v[8];
                    tc::tmem_ld8(taddr, v);
#pragma unroll
                    for (int g = 0; g < CHAIN_GROUPS; ++g) {
                        tc::wait_ld_tie<8>(v);
                        float w[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) w[e] = v[e];
                        if (g < CHAIN_GROUPS - 1) tc::tmem_ld8(taddr + (uint32_t)(8 * g + 8), v);
                        float4 r0 = acc[2 * g], r1 = acc[2 * g + 1];
                        r0.x += w[0]; r0.y += w[1]; r0.z += w[2]; r0.w += w[3];
                        r1.x += w[4]; r1.y += w[5]; r1.z += w[6]; r1.w += w[7];
                        reinterpret_cast<float4*>(arow)[2 * g] = r0;
                        reinterpret_cast<float4*>(arow)[2 * g + 1] = r1;
                    }
                }
                tc::fence_before();
                mbar_arrive(&s_aready);
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 64u);
}

static cudaError_t launch_dgrad_chain(const ChainArgs& a, int sm_count, cudaStream_t stream) {
    const size_t smem = A_BLOCK + (size_t)a.n_layers * CHAIN_W_BYTES;
    static size_t attr = 0;
    if (attr < smem) {
        cudaError_t e = cudaFuncSetAttribute(dgrad_chain_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    const int per_sm = smem <= 110 * 1024 ? 2 : 1;
    int grid = a.m_tiles < sm_count * per_sm ? a.m_tiles : sm_count * per_sm;
    if (grid < 1) grid = 1;
    dgrad_chain_kernel<<<grid, CHAIN_THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}

// column sums of a delta image: bias gradients, and per-time-step sums for d loss / d emb
__global__ void __launch_bounds__(128) colsum_kernel(const uint8_t* __restrict__ img, int n_chunks, int n_valid, float* __restrict__ out,
                                                     int mt_div, int64_t mt_stride, int planar, int Hp) {
    __shared__ float s_part[16][64];
    const int mt = blockIdx.x, tid = threadIdx.x, kg = tid & 7, rs = tid >> 3;
    float* dst = out + (mt_div > 0 ? (int64_t)(mt / mt_div) * mt_stride : 0);
    for (int ch = 0; ch < n_chunks; ++ch) {
        const uint8_t* blk = img + ((int64_t)mt * n_chunks + ch) * A_BLOCK;
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int r = rs; r < 128; r += 16) {
            float h[8], l[8];
            unpack8(*reinterpret_cast<const uint4*>(blk + (kg * 128 + r) * 16), h);
            unpack8(*reinterpret_cast<const uint4*>(blk + A_HALF + (kg * 128 + r) * 16), l);
#pragma unroll
            for (int q = 0; q < 8; ++q) acc[q] += h[q] + l[q];
        }
#pragma unroll
        for (int q = 0; q < 8; ++q) s_part[rs][kg * 8 + q] = acc[q];
        __syncthreads();
        if (tid < 64) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) s += s_part[i][tid];
            const int f = to_natural(ch * 64 + tid, planar, Hp);
            if (f < n_valid && s != 0.f) atomicAdd(dst + f, s);
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------ wgrad GEMM
// dW[n][k] += sum_rows delta[row][n] a[row][k] for one 64 x 64 block (delta chunk cd, activation chunk ca): the
// contraction runs over the ROWS, so the activation images are read as MN-major operands (inside a block the
// element (row r, feature f) sits at ((f/8)*128 + r)*16 + (f%8)*2 bytes: 8 features contiguous, rows 16 B apart,
// row groups 128 B apart (LBO), feature groups 2048 B apart (SBO)).  A = [delta_hi ; delta_lo] spans the whole
// 32 KB block as M = 128 "features", so D rows 0-63 and 64-127 are the hi and lo parts of the same gradient block;
// B = a_hi, then a_lo.  Each CTA sums its share of the row tiles in TMEM and adds the block to global memory once.
struct WgradArgs {
    const uint8_t* d_img; int d_chunks;   // delta image, [m_tile][d_chunks] blocks
    const uint8_t* a_img; int a_chunks;   // activation image
    int m_tiles;
    float* dw; int ldw;                   // row-major (out, in) gradient, += via atomics
    float* db;                            // bias gradient (column sums of delta), += via atomics; NULL = not wanted.  Comes
                                          // out of the same pass as one extra N = 16 MMA per k-step against a block of ones
    int n_valid, k_valid;
    int n_planar, k_planar, Hp;           // wide engine: image feature index -> natural index (2 (p % Hp) + p / Hp)
};

__device__ __forceinline__ uint32_t idesc_bf16_mn(int M, int N) {
    return tc::idesc_bf16(M, N) | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}

constexpr int WG_THREADS = 192;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2-5: one TMEM lane quadrant each
constexpr uint32_t WG_ONES_BYTES = 4096u;  // MN-major B operand of ones: 2 feature groups x 2048 B
constexpr size_t WG_SMEM = 3 * 2 * (size_t)A_BLOCK + WG_ONES_BYTES;

__global__ void __launch_bounds__(WG_THREADS, 1) wgrad_mma_kernel(const __grid_constant__ WgradArgs a) {
    extern __shared__ __align__(1024) uint8_t smem_b[];
    uint8_t* smem = smem_b;
    __shared__ uint64_t s_full[3], s_empty[3], s_acc;
    __shared__ uint32_t s_tmem;
    constexpr int STAGES = 3;
    constexpr uint32_t STAGE_BYTES = 2u * A_BLOCK;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int pair = blockIdx.x, cd = pair / a.a_chunks, ca = pair % a.a_chunks;
    const int split = blockIdx.y, n_split = gridDim.y;
    const int n_my = a.m_tiles > split ? (a.m_tiles - split + n_split - 1) / n_split : 0;
    const bool want_db = a.db != nullptr && ca == 0;
    uint8_t* s_ones = smem + (size_t)STAGES * STAGE_BYTES;
    if (want_db) {
        for (int e = tid; e < (int)(WG_ONES_BYTES / 4); e += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[e] = 0x3F803F80u;  // bf16 1.0 pairs
        tc::fence_proxy_async();
    }

    if (warp == 1) {
        tc::tmem_alloc(&s_tmem, 128);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
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
    if (n_my > 0) {
        if (warp == 0) {
            if (lane == 0) {
                for (int i = 0; i < n_my; ++i) {
                    const int s = i % STAGES, it = i / STAGES;
                    const int64_t mt = split + (int64_t)i * n_split;
                    if (it > 0) tc::mbar_wait(&s_empty[s], (uint32_t)((it - 1) & 1));
                    tc::mbar_arrive_expect_tx(&s_full[s], STAGE_BYTES);
                    uint8_t* dst = smem + (size_t)s * STAGE_BYTES;
                    const uint8_t* dp = a.d_img + (mt * a.d_chunks + cd) * A_BLOCK;
                    const uint8_t* ap = a.a_img + (mt * a.a_chunks + ca) * A_BLOCK;
                    tc::bulk_g2s(dst, dp, A_HALF, &s_full[s]);
                    tc::bulk_g2s(dst + A_HALF, dp + A_HALF, A_HALF, &s_full[s]);
                    tc::bulk_g2s(dst + A_BLOCK, ap, A_HALF, &s_full[s]);
                    tc::bulk_g2s(dst + A_BLOCK + A_HALF, ap + A_HALF, A_HALF, &s_full[s]);
                }
            }
        } else if (warp == 1) {
            if (lane == 0) {
                const uint32_t idesc = idesc_bf16_mn(128, 64), idesc1 = idesc_bf16_mn(128, 16);
                const uint32_t ones = tc::smem_u32(s_ones);
                for (int i = 0; i < n_my; ++i) {
                    const int s = i % STAGES, it = i / STAGES;
                    tc::mbar_wait(&s_full[s], (uint32_t)(it & 1));
                    tc::fence_after();
                    const uint32_t dl = tc::smem_u32(smem + (size_t)s * STAGE_BYTES), ah = dl + A_BLOCK, al = ah + A_HALF;
#pragma unroll
                    for (int ks = 0; ks < 8; ++ks) {  // 16 rows per MMA
                        const uint64_t da = tc::smem_desc_kmajor(dl + (uint32_t)ks * 256u, 128u, 2048u);
                        const uint64_t dbh = tc::smem_desc_kmajor(ah + (uint32_t)ks * 256u, 128u, 2048u);
                        const uint64_t dbl = tc::smem_desc_kmajor(al + (uint32_t)ks * 256u, 128u, 2048u);
                        mma_f16_ss(tmem_d, da, dbl, idesc, (i > 0 || ks > 0) ? 1u : 0u);
                        mma_f16_ss(tmem_d, da, dbh, idesc, 1u);
                        if (want_db)  // column sums of delta over these 16 rows: every column of the extra tile holds them
                            mma_f16_ss(tmem_d + 64u, da, tc::smem_desc_kmajor(ones, 128u, 2048u), idesc1, (i > 0 || ks > 0) ? 1u : 0u);
                    }
                    tc::mma_commit(&s_empty[s]);
                }
                tc::mma_commit(&s_acc);
            }
        } else {
            const int q = warp & 3, r = q * 32 + lane;  // D row = feature (hi part for r < 64, lo part above)
            tc::mbar_wait(&s_acc, 0);
            tc::fence_after();
            const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
            const int n = to_natural(cd * 64 + (r & 63), a.n_planar, a.Hp);
            for (int c0 = 0; c0 < 64; c0 += 8) {
                float v[8];
                tc::tmem_ld8(taddr + (uint32_t)c0, v);
                tc::wait_ld_tie<8>(v);
                if (n < a.n_valid) {
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        const int k = to_natural(ca * 64 + c0 + e, a.k_planar, a.Hp);
                        if (k < a.k_valid && v[e] != 0.f) atomicAdd(a.dw + (int64_t)n * a.ldw + k, v[e]);
                    }
                }
            }
            if (want_db) {
                float v[8];
                tc::tmem_ld8(taddr + 64u, v);
                tc::wait_ld_tie<8>(v);
                if (n < a.n_valid && v[0] != 0.f) atomicAdd(a.db + n, v[0]);  // hi row (r < 64) and lo row (r >= 64) both add
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_d, 128);
}

// the same contraction on the CUDA cores (SDES_F_MLP_SIMT cross-check)
__global__ void __launch_bounds__(64) wgrad_simt_kernel(const WgradArgs a) {
    const int pair = blockIdx.x, cd = pair / a.a_chunks, ca = pair % a.a_chunks;
    const int nl = threadIdx.x, n = to_natural(cd * 64 + nl, a.n_planar, a.Hp);
    for (int k0 = 0; k0 < 64; k0 += 8) {
        float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int64_t mt = blockIdx.y; mt < a.m_tiles; mt += gridDim.y) {
            const uint8_t* dp = a.d_img + (mt * a.d_chunks + cd) * A_BLOCK + (nl >> 3) * 2048 + (nl & 7) * 2;
            const uint8_t* ap = a.a_img + (mt * a.a_chunks + ca) * A_BLOCK + (k0 >> 3) * 2048;
            for (int r = 0; r < 128; ++r) {
                const uint16_t dh = *reinterpret_cast<const uint16_t*>(dp + r * 16), dl = *reinterpret_cast<const uint16_t*>(dp + A_HALF + r * 16);
                const float dv = __uint_as_float((uint32_t)dh << 16) + __uint_as_float((uint32_t)dl << 16);
                float ah[8], al[8];
                unpack8(*reinterpret_cast<const uint4*>(ap + r * 16), ah);
                unpack8(*reinterpret_cast<const uint4*>(ap + A_HALF + r * 16), al);
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] = fmaf(dv, ah[e] + al[e], acc[e]);
            }
        }
        if (n < a.n_valid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int k = to_natural(ca * 64 + k0 + e, a.k_planar, a.Hp);
                if (k < a.k_valid && acc[e] != 0.f) atomicAdd(a.dw + (int64_t)n * a.ldw + k, acc[e]);
            }
        }
    }
}

// ---- wide engine (planar state, one warp per row): output cotangent of the control network
struct CotWideArgs {
    SdesRolloutDesc d;
    const float* tab;
    const float* w;
    const float* nn;      // (rows, P) planar
    uint8_t* dnn_img;     // [m_tile][pc] blocks, planar features
    int Hp, P, pc, s0;
    int64_t Bp;
};

__device__ __forceinline__ void img_store_pair_g(uint8_t* img, int pc, int Hp, int64_t row, int plane, int k, float v0, float v1) {
    const int mt = (int)(row >> 7), r = (int)(row & 127), kk = plane * Hp + k;
    uint32_t hi, lo;
    split_pair(v0, v1, hi, lo);
    uint8_t* o = img + (int64_t)mt * pc * A_BLOCK + img_group_offset(r, kk) + (kk & 7) * 2;
    *reinterpret_cast<uint32_t*>(o) = hi;
    *reinterpret_cast<uint32_t*>(o + A_HALF) = lo;
}

__global__ void __launch_bounds__(256) cotangent_wide_kernel(const CotWideArgs a) {
    const SdesRolloutDesc& d = a.d;
    const int lane = threadIdx.x & 31;
    const int64_t rr = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    const int s = a.s0 + (int)(rr / a.Bp);
    const int64_t b = rr % a.Bp, B = d.batch;
    const bool valid = b < B;
    const int64_t bb = valid ? b : 0;
    const int dim = d.dim, Hp = a.Hp;
    const StepCoef c = make_step_coef(d, a.tab + (int64_t)s * TAB_STRIDE);
    const float cscale = (valid ? a.w[bb] : 0.f) * (c.exp_int ? c.sg * c.beta_k : c.sqrt_dt);
    const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)bb);
    const float* noise = c.from_hbm ? d.noise + ((int64_t)s * B + bb) * dim : nullptr;
    const float* nr = a.nn + rr * a.P;
    for (int q = lane; q < Hp / 2; q += 32) {
        float cot[4] = {0.f, 0.f, 0.f, 0.f};
        if (4 * q < dim) {
            float e[4];
            if (c.from_hbm) {
#pragma unroll
                for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? noise[4 * q + r] : 0.f;
            } else {
                const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)s, (uint32_t)q);
                e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
            }
            const float2 ne = *reinterpret_cast<const float2*>(nr + 2 * q), no = *reinterpret_cast<const float2*>(nr + Hp + 2 * q);
            const float nn4[4] = {ne.x, no.x, ne.y, no.y};  // natural order within the quad
#pragma unroll
            for (int r = 0; r < 4; ++r) cot[r] = (4 * q + r < dim && fabsf(nn4[r]) <= c.cm) ? cscale * e[r] : 0.f;
        }
        img_store_pair_g(a.dnn_img, a.pc, Hp, rr, 0, 2 * q, cot[0], cot[2]);
        img_store_pair_g(a.dnn_img, a.pc, Hp, rr, 1, 2 * q, cot[1], cot[3]);
    }
}

// ---- kl / kl_ito on the wide engine: the reverse sweep's elementwise step (planar state, one warp per row).
// With a_{s+1} = d loss / d x_{s+1}, w_b = d loss / d rnd_b, g = generative_ctrl(s, x_s) = clip(NN) + part(x_s):
//   q = w (cq g_m + ci eps),  delta_s = q + a_{s+1} Bc  (cotangent of the control),  a_s = A a_{s+1} + (d part / d x)^T delta_s
//   [+ q sigma / scale_prior^2 for the Euler-DDS reference control]  + J_x NN^T (delta_s 1[|NN| <= clip_model])  — the last term
//   is added by the dgrad GEMM chain that follows this kernel (it reads the masked delta image written here).
//   Euler-Maruyama (losses/oc.py:204-219, :316-331): A = 1 + mu dt, Bc = sigma dt, cq = dt, ci = sqrt(dt) [ito]
//   exponential integrator (:429-443):                A = alpha_k,   Bc = beta_k^2 sigma^2, cq = Bc, ci = sigma beta_k [ito]
// The target score inside `part` is the value the forward kept (SDES_F_KEEP_SCORE): a NICE / multi-component GMM score is an
// autograd score without create_graph in the reference (distr/base.py:130-137) and enters as a constant; a single Gaussian is
// differentiated (-1 / scale^2); the prior score of the Lerp controls always is.
struct AdjWideArgs {
    SdesRolloutDesc d;
    const float *tab, *gate, *w, *nn, *sc, *vec_prior, *gmm_h;
    const uint8_t* ximg;   // state image of this step
    float* adj;            // (Bp, P) fp32 planar, in: a_{s+1}, out: A a_{s+1} + score-term part of a_s
    uint8_t* dnn_img;      // masked delta of this step's rows: [m_tile][pc] blocks
    float* qg;             // (Bp) sum_j delta_j part_j / gate, or NULL
    int Hp, P, pc, step;
    int64_t Bp;
    uint32_t gflags;
};

__device__ __forceinline__ void img_load_pair(const uint8_t* img, int pc, int Hp, int64_t row, int plane, int k, float& v0, float& v1) {
    const int mt = (int)(row >> 7), r = (int)(row & 127), kk = plane * Hp + k;
    const uint8_t* o = img + (int64_t)mt * pc * A_BLOCK + img_group_offset(r, kk) + (kk & 7) * 2;
    const uint32_t hi = *reinterpret_cast<const uint32_t*>(o), lo = *reinterpret_cast<const uint32_t*>(o + A_HALF);
    v0 = __uint_as_float(hi << 16) + __uint_as_float(lo << 16);
    v1 = __uint_as_float(hi & 0xFFFF0000u) + __uint_as_float(lo & 0xFFFF0000u);
}

__global__ void __launch_bounds__(256) adj_wide_step_kernel(const AdjWideArgs a) {
    const SdesRolloutDesc& d = a.d;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= a.Bp) return;
    const int64_t B = d.batch;
    const bool valid = row < B;
    const int64_t bb = valid ? row : 0;
    const int dim = d.dim, Hp = a.Hp, P = a.P, s = a.step;
    const float* tab = a.tab + (int64_t)s * TAB_STRIDE;
    const StepCoef c = make_step_coef(d, tab);
    const bool ito = (d.flags & SDES_F_COMPUTE_ITO) != 0;
    const float wb = valid ? a.w[bb] : 0.f;
    const float A = c.exp_int ? c.alpha_k : fmaf(c.mu, c.dt, 1.0f), Bc = c.exp_int ? c.bb_ss : c.sigma * c.dt;
    const float cq = c.exp_int ? c.bb_ss : c.dt, ci = ito ? (c.exp_int ? c.sg * c.beta_k : c.sqrt_dt) : 0.f;
    const float gate = a.gate[s], lerp_w = tab[TAB_LERP_W], cs = d.clip_score;
    const float outer = (d.ctrl_kind == SDES_CTRL_SCORE ? 1.0f : c.sigma) * d.scale_score;
    const bool has_part = d.ctrl_kind != SDES_CTRL_CLIPPED;
    const bool use_sc = has_part && d.ctrl_kind != SDES_CTRL_LERP_PRIOR;
    const bool score_detached = (a.gflags & SDES_GRAD_SCORE_DETACHED) != 0;
    const bool target_diff = !(a.gflags & SDES_GRAD_TARGET_SCORE_CONST) && d.target_kind == SDES_TARGET_GMM && d.n_components == 1;
    const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)bb);
    const float* noise = (c.from_hbm && ito) ? d.noise + ((int64_t)s * B + bb) * dim : nullptr;
    const float* nr = a.nn + row * P;
    const float* sr = use_sc ? a.sc + row * P : nullptr;
    float* ar = a.adj + row * P;
    float qg = 0.f;
    for (int q = lane; q < Hp / 2; q += 32) {
        float dm[4] = {0.f, 0.f, 0.f, 0.f}, an[4] = {0.f, 0.f, 0.f, 0.f};
        if (4 * q < dim) {
            float e[4] = {0.f, 0.f, 0.f, 0.f};
            if (ito) {
                if (c.from_hbm) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) e[r] = (4 * q + r < dim) ? noise[4 * q + r] : 0.f;
                } else {
                    const float4 n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)s, (uint32_t)q);
                    e[0] = n4.x; e[1] = n4.y; e[2] = n4.z; e[3] = n4.w;
                }
            }
            float xe0, xe1, xo0, xo1;
            img_load_pair(a.ximg, a.pc, Hp, row, 0, 2 * q, xe0, xe1);
            img_load_pair(a.ximg, a.pc, Hp, row, 1, 2 * q, xo0, xo1);
            const float2 ne = *reinterpret_cast<const float2*>(nr + 2 * q), no = *reinterpret_cast<const float2*>(nr + Hp + 2 * q);
            float2 se = make_float2(0.f, 0.f), so = se, he = se, ho = se;
            if (use_sc) { se = *reinterpret_cast<const float2*>(sr + 2 * q); so = *reinterpret_cast<const float2*>(sr + Hp + 2 * q); }
            if (target_diff) { he = *reinterpret_cast<const float2*>(a.gmm_h + 2 * q); ho = *reinterpret_cast<const float2*>(a.gmm_h + Hp + 2 * q); }
            const float2 ae = *reinterpret_cast<const float2*>(ar + 2 * q), ao = *reinterpret_cast<const float2*>(ar + Hp + 2 * q);
            const float2 ple = *reinterpret_cast<const float2*>(a.vec_prior + 2 * q), plo = *reinterpret_cast<const float2*>(a.vec_prior + Hp + 2 * q);
            const float2 pie = *reinterpret_cast<const float2*>(a.vec_prior + P + 2 * q), pio = *reinterpret_cast<const float2*>(a.vec_prior + P + Hp + 2 * q);
            const float x4[4] = {xe0, xo0, xe1, xo1}, nn4[4] = {ne.x, no.x, ne.y, no.y}, sc4[4] = {se.x, so.x, se.y, so.y};
            const float a4[4] = {ae.x, ao.x, ae.y, ao.y}, pl4[4] = {ple.x, plo.x, ple.y, plo.y}, pi4[4] = {pie.x, pio.x, pie.y, pio.y};
            const float h4[4] = {he.x, ho.x, he.y, ho.y};
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (4 * q + r >= dim) continue;
                const float pscore = (pl4[r] - x4[r]) * pi4[r];
                const float tsd = target_diff ? -2.0f * h4[r] : 0.f;  // d target score / d x (diagonal)
                float inner = 0.f, dinner = 0.f;
                if (d.ctrl_kind == SDES_CTRL_SCORE) { inner = sc4[r]; dinner = tsd; }
                else if (d.ctrl_kind == SDES_CTRL_LERP) { inner = torch_lerp(pscore, sc4[r], lerp_w); dinner = (1.0f - lerp_w) * (-pi4[r]) + lerp_w * tsd; }
                else if (d.ctrl_kind == SDES_CTRL_LERP_PRIOR) { inner = (1.0f - lerp_w) * pscore; dinner = (1.0f - lerp_w) * (-pi4[r]); }
                else if (d.ctrl_kind == SDES_CTRL_LERP_TARGET) { inner = lerp_w * sc4[r]; dinner = lerp_w * tsd; }
                const float ungated = has_part ? outer * clipf(inner, cs) : 0.f;
                const float part = ungated * gate;
                const float dpart = (has_part && !score_detached && fabsf(inner) <= cs) ? outer * gate * dinner : 0.f;
                const float g = clipf(nn4[r], c.cm) + part;
                const float gm = c.ref_ctrl ? g - c.sigma * pscore : g;
                const float qv = wb * (cq * gm + ci * e[r]);
                const float delta = qv + a4[r] * Bc;
                an[r] = a4[r] * A + dpart * delta + (c.ref_ctrl ? qv * c.sigma * pi4[r] : 0.f);
                dm[r] = fabsf(nn4[r]) <= c.cm ? delta : 0.f;
                qg = fmaf(delta, ungated, qg);
            }
        }
        *reinterpret_cast<float2*>(ar + 2 * q) = make_float2(an[0], an[2]);
        *reinterpret_cast<float2*>(ar + Hp + 2 * q) = make_float2(an[1], an[3]);
        img_store_pair_g(a.dnn_img, a.pc, Hp, row, 0, 2 * q, dm[0], dm[2]);
        img_store_pair_g(a.dnn_img, a.pc, Hp, row, 1, 2 * q, dm[1], dm[3]);
    }
    if (a.qg != nullptr) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) qg += __shfl_xor_sync(0xffffffffu, qg, o);
        if (lane == 0) a.qg[row] = valid ? qg : 0.f;
    }
}

// terminal adjoint a_T = w (grad log p_ref(x_T) - 1[|log rho(x_T)| <= clip_target] grad log rho(x_T))   (losses/oc.py:225, :337, :449-450)
struct AdjWideInitArgs {
    SdesRolloutDesc d;
    const float *w, *xst, *logp, *sc_T, *vec_ref;
    float* adj;
    int P;
    int64_t Bp;
};

__global__ void __launch_bounds__(256) adj_wide_init_kernel(const AdjWideInitArgs a) {
    const SdesRolloutDesc& d = a.d;
    const int64_t n = a.Bp * a.P;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / a.P;
        const int p = (int)(e - row * a.P);
        float v = 0.f;
        if (row < d.batch) {
            const float lp = a.logp[row];
            float t = fabsf(lp) <= d.clip_target ? -a.sc_T[e] : 0.f;   // padded planar dims hold score 0
            if (d.loss_kind != SDES_LOSS_TIME_REVERSAL) t += (a.vec_ref[p] - a.xst[e]) * a.vec_ref[a.P + p];
            v = a.w[row] * t;
        }
        a.adj[e] = v;
    }
}

// d loss / d gate(s) = 1[|gate| < clip] sum_b w_b q[s][b]   (q written by the wide forward in keep mode; w = NULL: plain sum)
__global__ void __launch_bounds__(256) qgate_reduce_kernel(const float* __restrict__ q, const float* __restrict__ w, const float* __restrict__ gate,
                                                           int64_t B, int64_t Bp, float clip_model, float* __restrict__ out, int gate_stride = 1) {
    __shared__ float s_red[8];
    const int s = blockIdx.x;
    float acc = 0.f;
    for (int64_t b = threadIdx.x; b < B; b += blockDim.x) acc = fmaf(w != nullptr ? w[b] : 1.0f, q[(int64_t)s * Bp + b], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; ++i) t += s_red[i];
        out[s] = fabsf(gate[(int64_t)s * gate_stride]) < clip_model ? t : 0.f;
    }
}

// ---- the two x-independent TimeEmbed networks (models/mlp.py:43-82): parameter gradients from the per-step
// cotangents d loss / d emb (T, 64) and d loss / d gate (T, gate_dim).  One block per time step re-evaluates the tiny
// network for its s_i keeping the pre-activations in shared memory, walks it backwards and adds into the gradient
// blob (T rows x ~17 k parameters of atomics: negligible next to the B*T-row passes above).
struct TeGradArgs {
    const float* blob;       // parameters
    float* grad;             // gradient blob (same layout)
    const float* ts;         // (T+1)
    const float* cot;        // (T, n_out) cotangent of the network output
    int64_t o_phase, o_hw[SDES_MAX_HIDDEN], o_hb[SDES_MAX_HIDDEN], o_ow, o_ob;
    int n_hidden, n_out;
    int64_t o_extra_bias;    // >= 0: also add the cotangent row to this bias (FourierMLP.input_embed.bias), n_out = C
};

__device__ __forceinline__ float gelu_erf_grad(float x) {
    // d/dx [x Phi(x)] = Phi(x) + x phi(x)
    return 0.5f * (1.0f + erff(x * 0.70710678118654752440f)) + x * 0.3989422804014327f * expf(-0.5f * x * x);
}

__global__ void __launch_bounds__(256) time_embed_grad_kernel(const TeGradArgs a) {
    __shared__ float feat[2 * C], arg_s[C], pre[SDES_MAX_HIDDEN][C], act[SDES_MAX_HIDDEN][C], delta[2 * C], delta2[2 * C];
    const int i = blockIdx.x, tid = threadIdx.x;
    const float s = a.ts[i];
    const float* blob = a.blob;
    if (tid < C) {
        const float arg = __fadd_rn(__fmul_rn(linspace_coeff(tid), s), blob[a.o_phase + tid]);
        arg_s[tid] = arg;
        feat[tid] = sinf(arg);
        feat[C + tid] = cosf(arg);
    }
    __syncthreads();
    // forward, keeping pre-activations
    for (int l = 0; l < a.n_hidden; ++l) {
        const int k_in = l == 0 ? 2 * C : C;
        const float* in = l == 0 ? feat : act[l - 1];
        if (tid < C) {
            const float* w = blob + a.o_hw[l] + (int64_t)tid * k_in;
            float acc = blob[a.o_hb[l] + tid];
            for (int k = 0; k < k_in; ++k) acc = fmaf(w[k], in[k], acc);
            pre[l][tid] = acc;
            act[l][tid] = gelu_erf(acc);
        }
        __syncthreads();
    }
    const float* h_last = act[a.n_hidden - 1];
    const float* cot = a.cot + (int64_t)i * a.n_out;
    // output layer: dW[n][k] += cot[n] h[k], db[n] += cot[n], delta[k] = sum_n cot[n] W[n][k]
    for (int e = tid; e < a.n_out * C; e += blockDim.x) {
        const int n = e / C, k = e % C;
        const float g = cot[n] * h_last[k];
        if (g != 0.f) atomicAdd(a.grad + a.o_ow + e, g);
    }
    for (int n = tid; n < a.n_out; n += blockDim.x) {
        if (cot[n] != 0.f) {
            atomicAdd(a.grad + a.o_ob + n, cot[n]);
            if (a.o_extra_bias >= 0) atomicAdd(a.grad + a.o_extra_bias + n, cot[n]);
        }
    }
    if (tid < C) {
        float acc = 0.f;
        for (int n = 0; n < a.n_out; ++n) acc = fmaf(cot[n], blob[a.o_ow + (int64_t)n * C + tid], acc);
        delta[tid] = acc;
    }
    __syncthreads();
    // hidden layers, last to first
    for (int l = a.n_hidden - 1; l >= 0; --l) {
        const int k_in = l == 0 ? 2 * C : C;
        const float* in = l == 0 ? feat : act[l - 1];
        if (tid < C) delta[tid] *= gelu_erf_grad(pre[l][tid]);
        __syncthreads();
        for (int e = tid; e < C * k_in; e += blockDim.x) {
            const int n = e / k_in, k = e % k_in;
            const float g = delta[n] * in[k];
            if (g != 0.f) atomicAdd(a.grad + a.o_hw[l] + e, g);
        }
        if (tid < C && delta[tid] != 0.f) atomicAdd(a.grad + a.o_hb[l] + tid, delta[tid]);
        if (tid < k_in) {
            float acc = 0.f;
            for (int n = 0; n < C; ++n) acc = fmaf(delta[n], blob[a.o_hw[l] + (int64_t)n * k_in + tid], acc);
            delta2[tid] = acc;
        }
        __syncthreads();
        if (tid < k_in) delta[tid] = delta2[tid];
        __syncthreads();
    }
    // features: [sin(arg), cos(arg)], arg = coeff * t + phase
    if (tid < C) {
        const float g = delta[tid] * cosf(arg_s[tid]) - delta[C + tid] * sinf(arg_s[tid]);
        if (g != 0.f) atomicAdd(a.grad + a.o_phase + tid, g);
    }
}

static cudaError_t launch_time_embed_grads(const KParams& kp, const SdesLvGradDesc& g, cudaStream_t stream, int64_t& launches) {
    const SdesRolloutDesc& d = kp.d;
    TeGradArgs a;
    a.blob = d.params; a.grad = g.grad_params; a.ts = d.ts;
    a.cot = g.grad_emb; a.o_phase = kp.bl.te_phase; a.o_ow = kp.bl.te_out_w; a.o_ob = kp.bl.te_out_b;
    for (int l = 0; l < SDES_MAX_HIDDEN; ++l) { a.o_hw[l] = kp.bl.te_h_w[l]; a.o_hb[l] = kp.bl.te_h_b[l]; }
    a.n_hidden = d.te_hidden; a.n_out = C; a.o_extra_bias = kp.bl.in_b;
    time_embed_grad_kernel<<<d.n_steps, 256, 0, stream>>>(a);
    ++launches;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    if ((d.flags & SDES_F_HAS_GATE) && g.grad_gate != nullptr && d.ctrl_kind != SDES_CTRL_CLIPPED) {
        a.cot = g.grad_gate; a.o_phase = kp.bl.g_phase; a.o_ow = kp.bl.g_out_w; a.o_ob = kp.bl.g_out_b;
        for (int l = 0; l < SDES_MAX_HIDDEN; ++l) { a.o_hw[l] = kp.bl.g_h_w[l]; a.o_hb[l] = kp.bl.g_h_b[l]; }
        a.n_hidden = d.gate_hidden; a.n_out = d.gate_dim; a.o_extra_bias = -1;
        time_embed_grad_kernel<<<d.n_steps, 256, 0, stream>>>(a);
        ++launches;
        e = cudaGetLastError();
    }
    return e;
}

template <int DPAD>
static cudaError_t launch_cot_t(const CotArgs& a, int m_tiles, cudaStream_t stream) {
    const int K2 = (a.kp.d.n_components + 1) & ~1;
    const size_t smem = (2 * (size_t)K2 * DPAD + 64 + 2 * DPAD + 4 + DPAD) * sizeof(float);
    cudaError_t e = cudaFuncSetAttribute(cotangent_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cotangent_kernel<DPAD><<<m_tiles, 128, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int DPAD>
static cudaError_t launch_adj_t(const CotArgs& a, int tiles_per_step, bool init, cudaStream_t stream) {
    const int K2 = (a.kp.d.n_components + 1) & ~1;
    const size_t smem = (2 * (size_t)K2 * DPAD + 64 + 2 * (2 * DPAD + 4) + DPAD + 2 * 128 * (DPAD + 1)) * sizeof(float);
    static size_t attr_bytes = 0;  // per DPAD instantiation; the footprint also depends on the number of GMM components
    if (smem > attr_bytes) {
        cudaError_t e = cudaFuncSetAttribute(adj_init_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(adj_step_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_bytes = smem;
    }
    if (init) adj_init_kernel<DPAD><<<tiles_per_step, 128, smem, stream>>>(a);
    else adj_step_kernel<DPAD><<<tiles_per_step, 128, smem, stream>>>(a);
    return cudaGetLastError();
}

static cudaError_t launch_adj(const CotArgs& a, int tiles_per_step, bool init, cudaStream_t stream) {
    switch (a.kp.ws.dpad) {
        case 4: return launch_adj_t<4>(a, tiles_per_step, init, stream);
        case 8: return launch_adj_t<8>(a, tiles_per_step, init, stream);
        case 12: return launch_adj_t<12>(a, tiles_per_step, init, stream);
        case 16: return launch_adj_t<16>(a, tiles_per_step, init, stream);
        case 32: return launch_adj_t<32>(a, tiles_per_step, init, stream);
        case 52: return launch_adj_t<52>(a, tiles_per_step, init, stream);
        case 64: return launch_adj_t<64>(a, tiles_per_step, init, stream);
    }
    return cudaErrorInvalidValue;
}

static cudaError_t launch_cot(const CotArgs& a, int m_tiles, cudaStream_t stream) {
    switch (a.kp.ws.dpad) {
        case 4: return launch_cot_t<4>(a, m_tiles, stream);
        case 8: return launch_cot_t<8>(a, m_tiles, stream);
        case 12: return launch_cot_t<12>(a, m_tiles, stream);
        case 16: return launch_cot_t<16>(a, m_tiles, stream);
        case 32: return launch_cot_t<32>(a, m_tiles, stream);
        case 52: return launch_cot_t<52>(a, m_tiles, stream);
        case 64: return launch_cot_t<64>(a, m_tiles, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace grad

// ------------------------------------------------------------------------------ host side
using namespace grad;

void launch_prepare(const KParams& p, cudaStream_t stream);

// ---- wide engine gradient: scratch layout after the forward plan (WideGradView::grad_base)
struct WideGradScratch {
    Lin b_h[SDES_MAX_HIDDEN], b_out, b_in;
    int chunk_steps;
    int64_t a_img[SDES_MAX_HIDDEN + 1], gp_img[SDES_MAX_HIDDEN + 1], nn, dnn_img, dh_img[2], total;
    int64_t dh_all[SDES_MAX_HIDDEN + 1], adj;  // kl sweep: one cotangent image per layer kept until the chunk's wgrad, fp32 adjoint
};

static void wide_scratch(const WideGradView& v, WideGradScratch& sc, bool bptt) {
    int64_t rows = ((int64_t)1 << 27) / v.P;          // keep nn (rows x P fp32) at <= 512 MB
    sc.chunk_steps = (int)(rows / v.Bp);
    if (sc.chunk_steps < 1) sc.chunk_steps = 1;
    if (sc.chunk_steps > v.T) sc.chunk_steps = v.T;
    const int64_t m_tiles = (int64_t)sc.chunk_steps * v.m_tiles;
    int64_t o = v.grad_base;
    auto take = [&](int64_t bytes) { int64_t r = o; o = align256(o + bytes); return r; };
    for (int l = 0; l < v.nh; ++l) {
        set_tiling(sc.b_h[l], C, C);
        sc.b_h[l].w_off = take(lin_image_bytes(sc.b_h[l]));
        sc.b_h[l].b_off = -1;
    }
    set_tiling(sc.b_out, C, v.P);
    sc.b_out.w_off = take(lin_image_bytes(sc.b_out));
    sc.b_out.b_off = -1;
    const int64_t img1 = m_tiles * A_BLOCK, imgp = img1 * v.pc;
    for (int l = 0; l <= v.nh; ++l) {
        sc.a_img[l] = take(img1);
        sc.gp_img[l] = take(img1);
    }
    sc.nn = take(m_tiles * 128 * (int64_t)v.P * 4);
    sc.dnn_img = take(imgp);
    sc.dh_img[0] = take(img1);
    sc.dh_img[1] = take(img1);
    if (bptt) {
        set_tiling(sc.b_in, v.P, C);   // W_in^T: the adjoint's share of the input layer (N = planar state width, K = 64)
        sc.b_in.w_off = take(lin_image_bytes(sc.b_in));
        sc.b_in.b_off = -1;
        for (int l = 0; l <= v.nh; ++l) sc.dh_all[l] = take(img1);
        sc.adj = take(v.Bp * (int64_t)v.P * 4);
    }
    sc.total = o;
}

int64_t lv_grad_wide_scratch_bytes(const WideGradView& v, bool bptt) {
    WideGradScratch sc;
    wide_scratch(v, sc, bptt);
    return sc.total - v.grad_base;
}

// Gradient for a wide-engine rollout that ran with SDES_F_KEEP_FOR_GRAD in the SAME workspace: the per-step state images
// are already tensor-core operands, the forward's weight images are still there; only W^T images are added.
int64_t launch_lv_grad_wide(const KParams& kp, const SdesLvGradDesc& g, const WideGradView& v, bool simt, cudaStream_t stream, cudaError_t* err,
                            bool bptt) {
    const SdesRolloutDesc& d = kp.d;
    WideGradScratch sc;
    wide_scratch(v, sc, bptt);
    uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
    auto F = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
    int64_t launches = 0;
    *err = cudaSuccess;
#define GRADW_CHECK(expr)                         \
    do {                                          \
        *err = (expr);                            \
        if (*err != cudaSuccess) return launches; \
    } while (0)
    const float* blob = d.params;
    auto image_t = [&](const Lin& l, const float* src, int src_ld, int N, int K, int k_planar) {
        ImgArgs ia;
        ia.src = src; ia.src_ld = src_ld; ia.N = N; ia.K = K; ia.transpose = 1; ia.n_planar = 0; ia.k_planar = k_planar; ia.Hp = v.Hp;
        ia.out = ws + l.w_off; ia.n_pad = l.n_pad; ia.tile_n = l.tile_n; ia.k_chunks = l.k_chunks;
        const int64_t groups = (int64_t)l.n_pad * l.k_chunks * 8;
        weight_image_kernel<<<(int)((groups + 255) / 256), 256, 0, stream>>>(ia);
        ++launches;
        return cudaGetLastError();
    };
    for (int l = 0; l < v.nh; ++l) GRADW_CHECK(image_t(sc.b_h[l], blob + kp.bl.h_w[l], C, C, C, 0));
    GRADW_CHECK(image_t(sc.b_out, blob + kp.bl.out_w, C, C, d.dim, 1));   // out[n][k = planar dim] = W_out[natural(k)][n]
    if (bptt) {  // W_in^T: out[n = planar dim][k] = W_in[k][natural(n)]
        ImgArgs ia;
        ia.src = blob + kp.bl.in_w; ia.src_ld = d.dim; ia.N = d.dim; ia.K = C; ia.transpose = 1; ia.n_planar = 1; ia.k_planar = 0; ia.Hp = v.Hp;
        ia.out = ws + sc.b_in.w_off; ia.n_pad = sc.b_in.n_pad; ia.tile_n = sc.b_in.tile_n; ia.k_chunks = sc.b_in.k_chunks;
        const int64_t groups = (int64_t)sc.b_in.n_pad * sc.b_in.k_chunks * 8;
        weight_image_kernel<<<(int)((groups + 255) / 256), 256, 0, stream>>>(ia);
        ++launches;
        GRADW_CHECK(cudaGetLastError());
    }
    GRADW_CHECK(cudaMemsetAsync(g.grad_params, 0, (size_t)d.n_params * 4, stream));
    GRADW_CHECK(cudaMemsetAsync(g.grad_emb, 0, (size_t)v.T * C * 4, stream));
    auto base_args = [&](const Lin& l) {
        LinArgs a;
        a.a_img = nullptr; a.a_mt_stride = A_BLOCK; a.w_img = ws + l.w_off; a.k_chunks = l.k_chunks; a.n_tiles = l.n_tiles; a.tile_n = l.tile_n;
        a.bias = l.b_off >= 0 ? F(l.b_off) : nullptr; a.bias_mt_div = 0; a.bias_mt_stride = 0; a.act = ACT_NONE;
        a.mask_img = nullptr; a.mask_mt_stride = 0; a.mul_img = nullptr; a.mul_mt_stride = 0; a.aux_img = nullptr; a.resid = nullptr;
        a.out_f32 = nullptr; a.ld_f32 = v.P; a.out_img = nullptr; a.out_mt_stride = A_BLOCK;
        return a;
    };
    static bool attr_set = false;
    if (!attr_set) {
        GRADW_CHECK(cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
        attr_set = true;
    }
    auto wgrad = [&](const uint8_t* d_img, int d_chunks, const uint8_t* a_img, int a_chunks, int m_tiles, float* dw, int ldw, int n_valid,
                     int k_valid, int n_planar, int k_planar) {
        WgradArgs wa;
        wa.d_img = d_img; wa.d_chunks = d_chunks; wa.a_img = a_img; wa.a_chunks = a_chunks; wa.m_tiles = m_tiles; wa.dw = dw; wa.ldw = ldw;
        wa.db = nullptr;
        wa.n_valid = n_valid; wa.k_valid = k_valid; wa.n_planar = n_planar; wa.k_planar = k_planar; wa.Hp = v.Hp;
        const int pairs = d_chunks * a_chunks;
        int splits = 296 / pairs;
        if (splits > m_tiles) splits = m_tiles;
        if (splits < 1) splits = 1;
        ++launches;
        if (simt) wgrad_simt_kernel<<<dim3(pairs, splits), 64, 0, stream>>>(wa);
        else wgrad_mma_kernel<<<dim3(pairs, splits), WG_THREADS, WG_SMEM, stream>>>(wa);
        return cudaGetLastError();
    };
    auto colsum = [&](const uint8_t* img, int n_chunks, int n_valid, float* out, int m_tiles, int mt_div, int64_t mt_stride, int planar) {
        colsum_kernel<<<m_tiles, 128, 0, stream>>>(img, n_chunks, n_valid, out, mt_div, mt_stride, planar, v.Hp);
        ++launches;
        return cudaGetLastError();
    };
    float* gp = g.grad_params;
    const int x_stride_blocks = v.pc;
    const bool want_gate = g.grad_gate != nullptr && (d.flags & SDES_F_HAS_GATE) && d.ctrl_kind != SDES_CTRL_CLIPPED;
    if (bptt) {  // terminal adjoint
        AdjWideInitArgs ia;
        ia.d = d; ia.w = g.w; ia.xst = F(v.xst); ia.logp = F(v.logp); ia.sc_T = F(v.sc_keep) + (int64_t)v.T * v.Bp * v.P;
        ia.vec_ref = F(v.vec_ref); ia.adj = F(sc.adj); ia.P = v.P; ia.Bp = v.Bp;
        adj_wide_init_kernel<<<148 * 4, 256, 0, stream>>>(ia);
        ++launches;
        GRADW_CHECK(cudaGetLastError());
    }
    const int n_chunks = (v.T + sc.chunk_steps - 1) / sc.chunk_steps;
    for (int ci = 0; ci < n_chunks; ++ci) {
        const int s0 = (bptt ? n_chunks - 1 - ci : ci) * sc.chunk_steps;   // the sweep walks the chunks backwards in time
        const int ns = (s0 + sc.chunk_steps <= v.T) ? sc.chunk_steps : v.T - s0;
        const int m_tiles = ns * v.m_tiles;
        const uint8_t* ximg = ws + v.ximg + (int64_t)s0 * v.ximg_slot;   // the chunk's state images are contiguous
        {
            LinArgs a = base_args(v.mlp_in);
            a.a_img = ximg; a.a_mt_stride = (int64_t)x_stride_blocks * A_BLOCK;
            a.bias = F(v.emb) + (int64_t)s0 * C; a.bias_mt_div = v.m_tiles; a.bias_mt_stride = C;
            a.act = ACT_GELU_GRAD; a.out_img = ws + sc.a_img[0]; a.aux_img = ws + sc.gp_img[0];
            GRADW_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
            for (int l = 0; l < v.nh; ++l) {
                a = base_args(v.mlp_h[l]);
                a.a_img = ws + sc.a_img[l]; a.act = ACT_GELU_GRAD; a.out_img = ws + sc.a_img[l + 1]; a.aux_img = ws + sc.gp_img[l + 1];
                GRADW_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
            }
            a = base_args(v.mlp_out);
            a.a_img = ws + sc.a_img[v.nh]; a.out_f32 = F(sc.nn);
            GRADW_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
        }
        if (!bptt) {
            CotWideArgs ca;
            ca.d = d; ca.tab = F(v.tab); ca.w = g.w; ca.nn = F(sc.nn); ca.dnn_img = ws + sc.dnn_img; ca.Hp = v.Hp; ca.P = v.P; ca.pc = v.pc;
            ca.s0 = s0; ca.Bp = v.Bp;
            cotangent_wide_kernel<<<m_tiles * 16, 256, 0, stream>>>(ca);
            ++launches;
            GRADW_CHECK(cudaGetLastError());
        } else {
            // reverse sweep over the chunk's steps: cotangent of the control from the adjoint, then the dgrad chain of THIS step's
            // row tiles (layer cotangent images kept per layer for the chunk's weight gradients), W_in^T accumulating into the adjoint
            for (int s = s0 + ns - 1; s >= s0; --s) {
                const int64_t t_off = (int64_t)(s - s0) * v.m_tiles;   // first row tile of this step inside the chunk buffers
                AdjWideArgs aa;
                aa.d = d; aa.tab = F(v.tab); aa.gate = F(v.gate); aa.w = g.w; aa.nn = F(sc.nn) + t_off * 128 * v.P;
                aa.sc = F(v.sc_keep) + (int64_t)s * v.Bp * v.P; aa.vec_prior = F(v.vec_prior); aa.gmm_h = F(v.gmm_h);
                aa.ximg = ws + v.ximg + (int64_t)s * v.ximg_slot; aa.adj = F(sc.adj);
                aa.dnn_img = ws + sc.dnn_img + t_off * v.pc * A_BLOCK;
                aa.qg = want_gate ? F(v.qgate) + (int64_t)s * v.Bp : nullptr;
                aa.Hp = v.Hp; aa.P = v.P; aa.pc = v.pc; aa.step = s; aa.Bp = v.Bp; aa.gflags = g.flags;
                adj_wide_step_kernel<<<(int)((v.Bp + 7) / 8), 256, 0, stream>>>(aa);
                ++launches;
                GRADW_CHECK(cudaGetLastError());
                LinArgs a = base_args(sc.b_out);
                a.a_img = aa.dnn_img; a.a_mt_stride = (int64_t)v.pc * A_BLOCK;
                a.mul_img = ws + sc.gp_img[v.nh] + t_off * A_BLOCK; a.mul_mt_stride = A_BLOCK;
                a.out_img = ws + sc.dh_all[v.nh] + t_off * A_BLOCK;
                GRADW_CHECK(launch_linear(a, v.m_tiles, simt, stream, launches));
                for (int l = v.nh - 1; l >= 0; --l) {
                    a = base_args(sc.b_h[l]);
                    a.a_img = ws + sc.dh_all[l + 1] + t_off * A_BLOCK;
                    a.mul_img = ws + sc.gp_img[l] + t_off * A_BLOCK; a.mul_mt_stride = A_BLOCK;
                    a.out_img = ws + sc.dh_all[l] + t_off * A_BLOCK;
                    GRADW_CHECK(launch_linear(a, v.m_tiles, simt, stream, launches));
                }
                a = base_args(sc.b_in);
                a.a_img = ws + sc.dh_all[0] + t_off * A_BLOCK;
                a.resid = F(sc.adj); a.out_f32 = F(sc.adj);
                GRADW_CHECK(launch_linear(a, v.m_tiles, simt, stream, launches));
            }
        }
        GRADW_CHECK(wgrad(ws + sc.dnn_img, v.pc, ws + sc.a_img[v.nh], 1, m_tiles, gp + kp.bl.out_w, C, d.dim, C, 1, 0));
        GRADW_CHECK(colsum(ws + sc.dnn_img, v.pc, d.dim, gp + kp.bl.out_b, m_tiles, 0, 0, 1));
        if (bptt) {
            for (int l = v.nh - 1; l >= 0; --l) {
                GRADW_CHECK(wgrad(ws + sc.dh_all[l + 1], 1, ws + sc.a_img[l], 1, m_tiles, gp + kp.bl.h_w[l], C, C, C, 0, 0));
                GRADW_CHECK(colsum(ws + sc.dh_all[l + 1], 1, C, gp + kp.bl.h_b[l], m_tiles, 0, 0, 0));
            }
            GRADW_CHECK(wgrad(ws + sc.dh_all[0], 1, ximg, v.pc, m_tiles, gp + kp.bl.in_w, d.dim, C, d.dim, 0, 1));
            GRADW_CHECK(colsum(ws + sc.dh_all[0], 1, C, g.grad_emb + (int64_t)s0 * C, m_tiles, v.m_tiles, C, 0));
            continue;
        }
        int cur = 0;
        {
            LinArgs a = base_args(sc.b_out);
            a.a_img = ws + sc.dnn_img; a.a_mt_stride = (int64_t)v.pc * A_BLOCK; a.mul_img = ws + sc.gp_img[v.nh]; a.mul_mt_stride = A_BLOCK;
            a.out_img = ws + sc.dh_img[cur];
            GRADW_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
        }
        for (int l = v.nh - 1; l >= 0; --l) {
            GRADW_CHECK(wgrad(ws + sc.dh_img[cur], 1, ws + sc.a_img[l], 1, m_tiles, gp + kp.bl.h_w[l], C, C, C, 0, 0));
            GRADW_CHECK(colsum(ws + sc.dh_img[cur], 1, C, gp + kp.bl.h_b[l], m_tiles, 0, 0, 0));
            LinArgs a = base_args(sc.b_h[l]);
            a.a_img = ws + sc.dh_img[cur]; a.mul_img = ws + sc.gp_img[l]; a.mul_mt_stride = A_BLOCK; a.out_img = ws + sc.dh_img[1 - cur];
            GRADW_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
            cur = 1 - cur;
        }
        GRADW_CHECK(wgrad(ws + sc.dh_img[cur], 1, ximg, v.pc, m_tiles, gp + kp.bl.in_w, d.dim, C, d.dim, 0, 1));
        GRADW_CHECK(colsum(ws + sc.dh_img[cur], 1, C, g.grad_emb + (int64_t)s0 * C, m_tiles, v.m_tiles, C, 0));
    }
    if (g.grad_gate != nullptr) {
        if ((d.flags & SDES_F_HAS_GATE) && d.ctrl_kind != SDES_CTRL_CLIPPED) {
            // lv: the forward left sum_j c_j part_j / gate per (step, row), weighted by w here; kl: the sweep left sum_j delta_j part_j / gate
            qgate_reduce_kernel<<<v.T, 256, 0, stream>>>(F(v.qgate), bptt ? nullptr : g.w, F(v.gate), v.B, v.Bp, d.clip_model, g.grad_gate);
            ++launches;
            GRADW_CHECK(cudaGetLastError());
        } else {
            GRADW_CHECK(cudaMemsetAsync(g.grad_gate, 0, (size_t)v.T * 4, stream));
        }
    }
    GRADW_CHECK(launch_time_embed_grads(kp, g, stream, launches));
#undef GRADW_CHECK
    return launches;
}

void wide_grad_view(const SdesRolloutDesc& d, wide::WideGradView& v);  // sdes_wide.cu

int64_t launch_lv_grad_wide_desc(const KParams& kp, const SdesLvGradDesc& g, bool simt, cudaStream_t stream, cudaError_t* err, bool bptt) {
    WideGradView v;
    wide_grad_view(kp.d, v);
    return launch_lv_grad_wide(kp, g, v, simt, stream, err, bptt);
}

size_t lv_grad_workspace_bytes(const SdesRolloutDesc& d, int64_t fused_bytes, int64_t chunk_rows) {
    GradPlan p;
    make_plan(d, chunk_rows, align256(fused_bytes), p);
    return (size_t)p.total;
}

// kp: descriptor with SDES_F_MLP_SIMT set (fp32 tables / target images of the fused prologue at the start of the
// workspace), its blob and workspace layouts.  Returns the number of kernel launches.
cudaError_t launch_kl_adjoint(const KParams& kp, const float* xs, const float* w, float* delta, uint32_t gflags, int sm_count,
                              cudaStream_t stream);  // sdes_adjoint.cu

static int64_t delta_floats(const SdesRolloutDesc& d) {
    if (d.flags & SDES_F_TRAJ_TILED) return (int64_t)d.n_steps * ((d.batch + 127) / 128) * mma_pad_dim(d.dim) * 128;
    return (int64_t)d.n_steps * d.batch * d.dim;
}

// simt: the thread-per-trajectory sweep of sdes_adjoint.cu (needs the delta buffer); otherwise the sweep runs on the
// tensor cores inside the chunk loop (needs the adjoint, W_in^T and one delta_h image per layer)
size_t kl_grad_workspace_bytes(const SdesRolloutDesc& d, int64_t fused_bytes, int64_t chunk_rows, bool simt) {
    GradPlan p;
    make_plan(d, chunk_rows, align256(fused_bytes), p, !simt);
    return (size_t)(align256(p.total) + (simt ? delta_floats(d) * 4 : 0));
}

// bptt = false: lv (closed-form cotangent).  bptt = true: kl / kl_ito.  With simt the reverse sweep of sdes_adjoint.cu first
// writes the control cotangent of every (trajectory, step) after the plan's scratch, then the same GEMM passes consume it;
// otherwise the sweep is part of the chunk loop (chunks in reverse time order): per step adj_step_kernel + the dgrad GEMMs
// on that step's row tiles, the weight gradients once per chunk.
int64_t launch_lv_grad(const KParams& kp, const SdesLvGradDesc& g, int64_t fused_bytes, bool simt, cudaStream_t stream, cudaError_t* err,
                       bool bptt, int sm_count) {
    const SdesRolloutDesc& d = kp.d;
    GradPlan p;
    const bool bptt_tc = bptt && !simt;
    make_plan(d, g.chunk_rows, align256(fused_bytes), p, bptt_tc);
    uint8_t* ws = reinterpret_cast<uint8_t*>(d.workspace);
    auto F = [&](int64_t off) { return reinterpret_cast<float*>(ws + off); };
    int64_t launches = 0;
    *err = cudaSuccess;
#define GRAD_CHECK(expr)                          \
    do {                                          \
        *err = (expr);                            \
        if (*err != cudaSuccess) return launches; \
    } while (0)
    // ---- prologue: per-step tables and target images (fused prologue), operand images of W and W^T
    launch_prepare(kp, stream);
    ++launches;
    GRAD_CHECK(cudaGetLastError());
    const float* delta = nullptr;
    if (bptt && simt) {
        float* dl = F(align256(p.total));
        GRAD_CHECK(launch_kl_adjoint(kp, g.xs, g.w, dl, g.flags, sm_count, stream));
        ++launches;
        delta = dl;
    }
    const float* blob = d.params;
    const float* fws = reinterpret_cast<const float*>(d.workspace);
    embb_kernel<<<(p.T * C + 255) / 256, 256, 0, stream>>>(fws + kp.ws.emb, blob + kp.bl.in_b, F(p.embb), p.T * C, F(p.ones));
    ++launches;
    GRAD_CHECK(cudaGetLastError());
    auto image = [&](const Lin& l, const float* src, int src_ld, int N, int K, int transpose) {
        ImgArgs ia;
        ia.src = src; ia.src_ld = src_ld; ia.N = N; ia.K = K; ia.transpose = transpose; ia.n_planar = 0; ia.k_planar = 0; ia.Hp = 64;
        ia.out = ws + l.w_off; ia.n_pad = l.n_pad; ia.tile_n = l.tile_n; ia.k_chunks = l.k_chunks;
        const int64_t groups = (int64_t)l.n_pad * l.k_chunks * 8;
        weight_image_kernel<<<(int)((groups + 255) / 256), 256, 0, stream>>>(ia);
        ++launches;
        return cudaGetLastError();
    };
    auto bias = [&](const Lin& l, const float* src, int n) {
        pad_bias_kernel<<<(l.n_pad + 255) / 256, 256, 0, stream>>>(src, n, F(l.b_off), l.n_pad);
        ++launches;
        return cudaGetLastError();
    };
    // lv with a scalar gate and the forward's gate_cot: the gate gradient is one reduction over (T, B), no target score here
    const bool gate_from_fwd = !bptt && g.gate_cot != nullptr && g.grad_gate != nullptr && (d.flags & SDES_F_HAS_GATE) &&
                               d.ctrl_kind != SDES_CTRL_CLIPPED && d.gate_dim == 1;
    const bool gate_wanted = g.grad_gate != nullptr && (d.flags & SDES_F_HAS_GATE) && d.ctrl_kind != SDES_CTRL_CLIPPED;
    // the whole pass as one persistent kernel (sdes_grad_fused.cuh) whenever its shapes allow; SDES_GRAD_LAYERWISE_SWEEP keeps
    // the layer-by-layer GEMM passes (A/B measurements, cross-check)
    const bool fused_lv = !bptt && !simt && lv_fused_supported(d) && !(g.flags & SDES_GRAD_LAYERWISE_SWEEP) && (gate_from_fwd || !gate_wanted);
    // kl / kl_ito: the same kernel walks each row tile backwards in time with the adjoint in registers, whenever the score
    // part's x-derivative is local per dimension (no Hessian-vector product of the target: constant / detached target score, or
    // a control without one), the forward kept the ungated score part (score_keep) and the gate is scalar or absent
    const bool kl_target_in_ctrl = d.ctrl_kind == SDES_CTRL_SCORE || d.ctrl_kind == SDES_CTRL_LERP || d.ctrl_kind == SDES_CTRL_LERP_TARGET;
    const bool kl_hvp = kl_target_in_ctrl && !(g.flags & (SDES_GRAD_TARGET_SCORE_CONST | SDES_GRAD_SCORE_DETACHED));
    // a Hessian that is diagonal (one Gaussian, multi-well) is local per dimension too; the funnel's needs the whole row in one
    // epilogue thread (d <= 16), and so does a per-dimension gate's reduction
    const bool row_in_thread = mma_pad_dim(d.dim) <= 16;
    const bool hvp_ok = !kl_hvp || (d.target_kind == SDES_TARGET_GMM && d.n_components == 1) || d.target_kind == SDES_TARGET_MULTIWELL ||
                        (d.target_kind == SDES_TARGET_FUNNEL && row_in_thread);
    const bool fused_kl = bptt_tc && lv_fused_supported(d) && !(g.flags & SDES_GRAD_LAYERWISE_SWEEP) && hvp_ok &&
                          (d.ctrl_kind == SDES_CTRL_CLIPPED || g.score_keep != nullptr) && (!gate_wanted || d.gate_dim == 1 || row_in_thread);
    const bool fused_any = fused_lv || fused_kl;
    GRAD_CHECK(image(p.f_in, blob + kp.bl.in_w, d.dim, C, d.dim, 0));
    for (int l = 0; l < p.nh; ++l) {
        GRAD_CHECK(image(p.f_h[l], blob + kp.bl.h_w[l], C, C, C, 0));
        GRAD_CHECK(bias(p.f_h[l], blob + kp.bl.h_b[l], C));
        if (!fused_any) GRAD_CHECK(image(p.b_h[l], blob + kp.bl.h_w[l], C, C, C, 1));
    }
    GRAD_CHECK(image(p.f_out, blob + kp.bl.out_w, C, d.dim, C, 0));
    GRAD_CHECK(bias(p.f_out, blob + kp.bl.out_b, d.dim));
    if (!fused_any) GRAD_CHECK(image(p.b_out, blob + kp.bl.out_w, C, C, d.dim, 1));
    if (bptt_tc && !fused_kl) GRAD_CHECK(image(p.b_in, blob + kp.bl.in_w, d.dim, d.dim, C, 1));
    GRAD_CHECK(cudaMemsetAsync(g.grad_params, 0, (size_t)d.n_params * 4, stream));
    GRAD_CHECK(cudaMemsetAsync(g.grad_emb, 0, (size_t)p.T * C * 4, stream));
    if (g.grad_gate != nullptr) GRAD_CHECK(cudaMemsetAsync(g.grad_gate, 0, (size_t)p.T * (d.gate_dim > 0 ? d.gate_dim : 1) * 4, stream));

    auto base_args = [&](const Lin& l) {
        LinArgs a;
        a.a_img = nullptr; a.a_mt_stride = A_BLOCK; a.w_img = ws + l.w_off; a.k_chunks = l.k_chunks; a.n_tiles = l.n_tiles; a.tile_n = l.tile_n;
        a.bias = l.b_off >= 0 ? F(l.b_off) : nullptr; a.bias_mt_div = 0; a.bias_mt_stride = 0; a.act = ACT_NONE;
        a.mask_img = nullptr; a.mask_mt_stride = 0; a.mul_img = nullptr; a.mul_mt_stride = 0; a.aux_img = nullptr; a.resid = nullptr;
        a.out_f32 = nullptr; a.ld_f32 = p.P; a.out_img = nullptr; a.out_mt_stride = A_BLOCK;
        return a;
    };
    static bool attr_set = false;
    if (!attr_set) {
        GRAD_CHECK(cudaFuncSetAttribute(wgrad_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)WG_SMEM));
        attr_set = true;
    }
    auto colsum = [&](const uint8_t* img, int n_chunks, int n_valid, float* out, int m_tiles, int mt_div, int64_t mt_stride) {
        colsum_kernel<<<m_tiles, 128, 0, stream>>>(img, n_chunks, n_valid, out, mt_div, mt_stride, 0, 64);
        ++launches;
        return cudaGetLastError();
    };
    // dW += delta^T a over the chunk's rows; db (may be NULL) += column sums of delta — on the tensor cores inside the same
    // kernel (one extra N = 16 MMA per k-step against a block of ones), as a separate pass on the CUDA-core engine
    auto wgrad = [&](const uint8_t* d_img, int d_chunks, const uint8_t* a_img, int a_chunks, int m_tiles, float* dw, int ldw, int n_valid,
                     int k_valid, float* db) {
        WgradArgs wa;
        wa.d_img = d_img; wa.d_chunks = d_chunks; wa.a_img = a_img; wa.a_chunks = a_chunks; wa.m_tiles = m_tiles; wa.dw = dw; wa.ldw = ldw;
        wa.db = simt ? nullptr : db;
        wa.n_valid = n_valid; wa.k_valid = k_valid; wa.n_planar = 0; wa.k_planar = 0; wa.Hp = 64;
        const int pairs = d_chunks * a_chunks;
        int splits = 296 / pairs;
        if (splits > m_tiles) splits = m_tiles;
        if (splits < 1) splits = 1;
        ++launches;
        if (simt) wgrad_simt_kernel<<<dim3(pairs, splits), 64, 0, stream>>>(wa);
        else wgrad_mma_kernel<<<dim3(pairs, splits), WG_THREADS, WG_SMEM, stream>>>(wa);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess && simt && db != nullptr) e = colsum(d_img, d_chunks, n_valid, db, m_tiles, 0, 0);
        return e;
    };

    const int tiles_per_step = (int)(p.Bp / 128);
    CotArgs ca;
    ca.kp = kp; ca.xs = g.xs; ca.w = g.w; ca.nn = F(p.nn); ca.ones = F(p.ones); ca.delta = delta; ca.dnn_img = ws + p.dnn_img;
    ca.grad_gate = g.grad_gate; ca.P = p.P; ca.pc = p.pc; ca.s0 = 0; ca.Bp = p.Bp;
    ca.adj = bptt_tc ? F(p.adj) : nullptr; ca.gflags = g.flags; ca.step = 0;
    ca.skip_gate = gate_from_fwd ? 1 : 0;
    // the fused dgrad chain serves the fused engines' shapes (P = 64: every transposed layer is one 64 x 64 block);
    // SDES_GRAD_LAYERWISE_SWEEP keeps the layer-by-layer launches (A/B measurements, cross-check)
    const bool use_chain = bptt_tc && !(g.flags & SDES_GRAD_LAYERWISE_SWEEP) && p.pc == 1 && p.P == 64;
    if (bptt_tc) {  // a_T: the terminal cost's gradient
        GRAD_CHECK(launch_adj(ca, tiles_per_step, true, stream));
        ++launches;
    }
    if (fused_any) {
        FusedLvArgs fa;
        fa.adj_init = fused_kl ? F(p.adj) : nullptr;
        fa.kl_flags = fused_kl ? reinterpret_cast<uint32_t*>(ws + p.kl_flags) : nullptr;
        if (fused_kl) GRAD_CHECK(cudaMemsetAsync(ws + p.kl_flags, 0, (size_t)tiles_per_step * 4 + 4096, stream));
        fa.score_keep = g.score_keep; fa.gate = fws + kp.ws.gate; fa.gate_stride = kp.ws.dpad;
        fa.prior_loc = fws + kp.ws.prior; fa.prior_iv = fws + kp.ws.prior + kp.ws.dpad;
        fa.grad_gate = (fused_kl && gate_wanted) ? g.grad_gate : nullptr; fa.gflags = g.flags;
        fa.watch_all = getenv("SDES_FL_DEBUG") != nullptr;
        fa.gmm_h = fws + kp.ws.gmm_h; fa.hvp = (fused_kl && kl_hvp) ? 1 : 0;
        fa.d = d; fa.tab = fws + kp.ws.tab; fa.xs = g.xs; fa.w = g.w; fa.embb = F(p.embb);
        fa.nh = p.nh; fa.T = p.T; fa.tiles_per_step = tiles_per_step; fa.grad_emb = g.grad_emb;
        float* gp = g.grad_params;
        const int Lf = p.nh + 2;
        for (int l = 0; l < FL_MAX_LAYERS; ++l) {
            fa.w_img[l] = nullptr; fa.bias[l] = nullptr; fa.dw[l] = nullptr; fa.db[l] = nullptr; fa.ldw[l] = fa.n_valid[l] = fa.k_valid[l] = 0;
        }
        fa.w_img[0] = ws + p.f_in.w_off; fa.dw[0] = gp + kp.bl.in_w; fa.ldw[0] = d.dim; fa.n_valid[0] = C; fa.k_valid[0] = d.dim;
        for (int l = 0; l < p.nh; ++l) {
            fa.w_img[1 + l] = ws + p.f_h[l].w_off; fa.bias[1 + l] = F(p.f_h[l].b_off);
            fa.dw[1 + l] = gp + kp.bl.h_w[l]; fa.db[1 + l] = gp + kp.bl.h_b[l]; fa.ldw[1 + l] = C; fa.n_valid[1 + l] = C; fa.k_valid[1 + l] = C;
        }
        fa.w_img[Lf - 1] = ws + p.f_out.w_off; fa.bias[Lf - 1] = F(p.f_out.b_off);
        fa.dw[Lf - 1] = gp + kp.bl.out_w; fa.db[Lf - 1] = gp + kp.bl.out_b; fa.ldw[Lf - 1] = C; fa.n_valid[Lf - 1] = d.dim; fa.k_valid[Lf - 1] = C;
        fa.timeline = nullptr;
        if (getenv("SDES_FL_TIMELINE")) {  // debugging: clock64 stamps of CTA 0 (control warp and one epilogue warp), printed per launch
            static unsigned long long* tl = nullptr;
            if (tl == nullptr) cudaMalloc(&tl, 1024 * sizeof(unsigned long long));
            cudaMemsetAsync(tl, 0, 1024 * sizeof(unsigned long long), stream);
            fa.timeline = tl;
        }
        GRAD_CHECK(launch_lv_fused(fa, fused_kl, sm_count > 0 ? sm_count : 148, stream));
        ++launches;
        if (fa.timeline != nullptr) {
            std::vector<unsigned long long> h(1024);
            cudaStreamSynchronize(stream);
            cudaMemcpy(h.data(), fa.timeline, 1024 * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
            const unsigned long long t0 = h[0];
            fprintf(stderr, "TIMELINE control:");
            for (int i = 0; i < 512 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - t0);
            fprintf(stderr, "\nTIMELINE epilogue:");
            for (int i = 512; i < 1024 && h[i]; ++i) fprintf(stderr, " %llu", h[i] - t0);
            fprintf(stderr, "\n");
        }
        if (fused_kl && getenv("SDES_FL_DEBUG")) {
            std::vector<uint32_t> rec(1024);
            cudaStreamSynchronize(stream);
            cudaMemcpy(rec.data(), ws + p.kl_flags + (size_t)tiles_per_step * 4, 4096, cudaMemcpyDeviceToHost);
            fprintf(stderr, "lv_fused kl watchdog records: %u\n", rec[0]);
            for (uint32_t i = 0; i < rec[0] && i < 200; ++i)
                    fprintf(stderr, "  REC cta %u who %u code %u item %u\n", rec[4 + 4 * i], rec[5 + 4 * i], rec[6 + 4 * i], rec[7 + 4 * i]);
        }
    }
    for (int chi = 0; chi < (fused_any ? 0 : p.n_chunks); ++chi) {
        const int ch = bptt_tc ? p.n_chunks - 1 - chi : chi;  // the sweep walks the chunks backwards in time
        const int s0 = ch * p.chunk_steps;
        const int ns = (s0 + p.chunk_steps <= p.T) ? p.chunk_steps : p.T - s0;
        const int m_tiles = ns * tiles_per_step;
        pack_rows_kernel<<<m_tiles, 128, 0, stream>>>(d, g.xs, p.Bp, p.pc, s0, ws + p.ximg);
        ++launches;
        GRAD_CHECK(cudaGetLastError());
        // ---- forward (models/mlp.py:114-122), keeping GELU(h) and GELU'(h) of every layer
        {
            LinArgs a = base_args(p.f_in);
            a.a_img = ws + p.ximg; a.a_mt_stride = (int64_t)p.pc * A_BLOCK;
            a.bias = F(p.embb) + (int64_t)s0 * C; a.bias_mt_div = tiles_per_step; a.bias_mt_stride = C;
            a.act = ACT_GELU_GRAD; a.out_img = ws + p.a_img[0]; a.aux_img = ws + p.gp_img[0];
            GRAD_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
            for (int l = 0; l < p.nh; ++l) {
                a = base_args(p.f_h[l]);
                a.a_img = ws + p.a_img[l]; a.act = ACT_GELU_GRAD; a.out_img = ws + p.a_img[l + 1]; a.aux_img = ws + p.gp_img[l + 1];
                GRAD_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
            }
            a = base_args(p.f_out);
            a.a_img = ws + p.a_img[p.nh]; a.out_f32 = F(p.nn);
            GRAD_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
        }
        float* gp = g.grad_params;
        ca.s0 = s0;
        if (bptt_tc) {
            // ---- reverse sweep over the chunk's steps: delta_s, then the dgrad chain on this step's row tiles down to x
            for (int s = s0 + ns - 1; s >= s0; --s) {
                ca.step = s;
                GRAD_CHECK(launch_adj(ca, tiles_per_step, false, stream));
                ++launches;
                const int64_t t0 = (int64_t)(s - s0) * tiles_per_step;
                if (use_chain) {  // the whole dgrad chain of this step in one launch
                    ChainArgs c;
                    c.dnn_img = ws + p.dnn_img + t0 * A_BLOCK;
                    c.n_layers = p.nh + 2;
                    c.m_tiles = tiles_per_step;
                    c.adj = F(p.adj);
                    c.w_img[0] = ws + p.b_out.w_off;
                    for (int l = p.nh - 1, i = 1; l >= 0; --l, ++i) c.w_img[i] = ws + p.b_h[l].w_off;
                    c.w_img[p.nh + 1] = ws + p.b_in.w_off;
                    for (int l = p.nh, i = 0; l >= 0; --l, ++i) {
                        c.gp_img[i] = ws + p.gp_img[l] + t0 * A_BLOCK;
                        c.dh_img[i] = ws + p.dh_all[l] + t0 * A_BLOCK;
                    }
                    GRAD_CHECK(launch_dgrad_chain(c, sm_count > 0 ? sm_count : 148, stream));
                    ++launches;
                    continue;
                }
                LinArgs a = base_args(p.b_out);
                a.a_img = ws + p.dnn_img + t0 * p.pc * A_BLOCK; a.a_mt_stride = (int64_t)p.pc * A_BLOCK;
                a.mul_img = ws + p.gp_img[p.nh] + t0 * A_BLOCK; a.mul_mt_stride = A_BLOCK;
                a.out_img = ws + p.dh_all[p.nh] + t0 * A_BLOCK;
                GRAD_CHECK(launch_linear(a, tiles_per_step, simt, stream, launches));
                for (int l = p.nh - 1; l >= 0; --l) {
                    a = base_args(p.b_h[l]);
                    a.a_img = ws + p.dh_all[l + 1] + t0 * A_BLOCK; a.mul_img = ws + p.gp_img[l] + t0 * A_BLOCK; a.mul_mt_stride = A_BLOCK;
                    a.out_img = ws + p.dh_all[l] + t0 * A_BLOCK;
                    GRAD_CHECK(launch_linear(a, tiles_per_step, simt, stream, launches));
                }
                a = base_args(p.b_in);  // a_s += delta_h0 W_in: fp32 accumulation into the adjoint
                a.a_img = ws + p.dh_all[0] + t0 * A_BLOCK; a.out_f32 = F(p.adj); a.resid = F(p.adj); a.ld_f32 = p.P;
                GRAD_CHECK(launch_linear(a, tiles_per_step, simt, stream, launches));
            }
            // ---- weight gradients of the chunk from the delta images the sweep left behind
            GRAD_CHECK(wgrad(ws + p.dnn_img, p.pc, ws + p.a_img[p.nh], 1, m_tiles, gp + kp.bl.out_w, C, d.dim, C, gp + kp.bl.out_b));
            for (int l = p.nh - 1; l >= 0; --l)
                GRAD_CHECK(wgrad(ws + p.dh_all[l + 1], 1, ws + p.a_img[l], 1, m_tiles, gp + kp.bl.h_w[l], C, C, C, gp + kp.bl.h_b[l]));
            GRAD_CHECK(wgrad(ws + p.dh_all[0], 1, ws + p.ximg, p.pc, m_tiles, gp + kp.bl.in_w, d.dim, C, d.dim, nullptr));
            GRAD_CHECK(colsum(ws + p.dh_all[0], 1, C, g.grad_emb + (int64_t)s0 * C, m_tiles, tiles_per_step, C));
            continue;
        }
        // ---- output cotangent and gate gradient
        GRAD_CHECK(launch_cot(ca, m_tiles, stream));
        ++launches;
        // ---- out layer: dW_out += delta_nn^T a_nh, db_out += colsum(delta_nn); delta_h[nh] = (delta_nn W_out) * GELU'(h_nh)
        GRAD_CHECK(wgrad(ws + p.dnn_img, p.pc, ws + p.a_img[p.nh], 1, m_tiles, gp + kp.bl.out_w, C, d.dim, C, gp + kp.bl.out_b));
        int cur = 0;
        {
            LinArgs a = base_args(p.b_out);
            a.a_img = ws + p.dnn_img; a.a_mt_stride = (int64_t)p.pc * A_BLOCK; a.mul_img = ws + p.gp_img[p.nh]; a.mul_mt_stride = A_BLOCK;
            a.out_img = ws + p.dh_img[cur];
            GRAD_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
        }
        for (int l = p.nh - 1; l >= 0; --l) {
            // delta_h[l+1] is in dh_img[cur]: gradients of hidden layer l, then delta_h[l]
            GRAD_CHECK(wgrad(ws + p.dh_img[cur], 1, ws + p.a_img[l], 1, m_tiles, gp + kp.bl.h_w[l], C, C, C, gp + kp.bl.h_b[l]));
            LinArgs a = base_args(p.b_h[l]);
            a.a_img = ws + p.dh_img[cur]; a.mul_img = ws + p.gp_img[l]; a.mul_mt_stride = A_BLOCK; a.out_img = ws + p.dh_img[1 - cur];
            GRAD_CHECK(launch_linear(a, m_tiles, simt, stream, launches));
            cur = 1 - cur;
        }
        // ---- input layer: dW_in += delta_h0^T x ; d emb[s] += per-step column sums of delta_h0
        GRAD_CHECK(wgrad(ws + p.dh_img[cur], 1, ws + p.ximg, p.pc, m_tiles, gp + kp.bl.in_w, d.dim, C, d.dim, nullptr));
        GRAD_CHECK(colsum(ws + p.dh_img[cur], 1, C, g.grad_emb + (int64_t)s0 * C, m_tiles, tiles_per_step, C));
    }
    if (gate_from_fwd) {
        qgate_reduce_kernel<<<p.T, 256, 0, stream>>>(g.gate_cot, g.w, fws + kp.ws.gate, d.batch, d.batch, d.clip_model, g.grad_gate, kp.ws.dpad);
        ++launches;
        GRAD_CHECK(cudaGetLastError());
    }
    GRAD_CHECK(launch_time_embed_grads(kp, g, stream, launches));
#undef GRAD_CHECK
    return launches;
}

}  // namespace sdes
