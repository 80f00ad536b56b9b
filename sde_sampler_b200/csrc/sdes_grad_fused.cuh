// sdes_grad_fused.cuh — the gradient of the control MLP as ONE persistent kernel (`loss.backward()` of
// `Trainable.step`, solver/base.py:404-407, loss.method = lv | kl | kl_ito).  lv: per 128-row tile of (step, trajectory) rows
//     replayed forward (models/mlp.py:114-122)  ->  output cotangent  ->  dgrad chain  ->  weight-gradient MMAs
// with every activation / delta operand living in shared memory only and the weight gradients accumulating in TMEM for
// the whole launch.  Nothing but the stored trajectory (200 B per row) is read from HBM and nothing but the final
// gradients is written — the layer-by-layer path of sdes_grad.cu moves ~45 GB of operand images per training step.
//
// One CTA per SM, 18 warps: warp 0 issues every MMA (converged, one elected lane), warp 1 owns the TMEM allocation, warps 2-17 are the
// epilogue: TMEM lane quadrant = warp % 4 (thread = row), and the four warps of a quadrant share the row's 64 columns
// (16 each) so that every hop of the per-tile dependency chain is short.  Shared memory (nh = 2, d <= 56):
//     X (input rows, bf16 hi | lo)  28 KB     A1, A2 (hidden activations, kept for their weight gradients)  64 KB
//     P, Q (ping-pong: last activation / cotangent / delta images)  64 KB     4 weight images  64 KB     ones  4 KB
// All images use the canonical no-swizzle layout of sdes_linear.cuh: element (row r, feature f) of a half at
// ((f / 8) * 128 + r) * 16 + (f % 8) * 2 bytes, which is at once
//   - a K-major A operand  [M = 128 rows,  K = features]   (forward and dgrad GEMMs),
//   - an MN-major operand  [MN = features, K = 128 rows]   (weight gradients: the contraction runs over the rows).
// The 64 x 64 weight images are the forward ones; W^T for the dgrad GEMMs is the SAME image read as an MN-major B operand.
// Per layer of the backward chain the pre-activation is recomputed by one more GEMM into a second accumulator (the
// tensor pipe is far from busy; storing GELU' would need another 96 KB), so  delta_l = (delta_{l+1} W^T) * GELU'(z_l).
// MMAs complete in issue order, so ONE commit also covers every MMA issued before it: per backward hop the control warp issues
// the dgrad GEMM, commits (the epilogue may start), and only then the layer's weight-gradient MMAs and the next hop's recompute,
// which run under the epilogue; the buffer the epilogue writes was last read by MMAs issued before this hop's dgrad, i.e.
// complete by the time the commit fires — which is what lets two ping-pong buffers serve the whole chain.
// The same kernel runs the kl / kl_ito reverse sweep (MODE 1-3): a CTA walks row tiles backwards in time with the adjoint in
// registers, cut into two units per tile that hand the adjoint over through HBM (see the kernel body); details in DESIGN §4.5.
// TMEM: 4 x 64 columns of weight-gradient accumulators (rows 0-63 / 64-127: contributions of the hi / lo half of delta),
// 4 x 16 columns of column sums (bias gradients; the input layer's are d loss / d emb(s), flushed whenever the step
// changes — items are step-major, so a CTA sees at most a few steps), 3 x 64 working accumulators (the dgrad result and
// two alternating recomputed pre-activations, so the next hop's recompute runs under the current hop's epilogue).
#pragma once

#include <cstdio>
#include <cstdlib>

#include "sdes_linear.cuh"
#include "sdes_step.cuh"

namespace sdes {
namespace grad {

using namespace wide;

constexpr int FL_MAX_LAYERS = 4;                 // W_in, up to two hidden layers, W_out
constexpr int FL_EPI_WARPS = 16;
constexpr int FL_THREADS = 64 + 32 * FL_EPI_WARPS;
constexpr uint32_t FL_W_BYTES = 16384u;          // one 64 x 64 weight image (hi | lo)
constexpr uint32_t FL_ONES_BYTES = 4096u;
constexpr uint32_t FL_COL_DW = 0u, FL_COL_DB = 256u, FL_COL_D = 320u, FL_COL_D2 = 384u;

struct FusedLvArgs {
    SdesRolloutDesc d;                        // dim, batch, flags (trajectory layout, noise source), seed, traj_offset, noise, clip_model
    const float* tab;                         // (T, TAB_STRIDE) per-step scalars of the fused prologue
    const float* xs;                          // stored trajectory of the forward rollout
    const float* w;                           // (B) d loss / d rnd_b
    const float* embb;                        // (T, 64) timestep_embed(s) + b_in
    const uint8_t* w_img[FL_MAX_LAYERS];      // forward images of W_in, W_h[0..nh-1], W_out
    const float* bias[FL_MAX_LAYERS];         // [1..nh]: b_h, [nh+1]: b_out (padded to 64); [0] unused
    float* dw[FL_MAX_LAYERS];                 // row-major (out, in) gradients, += via atomics
    int ldw[FL_MAX_LAYERS], n_valid[FL_MAX_LAYERS], k_valid[FL_MAX_LAYERS];
    float* db[FL_MAX_LAYERS];                 // bias gradients of layers 1..nh+1 ([0]: NULL — the input layer's go to grad_emb)
    float* grad_emb;                          // (T, 64)
    int nh, T, tiles_per_step;
    // kl / kl_ito (BPTT instantiation): the reverse sweep of sdes_grad.cu's adj_step_kernel inside the same kernel
    float* adj_init;                          // (Bp, 64) a_T = d loss / d x_T (adj_init_kernel); a tile's adjoint at the half-way step
                                              // is handed from its late to its early half through the same rows
    uint32_t* kl_flags;                       // (tiles) zeroed by the launcher: epilogue warps that have parked the tile's adjoint
    const float* score_keep;                  // forward's ungated score part, layout of xs (NULL: control without a score part)
    const float* gate;                        // (T, gate_stride) clip(score_model(s)) table of the prologue
    const float* prior_loc;                   // prior / reference Gaussian of the control: loc | 1 / scale^2
    const float* prior_iv;
    float* grad_gate;                         // (T) scalar gate gradient or NULL
    int gate_stride;
    uint32_t gflags;                          // SDES_GRAD_*
    const float* gmm_h;                       // MODE 3, single-Gaussian target: 0.5 / scale^2 per dimension
    int hvp;                                  // MODE 3: the target's Hessian enters (adj_step_kernel's target_hvp)
    unsigned long long* timeline;             // debugging (build -DSDES_FL_TIMELINE, env SDES_FL_TIMELINE=1): clock64 stamps of CTA 0, or NULL
    int watch_all;                            // debugging (SDES_FL_DEBUG): watchdog on the CTA-internal waits too (fl_wait)
};

// GELU'(x) = Phi(x) + x phi(x) of the exact-erf GELU, two values at a time: Phi from the logistic fit of gelu_fast2
// (sdes_common.cuh), phi(x) = 2^(-x^2 log2(e) / 2) / sqrt(2 pi).
__device__ __forceinline__ float2 gelu_grad2(float2 x) {
    const float2 u = __fmul2_rn(x, x);
    float2 p = __ffma2_rn(make_float2(-5.42691260e-09f, -5.42691260e-09f), u, make_float2(3.93527977e-07f, 3.93527977e-07f));
    p = __ffma2_rn(p, u, make_float2(-1.15760618e-05f, -1.15760618e-05f));
    p = __ffma2_rn(p, u, make_float2(1.60239457e-04f, 1.60239457e-04f));
    p = __ffma2_rn(p, u, make_float2(9.27478302e-05f, 9.27478302e-05f));
    p = __ffma2_rn(p, u, make_float2(-1.04834383e-01f, -1.04834383e-01f));
    p = __ffma2_rn(p, u, make_float2(-2.30220913e+00f, -2.30220913e+00f));
    const float2 q = __fmul2_rn(x, p);
    const float2 t = __fmul2_rn(u, make_float2(-0.72134752044448170f, -0.72134752044448170f));
    float2 e, r, g;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g.x) : "f"(t.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g.y) : "f"(t.y));
    const float2 dn = __fadd2_rn(e, make_float2(1.f, 1.f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(dn.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(dn.y));
    return __ffma2_rn(__fmul2_rn(x, make_float2(0.3989422804014327f, 0.3989422804014327f)), g, r);
}

// ---- MMA issue by a CONVERGED warp: every lane runs the (warp-uniform) descriptor arithmetic, one elected lane issues.
// Issued from a divergent `if (lane == 0)` the compiler wraps every tcgen05.mma in an elect / branch loop and moves each
// descriptor through R2UR (~10 instructions and a branch per MMA); here the operands stay in uniform registers.
__device__ __forceinline__ void fl_mma(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void fl_commit(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(tc::smem_u32(bar))
        : "memory");
}

// mbarrier wait with a watchdog for the kl instantiations (their CTAs wait on each other): after ~3 s of failed polls the
// waiter records (CTA, warp, wait site, item), raises the launch-wide abort word and every wait returns at once — the kernel
// ends (gradients poisoned with NaN; details with SDES_FL_DEBUG) instead of hanging the device.
__device__ __forceinline__ void fl_wait(uint64_t* bar, uint32_t parity, uint32_t* dbg, uint32_t code, uint32_t item) {
    if (dbg == nullptr) {
        tc::mbar_wait(bar, parity);
        return;
    }
    uint64_t t0 = 0;
    for (uint32_t spins = 0;; ++spins) {
        uint32_t ok;
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}" : "=r"(ok) : "r"(tc::smem_u32(bar)), "r"(parity) : "memory");
        if (ok) return;
        if ((spins & 255u) == 255u) {
            uint64_t now;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
            if (t0 == 0) t0 = now;
            const bool aborted = *reinterpret_cast<volatile uint32_t*>(dbg + 1) != 0u;
            if (aborted || now - t0 > 3000000000ull) {
                if (!aborted && (threadIdx.x & 31) == 0) {
                    const uint32_t slot = atomicAdd(dbg, 1u);
                    if (slot < 200u) {
                        uint32_t* rec = dbg + 4 + 4 * slot;
                        rec[0] = blockIdx.x; rec[1] = 1000u + (threadIdx.x >> 5); rec[2] = code; rec[3] = item;
                    }
                    *reinterpret_cast<volatile uint32_t*>(dbg + 1) = 1u;
                }
                return;
            }
        }
    }
}

// ---- MMA issue, lean form.  A shared-memory matrix descriptor (sdes_tc.cuh smem_desc_kmajor) is two 32-bit words:
//   low  = (address >> 4) | (LBO >> 4) << 16     — the only part that changes from k-step to k-step (one integer add),
//   high = (SBO >> 4) | 1 << 14 (descriptor version) — a constant per operand kind.
// Building the 64-bit values with shifts and masks per MMA, one elect per MMA and one R2UR per word cost ~20 instructions per
// MMA on the single issuing warp — 216 MMAs per item made that warp, not the tensor pipe or the epilogue, the bound of the
// kernel.  Here one asm block issues the three MMAs of a k-step behind ONE elect, from low words that are base + immediate.
constexpr uint32_t FL_TOP_KMAJOR = (128u >> 4) | (1u << 14);      // SBO 128 B: row images as A, weight images as B (K-major)
constexpr uint32_t FL_TOP_WT = (1024u >> 4) | (1u << 14);         // SBO 1024 B: weight image read MN-major (W^T)
constexpr uint32_t FL_TOP_ROWS = (2048u >> 4) | (1u << 14);       // SBO 2048 B: row images read MN-major (K = rows)
__device__ __forceinline__ uint32_t fl_desc_lo(uint32_t saddr, uint32_t lbo_bytes) { return ((saddr & 0x3FFFFu) >> 4) | ((lbo_bytes >> 4) << 16); }

// D (+)= A_lo B_hi + A_hi B_lo + A_hi B_hi for one k-step (small terms first).  The descriptors' high words and the
// instruction descriptor are compile-time immediates of the asm block (no register moves for constants).
template <uint32_t ATOP, uint32_t BTOP, uint32_t IDESC>
__device__ __forceinline__ void fl_mma3(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b64 dah, dal, dbh, dbl;\n\t"
        ".reg .b32 id;\n\t"
        "mov.b64 dah, {%1, %6};\n\t"
        "mov.b64 dal, {%2, %6};\n\t"
        "mov.b64 dbh, {%3, %7};\n\t"
        "mov.b64 dbl, {%4, %7};\n\t"
        "mov.b32 id, %8;\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "setp.eq.u32 t, %0, %0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dal, dbh, id, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbl, id, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dah, dbh, id, t;\n\t"
        "}" ::"r"(tmem_d), "r"(a_hi), "r"(a_lo), "r"(b_hi), "r"(b_lo), "r"(accumulate), "n"(ATOP), "n"(BTOP), "n"(IDESC)
        : "memory");
}
// one k-step (16 rows) of a weight gradient: dW (+)= delta^T a_lo + delta^T a_hi, column sums (+)= delta^T ones
template <uint32_t IDESC, uint32_t IDESC1>
__device__ __forceinline__ void fl_mma_wg3(uint32_t tmem_dw, uint32_t tmem_db, uint32_t da, uint32_t b_hi, uint32_t b_lo, uint32_t ones,
                                           uint32_t acc_dw, uint32_t acc_db) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, r, q, t;\n\t"
        ".reg .b64 dda, dbh, dbl, don;\n\t"
        ".reg .b32 id, id1;\n\t"
        "mov.b64 dda, {%2, %8};\n\t"
        "mov.b64 dbh, {%3, %8};\n\t"
        "mov.b64 dbl, {%4, %8};\n\t"
        "mov.b64 don, {%5, %8};\n\t"
        "mov.b32 id, %9;\n\t"
        "mov.b32 id1, %10;\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.ne.b32 r, %7, 0;\n\t"
        "setp.eq.u32 t, %0, %0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dda, dbl, id, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], dda, dbh, id, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%1], dda, don, id1, r;\n\t"
        "}" ::"r"(tmem_dw), "r"(tmem_db), "r"(da), "r"(b_hi), "r"(b_lo), "r"(ones), "r"(acc_dw), "r"(acc_db), "n"(FL_TOP_ROWS), "n"(IDESC), "n"(IDESC1)
        : "memory");
}

// D[128 rows, 64] = A (row image, K-major) x W image (N = 64 out features, K-major).  Arguments are shared-memory addresses.
__device__ __forceinline__ void fl_mma_rows_w(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, int nks) {
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64);
    const uint32_t ah = fl_desc_lo(a_hi, 2048u), al = fl_desc_lo(a_lo, 2048u), bh = fl_desc_lo(w_hi, 1024u), bl = fl_desc_lo(w_hi + 8192u, 1024u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        if (ks < nks)
            fl_mma3<FL_TOP_KMAJOR, FL_TOP_KMAJOR, idesc>(tmem_d, ah + (uint32_t)ks * 256u, al + (uint32_t)ks * 256u, bh + (uint32_t)ks * 128u,
                                                         bl + (uint32_t)ks * 128u, ks > 0 ? 1u : 0u);
}
// D[128 rows, 64 in features] = A (delta image, K = out features) x W: the forward image (n = out, k = in) read as an
// MN-major B operand — 8 in-features contiguous, out-features 16 B apart, groups of 8 out-features 128 B apart, groups of
// 8 in-features 1024 B apart; one k-step = 16 out-features = 256 B.
__device__ __forceinline__ void fl_mma_rows_wt(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t w_hi, int nks) {
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64) | (1u << 16);  // b_major = MN
    const uint32_t ah = fl_desc_lo(a_hi, 2048u), al = fl_desc_lo(a_lo, 2048u), bh = fl_desc_lo(w_hi, 128u), bl = fl_desc_lo(w_hi + 8192u, 128u);
#pragma unroll
    for (int ks = 0; ks < 4; ++ks)
        if (ks < nks)
            fl_mma3<FL_TOP_KMAJOR, FL_TOP_WT, idesc>(tmem_d, ah + (uint32_t)ks * 256u, al + (uint32_t)ks * 256u, bh + (uint32_t)ks * 16u,
                                                     bl + (uint32_t)ks * 16u, ks > 0 ? 1u : 0u);
}
// dW[(hi | lo) out features, in features] += delta^T a over the tile's 128 rows, and the column sums of delta (wgrad_mma_kernel)
__device__ __forceinline__ void fl_mma_wgrad(uint32_t tmem_dw, uint32_t tmem_db, uint32_t delta, uint32_t act_hi, uint32_t act_lo,
                                             uint32_t ones, bool acc_dw, bool acc_db) {
    constexpr uint32_t idesc = tc::idesc_bf16(128, 64) | (1u << 15) | (1u << 16), idesc1 = tc::idesc_bf16(128, 16) | (1u << 15) | (1u << 16);
    const uint32_t da = fl_desc_lo(delta, 128u), bh = fl_desc_lo(act_hi, 128u), bl = fl_desc_lo(act_lo, 128u), on = fl_desc_lo(ones, 128u);
#pragma unroll
    for (int ks = 0; ks < 8; ++ks)  // 16 rows per MMA
        fl_mma_wg3<idesc, idesc1>(tmem_dw, tmem_db, da + (uint32_t)ks * 16u, bh + (uint32_t)ks * 16u, bl + (uint32_t)ks * 16u, on,
                                  (acc_dw || ks > 0) ? 1u : 0u, (acc_db || ks > 0) ? 1u : 0u);
}

// 16 values of one row -> the two 16-byte groups (c_lo .. c_lo + 15) of both halves of an image
__device__ __forceinline__ void fl_store16(uint8_t* img, uint32_t half_bytes, int r, int c_lo, const float (&v)[16]) {
#pragma unroll
    for (int g = 0; g < 2; ++g) {
        uint4 hi, lo;
        tc::split_bf16_pair2(make_float2(v[8 * g + 0], v[8 * g + 1]), hi.x, lo.x);
        tc::split_bf16_pair2(make_float2(v[8 * g + 2], v[8 * g + 3]), hi.y, lo.y);
        tc::split_bf16_pair2(make_float2(v[8 * g + 4], v[8 * g + 5]), hi.z, lo.z);
        tc::split_bf16_pair2(make_float2(v[8 * g + 6], v[8 * g + 7]), hi.w, lo.w);
        uint8_t* o = img + (uint32_t)((((c_lo >> 3) + g) * 128 + r) * 16);
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + half_bytes) = lo;
    }
}
__device__ __forceinline__ void fl_ld16(uint32_t taddr, float (&v)[16]) {
    tc::tmem_ld8(taddr, &v[0]);
    tc::tmem_ld8(taddr + 8u, &v[8]);
    tc::wait_ld_tie<16>(v);
}

// MODE 0: lv.  MODE 1: kl (no noise in the cotangent).  MODE 2: kl_ito.  MODE 3: kl / kl_ito whose score term brings the target's
// Hessian into the sweep (Gauss / multi-well: diagonal, any d; funnel: the whole row must sit in one thread, d <= 16) and / or a
// per-dimension gate (d <= 16) — kept apart so that its extra registers do not spill the common instantiations.
template <int DPAD, int MODE>
__global__ void __launch_bounds__(FL_THREADS, 1) lv_fused_kernel(const __grid_constant__ FusedLvArgs a) {
    constexpr bool BPTT = MODE != 0;
    constexpr bool HVP = MODE == 3;
    extern __shared__ __align__(128) uint8_t fl_smem[];
    __shared__ uint64_t s_wfull, s_acc, s_aready, s_wdone, s_z[2];
    __shared__ uint32_t s_tmem;
    constexpr uint32_t XHALF = (uint32_t)(DPAD / 8) * 2048u, XBYTES = 2u * XHALF;
    constexpr int NKS_IN = (DPAD + 15) / 16;
    const SdesRolloutDesc& d = a.d;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nh = a.nh, L = nh + 2, dim = d.dim;
    uint8_t* s_x = fl_smem;
    uint8_t* s_act = s_x + XBYTES;                       // a_1 .. a_nh
    uint8_t* s_p = s_act + (size_t)nh * A_BLOCK;
    uint8_t* s_q = s_p + A_BLOCK;
    uint8_t* s_w = s_q + A_BLOCK;                        // L weight images
    uint8_t* s_ones = s_w + (size_t)L * FL_W_BYTES;
    float* s_bias = reinterpret_cast<float*>(s_ones + FL_ONES_BYTES);  // [l - 1][64], l = 1 .. nh + 1
    float* s_prior = s_bias + (nh + 1) * 64;                           // kl: prior loc[64] | 1 / scale^2 [64]

    // every operand buffer starts finite: padded k-steps multiply whatever lies there by zero weights
    for (uint32_t e = tid; e < (uint32_t)(s_w - fl_smem) / 16u; e += blockDim.x) reinterpret_cast<uint4*>(fl_smem)[e] = make_uint4(0u, 0u, 0u, 0u);
    for (uint32_t e = tid; e < FL_ONES_BYTES / 4u; e += blockDim.x) reinterpret_cast<uint32_t*>(s_ones)[e] = 0x3F803F80u;  // bf16 1.0 pairs
    for (int e = tid; e < (nh + 1) * 64; e += blockDim.x) s_bias[e] = a.bias[1 + e / 64][e % 64];
    if (BPTT)
        for (int e = tid; e < 128; e += blockDim.x) s_prior[e] = (e & 63) < dim ? (e < 64 ? a.prior_loc[e] : a.prior_iv[e - 64]) : 0.f;
    tc::fence_proxy_async();
    if (warp == 1) {
        tc::tmem_alloc(&s_tmem, 512u);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(&s_wfull, 1);
        tc::mbar_init(&s_acc, 1);
        tc::mbar_init(&s_aready, FL_EPI_WARPS);  // one arrive per epilogue warp
        tc::mbar_init(&s_wdone, 1);
        tc::mbar_init(&s_z[0], 1);
        tc::mbar_init(&s_z[1], 1);
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = s_tmem;
    uint32_t* dbg = BPTT ? a.kl_flags + a.tiles_per_step : nullptr;  // watchdog record area: always for the inter-CTA hand-off,
    uint32_t* dbg_in = (BPTT && a.watch_all) ? dbg : nullptr;        // for the CTA-internal mbarrier waits only on request
    // lv: the (step, tile) items are independent — each CTA takes a contiguous step-major range.
    // kl: a tile's steps are a chain (the adjoint lives in registers while the tile walks backwards in time).  512 tiles do not
    // divide by 148 CTAs, so each chain is cut in two UNITS — the late half (s = T-1 .. Th) and the early half (Th-1 .. 0) —
    // and unit u = half * tiles + tile goes to CTA u mod gridDim.x: every CTA runs its late units first, then its early ones,
    // which wait (flag, acquire) for the adjoint the tile's late half parked in HBM.  A late unit never waits and all CTAs are
    // resident (grid <= SM count, one CTA per SM), so there is no deadlock.
    const int64_t n_items = (int64_t)a.T * a.tiles_per_step;
    const int G = (int)gridDim.x, bx = (int)blockIdx.x, tiles = a.tiles_per_step;
    const int Th = a.T / 2, L0 = a.T - Th, L1 = Th;  // steps of a late / early unit
    const int n_units = BPTT ? (2 * tiles > bx ? (2 * tiles - bx + G - 1) / G : 0) : 0;
    const int n_late = BPTT ? (tiles > bx ? (tiles - bx + G - 1) / G : 0) : 0;
    int64_t i0, i1;
    if (BPTT) {
        i0 = 0;
        i1 = (int64_t)n_late * L0 + (int64_t)(n_units - n_late) * L1;
    } else {
        i0 = (int64_t)bx * n_items / G;
        i1 = (int64_t)(bx + 1) * n_items / G;
    }
    struct Item { int s, tile; bool first, last, late; };
    auto decode = [&](int64_t item) -> Item {
        Item it;
        if (!BPTT) {
            it.s = (int)(item / tiles);
            it.tile = (int)(item - (int64_t)it.s * tiles);
            it.first = it.last = it.late = false;
            return it;
        }
        const int64_t late_items = (int64_t)n_late * L0;
        int k, pos, len;
        if (item < late_items) {
            k = (int)(item / L0); pos = (int)(item - (int64_t)k * L0); len = L0; it.late = true;
            it.s = a.T - 1 - pos;
        } else {
            const int64_t e = item - late_items;
            const int ke = (int)(e / L1);
            pos = (int)(e - (int64_t)ke * L1); len = L1; it.late = false;
            k = n_late + ke;
            it.s = Th - 1 - pos;
        }
        const int u = bx + k * G;
        it.tile = it.late ? u : u - tiles;
        it.first = pos == 0;
        it.last = pos == len - 1;
        return it;
    };
    auto item_step = [&](int64_t item) -> int { return decode(item).s; };

    if (warp == 0) {
        if (i1 > i0) {  // ---- control warp (converged: see fl_mma)
            if (lane == 0) {
                tc::mbar_arrive_expect_tx(&s_wfull, (uint32_t)L * FL_W_BYTES);
                for (int l = 0; l < L; ++l) tc::bulk_g2s(s_w + (size_t)l * FL_W_BYTES, a.w_img[l], FL_W_BYTES, &s_wfull);
            }
            __syncwarp();
            tc::mbar_wait(&s_wfull, 0u);
            const uint32_t x_hi = tc::smem_u32(s_x), x_lo = x_hi + XHALF;
            const uint32_t act0 = tc::smem_u32(s_act), pb = tc::smem_u32(s_p), qb = tc::smem_u32(s_q), wb = tc::smem_u32(s_w);
            const uint32_t ones = tc::smem_u32(s_ones);
            const uint32_t tD = tmem_base + FL_COL_D, tD2 = tmem_base + FL_COL_D2;
            uint32_t ph = 0u;
            bool first = true;
            int prev_s = -1;
            int64_t cur_ctl_item = 0;
            // The recompute of hop h + 1 is committed while the epilogue may not have looked at hop h's commit yet: on ONE
            // mbarrier a slow warp would then see the phase flip twice and wait for a third commit that needs its own arrival
            // (observed as a hang under load).  Consecutive commits therefore alternate between two barriers; a barrier is
            // reused only after a `ready()` that every warp can reach only past its wait on that barrier's previous phase.
            uint32_t zbar = 0u;
            int tl_c = 0;  // timeline: [0, 512): control warp — (before, after) every wait for the epilogue
            auto stamp_c = [&]() {
#ifdef SDES_FL_TIMELINE  // compile-time switch (build with -DSDES_FL_TIMELINE, run with SDES_FL_TIMELINE=1): no cost otherwise
                if (a.timeline != nullptr && bx == 0 && lane == 0 && tl_c < 512) a.timeline[tl_c++] = (unsigned long long)clock64();
#endif
            };
            auto ready = [&]() {
                stamp_c();
                fl_wait(&s_aready, ph, dbg_in, 1u, (uint32_t)cur_ctl_item);
                ph ^= 1u;
                tc::fence_after();
                stamp_c();
            };
            for (int64_t item = i0; item < i1; ++item) {
                cur_ctl_item = item;
                const int s = item_step(item);
                const bool new_step = s != prev_s;
                prev_s = s;
                // ---- replayed forward
                ready();
                fl_mma_rows_w(tD, x_hi, x_lo, wb, NKS_IN);
                fl_commit(&s_acc);
                for (int l = 0; l < nh; ++l) {
                    ready();
                    const uint32_t ab = act0 + (uint32_t)l * A_BLOCK;
                    fl_mma_rows_w(tD, ab, ab + A_HALF, wb + (uint32_t)(1 + l) * FL_W_BYTES, 4);
                    fl_commit(&s_acc);
                }
                ready();
                fl_mma_rows_w(tD, pb, pb + A_HALF, wb + (uint32_t)(nh + 1) * FL_W_BYTES, 4);
                fl_commit(&s_acc);
                // z_{nh+1} = a_nh W_h[nh-1] for the first backward hop: runs while the epilogue builds the cotangent
                {
                    const uint32_t ab = act0 + (uint32_t)(nh - 1) * A_BLOCK;
                    fl_mma_rows_w(tD2, ab, ab + A_HALF, wb + (uint32_t)nh * FL_W_BYTES, 4);
                    fl_commit(&s_z[zbar]);
                    zbar ^= 1u;
                }
                // ---- backward: output layer (cotangent in Q, a_{nh+1} in P).  Its weight gradient goes first: the epilogue
                //      overwrites P with delta_{nh+1}, and MMAs complete in order
                ready();
                fl_mma_wgrad(tmem_base + FL_COL_DW + 64u * (uint32_t)(nh + 1), tmem_base + FL_COL_DB + 16u * (uint32_t)(nh + 1), qb, pb, pb + A_HALF, ones,
                             !first, !first);
                fl_mma_rows_wt(tD, qb, qb + A_HALF, wb + (uint32_t)(nh + 1) * FL_W_BYTES, NKS_IN);
                fl_commit(&s_acc);
                uint32_t z_next = 1u;  // which recompute accumulator the NEXT hop reads
                auto recompute = [&](int l) {  // z_{l+1} for the hop that produces delta_{l+1}
                    const uint32_t tz = tD2 + 64u * z_next;
                    if (l > 0) {
                        const uint32_t pa = act0 + (uint32_t)(l - 1) * A_BLOCK;
                        fl_mma_rows_w(tz, pa, pa + A_HALF, wb + (uint32_t)l * FL_W_BYTES, 4);
                    } else {
                        fl_mma_rows_w(tz, x_hi, x_lo, wb, NKS_IN);
                    }
                    fl_commit(&s_z[zbar]);
                    zbar ^= 1u;
                    z_next ^= 1u;
                };
                recompute(nh - 1);
                uint32_t cur = pb, other = qb;  // the epilogue writes delta_{nh+1} into P
                for (int l = nh - 1; l >= 0; --l) {
                    ready();  // delta_{l+2} is in `cur`; `other` was last read by MMAs issued before this hop's dgrad
                    const uint32_t ab = act0 + (uint32_t)l * A_BLOCK;  // a_{l+1}
                    fl_mma_rows_wt(tD, cur, cur + A_HALF, wb + (uint32_t)(1 + l) * FL_W_BYTES, 4);
                    fl_commit(&s_acc);
                    // off the critical path (they run while the epilogue works on this hop): the weight gradient of this
                    // layer and the next hop's pre-activation
                    fl_mma_wgrad(tmem_base + FL_COL_DW + 64u * (uint32_t)(1 + l), tmem_base + FL_COL_DB + 16u * (uint32_t)(1 + l), cur, ab, ab + A_HALF, ones,
                                 !first, !first);
                    if (l > 0) recompute(l - 1);
                    const uint32_t t = cur;
                    cur = other;
                    other = t;
                }
                ready();  // delta_1 is in `cur`
                if (BPTT) {  // J_x NN^T delta: one more transposed layer, added to the adjoint by the epilogue
                    fl_mma_rows_wt(tD, cur, cur + A_HALF, wb, 4);
                    fl_commit(&s_acc);
                }
                fl_mma_wgrad(tmem_base + FL_COL_DW, tmem_base + FL_COL_DB, cur, x_hi, x_lo, ones, !first, !new_step);
                fl_commit(&s_wdone);
                first = false;
            }
        }
    } else if (warp >= 2 && i1 > i0) {  // ---- epilogue warps
        const int q = warp & 3, r = q * 32 + lane, cw = (warp - 2) >> 2, c_lo = 16 * cw;
        const uint32_t lane_base = tmem_base + ((uint32_t)(q * 32) << 16);
        const uint32_t tD = lane_base + FL_COL_D + (uint32_t)c_lo, tD2 = lane_base + FL_COL_D2 + (uint32_t)c_lo;
        const int64_t B = d.batch;
        uint32_t ph_acc = 0u, ph_w = 0u, ph_zbits = 0u, zbar = 0u;  // ph_zbits: one parity bit per s_z barrier
        int prev_s = -1;
        float adj[16];  // kl: this thread's 16 dimensions of a_{s+1} = d loss / d x_{s+1}
#pragma unroll
        for (int e = 0; e < 16; ++e) adj[e] = 0.f;
        int tl_e = 0;  // timeline: [512, 1024): epilogue warp 2 — (before, after) every accumulator wait, and every arrive
        auto stamp_e = [&]() {
#ifdef SDES_FL_TIMELINE
            if (a.timeline != nullptr && bx == 0 && warp == 2 && lane == 0 && tl_e < 512) a.timeline[512 + tl_e++] = (unsigned long long)clock64();
#endif
        };
        auto arrive = [&]() {
            stamp_e();
            tc::fence_proxy_async();  // generic-proxy writes of the operand -> visible to the tensor-core (async) proxy
            tc::fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_aready);  // 16 arrivals per hop instead of 512 serialised ones
        };
        int64_t cur_e_item = 0;
        auto wait_acc = [&]() {
            stamp_e();
            fl_wait(&s_acc, ph_acc, dbg_in, 2u, (uint32_t)cur_e_item);
            ph_acc ^= 1u;
            tc::fence_after();
            stamp_e();
        };
        auto flush_emb = [&](int s_prev) {  // d loss / d emb(s_prev): column sums of delta_1 over the step's tiles of this CTA
            if (cw == 0) {
                float v[8];
                tc::tmem_ld8(lane_base + FL_COL_DB, v);
                tc::wait_ld_tie<8>(v);
                if (v[0] != 0.f) atomicAdd(a.grad_emb + (int64_t)s_prev * 64 + (r & 63), v[0]);  // hi row and lo row both add
            }
        };
        float xnext[16];
        auto load_x = [&](int64_t it, float (&xv)[16]) {
            const Item i2 = decode(it);
            const int s2 = i2.s, tile2 = i2.tile;
            const int64_t b2 = (int64_t)tile2 * 128 + r;
            const bool valid2 = b2 < B;
            const TrajRef xr = traj_ref(d, const_cast<float*>(a.xs), s2, valid2 ? b2 : 0);
#pragma unroll
            for (int e = 0; e < 16; ++e) xv[e] = (valid2 && c_lo + e < dim) ? __ldg(xr.p + (int64_t)(c_lo + e) * xr.stride) : 0.f;
        };
        for (int64_t item = i0; item < i1; ++item) {
            cur_e_item = item;
            const Item cur_item = decode(item);
            const int s = cur_item.s, tile = cur_item.tile;
            const int64_t b = (int64_t)tile * 128 + r;
            const bool valid = b < B;
            const int64_t bb = valid ? b : 0;
            if (BPTT && cur_item.first) {  // a new unit: the tile's terminal adjoint, or the one its late half parked
                if (!cur_item.late) {
                    if (lane == 0) {
                        uint32_t seen;
                        for (uint32_t spins = 0;; ++spins) {
                            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(a.kl_flags + tile) : "memory");
                            if (seen >= (uint32_t)FL_EPI_WARPS) break;
                            __nanosleep(1000);
                            if (spins > 2000000u || *reinterpret_cast<volatile uint32_t*>(dbg + 1) != 0u) {  // seconds: the producer is gone
                                if (warp == 2) {
                                    const uint32_t slot = atomicAdd(dbg, 1u);
                                    if (slot < 200u) {
                                        uint32_t* rec = dbg + 4 + 4 * slot;
                                        rec[0] = (uint32_t)bx; rec[1] = (uint32_t)tile; rec[2] = 100u + seen; rec[3] = (uint32_t)item;
                                    }
                                    // the result is wrong from here on: poison it so that nobody can use it by accident (the
                                    // trainer's finite check skips the step, the tests fail) — later atomic adds keep the NaN
                                    atomicExch(reinterpret_cast<unsigned int*>(a.grad_emb), 0x7fc00000u);
                                    atomicExch(reinterpret_cast<unsigned int*>(a.dw[0]), 0x7fc00000u);
                                }
                                break;
                            }
                        }
                    }
                    __syncwarp();
                }
#pragma unroll
                for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(&adj[e]) = __ldcg(reinterpret_cast<const float4*>(a.adj_init + b * 64 + c_lo + e));
            }
            // kl: the kept score part of this row and the step's (scalar) gate, fetched now — the cotangent math after the output
            // layer sits on the item's critical path and must not wait for HBM
            float gate0 = 0.f;
            const bool has_score = BPTT && d.ctrl_kind != SDES_CTRL_CLIPPED && a.score_keep != nullptr;
            const TrajRef kr = traj_ref(d, const_cast<float*>(has_score ? a.score_keep : a.xs), s, bb);
            if (BPTT) {
                gate0 = __ldg(a.gate + (int64_t)s * a.gate_stride);
                if (s > 0 && valid) {  // the next item of this tile: rows of x_{s-1}
                    const TrajRef xp = traj_ref(d, const_cast<float*>(a.xs), s - 1, bb);
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        if (c_lo + e < dim) asm volatile("prefetch.global.L2 [%0];" ::"l"(xp.p + (int64_t)(c_lo + e) * xp.stride));
                }
            }
            // ---- this thread's 16 features of the row: fetched at the END of the previous item (load_x below), so that the HBM
            //      latency runs under that item's last tensor phases
            float v[16];
            if (BPTT || item == i0) load_x(item, xnext);  // (kl: 16 more live registers across the last hop would spill — L2 prefetch instead)
#pragma unroll
            for (int e = 0; e < 16; ++e) v[e] = xnext[e];
            if (item > i0) {
                fl_wait(&s_wdone, ph_w, dbg_in, 3u, (uint32_t)item);  // X (and P / Q) are free again
                ph_w ^= 1u;
                tc::fence_after();
                if (s != prev_s) flush_emb(prev_s);
            }
            prev_s = s;
            if (BPTT && has_score && valid) {
                // the kept score part of this row goes global -> shared asynchronously, into the 64 bytes of Q this thread will
                // overwrite with its cotangent anyway (its two 16-byte groups of both halves): no registers held across the
                // forward hops, no exposed HBM latency at the cotangent
#pragma unroll
                for (int e = 0; e < 16; ++e) {
                    if (c_lo + e < dim) {
                        const uint32_t dst = tc::smem_u32(s_q + (uint32_t)((((c_lo >> 3) + ((e >> 2) & 1)) * 128 + r) * 16) + ((e >> 3) ? A_HALF : 0u) + 4u * (e & 3));
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(kr.p + (int64_t)(c_lo + e) * kr.stride) : "memory");
                    }
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            }
            if (c_lo < DPAD) {
#pragma unroll
                for (int g = 0; g < 2; ++g) {
                    if (c_lo + 8 * g < DPAD) {
                        uint4 hi, lo;
                        tc::split_bf16_pair2(make_float2(v[8 * g + 0], v[8 * g + 1]), hi.x, lo.x);
                        tc::split_bf16_pair2(make_float2(v[8 * g + 2], v[8 * g + 3]), hi.y, lo.y);
                        tc::split_bf16_pair2(make_float2(v[8 * g + 4], v[8 * g + 5]), hi.z, lo.z);
                        tc::split_bf16_pair2(make_float2(v[8 * g + 6], v[8 * g + 7]), hi.w, lo.w);
                        uint8_t* o = s_x + (uint32_t)((((c_lo >> 3) + g) * 128 + r) * 16);
                        *reinterpret_cast<uint4*>(o) = hi;
                        *reinterpret_cast<uint4*>(o + XHALF) = lo;
                    }
                }
            }
            arrive();
            const float* embb = a.embb + (int64_t)s * 64 + c_lo;
            // the noise of the cotangent is drawn one quad of dimensions per forward hop, BEFORE that hop's wait: the ~130
            // Philox / Box-Muller instructions run while the tensor pipe works on the layer
            const float* tab = a.tab + (int64_t)s * TAB_STRIDE;
            const StepCoef c = make_step_coef(d, tab);
            const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)bb);
            const float* nrow = c.from_hbm ? d.noise + ((int64_t)s * B + bb) * dim : nullptr;
            float eps[16];
            const bool need_eps = MODE == 0 || MODE == 2 || (MODE == 3 && (d.flags & SDES_F_COMPUTE_ITO) != 0);
            auto draw = [&](int hop) {
#pragma unroll
                for (int qd = 0; qd < 4; ++qd) {
                    if (!need_eps) {
                        eps[4 * qd] = eps[4 * qd + 1] = eps[4 * qd + 2] = eps[4 * qd + 3] = 0.f;
                        continue;
                    }
                    // lv: quad -> forward hop, round robin.  kl: all four at the output hop (hop < 0) — the adjoint already
                    // occupies 16 registers across the forward hops, and only kl_ito draws noise at all
                    if (hop >= 0 && (BPTT || (qd < nh + 2 ? qd : qd - (nh + 2)) != hop)) continue;
                    const int j0 = c_lo + 4 * qd;
                    float4 n4 = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (j0 < dim) {
                        if (c.from_hbm) {
                            n4.x = nrow[j0];
                            n4.y = j0 + 1 < dim ? nrow[j0 + 1] : 0.f;
                            n4.z = j0 + 2 < dim ? nrow[j0 + 2] : 0.f;
                            n4.w = j0 + 3 < dim ? nrow[j0 + 3] : 0.f;
                        } else {
                            n4 = normal4_call(c.k0, c.k1, traj, (uint32_t)s, (uint32_t)(j0 >> 2));
                        }
                    }
                    eps[4 * qd] = n4.x; eps[4 * qd + 1] = n4.y; eps[4 * qd + 2] = n4.z; eps[4 * qd + 3] = n4.w;
                }
            };
            // ---- replayed forward: input layer, hidden layers
            for (int l = 0; l <= nh; ++l) {
                float bv[16];
                if (l == 0) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(&bv[e]) = __ldg(reinterpret_cast<const float4*>(embb + e));
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(&bv[e]) = *reinterpret_cast<const float4*>(s_bias + (l - 1) * 64 + c_lo + e);
                }
                draw(l);
                wait_acc();
                fl_ld16(tD, v);
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const float2 y = gelu_fast2(__fadd2_rn(make_float2(v[e], v[e + 1]), make_float2(bv[e], bv[e + 1])));
                    v[e] = y.x;
                    v[e + 1] = y.y;
                }
                fl_store16(l < nh ? s_act + (size_t)l * A_BLOCK : s_p, A_HALF, r, c_lo, v);
                arrive();
            }
            // ---- output layer -> cotangent of the network output: w_b c_j 1[|NN_j| <= clip_model], c = eps sqrt(dt) (sigma beta_k eps)
            {
                const float wb = valid ? __ldg(a.w + bb) : 0.f;
                const float cscale = wb * (c.exp_int ? c.sg * c.beta_k : c.sqrt_dt);
                const float* bo = s_bias + nh * 64 + c_lo;
                draw(BPTT ? -1 : nh + 1);
                float kbase[16];
                if (BPTT) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) kbase[e] = 0.f;
                    if (has_score && valid) {
                        asm volatile("cp.async.wait_all;" ::: "memory");
#pragma unroll
                        for (int e = 0; e < 16; e += 4) {
                            const float4 t4 = *reinterpret_cast<const float4*>(s_q + (uint32_t)((((c_lo >> 3) + ((e >> 2) & 1)) * 128 + r) * 16) + ((e >> 3) ? A_HALF : 0u));
                            kbase[e] = t4.x; kbase[e + 1] = t4.y; kbase[e + 2] = t4.z; kbase[e + 3] = t4.w;
                        }
#pragma unroll
                        for (int e = 0; e < 16; ++e) kbase[e] = (c_lo + e < dim) ? kbase[e] : 0.f;
                    }
                }
                wait_acc();
                fl_ld16(tD, v);
                if (!BPTT) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const float nn = v[e] + bo[e];
                        v[e] = (c_lo + e < dim && fabsf(nn) <= c.cm) ? cscale * eps[e] : 0.f;  // d clip(NN) / d NN: 1 on [-c, c]
                    }
                } else {
                    // the reverse sweep's elementwise step (adj_step_kernel of sdes_grad.cu, per dimension): with a = a_{s+1},
                    //   g = clip(NN) + gate * base,   q = w (cq g_m + ci eps),   delta = q + a Bc   (cotangent of the control),
                    //   a <- A a  [+ q sigma / scale_prior^2: Euler-DDS reference control]  + (d score part / d x)^T delta
                    // (the score part's own x-derivative is local here: prior score of the Lerp controls; a target score that is a
                    // constant of the graph or detached) and + J_x NN^T (delta 1[|NN| <= clip_model]) after the last hop.
                    const int ck = d.ctrl_kind;
                    const bool dead = !(wb != 0.f);
                    const bool score_detached = (a.gflags & SDES_GRAD_SCORE_DETACHED) != 0 || ck == SDES_CTRL_CLIPPED;
                    const bool prior_in_ctrl = ck == SDES_CTRL_LERP || ck == SDES_CTRL_LERP_PRIOR;
                    const float outer = (ck == SDES_CTRL_SCORE ? 1.0f : c.sigma) * d.scale_score;
                    const float clip_edge = fabsf(outer) * d.clip_score;  // |base| of a clipped inner value
                    const float a_mul = c.exp_int ? c.alpha_k : fmaf(c.mu, c.dt, 1.0f);
                    const float wp = 1.0f - tab[TAB_LERP_W];
                    const TrajRef xr = traj_ref(d, const_cast<float*>(a.xs), s, bb);
                    float gsum = 0.f;
                    float xv16[HVP ? 16 : 1], facv[HVP ? 16 : 1], gdim[HVP ? 16 : 1];
                    const bool gate_per_dim = HVP && d.gate_dim > 1;
#pragma unroll
                    for (int e = 0; e < 16; ++e) {
                        const int j = c_lo + e;
                        const bool in = j < dim;
                        const float nn = v[e] + bo[e];
                        const float base = kbase[e];
                        // scalar gate (or the constant 1 of a control without one), 0 on padding; MODE 3: the table's own column
                        const float gt = in ? (gate_per_dim ? __ldg(a.gate + (int64_t)s * a.gate_stride + j) : gate0) : 0.f;
                        const float iv = s_prior[64 + j];
                        if (HVP) xv16[e] = (in && valid) ? __ldg(xr.p + (int64_t)j * xr.stride) : 0.f;
                        const float g = clipf(nn, c.cm) + base * gt;
                        const float ap = adj[e];
                        float dg, nx;
                        if (c.exp_int) {
                            dg = wb * (c.bb_ss * g + c.s_bk * eps[e]) + ap * c.bb_ss;
                            nx = ap * a_mul;
                        } else {
                            float gm = g;
                            if (c.ref_ctrl) {
                                const float xj = (in && valid) ? __ldg(xr.p + (int64_t)j * xr.stride) : 0.f;
                                gm = g - c.sigma * ((s_prior[j] - xj) * iv);
                            }
                            const float qj = wb * (gm * c.dt + eps[e] * c.sqrt_dt);
                            dg = fmaf(ap, c.sigma * c.dt, qj);
                            nx = ap * a_mul;
                            if (c.ref_ctrl) nx = fmaf(qj, c.sigma * iv, nx);
                        }
                        if (dead || !in) dg = 0.f;
                        gsum = fmaf(dg, base, gsum);
                        const float fac = (fabsf(base) != clip_edge) ? outer * gt * dg : 0.f;  // cotangent of inner_j (1 inside the clip)
                        if (!score_detached && prior_in_ctrl) nx = fmaf(-wp * iv, fac, nx);
                        if (HVP) {
                            facv[e] = fac;
                            gdim[e] = dg * base;
                        }
                        adj[e] = (dead || !valid || !in) ? 0.f : nx;
                        v[e] = fabsf(nn) <= c.cm ? dg : 0.f;  // d clip(NN) / d NN
                    }
                    if (HVP && a.hvp && !score_detached && !dead && valid) {
                        // + H(x_s) (cotangent of the target score): sdes_step.cuh target_hvp_add, on this thread's 16 dimensions
                        const float lw = (ck == SDES_CTRL_SCORE) ? 1.0f : tab[TAB_LERP_W];
                        if (d.target_kind == SDES_TARGET_GMM) {  // one component: score = (loc - x) / scale^2
#pragma unroll
                            for (int e = 0; e < 16; ++e)
                                if (c_lo + e < dim) adj[e] = fmaf(-2.0f * __ldg(a.gmm_h + c_lo + e), lw * facv[e], adj[e]);
                        } else if (d.target_kind == SDES_TARGET_MULTIWELL) {
#pragma unroll
                            for (int e = 0; e < 16; ++e) {
                                const int j = c_lo + e;
                                const float y = xv16[e] - d.shift;
                                if (j < d.n_double_wells) adj[e] = fmaf(4.0f * d.separation - 12.0f * y * y, lw * facv[e], adj[e]);
                                else if (j < dim) adj[e] -= lw * facv[e];
                            }
                        } else if (c_lo == 0) {  // funnel (d <= 16: the whole row is here); distr/funnel.py:71-80
                            float sq = 0.f, xf = 0.f;
#pragma unroll
                            for (int e = 1; e < 16; ++e) {
                                sq = fmaf(xv16[e], xv16[e], sq);
                                xf = fmaf(xv16[e], lw * facv[e], xf);
                            }
                            const float inv = expf(-xv16[0]), f0 = lw * facv[0];
                            adj[0] += f0 * (-1.0f / d.variance - 0.5f * sq * inv) + inv * xf;
#pragma unroll
                            for (int e = 1; e < 16; ++e)
                                if (e < dim) adj[e] += inv * (xv16[e] * f0 - lw * facv[e]);
                        }
                    }
                    if (gate_per_dim && a.grad_gate != nullptr && c_lo < dim) {  // per-dimension gate (d <= 16): one reduction per column
#pragma unroll
                        for (int e = 0; e < 16; ++e) {
                            float t = gdim[e];
#pragma unroll
                            for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
                            const int j = c_lo + e;
                            if (lane == 0 && j < dim && t != 0.f && fabsf(__ldg(a.gate + (int64_t)s * a.gate_stride + j)) < c.cm)
                                atomicAdd(a.grad_gate + (int64_t)s * dim + j, t);
                        }
                    } else if (a.grad_gate != nullptr) {  // scalar gate: d loss / d gate(s) = 1[|gate| < clip] sum_b sum_j delta_j base_j
#pragma unroll
                        for (int o = 16; o > 0; o >>= 1) gsum += __shfl_xor_sync(0xffffffffu, gsum, o);
                        if (lane == 0 && gsum != 0.f && fabsf(gate0) < c.cm) atomicAdd(a.grad_gate + s, gsum);
                    }
                }
                fl_store16(s_q, A_HALF, r, c_lo, v);
                arrive();
            }
            // ---- backward chain: delta_l = (delta_{l+1} W^T) * GELU'(z_l), z_l recomputed into the second accumulator
            uint8_t* dst = s_p;
            uint8_t* nxt = s_q;
            uint32_t zsel = 0u;  // the two recompute accumulators alternate hop by hop
            for (int l = nh; l >= 0; --l) {  // produces delta_{l+1}
                float bv[16], z[16];
                if (l == 0) {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(&bv[e]) = __ldg(reinterpret_cast<const float4*>(embb + e));
                } else {
#pragma unroll
                    for (int e = 0; e < 16; e += 4) *reinterpret_cast<float4*>(&bv[e]) = *reinterpret_cast<const float4*>(s_bias + (l - 1) * 64 + c_lo + e);
                }
                // the recomputed pre-activation was committed a whole hop ago: GELU'(z) — most of this hop's arithmetic — is
                // evaluated while the tensor pipe still runs the hop's dgrad GEMM
                fl_wait(&s_z[zbar], (ph_zbits >> zbar) & 1u, dbg_in, 4u, (uint32_t)item);
                ph_zbits ^= 1u << zbar;
                zbar ^= 1u;
                tc::fence_after();
                tc::tmem_ld8(tD2 + 64u * zsel, &z[0]);
                tc::tmem_ld8(tD2 + 64u * zsel + 8u, &z[8]);
                zsel ^= 1u;
                tc::wait_ld_tie<16>(z);
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const float2 g = gelu_grad2(__fadd2_rn(make_float2(z[e], z[e + 1]), make_float2(bv[e], bv[e + 1])));
                    z[e] = g.x;
                    z[e + 1] = g.y;
                }
                wait_acc();
                fl_ld16(tD, v);
#pragma unroll
                for (int e = 0; e < 16; e += 2) {
                    const float2 y = __fmul2_rn(make_float2(v[e], v[e + 1]), make_float2(z[e], z[e + 1]));
                    v[e] = y.x;
                    v[e + 1] = y.y;
                }
                fl_store16(dst, A_HALF, r, c_lo, v);
                arrive();
                uint8_t* t = dst;
                dst = nxt;
                nxt = t;
            }
            if (!BPTT && item + 1 < i1) load_x(item + 1, xnext);
            if (BPTT) {  // a_s += J_x NN^T delta  (dimension columns of delta_1 W_in)
                wait_acc();
                fl_ld16(tD, v);
                const bool dead = !(valid && __ldg(a.w + bb) != 0.f);
#pragma unroll
                for (int e = 0; e < 16; ++e) adj[e] = (dead || c_lo + e >= dim) ? 0.f : adj[e] + v[e];
                tc::fence_before();  // the next item's first GEMM overwrites this accumulator (ordered through the X hand-off)
                if (cur_item.late && cur_item.last && L1 > 0) {  // park the adjoint for the tile's early half (another CTA, later)
#pragma unroll
                    for (int e = 0; e < 16; e += 4) __stcg(reinterpret_cast<float4*>(a.adj_init + b * 64 + c_lo + e), *reinterpret_cast<const float4*>(&adj[e]));
                    __threadfence();
                    __syncwarp();
                    if (lane == 0) asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.kl_flags + tile) : "memory");
                }
            }
        }
        // ---- the launch's accumulators -> global memory
        fl_wait(&s_wdone, ph_w, dbg_in, 5u, 0xffffffffu);
        tc::fence_after();
        flush_emb(prev_s);
        for (int l = 0; l < L; ++l) {
            float v[16];
            fl_ld16(lane_base + FL_COL_DW + 64u * (uint32_t)l + (uint32_t)c_lo, v);
            const int n = r & 63;  // D rows 0-63: contributions of delta_hi, 64-127: of delta_lo
            if (n < a.n_valid[l]) {
#pragma unroll
                for (int e = 0; e < 16; ++e)
                    if (c_lo + e < a.k_valid[l] && v[e] != 0.f) atomicAdd(a.dw[l] + (int64_t)n * a.ldw[l] + c_lo + e, v[e]);
            }
            if (l > 0 && cw == 0 && a.db[l] != nullptr) {
                float u[8];
                tc::tmem_ld8(lane_base + FL_COL_DB + 16u * (uint32_t)l, u);
                tc::wait_ld_tie<8>(u);
                if (n < a.n_valid[l] && u[0] != 0.f) atomicAdd(a.db[l] + n, u[0]);
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 512u);
}

static size_t lv_fused_smem_bytes(int dpad, int nh) {
    return (size_t)2 * (dpad / 8) * 2048 + (size_t)(nh + 2) * A_BLOCK + (size_t)(nh + 2) * FL_W_BYTES + FL_ONES_BYTES + (size_t)(nh + 1) * 256 + 512;
}
// the shapes the fused kernel serves: d <= 56 (input image + buffers fit 227 KB), one or two hidden layers
static bool lv_fused_supported(const SdesRolloutDesc& d) {
    return d.dim <= 56 && d.n_hidden >= 1 && d.n_hidden <= 2 && lv_fused_smem_bytes(mma_pad_dim(d.dim), d.n_hidden) <= 232448u - 64u;
}

template <int DPAD, int MODE>
static cudaError_t launch_lv_fused_t(const FusedLvArgs& a, int sm_count, cudaStream_t stream) {
    const size_t smem = lv_fused_smem_bytes(DPAD, a.nh);
    static size_t attr = 0;
    if (attr < smem) {
        cudaError_t e = cudaFuncSetAttribute(lv_fused_kernel<DPAD, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr = smem;
    }
    const int64_t n_items = MODE != 0 ? 2 * (int64_t)a.tiles_per_step : (int64_t)a.T * a.tiles_per_step;  // kl: units (half chains)
    int grid = (int)(n_items < sm_count ? n_items : sm_count);
    if (MODE != 0) {
        // the kl units wait on each other: every CTA of the grid must be resident
        int occ = 0;
        cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, lv_fused_kernel<DPAD, MODE>, FL_THREADS, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        if (const char* g = getenv("SDES_FL_GRID")) grid = atoi(g) < grid ? atoi(g) : grid;  // debugging aid
        if (getenv("SDES_FL_DEBUG")) fprintf(stderr, "lv_fused kl: occ/SM %d, sm_count %d, grid %d, units %lld\n", occ, sm_count, grid, (long long)n_items);
    }
    if (grid < 1) grid = 1;
    lv_fused_kernel<DPAD, MODE><<<grid, FL_THREADS, smem, stream>>>(a);
    return cudaGetLastError();
}

template <int MODE>
static cudaError_t launch_lv_fused_m(const FusedLvArgs& a, int sm_count, cudaStream_t stream) {
    switch (mma_pad_dim(a.d.dim)) {
        case 8: return launch_lv_fused_t<8, MODE>(a, sm_count, stream);
        case 16: return launch_lv_fused_t<16, MODE>(a, sm_count, stream);
        case 32: return launch_lv_fused_t<32, MODE>(a, sm_count, stream);
        case 48: return launch_lv_fused_t<48, MODE>(a, sm_count, stream);
        case 56: return launch_lv_fused_t<56, MODE>(a, sm_count, stream);
        default: return cudaErrorInvalidValue;
    }
}
static cudaError_t launch_lv_fused(const FusedLvArgs& a, bool bptt, int sm_count, cudaStream_t stream) {
    if (!bptt) return launch_lv_fused_m<0>(a, sm_count, stream);
    if (a.hvp || a.d.gate_dim > 1) return launch_lv_fused_m<3>(a, sm_count, stream);
    return (a.d.flags & SDES_F_COMPUTE_ITO) ? launch_lv_fused_m<2>(a, sm_count, stream) : launch_lv_fused_m<1>(a, sm_count, stream);
}

}  // namespace grad
}  // namespace sdes
