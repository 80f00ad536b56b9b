// sdes_rollout_mma.cu — persistent rollout kernel, control MLP on tcgen05 tensor cores.
//
// One CTA per SM, resident for the whole rollout.  It holds the bf16 hi/lo weight images of every
// layer in shared memory (loaded once with TMA bulk copies), owns all 512 TMEM columns, and runs
// FOUR independent groups of 4 warps.  A group pulls work items (time chunk, 128-trajectory tile) from
// a global counter and carries the tile through the chunk's time steps: thread r owns trajectory r —
// its state x lives in shared memory as dimension pairs ([pair][128] float2, conflict free), its
// network activations in TMEM lane r.  Per step and layer the group writes the activations (bf16 hi, lo)
// into TMEM as the A operand, one thread issues the layer's 3 x K/16 kind::f16 MMAs against the
// shared-memory weights and commits to the group's mbarrier, and the threads stream the fp32 accumulator
// back with tcgen05.ld, 8 columns at a time, through the fused epilogue (bias, exact-erf GELU, hi/lo split)
// into the next layer's A operand.  The last accumulator (the network output) streams 8 columns at a time
// into the control / noise / cost / state update (one pass over the state per step).
//
// Round-2 rewrite ("lean" step).  The kernel is bound by FP32-pipe ISSUE slots and MUFU throughput, not by the
// tensor pipe, so the step is written to minimise issued instructions:
//   * every elementwise FP32 operation works on a register pair (packed FFMA2 / FMUL2 / FADD2: one issue slot
//     for two lanes), the GELU in its logistic form (sdes_common.cuh gelu_fast2);
//   * the score part of the control is evaluated inside the update loop from the pair of x it is about to
//     advance (no sc[DPAD] array held across the MLP, no spills), the only pre-pass being the target's
//     "global" quantities: the softmax over mixture components on the dimensions that differ between
//     components (<= 8 leading dims, else the DENSE kernel below), the funnel's two sums;
//   * Philox round keys are kernel parameters (constant-bank operands), the generator is inlined;
//   * control kind x target kind are compile-time arguments of the update loop, selected by one
//     warp-uniform switch per step;
//   * the Euler-Maruyama and exponential-integrator updates share one form x' = A x + Bc g + Cc eps.
// DENSE = true is the same kernel with the target score of ALL dimensions precomputed into registers per step
// (a GMM whose components differ beyond the first 8 dims); the host launches both instantiations and the
// one that does not match the prologue's dimension mask returns at once.
#include <cuda_bf16.h>

#include "sdes_step.cuh"
#include <cstdio>
#include "sdes_tc.cuh"

namespace sdes {

constexpr int TC_GROUPS = 4;
constexpr int TC_THREADS = TC_GROUPS * 128;
constexpr int TMEM_COLS = 512;
constexpr int GROUP_COLS = 128;  // D[64] | A_hi bf16x2 [32] | A_lo bf16x2 [32]
constexpr int GMM_ACT = 8;       // leading dimensions of every mixture component kept in shared memory (lean kernel)
constexpr float LOG2E = 1.4426950408889634f;

__host__ __device__ inline int mma_nout(int dpad) { return (dpad + 15) / 16 * 16; }

// dynamic shared memory of rollout_tc_kernel (carve-up in the kernel)
inline size_t tc_smem_bytes(const KParams& p) {
    const int dpad = p.ws.dpad, K = p.d.target_kind == SDES_TARGET_GMM ? p.d.n_components : 0;
    const size_t fl = (size_t)((p.ws.w_mma4_len + 31) & ~31ll) + 2 * (size_t)((K + 3) & ~3) * GMM_ACT + 64 + 2 * (size_t)dpad + 2 * (2 * dpad + 8) +
                      (size_t)TC_GROUPS * dpad * 128;
    return fl * sizeof(float);
}
// DENSE instantiation: the full mixture image ([K4][DPAD / 2] float4 = {-mu pair, h pair}) behind everything else, where it fits
inline size_t tc_dense_image_bytes(const KParams& p) { return (size_t)((p.d.n_components + 3) & ~3) * p.ws.dpad * 8; }
inline bool tc_dense_image_fits(const KParams& p) { return tc_smem_bytes(p) + tc_dense_image_bytes(p) <= 227u * 1024u; }

// ---------------------------------------------------------------------------- the kernel
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

struct GroupCtx {
    int g;                     // group index
    int gtid;                  // thread index in the group = row in the tile = TMEM lane
    int issuer;                // gtid of the thread that issues this group's MMAs: lane 0 of warp (g mod 4), so that each of the
                               // SM's four schedulers hosts exactly one issuing warp
    uint32_t t_d, t_hi, t_lo;  // TMEM addresses, lane 0 of the tile (for the issuing thread)
    uint32_t l_d, l_hi, l_lo;  // same, at this thread's warp lane window (for ld / st)
    uint64_t* bar;
    uint32_t phase;
};

// The state of one trajectory in shared memory: dimensions (2r, 2r+1) as one float2 at p[r * 128] (the 128 threads of
// a group read / write consecutive 8-byte words: conflict free, and a pair is one LDS.64 into an aligned register pair).
struct XPair {
    float2* p;
    __device__ __forceinline__ float operator[](int j) const { return reinterpret_cast<const float*>(p + (j >> 1) * 128)[j & 1]; }
    __device__ __forceinline__ float2 pair(int r) const { return p[r * 128]; }
    __device__ __forceinline__ void set_pair(int r, float2 v) const { p[r * 128] = v; }
};

// state -> A operand of the input layer (bf16 hi / lo, two values per TMEM column)
template <int DPAD>
__device__ __forceinline__ void store_a_from_x(uint32_t addr_hi, uint32_t addr_lo, const XPair& x) {
#pragma unroll
    for (int c = 0; c < DPAD / 8; ++c) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) tc::split_bf16_pair2(x.pair(4 * c + q), hi[q], lo[q]);
        tc::tmem_st4(addr_hi + 4 * c, hi);
        tc::tmem_st4(addr_lo + 4 * c, lo);
    }
    if (DPAD % 16 == 8) {  // K of the MMA is a multiple of 16: clear the upper half of the last k-step
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
        tc::tmem_st4(addr_hi + DPAD / 2, z);
        tc::tmem_st4(addr_lo + DPAD / 2, z);
    }
}

// 16 accumulator columns -> + bias -> exact GELU -> bf16 hi/lo split -> A operand of the next layer.  All eight register
// pairs are processed in one basic block (eight independent dependency chains for the scheduler to interleave: the kernel
// runs four warps per scheduler, so instruction-level parallelism inside the warp is what hides the FMA / MUFU latencies).
__device__ __forceinline__ void gelu_split_store16(uint32_t addr_hi, uint32_t addr_lo, const float (&v)[16], const float4 (&b)[4]) {
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 bq = b[q >> 1];
        const float2 bb = (q & 1) ? make_float2(bq.z, bq.w) : make_float2(bq.x, bq.y);
        const float2 a = gelu_fast2(__fadd2_rn(make_float2(v[2 * q], v[2 * q + 1]), bb));
        tc::split_bf16_pair2(a, hi[q], lo[q]);
    }
    tc::tmem_st4(addr_hi, &hi[0]);
    tc::tmem_st4(addr_hi + 4u, &hi[4]);
    tc::tmem_st4(addr_lo, &lo[0]);
    tc::tmem_st4(addr_lo + 4u, &lo[4]);
}

// A operand is in TMEM: hand the layer to the tensor core (one thread issues, commit -> the group's mbarrier) ...
__device__ __forceinline__ void issue_layer(const GroupCtx& c, uint32_t w_hi_saddr, uint32_t w_lo_saddr, int K16, int N) {
    tc::wait_st();
    tc::fence_before();
    group_bar(c.g);
    if ((c.gtid >> 5) == (c.issuer >> 5)) {  // the whole issuing warp, converged: one elected lane issues (sdes_tc.cuh)
        tc::fence_after();
        tc::issue_layer_bf16x3_warp(c.t_d, c.t_hi, c.t_lo, w_hi_saddr, w_lo_saddr, K16, N);
        tc::mma_commit_elect(c.bar);
    }
}
// ... and wait until the accumulator can be read
__device__ __forceinline__ void wait_layer(GroupCtx& c) {
    tc::mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    tc::fence_after();
}

// The fused epilogue between two layers, streamed 8 columns at a time straight from the accumulator (TMEM) into the
// next A operand (TMEM): the 64-wide activation row never sits in registers, and one compact loop serves every layer.
// `bias` may point to global (time-embedding row, input layer) or shared memory (hidden biases).
__device__ __forceinline__ void layer_epilogue(const GroupCtx& c, const float* __restrict__ bias) {
    float a[16], b[16];
    tc::tmem_ld8(c.l_d, &a[0]);
    tc::tmem_ld8(c.l_d + 8u, &a[8]);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll 1
    for (int ch = 0; ch < 4; ch += 2) {  // 16 columns per block, two blocks per iteration (double-buffered TMEM loads)
        float4 pa[4], pb[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {  // bias loads are issued before the wait so their latency hides behind the TMEM load
            pa[q] = b4[4 * ch + q];
            pb[q] = b4[4 * ch + 4 + q];
        }
        tc::wait_ld_tie<16>(a);
        tc::tmem_ld8(c.l_d + 16u * (ch + 1), &b[0]);
        tc::tmem_ld8(c.l_d + 16u * (ch + 1) + 8u, &b[8]);
        gelu_split_store16(c.l_hi + 8u * ch, c.l_lo + 8u * ch, a, pa);
        tc::wait_ld_tie<16>(b);
        if (ch + 2 < 4) {
            tc::tmem_ld8(c.l_d + 16u * (ch + 2), &a[0]);
            tc::tmem_ld8(c.l_d + 16u * (ch + 2) + 8u, &a[8]);
        }
        gelu_split_store16(c.l_hi + 8u * (ch + 1), c.l_lo + 8u * (ch + 1), b, pb);
    }
}

// ------------------------------------------------------------------- per-step scalars
// Both integrators advance the state as x' = A x + Bc g + Cc eps and accumulate sum gm^2 and sum gm eps:
//   Euler-Maruyama (losses/oc.py:204-219, :316-331):  A = 1 + mu dt, Bc = sigma dt, Cc = sigma sqrt(dt);
//       rnd += 1/2 dt sum gm^2 (+ sqrt(dt) sum gm eps)
//   exponential integrator (:429-443):                A = alpha_k, Bc = beta_k^2 sigma^2, Cc = sigma beta_k;
//       rnd += 1/2 beta_k^2 sigma^2 sum g^2 (+ sigma beta_k sum g eps)
struct StepK {
    float A, Bc, Cc, cost_scale, ito_scale;
    float sigma, lerp_w, one_m_w, outer, cm, cs;
    bool w_lt_half, ref_ctrl, from_hbm;
    int dim;
};

template <int CTRL>
__device__ __forceinline__ StepK make_step_k(const SdesRolloutDesc& d, const float* __restrict__ tab) {
    StepK k;
    const float dt = tab[TAB_DT], sqrt_dt = tab[TAB_SQRT_DT], mu = tab[TAB_MU], sigma = tab[TAB_SIGMA];
    k.sigma = sigma;
    if (d.loss_kind == SDES_LOSS_EXP_INTEGRATOR) {
        const float beta_k = tab[TAB_BETA_K], sg = d.sigma;
        const float bb_ss = (beta_k * beta_k) * (sg * sg);
        k.A = tab[TAB_ALPHA_K];
        k.Bc = bb_ss;
        k.Cc = sg * beta_k;
        k.cost_scale = 0.5f * bb_ss;
        k.ito_scale = sg * beta_k;
    } else {
        k.A = fmaf(mu, dt, 1.0f);
        k.Bc = sigma * dt;
        k.Cc = sigma * sqrt_dt;
        k.cost_scale = 0.5f * dt;
        k.ito_scale = sqrt_dt;
    }
    k.lerp_w = tab[TAB_LERP_W];
    k.one_m_w = 1.0f - k.lerp_w;
    k.w_lt_half = k.lerp_w < 0.5f;
    k.outer = (CTRL == SDES_CTRL_SCORE ? 1.0f : sigma) * d.scale_score;
    k.cm = d.clip_model;
    k.cs = d.clip_score;
    k.ref_ctrl = (d.flags & SDES_F_REFERENCE_CTRL) != 0;
    k.from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    k.dim = d.dim;
    return k;
}

// ------------------------------------------------------------------- noise
// Four standard normals for (trajectory, step, dim chunk): Philox4x32-10 with the round keys read from the kernel
// parameters, then two Box-Muller pairs.  Same stream as normal4_call / oracle/philox.py.
__device__ __forceinline__ float4 normal4_rk(const uint32_t (&rk)[20], uint32_t traj, uint32_t step, uint32_t chunk) {
    uint32_t c0 = traj, c1 = step, c2 = chunk, c3 = PHILOX_STREAM;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ rk[2 * r];
        c1 = lo1;
        c2 = hi0 ^ c3 ^ rk[2 * r + 1];
        c3 = lo0;
    }
    float4 e;
    box_muller(c0, c1, e.x, e.y);
    box_muller(c2, c3, e.z, e.w);
    return e;
}

// Eight normals for dimension chunks (chunk, chunk + 1): the two Philox streams advance in lock step so that the
// scheduler always has two independent multiply chains in flight.
__device__ __forceinline__ void normal8_rk(const uint32_t (&rk)[20], uint32_t traj, uint32_t step, uint32_t chunk, float (&e)[8]) {
    uint32_t a0 = traj, a1 = step, a2 = chunk, a3 = PHILOX_STREAM;
    uint32_t b0 = traj, b1 = step, b2 = chunk + 1u, b3 = PHILOX_STREAM;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t ah0 = __umulhi(PHILOX_M0, a0), al0 = PHILOX_M0 * a0, ah1 = __umulhi(PHILOX_M1, a2), al1 = PHILOX_M1 * a2;
        const uint32_t bh0 = __umulhi(PHILOX_M0, b0), bl0 = PHILOX_M0 * b0, bh1 = __umulhi(PHILOX_M1, b2), bl1 = PHILOX_M1 * b2;
        a0 = ah1 ^ a1 ^ rk[2 * r];
        b0 = bh1 ^ b1 ^ rk[2 * r];
        a1 = al1;
        b1 = bl1;
        a2 = ah0 ^ a3 ^ rk[2 * r + 1];
        b2 = bh0 ^ b3 ^ rk[2 * r + 1];
        a3 = al0;
        b3 = bl0;
    }
    box_muller(a0, a1, e[0], e[1]);
    box_muller(b0, b1, e[4], e[5]);
    box_muller(a2, a3, e[2], e[3]);
    box_muller(b2, b3, e[6], e[7]);
}

// ------------------------------------------------------------------- target "globals"
// What the per-dimension score of the target needs from the WHOLE state, evaluated once per step before the MLP
// (in the shadow of the input layer's MMAs).
struct TgtGlobals {
    float2 gs[GMM_ACT / 2];  // GMM: score on the leading dimension pairs that differ between components
    int np;                  // GMM: number of such pairs (1, 2 or 4)
    float f_ninv, f_s0;      // funnel: -exp(-x_0), score of dimension 0
};

// Shared-memory parameter images of the lean kernel
struct LeanSmem {
    const float* gmu;    // [K4][GMM_ACT]  -mu_k on the leading dims (K padded to a multiple of 4 with h = 0, c = -inf)
    const float* gh;     // [K4][GMM_ACT]  h_k = 1/2 sigma_k^-2
    const float* c2;     // [64]           log2(e) (log w_k - sum_j log sigma_kj - d/2 log 2 pi)
    const float* nmu0;   // [DPAD]         -mu of component 0 (dims shared by all components)
    const float* nh20;   // [DPAD]         -2 h of component 0
    const float* prior;  // loc[DPAD] | inv_var[DPAD] | lognorm
    const float* bo;     // output-layer bias [NOUT]
    int K4;
};

// Responsibility-weighted score on the first 2 NP dims (distr/gauss.py:119-140 through autograd, distr/base.py:130-137;
// analytic form SURVEY App. A.4): online softmax over the components, two per iteration, logits in log2 units,
// direct (x - mu)^2 form (the expanded form cancels catastrophically for modes at |mu| ~ 40).
template <int NP>
__device__ __forceinline__ void gmm_active_score(const XPair& xs, const LeanSmem& sm, float2 (&gs)[GMM_ACT / 2]) {
    float2 xv[NP], acc[NP];
#pragma unroll
    for (int r = 0; r < NP; ++r) {
        xv[r] = xs.pair(r);
        acc[r] = make_float2(0.f, 0.f);
    }
    float m = -INFINITY, ssum = 0.f;
#pragma unroll 1
    for (int k = 0; k < sm.K4; k += 4) {  // four components per iteration: independent chains for the scheduler
        const float2* mu = reinterpret_cast<const float2*>(sm.gmu + k * GMM_ACT);
        const float2* hh = reinterpret_cast<const float2*>(sm.gh + k * GMM_ACT);
        float q[4] = {0.f, 0.f, 0.f, 0.f};
        float2 t[4][NP];
#pragma unroll
        for (int r = 0; r < NP; ++r) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float2 dd = __fadd2_rn(xv[r], mu[r + u * (GMM_ACT / 2)]);
                t[u][r] = __fmul2_rn(dd, hh[r + u * (GMM_ACT / 2)]);
                q[u] = fmaf(dd.x, t[u][r].x, q[u]);
                q[u] = fmaf(dd.y, t[u][r].y, q[u]);
            }
        }
        const float4 c4 = *reinterpret_cast<const float4*>(sm.c2 + k);
        const float l[4] = {fmaf(q[0], -LOG2E, c4.x), fmaf(q[1], -LOG2E, c4.y), fmaf(q[2], -LOG2E, c4.z), fmaf(q[3], -LOG2E, c4.w)};
        const float m_new = fmaxf(fmaxf(m, fmaxf(l[0], l[1])), fmaxf(l[2], l[3]));
        float resc, e[4];
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(resc) : "f"(m - m_new));  // 1 when the max did not move, 0 on the first group
#pragma unroll
        for (int u = 0; u < 4; ++u) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[u]) : "f"(l[u] - m_new));
        m = m_new;
        ssum = fmaf(ssum, resc, (e[0] + e[1]) + (e[2] + e[3]));
#pragma unroll
        for (int r = 0; r < NP; ++r) {
            acc[r] = __fmul2_rn(acc[r], make_float2(resc, resc));
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[r] = __ffma2_rn(make_float2(e[u], e[u]), t[u][r], acc[r]);
        }
    }
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(ssum));
    inv *= -2.0f;  // score_j = -sum_k p_k (x_j - mu_kj) / sigma_kj^2 = -2 sum_k p_k h_kj (x_j - mu_kj)
#pragma unroll
    for (int r = 0; r < NP; ++r) gs[r] = __fmul2_rn(acc[r], make_float2(inv, inv));
}

// DENSE mixtures: score of EVERY dimension, components in groups of four, two sweeps over the dimension pairs per group —
// (1) the four logits, (2) after the online-softmax update, the responsibility-weighted (x - mu) h terms — with the
// mixture image in shared memory as one float4 {-mu pair, h pair} per (component, dimension pair) and packed math:
// 8 instructions per (component, pair).  Same direct (x - mu)^2 form and log2-unit softmax as gmm_active_score.
template <int DPAD>
__device__ __forceinline__ void gmm_dense_score(const XPair& xs, const float4* __restrict__ img, const float* __restrict__ c2, const int K4,
                                                float (&sc)[DPAD]) {
    constexpr int NPAIR = DPAD / 2;
    float2 acc[NPAIR];
#pragma unroll
    for (int r = 0; r < NPAIR; ++r) acc[r] = make_float2(0.f, 0.f);
    float m = -INFINITY, ssum = 0.f;
#pragma unroll 1
    for (int k = 0; k < K4; k += 4) {
        const float4* g = img + k * NPAIR;
        float2 q2[4] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
#pragma unroll
        for (int r = 0; r < NPAIR; ++r) {
            const float2 x2 = xs.pair(r);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 v = g[u * NPAIR + r];
                const float2 dd = __fadd2_rn(x2, make_float2(v.x, v.y));
                q2[u] = __ffma2_rn(dd, __fmul2_rn(dd, make_float2(v.z, v.w)), q2[u]);
            }
        }
        const float4 c4 = *reinterpret_cast<const float4*>(c2 + k);
        const float l[4] = {fmaf(q2[0].x + q2[0].y, -LOG2E, c4.x), fmaf(q2[1].x + q2[1].y, -LOG2E, c4.y),
                            fmaf(q2[2].x + q2[2].y, -LOG2E, c4.z), fmaf(q2[3].x + q2[3].y, -LOG2E, c4.w)};
        const float m_new = fmaxf(fmaxf(m, fmaxf(l[0], l[1])), fmaxf(l[2], l[3]));
        float resc, e[4];
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(resc) : "f"(m - m_new));
#pragma unroll
        for (int u = 0; u < 4; ++u) asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e[u]) : "f"(l[u] - m_new));
        m = m_new;
        ssum = fmaf(ssum, resc, (e[0] + e[1]) + (e[2] + e[3]));
        if (__any_sync(0xffffffffu, resc != 1.0f)) {  // the running maximum moved for some trajectory of the warp
#pragma unroll
            for (int r = 0; r < NPAIR; ++r) acc[r] = __fmul2_rn(acc[r], make_float2(resc, resc));
        }
#pragma unroll
        for (int r = 0; r < NPAIR; ++r) {
            const float2 x2 = xs.pair(r);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 v = g[u * NPAIR + r];
                const float2 t = __fmul2_rn(__fadd2_rn(x2, make_float2(v.x, v.y)), make_float2(v.z, v.w));
                acc[r] = __ffma2_rn(make_float2(e[u], e[u]), t, acc[r]);
            }
        }
    }
    float inv;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(ssum));
    inv *= -2.0f;
#pragma unroll
    for (int r = 0; r < NPAIR; ++r) {
        sc[2 * r] = acc[r].x * inv;
        sc[2 * r + 1] = acc[r].y * inv;
    }
}

template <int DPAD, int TGT>
__device__ __forceinline__ void target_globals(const SdesRolloutDesc& d, const XPair& xs, const LeanSmem& sm, uint32_t gmm_mask, TgtGlobals& tg) {
    if (TGT == SDES_TARGET_GMM) {
        if (gmm_mask < 2u) {
            tg.np = 1;
            gmm_active_score<1>(xs, sm, tg.gs);
        } else if (gmm_mask < 4u) {
            tg.np = 2;
            gmm_active_score<2>(xs, sm, tg.gs);
        } else {
            tg.np = 4;
            gmm_active_score<4>(xs, sm, tg.gs);
        }
    } else if (TGT == SDES_TARGET_FUNNEL) {
        // distr/funnel.py:71-80
        float2 sq2 = make_float2(0.f, 0.f);
        const float2 x01 = xs.pair(0);
        sq2.y = x01.y * x01.y;
#pragma unroll
        for (int r = 1; r < DPAD / 2; ++r) {
            const float2 v = xs.pair(r);
            sq2 = __ffma2_rn(v, v, sq2);
        }
        const float inv = expf(-x01.x);
        tg.f_ninv = -inv;
        tg.f_s0 = -x01.x / d.variance - 0.5f * (float)(d.dim - 1) + 0.5f * (sq2.x + sq2.y) * inv;
    }
}

constexpr bool ctrl_needs_target(int ctrl) { return ctrl == SDES_CTRL_SCORE || ctrl == SDES_CTRL_LERP || ctrl == SDES_CTRL_LERP_TARGET || ctrl == 100; }
constexpr bool ctrl_needs_prior(int ctrl) { return ctrl == SDES_CTRL_LERP || ctrl == SDES_CTRL_LERP_PRIOR || ctrl == 100; }

// ------------------------------------------------------------------- the update loop
// Eight dimensions (four pairs) of one trajectory: network output chunk from TMEM, control = clip(NN) + score part
// (models/reparam.py: ClippedCtrl :35-36, ScoreCtrl :78-83, LerpCtrl :131-162, LerpPriorCtrl :165-181, LerpTargetCtrl :184-200),
// noise, cost / Ito sums, state update.  MODE 0: first chunk (may hold the GMM's active pairs and the funnel's
// dimension 0), MODE 1: any later chunk, MODE 2: DENSE (target score of every dimension in scd[]).
constexpr int CTRL_LERP_HI = 100;  // LerpCtrl with s / T >= 0.5: the other branch of torch.lerp (compile-time so that no select is issued per pair)

template <int DPAD, int CTRL, int TGT, int MODE, bool QG>
__device__ __forceinline__ void chunk_update(const KParams& p, const StepK& k, const GroupCtx& c, const XPair& xs, const int q, const int step,
                                             const uint32_t traj, const LeanSmem& sm, const TgtGlobals& tg, const float* scd,
                                             const float* __restrict__ gate_row, const float* __restrict__ noise_row, float* xo,
                                             const int xo_st, float2& cost2, float2& ito2, float2& qs2, float* sko) {
    const SdesRolloutDesc& d = p.d;
    float nn[8];
    tc::tmem_ld8(c.l_d + 8u * q, nn);
    const int j0 = 8 * q;
    float e[8];
    if (k.from_hbm) {
#pragma unroll
        for (int r = 0; r < 8; ++r) e[r] = (j0 + r < k.dim) ? noise_row[j0 + r] : 0.f;
    } else if (j0 + 8 <= k.dim) {
        normal8_rk(p.philox_rk, traj, (uint32_t)step, (uint32_t)(2 * q), e);
    } else {
        const float4 n0 = normal4_rk(p.philox_rk, traj, (uint32_t)step, (uint32_t)(2 * q));
        e[0] = n0.x; e[1] = n0.y; e[2] = n0.z; e[3] = n0.w;
        if (j0 + 4 < k.dim) {
            const float4 n1 = normal4_rk(p.philox_rk, traj, (uint32_t)step, (uint32_t)(2 * q + 1));
            e[4] = n1.x; e[5] = n1.y; e[6] = n1.z; e[7] = n1.w;
        } else {
            e[4] = e[5] = e[6] = e[7] = 0.f;
        }
        if (j0 + 8 > k.dim) {  // the chunk that holds the last dimension: padded dimensions get no noise (their state stays 0)
#pragma unroll
            for (int r = 0; r < 8; ++r) e[r] = (j0 + r < k.dim) ? e[r] : 0.f;
        }
    }
    tc::wait_ld_tie<8>(nn);
    float2 xn[4];
#pragma unroll
    for (int pp = 0; pp < 4; ++pp) {  // one basic block for the four pairs (the trajectory stores follow the loop)
        const int r = 4 * q + pp;  // pair index
        float2 x2 = xs.pair(r);
        const float2 b2 = *reinterpret_cast<const float2*>(sm.bo + 2 * r);
        float2 g2 = __fadd2_rn(make_float2(nn[2 * pp], nn[2 * pp + 1]), b2);
        g2.x = clipf(g2.x, k.cm);
        g2.y = clipf(g2.y, k.cm);
        float2 ps2 = make_float2(0.f, 0.f);
        if (ctrl_needs_prior(CTRL) || k.ref_ctrl) {  // (loc - x) / scale^2   distr/gauss.py:222-223
            const float2 pl = *reinterpret_cast<const float2*>(sm.prior + 2 * r), iv = *reinterpret_cast<const float2*>(sm.prior + DPAD + 2 * r);
            ps2 = __fmul2_rn(__ffma2_rn(x2, make_float2(-1.f, -1.f), pl), iv);
        }
        if (CTRL != SDES_CTRL_CLIPPED) {
            float2 ts2 = make_float2(0.f, 0.f);
            if (ctrl_needs_target(CTRL)) {
                if (TGT == SDES_TARGET_GMM) {
                    if (MODE == 2) {
                        ts2 = make_float2(scd[2 * r], scd[2 * r + 1]);
                    } else {
                        // dimensions shared by all components factor out of the mixture: Gaussian score 2 h (mu - x)
                        const float2 nmu = *reinterpret_cast<const float2*>(sm.nmu0 + 2 * r), nh2 = *reinterpret_cast<const float2*>(sm.nh20 + 2 * r);
                        ts2 = __fmul2_rn(__fadd2_rn(x2, nmu), nh2);
                        if (MODE == 0 && pp < tg.np) ts2 = tg.gs[pp];
                    }
                } else if (TGT == SDES_TARGET_MULTIWELL) {
                    // distr/double_well.py:39-45, :165-179
                    const float2 y = __fadd2_rn(x2, make_float2(-d.shift, -d.shift));
                    const float2 a = __ffma2_rn(y, y, make_float2(-d.separation, -d.separation));
                    const float2 dw = __fmul2_rn(__fmul2_rn(a, y), make_float2(-4.f, -4.f));
                    ts2.x = (2 * r < d.n_double_wells) ? dw.x : (2 * r < k.dim ? -y.x : 0.f);
                    ts2.y = (2 * r + 1 < d.n_double_wells) ? dw.y : (2 * r + 1 < k.dim ? -y.y : 0.f);
                } else {
                    // distr/funnel.py:71-80: -x_j exp(-x_0) for j >= 1
                    ts2 = __fmul2_rn(x2, make_float2(tg.f_ninv, tg.f_ninv));
                    if (MODE == 0 && pp == 0) ts2.x = tg.f_s0;
                }
            }
            float2 inner;
            if (CTRL == SDES_CTRL_LERP || CTRL == CTRL_LERP_HI) {
                // torch.lerp: w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
                const float2 diff = __ffma2_rn(ps2, make_float2(-1.f, -1.f), ts2);
                if (CTRL == SDES_CTRL_LERP) inner = __ffma2_rn(make_float2(k.lerp_w, k.lerp_w), diff, ps2);
                else inner = __ffma2_rn(diff, make_float2(-k.one_m_w, -k.one_m_w), ts2);
            } else if (CTRL == SDES_CTRL_LERP_PRIOR) {
                inner = __fmul2_rn(ps2, make_float2(k.one_m_w, k.one_m_w));
            } else if (CTRL == SDES_CTRL_LERP_TARGET) {
                inner = __fmul2_rn(ts2, make_float2(k.lerp_w, k.lerp_w));
            } else {
                inner = ts2;
            }
            inner.x = clipf(inner.x, k.cs);
            inner.y = clipf(inner.y, k.cs);
            // gate row of the prologue's table: clip(score_model(s)) per dimension (a scalar gate replicated, 0 on padding)
            const float2 gt = *reinterpret_cast<const float2*>(gate_row + 2 * r);
            const float2 u2 = __fmul2_rn(inner, make_float2(k.outer, k.outer));  // the ungated score part
            g2 = __ffma2_rn(u2, gt, g2);
            if (QG) {
                if (sko != nullptr) {  // kl / kl_ito training forward: the ungated score part, kept for the reverse sweep (score_keep)
                    if (2 * r < k.dim) sko[(2 * r) * xo_st] = u2.x;
                    if (2 * r + 1 < k.dim) sko[(2 * r + 1) * xo_st] = u2.y;
                } else {
                    qs2 = __ffma2_rn(u2, make_float2(e[2 * pp], e[2 * pp + 1]), qs2);  // d rnd / d gate of the lv losses (SdesRolloutDesc.gate_cot)
                }
            }
        }
        float2 gm2 = g2;
        if (k.ref_ctrl) gm2 = __ffma2_rn(ps2, make_float2(-k.sigma, -k.sigma), g2);  // g - sigma * prior score  (solver/oc.py:305-306)
        const float2 e2 = make_float2(e[2 * pp], e[2 * pp + 1]);
        cost2 = __ffma2_rn(gm2, gm2, cost2);
        ito2 = __ffma2_rn(gm2, e2, ito2);
        x2 = __ffma2_rn(make_float2(k.A, k.A), x2, __ffma2_rn(make_float2(k.Bc, k.Bc), g2, __fmul2_rn(e2, make_float2(k.Cc, k.Cc))));
        xs.set_pair(r, x2);
        xn[pp] = x2;
    }
    if (xo != nullptr) {
#pragma unroll
        for (int pp = 0; pp < 4; ++pp) {
            const int r = 4 * q + pp;
            if (2 * r < k.dim) xo[(2 * r) * xo_st] = xn[pp].x;
            if (2 * r + 1 < k.dim) xo[(2 * r + 1) * xo_st] = xn[pp].y;
        }
    }
}

template <int DPAD, int CTRL, int TGT, bool DENSE, bool QG>
__device__ __forceinline__ void update_phase(const KParams& p, const StepK& k, const GroupCtx& c, const XPair& xs, const int step, const uint32_t traj,
                                             const LeanSmem& sm, const TgtGlobals& tg, const float* scd, const float* __restrict__ gate_row,
                                             const float* __restrict__ noise_row, float* xo, const int xo_st, float2& cost2, float2& ito2, float2& qs2,
                                             float* sko) {
    if (DENSE) {
#pragma unroll
        for (int q = 0; q < DPAD / 8; ++q)
            if (8 * q < k.dim) chunk_update<DPAD, CTRL, TGT, 2, QG>(p, k, c, xs, q, step, traj, sm, tg, scd, gate_row, noise_row, xo, xo_st, cost2, ito2, qs2, sko);
    } else {
        chunk_update<DPAD, CTRL, TGT, 0, QG>(p, k, c, xs, 0, step, traj, sm, tg, scd, gate_row, noise_row, xo, xo_st, cost2, ito2, qs2, sko);
        const int nq = (k.dim + 7) >> 3;
#pragma unroll 1
        for (int q = 1; q < nq; ++q) chunk_update<DPAD, CTRL, TGT, 1, QG>(p, k, c, xs, q, step, traj, sm, tg, scd, gate_row, noise_row, xo, xo_st, cost2, ito2, qs2, sko);
    }
}

// CTRL = the descriptor's ctrl_kind, TGT = its target_kind where the step depends on it (controls with a target score), else
// SDES_TARGET_GMM as a placeholder: the host dispatches on both (launch_rollout_tc_dpad below), so each kernel is one
// straight-line step with at most two update loops (the two branches of torch.lerp).
template <int DPAD, bool DENSE, int CTRL, int TGT>
__global__ void __launch_bounds__(TC_THREADS, 1) rollout_tc_kernel(const __grid_constant__ KParams p) {
    constexpr int NOUT = (DPAD + 15) / 16 * 16;
    constexpr uint32_t K0B = (DPAD + 15) & ~15;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t s_wbar;
    __shared__ uint64_t s_mbar[TC_GROUPS];
    __shared__ uint64_t s_xbar[TC_GROUPS];  // arrival of a tile's x0 rows (TMA bulk copy)
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_tile[TC_GROUPS];

    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, T = d.n_steps, K = d.target_kind == SDES_TARGET_GMM ? d.n_components : 0, nh = d.n_hidden;
    const int tid = threadIdx.x, warp = tid >> 5;

    // Which instantiation runs: the prologue kernel left the GMM's dimension-pair mask in the workspace.  A mixture
    // whose components differ beyond the first GMM_ACT dims needs the per-step score of every dimension (DENSE).
    const uint32_t gmm_mask = reinterpret_cast<const uint32_t*>(ws + p.ws.counter)[1];
    const bool want_dense = TGT == SDES_TARGET_GMM && ctrl_needs_target(CTRL) && gmm_mask >= (1u << (GMM_ACT / 2));
    if (want_dense != DENSE) return;

    // ---- shared memory: [bf16 weight image + biases | gmm -mu, h (leading dims) | c2 | -mu0 | -2 h0 | prior | ref | state x]
    float* s_w = smem;
    const int K2 = (K + 1) & ~1, K4 = (K + 3) & ~3;
    float* s_gmu = s_w + ((p.ws.w_mma4_len + 31) & ~31ll);
    float* s_gh = s_gmu + K4 * GMM_ACT;
    float* s_c2 = s_gh + K4 * GMM_ACT;
    float* s_nmu0 = s_c2 + 64;
    float* s_nh20 = s_nmu0 + DPAD;
    float* s_prior = s_nh20 + DPAD;
    float* s_ref = s_prior + 2 * DPAD + 8;
    float* s_x = s_ref + 2 * DPAD + 8;
    float4* s_gimg = reinterpret_cast<float4*>(s_x + TC_GROUPS * DPAD * 128);  // DENSE: full mixture image (when it fits)
    const bool dense_smem = DENSE && p.dense_smem != 0;

    if (warp == 0) {
        tc::tmem_alloc(&s_tmem, TMEM_COLS);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(&s_wbar, 1);
        for (int g = 0; g < TC_GROUPS; ++g) {
            tc::mbar_init(&s_mbar[g], 1);
            tc::mbar_init(&s_xbar[g], 1);
        }
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (tid == 0) {
        // weights: TMA bulk copies global -> shared, all counted on one mbarrier
        const uint32_t total = (uint32_t)(p.ws.w_mma4_len * sizeof(float));
        tc::mbar_arrive_expect_tx(&s_wbar, total);
        const char* src = reinterpret_cast<const char*>(ws + p.ws.w_mma4);
        char* dst = reinterpret_cast<char*>(s_w);
        for (uint32_t off = 0; off < total; off += 16384u) {
            const uint32_t n = total - off < 16384u ? total - off : 16384u;
            tc::bulk_g2s(dst + off, src + off, n, &s_wbar);
        }
    }
    for (int e = tid; e < K4 * GMM_ACT; e += blockDim.x) {
        const int k = e / GMM_ACT, j = e % GMM_ACT;
        s_gmu[e] = k < K2 ? -ws[p.ws.gmm_mu + k * DPAD + j] : 0.f;
        s_gh[e] = k < K2 ? ws[p.ws.gmm_h + k * DPAD + j] : 0.f;
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c2[e] = (e < K2 ? ws[p.ws.gmm_c + e] : -INFINITY) * LOG2E;
    if (dense_smem) {
        for (int e = tid; e < K4 * (DPAD / 2); e += blockDim.x) {
            const int k = e / (DPAD / 2), r = e % (DPAD / 2);
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < K2) v = make_float4(-ws[p.ws.gmm_mu + k * DPAD + 2 * r], -ws[p.ws.gmm_mu + k * DPAD + 2 * r + 1],
                                        ws[p.ws.gmm_h + k * DPAD + 2 * r], ws[p.ws.gmm_h + k * DPAD + 2 * r + 1]);
            s_gimg[e] = v;
        }
    }
    for (int e = tid; e < DPAD; e += blockDim.x) {
        s_nmu0[e] = K > 0 ? -ws[p.ws.gmm_mu + e] : 0.f;
        s_nh20[e] = K > 0 ? -2.0f * ws[p.ws.gmm_h + e] : 0.f;
    }
    for (int e = tid; e < 2 * DPAD + 8; e += blockDim.x) {
        s_prior[e] = e <= 2 * DPAD ? ws[p.ws.prior + e] : 0.f;
        s_ref[e] = e <= 2 * DPAD ? ws[p.ws.ref + e] : 0.f;
    }
    tc::mbar_wait(&s_wbar, 0);
    __syncthreads();

    // weight image addresses (bytes): per layer hi then lo
    const uint32_t w_base = tc::smem_u32(s_w);
    constexpr uint32_t L0_HALF = 64u * K0B * 2u, LH_HALF = 64u * 64u * 2u, LO_HALF = (uint32_t)NOUT * 64u * 2u;
    const uint32_t l0_hi = w_base, l0_lo = l0_hi + L0_HALF;
    const uint32_t lh_base = l0_lo + L0_HALF;
    const uint32_t lo_hi = lh_base + (uint32_t)nh * 2u * LH_HALF, lo_lo = lo_hi + LO_HALF;
    const float* s_bias = reinterpret_cast<const float*>(reinterpret_cast<const char*>(s_w) + 2 * L0_HALF + (size_t)nh * 2 * LH_HALF + 2 * LO_HALF);

    // generic (once-per-tile) evaluations read the full-width target images in the workspace
    const TargetSmem tsm{ws + p.ws.gmm_mu, ws + p.ws.gmm_h, ws + p.ws.gmm_c, gmm_mask, s_prior, s_ref};
    const LeanSmem lsm{s_gmu, s_gh, s_c2, s_nmu0, s_nh20, s_prior, s_bias + nh * C, K4};
    GroupCtx c;
    c.g = warp >> 2;
    c.gtid = tid & 127;
    c.issuer = (c.g & 3) * 32;
    const uint32_t tbase = s_tmem + (uint32_t)(c.g * GROUP_COLS);
    c.t_d = tbase;
    c.t_hi = tbase + 64;
    c.t_lo = tbase + 96;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    c.l_d = c.t_d + lane_off;
    c.l_hi = c.t_hi + lane_off;
    c.l_lo = c.t_lo + lane_off;
    c.bar = &s_mbar[c.g];
    c.phase = 0;
    const XPair xs{reinterpret_cast<float2*>(s_x + c.g * (DPAD * 128)) + c.gtid};
    // The same region, seen row-major [row][dim], stages a tile's x0 / x_T between HBM and the pair layout: a tile's rows are
    // one contiguous span of the caller's (B, d) tensors, moved by ONE TMA bulk copy each way (cp.async.bulk, SASS UBLKCP)
    // instead of 128 threads striding through global memory 4 bytes at a time.
    float* xrow = s_x + c.g * (DPAD * 128);
    uint32_t xphase = 0;
    const bool io_aligned = ((reinterpret_cast<uintptr_t>(d.x0) | reinterpret_cast<uintptr_t>(d.x_T)) & 15) == 0;

    uint32_t* counter = reinterpret_cast<uint32_t*>(const_cast<float*>(ws) + p.ws.counter);
    const bool from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    const bool ret_traj = (d.flags & SDES_F_RETURN_TRAJ) != 0;
    const int64_t B = d.batch;
    const uint32_t n_tiles = (uint32_t)((B + 127) / 128);

    // Work items are (time chunk, tile) pairs handed out chunk-major from one counter: a tile's T steps are cut into
    // n_chunks pieces so that the tiles do not quantise into rounds with a mostly idle last one.  Between chunks the
    // tile's state (x, rnd) parks in the workspace ([tile][j][128] so warps read and write whole 128-byte lines, L2
    // resident) and a per-tile progress word orders producer and consumer.  An item only ever waits on an item handed
    // out earlier, so there is no deadlock whatever the residency.
    const int n_chunks = p.n_chunks, chunk_steps = p.chunk_steps;
    const uint32_t n_items = n_tiles * (uint32_t)n_chunks;
    float* state = const_cast<float*>(ws) + p.ws.state;
    uint32_t* progress = reinterpret_cast<uint32_t*>(const_cast<float*>(ws) + p.ws.progress);
    for (;;) {
        if (c.gtid == 0) s_tile[c.g] = atomicAdd(counter, 1u);
        group_bar(c.g);
        const uint32_t item = s_tile[c.g];
        if (item >= n_items) break;
        const uint32_t chunk = item / n_tiles, tile = item - chunk * n_tiles;
        const int64_t row = (int64_t)tile * 128 + c.gtid;
        const bool valid = row < B;
        const int64_t rrow = valid ? row : (B - 1);
        float* st = state + (int64_t)tile * (DPAD + 1) * 128 + c.gtid;

        float rnd;
        const int tile_rows = (int)((B - (int64_t)tile * 128 < 128) ? (B - (int64_t)tile * 128) : 128);
        const uint32_t tile_bytes = (uint32_t)tile_rows * (uint32_t)dim * 4u;
        const bool tile_tma = io_aligned && (tile_bytes & 15u) == 0u;  // bulk copies move multiples of 16 bytes (ragged last tile: plain loads)
        if (chunk == 0) {
            const TrajRef o0 = traj_ref(d, d.xs, 0, rrow);
            float xv[DPAD];
            if (tile_tma) {
                if (c.gtid == 0) {
                    tc::fence_proxy_async();  // the region was last touched through the generic proxy (previous item of this group)
                    tc::mbar_arrive_expect_tx(&s_xbar[c.g], tile_bytes);
                    tc::bulk_g2s(xrow, d.x0 + (int64_t)tile * 128 * dim, tile_bytes, &s_xbar[c.g]);
                }
                tc::mbar_wait(&s_xbar[c.g], xphase);
                xphase ^= 1u;
                const float* my = xrow + (valid ? c.gtid : tile_rows - 1) * dim;
#pragma unroll
                for (int j = 0; j < DPAD; ++j) xv[j] = (j < dim) ? my[j] : 0.f;
                group_bar(c.g);  // every row is in registers before the region is rewritten in the pair layout
            } else {
#pragma unroll
                for (int j = 0; j < DPAD; ++j) xv[j] = (j < dim) ? __ldg(d.x0 + rrow * dim + j) : 0.f;
            }
#pragma unroll
            for (int r = 0; r < DPAD / 2; ++r) {
                const float2 v = make_float2(xv[2 * r], xv[2 * r + 1]);
                xs.set_pair(r, v);
                if (ret_traj && valid) {
                    if (2 * r < dim) o0.p[(2 * r) * o0.stride] = v.x;
                    if (2 * r + 1 < dim) o0.p[(2 * r + 1) * o0.stride] = v.y;
                }
            }
            rnd = initial_rnd<DPAD>(d, xs, tsm);
        } else {
            if (c.gtid == 0) {
                uint32_t seen;
                for (;;) {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + tile) : "memory");
                    if (seen >= chunk) break;
                    __nanosleep(2000);  // the predecessor chunk runs for tens of microseconds: do not spend issue slots polling
                }
            }
            group_bar(c.g);
#pragma unroll
            for (int r = 0; r < DPAD / 2; ++r) xs.set_pair(r, make_float2(__ldcg(st + (2 * r) * 128), __ldcg(st + (2 * r + 1) * 128)));  // L2 reads: another SM wrote them
            rnd = __ldcg(st + DPAD * 128);
        }
        const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)rrow);
        const int i_begin = (int)chunk * chunk_steps;
        const int i_end = (i_begin + chunk_steps < T) ? i_begin + chunk_steps : T;

        for (int i = i_begin; i < i_end; ++i) {
            const float* tab = ws + p.ws.tab + (int64_t)i * TAB_STRIDE;
            const float* gate_row = ws + p.ws.gate + (int64_t)i * DPAD;
            // ---- control MLP on the tensor cores (models/mlp.py:114-122): input layer first, so that the target's global
            //      quantities are evaluated while its MMAs run
#ifdef SDES_TC_TIMELINE
            unsigned long long tl[16];
            int tli = 0;
#define TL() do { if (tli < 16) tl[tli++] = (unsigned long long)clock64(); } while (0)
#else
#define TL() do { } while (0)
#endif
            TL();
            store_a_from_x<DPAD>(c.l_hi, c.l_lo, xs);
            TL();
            issue_layer(c, l0_hi, l0_lo, (int)K0B, C);
            TL();
            TgtGlobals tg;
            float scd[DENSE ? DPAD : 1];
            if constexpr (DENSE) {
                // score of every dimension, components differing on the first 8 / 16 / all dimension pairs (sdes_step.cuh)
                constexpr int NPAIR = DPAD / 2;
                if (dense_smem) gmm_dense_score<DPAD>(xs, s_gimg, s_c2, K4, scd);
                else if (NPAIR > 8 && gmm_mask < 256u) gmm_eval_na<DPAD, (NPAIR > 8 ? 8 : NPAIR), true>(xs, scd, tsm, K);
                else if (NPAIR > 16 && gmm_mask < 65536u) gmm_eval_na<DPAD, (NPAIR > 16 ? 16 : NPAIR), true>(xs, scd, tsm, K);
                else gmm_eval_na<DPAD, NPAIR, true>(xs, scd, tsm, K);
            } else if (ctrl_needs_target(CTRL)) {
                target_globals<DPAD, TGT>(d, xs, lsm, gmm_mask, tg);
            }
            TL();
            wait_layer(c);
            TL();
            layer_epilogue(c, ws + p.ws.emb + (int64_t)i * C);  // + (emb_t + b_in), GELU
            TL();
#pragma unroll 1
            for (int l = 0; l < nh; ++l) {
                issue_layer(c, lh_base + (uint32_t)l * 2u * LH_HALF, lh_base + (uint32_t)l * 2u * LH_HALF + LH_HALF, C, C);
                TL();
                wait_layer(c);
                TL();
                layer_epilogue(c, s_bias + l * C);
                TL();
            }
            issue_layer(c, lo_hi, lo_lo, C, NOUT);
            TL();
            const StepK k = make_step_k<CTRL>(d, tab);
            const float* nrow = from_hbm ? d.noise + ((int64_t)i * B + rrow) * dim : nullptr;
            const TrajRef xo_ref = traj_ref(d, d.xs, i + 1, rrow);
            float* xo = (ret_traj && valid) ? xo_ref.p : nullptr;
            float2 cost2 = make_float2(0.f, 0.f), ito2 = make_float2(0.f, 0.f);
            wait_layer(c);
            TL();
            // ---- network output streamed from TMEM into the control / cost / state update
            float2 qs2 = make_float2(0.f, 0.f);
            const bool want_q = CTRL != SDES_CTRL_CLIPPED && (d.gate_cot != nullptr || d.score_keep != nullptr);
            float* sko = (d.score_keep != nullptr && valid) ? traj_ref(d, d.score_keep, i, rrow).p : nullptr;
#define SDES_UPD(C_, Q_) update_phase<DPAD, C_, TGT, DENSE, Q_>(p, k, c, xs, i, traj, lsm, tg, scd, gate_row, nrow, xo, xo_ref.stride, cost2, ito2, qs2, sko)
            if (CTRL == SDES_CTRL_LERP && !k.w_lt_half) {
                if (want_q) SDES_UPD(CTRL_LERP_HI, true);
                else SDES_UPD(CTRL_LERP_HI, false);
            } else {
                if (want_q) SDES_UPD(CTRL, true);
                else SDES_UPD(CTRL, false);
            }
#undef SDES_UPD
            if (want_q && valid && d.gate_cot != nullptr) d.gate_cot[(int64_t)i * B + rrow] = k.ito_scale * (qs2.x + qs2.y);
            rnd = fmaf(k.cost_scale, cost2.x + cost2.y, rnd);
            if (d.flags & SDES_F_SUB_DIV_INT) rnd -= tab[TAB_DIV_INT];
            if (d.flags & SDES_F_COMPUTE_ITO) rnd = fmaf(k.ito_scale, ito2.x + ito2.y, rnd);
            TL();
#ifdef SDES_TC_TIMELINE
            if (blockIdx.x < 4 && (tid & 127) == 0 && i >= 42 && i < 44) {
                for (int q = tli; q < 16; ++q) tl[q] = tl[0];
                printf("TLSTEP %d cta %d g %d: %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu %llu\n", i, (int)blockIdx.x, c.g,
                       tl[1] - tl[0], tl[2] - tl[0], tl[3] - tl[0], tl[4] - tl[0], tl[5] - tl[0], tl[6] - tl[0], tl[7] - tl[0], tl[8] - tl[0],
                       tl[9] - tl[0], tl[10] - tl[0], tl[11] - tl[0], tl[12] - tl[0], tl[13] - tl[0], tl[14] - tl[0], tl[15] - tl[0]);
            }
#endif
        }
        if (i_end == T) {
            rnd += terminal_rnd<DPAD>(d, xs, tsm);
            if (tile_tma) {
                float xv[DPAD];
#pragma unroll
                for (int j = 0; j < DPAD; ++j) xv[j] = xs[j];
                group_bar(c.g);  // all pair-layout reads done before the region becomes the row-major x_T tile
                if (valid) {
                    float* my = xrow + c.gtid * dim;
#pragma unroll
                    for (int j = 0; j < DPAD; ++j)
                        if (j < dim) my[j] = xv[j];
                    d.rnd[rrow] = rnd;
                }
                tc::fence_proxy_async();  // generic-proxy writes -> visible to the bulk copy (async proxy)
                group_bar(c.g);
                if (c.gtid == 0) {
                    tc::bulk_s2g(d.x_T + (int64_t)tile * 128 * dim, xrow, tile_bytes);
                    tc::bulk_wait_read();  // the region is reused by this group's next item
                }
            } else if (valid) {
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) d.x_T[rrow * dim + j] = xs[j];
                d.rnd[rrow] = rnd;
            }
        } else {
#pragma unroll
            for (int r = 0; r < DPAD / 2; ++r) {
                const float2 v = xs.pair(r);
                __stcg(st + (2 * r) * 128, v.x);
                __stcg(st + (2 * r + 1) * 128, v.y);
            }
            __stcg(st + DPAD * 128, rnd);
            __threadfence();
            group_bar(c.g);
            if (c.gtid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(progress + tile), "r"(chunk + 1u) : "memory");
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(s_tmem, TMEM_COLS);
}

// ---------------------------------------------------------------------------- launch
// This translation unit is compiled once per padded state dimension (-DSDES_TC_DPAD=8|16|32|48|56|64, sde_sampler_b200/build.py)
// so that the instantiations build in parallel; sdes_rollout_tc_api.cu dispatches on ws.dpad.
#ifndef SDES_TC_DPAD
#define SDES_TC_DPAD 56
#endif

template <int DPAD, bool DENSE, int CTRL, int TGT>
static cudaError_t launch_k(const KParams& p, int sm_count, cudaStream_t stream) {
    KParams q = p;
    q.dense_smem = (DENSE && tc_dense_image_fits(p)) ? 1 : 0;
    const size_t smem = tc_smem_bytes(p) + (q.dense_smem ? tc_dense_image_bytes(p) : 0);
    cudaError_t e = cudaFuncSetAttribute(rollout_tc_kernel<DPAD, DENSE, CTRL, TGT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // work items = tiles x time chunks: a group that finds no first-chunk tile left starts on a second chunk and waits
    // for its predecessor, so every SM is used even when there are fewer tiles than resident groups
    const int64_t items = ((p.d.batch + 127) / 128) * (int64_t)(p.n_chunks > 0 ? p.n_chunks : 1);
    int grid = (int)((items + TC_GROUPS - 1) / TC_GROUPS);
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    rollout_tc_kernel<DPAD, DENSE, CTRL, TGT><<<grid, TC_THREADS, smem, stream>>>(q);
    return cudaGetLastError();
}

// a control with a target score: the lean kernel and, when the descriptor can need it (a multi-component GMM in d > GMM_ACT),
// the DENSE one as well — whichever does not match the prologue's dimension mask exits at once
template <int DPAD, int CTRL>
static cudaError_t launch_target_ctrl(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches) {
    if (p.d.target_kind == SDES_TARGET_MULTIWELL) return launch_k<DPAD, false, CTRL, SDES_TARGET_MULTIWELL>(p, sm_count, stream);
    if (p.d.target_kind == SDES_TARGET_FUNNEL) return launch_k<DPAD, false, CTRL, SDES_TARGET_FUNNEL>(p, sm_count, stream);
    cudaError_t e = launch_k<DPAD, false, CTRL, SDES_TARGET_GMM>(p, sm_count, stream);
    if constexpr (DPAD > GMM_ACT) {
        if (e == cudaSuccess && p.d.n_components > 1 && p.d.dim > GMM_ACT) {
            *n_launches = 2;
            e = launch_k<DPAD, true, CTRL, SDES_TARGET_GMM>(p, sm_count, stream);
        }
    }
    return e;
}

#define SDES_TC_CAT2(a, b) a##b
#define SDES_TC_CAT(a, b) SDES_TC_CAT2(a, b)
cudaError_t SDES_TC_CAT(launch_rollout_tc_, SDES_TC_DPAD)(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches) {
    constexpr int DPAD = SDES_TC_DPAD;
    *n_launches = 1;
    switch (p.d.ctrl_kind) {
        case SDES_CTRL_CLIPPED: return launch_k<DPAD, false, SDES_CTRL_CLIPPED, SDES_TARGET_GMM>(p, sm_count, stream);
        case SDES_CTRL_LERP_PRIOR: return launch_k<DPAD, false, SDES_CTRL_LERP_PRIOR, SDES_TARGET_GMM>(p, sm_count, stream);
        case SDES_CTRL_SCORE: return launch_target_ctrl<DPAD, SDES_CTRL_SCORE>(p, sm_count, stream, n_launches);
        case SDES_CTRL_LERP: return launch_target_ctrl<DPAD, SDES_CTRL_LERP>(p, sm_count, stream, n_launches);
        case SDES_CTRL_LERP_TARGET: return launch_target_ctrl<DPAD, SDES_CTRL_LERP_TARGET>(p, sm_count, stream, n_launches);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sdes
