// sdes_common.cuh — shared device helpers and the host<->kernel parameter block.
// Part of the B200-native rollout behind include/sdes_b200.h.  sm_100a only.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sdes_b200.h"

namespace sdes {

constexpr int C = SDES_CHANNELS;  // network width
constexpr int TAB_STRIDE = 8;     // floats per step in the scalar table
// scalar table columns (one row per time step i, s = ts[i], t = ts[i+1])
enum { TAB_DT = 0, TAB_SQRT_DT = 1, TAB_MU = 2, TAB_SIGMA = 3, TAB_DIV_INT = 4, TAB_LERP_W = 5,
       TAB_BETA_K = 6, TAB_ALPHA_K = 7 };

constexpr float LOG_2PI = 1.8378770664093453f;

// Offsets (in floats) of everything the prologue kernel writes into the workspace.
struct WsLayout {
    int dpad;          // padded state dimension (template parameter of the rollout kernel)
    int64_t tab;       // T * TAB_STRIDE
    int64_t emb;       // T * C        FourierMLP.timestep_embed(s)        (models/mlp.py:116)
    int64_t gate;      // T * dpad     clip(score_model(s), clip_model)    (models/reparam.py:68-76)
    int64_t gmm_mu;    // K * dpad
    int64_t gmm_h;     // K * dpad     0.5 / scale^2
    int64_t gmm_c;     // K (padded to 64)   log w_k - sum_j log scale_kj - d/2 log 2pi
    int64_t prior;     // 2 * dpad     loc | 1/scale^2           , + [2*dpad] = log-normaliser
    int64_t ref;       // 2 * dpad + 4
    int64_t w_simt;    // SIMT weights: WtIn[d][C] bIn[C] {Wt[C][C] b[C]} x nh  WtOut[C][dpad] bOut[dpad]
    int64_t w_simt_len;
    int64_t w_mma4;    // bf16 hi/lo operand images of the tcgen05 engine (see sdes_rollout_mma.cu)
    int64_t w_mma4_len;
    int64_t counter;   // 4 uint32: dynamic work counter, GMM chunk mask
    int64_t progress;  // n_tiles128 uint32: completed time chunks per tile   (tcgen05 engine)
    int64_t state;     // n_tiles128 * (dpad+1) * 128: parked tile state between time chunks
    int64_t total;     // floats
};

// Offsets (in floats) inside the caller's parameter blob (layout: include/sdes_b200.h).
struct BlobLayout {
    int64_t in_w, in_b, te_phase, te_h_w[SDES_MAX_HIDDEN], te_h_b[SDES_MAX_HIDDEN], te_out_w, te_out_b;
    int64_t h_w[SDES_MAX_HIDDEN], h_b[SDES_MAX_HIDDEN], out_w, out_b;
    int64_t g_phase, g_h_w[SDES_MAX_HIDDEN], g_h_b[SDES_MAX_HIDDEN], g_out_w, g_out_b;
    int64_t total;
};

// Everything a kernel needs, passed by value.
struct KParams {
    SdesRolloutDesc d;
    WsLayout ws;
    BlobLayout bl;
    int n_tiles;      // warp tiles of 32 trajectories
    int n_chunks;     // tcgen05 engine: time chunks per tile
    int chunk_steps;  // steps per chunk
    int dense_smem;   // tcgen05 engine, DENSE instantiation: the full mixture image sits in shared memory (set by the launcher)
    uint32_t philox_rk[20];  // Philox4x32-10 round keys (k0 + r W0, k1 + r W1), r = 0..9, of d.seed: read by the tensor-core
                             // rollout as constant-bank operands instead of being re-derived per draw
};

// Parameters of the Langevin / Euler integrator kernel (sdes_integrate.cu)
struct IntegrateParams {
    SdesRolloutDesc d;       // target fields, dim, batch, seed, traj_offset, noise (optional), workspace
    WsLayout ws;
    const float* timesteps;  // (n_steps + 1) integration grid
    const float* out_ts;     // (n_out) output times
    int n_steps, n_out;
    float diff_coeff, clip_score, eps;
    const float* x_init;     // (B, d)
    float* xs_out;           // (n_out, B, d)
    int noise_is_increment;  // noise (HBM) holds Brownian increments instead of standard normals
};

__host__ __device__ inline int pad_dim(int d) {
    if (d <= 4) return 4;
    if (d <= 8) return 8;
    if (d <= 12) return 12;
    if (d <= 16) return 16;
    if (d <= 32) return 32;
    if (d <= 52) return 52;
    return 64;
}

// padded state dimension of the tcgen05 engine: K of the input layer (multiple of 8)
__host__ __device__ inline int mma_pad_dim(int d) { return d <= 8 ? 8 : d <= 16 ? 16 : d <= 32 ? 32 : d <= 48 ? 48 : d <= 56 ? 56 : 64; }

// Where trajectory point (step, row) lives in `xs`: element j at p[j * stride].  Reference layout (T+1, B, d), or the
// row-tiled layout of SDES_F_TRAJ_TILED, [step][row / 128][j][row % 128], that thread-per-trajectory kernels read
// and write with fully coalesced 128-byte lines.
struct TrajRef {
    float* p;
    int stride;
};
__device__ __forceinline__ TrajRef traj_ref(const SdesRolloutDesc& d, float* xs, int step, int64_t row) {
    if (d.flags & SDES_F_TRAJ_TILED) {
        const int64_t nt = (d.batch + 127) / 128;
        return TrajRef{xs + (((int64_t)step * nt + (row >> 7)) * mma_pad_dim(d.dim)) * 128 + (row & 127), 128};
    }
    return TrajRef{xs + ((int64_t)step * d.batch + row) * d.dim, 1};
}

// ------------------------------------------------------------------------------------ math
__device__ __forceinline__ float clipf(float v, float c) {
    // utils/common.py:83-84 `tensor.clip(-max_norm, max_norm)`; c = +inf means no clip. NaN propagates
    // (min/max.NaN, two FMNMX instead of the five of fminf/fmaxf plus an explicit NaN fix-up).
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(v), "f"(-c));
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(c));
    return r;
}

__device__ __forceinline__ float gelu_erf(float x) {
    // torch.nn.GELU() default (exact erf; conf/model/base/fouriermlp.yaml:5-6)
    return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f));
}

// The same function for the hot loop: GELU(x) = relu(x) - |x| Phi(-|x|) with Phi(-|x|) = 0.5 erfc(|x|/sqrt 2)
// from Abramowitz & Stegun 7.1.26 (|error| <= 1.5e-7 on erf), evaluated with MUFU.RCP / MUFU.EX2:
// 13 instructions instead of ~25 for erff and no select.  Measured max |error| vs float64 over
// [-8, 8]: < 6e-7 (torch's own fp32 GELU: 1.2e-6).
__device__ __forceinline__ float gelu_fast(float x) {
    // z = |x| sqrt(log2(e)/2): exp(-x^2/2) = 2^(-z z), and A&S's 1 + 0.3275911 |x|/sqrt(2) = 1 + 0.27273748 z
    const float ax = fabsf(x);
    const float z = ax * 0.8493218002880191f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2727374808792225f, z, 1.0f)));
    // 0.5 * (a1 t + a2 t^2 + a3 t^3 + a4 t^4 + a5 t^5)
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    p *= t * ax;                                // |x| * 0.5 erfc-prefactor
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z));  // exp(-x^2/2)
    return fmaf(-p, e, fmaxf(x, 0.0f));         // relu(x) - |x| Phi(-|x|)
}

// Two values at a time for the tensor-core rollout's epilogue, every FP32 operation as a packed f32x2 instruction
// (FFMA2 / FMUL2 / FADD2 take one issue slot for both lanes — the epilogue is issue-bound, not pipe-bound).
// Logistic form of the exact-erf GELU:  GELU(x) = x Phi(x) = x / (1 + 2^(-x P(x^2))),  x P(x^2) = log2(Phi(x) / Phi(-x)),
// P a degree-6 minimax fit (LP fit of the GELU error on [-6.5, 6.5], positive leading coefficient so the tails saturate
// to x and -0; tools/fit_gelu_logistic.py).  10 packed instructions + 4 MUFU per pair instead of 14 + 14 scalar ones;
// max |error| vs float64 exact-erf GELU 7e-7 + 1.2e-7 |x| (tests/test_gpu_tcgen05.py; torch's own fp32 GELU: 1.2e-6).
__device__ __forceinline__ float2 gelu_fast2(float2 x) {
    const float2 u = __fmul2_rn(x, x);
    // coefficients of -P: q = -x P(x^2) comes out of the last multiply with no separate negation
    float2 p = __ffma2_rn(make_float2(-5.42691260e-09f, -5.42691260e-09f), u, make_float2(3.93527977e-07f, 3.93527977e-07f));
    p = __ffma2_rn(p, u, make_float2(-1.15760618e-05f, -1.15760618e-05f));
    p = __ffma2_rn(p, u, make_float2(1.60239457e-04f, 1.60239457e-04f));
    p = __ffma2_rn(p, u, make_float2(9.27478302e-05f, 9.27478302e-05f));
    p = __ffma2_rn(p, u, make_float2(-1.04834383e-01f, -1.04834383e-01f));
    p = __ffma2_rn(p, u, make_float2(-2.30220913e+00f, -2.30220913e+00f));
    const float2 q = __fmul2_rn(x, p);
    float2 e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
    const float2 d = __fadd2_rn(e, make_float2(1.f, 1.f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    return __fmul2_rn(x, r);
}

// exact-erf GELU and its derivative Phi(x) + x phi(x) (autograd of torch.nn.GELU()), from the same A&S erfc
__device__ __forceinline__ void gelu_and_grad(float x, float& y, float& dy) {
    const float ax = fabsf(x);
    const float z = ax * 0.8493218002880191f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2727374808792225f, z, 1.0f)));
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    p *= t;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z));  // exp(-x^2/2)
    const float q = p * e;                                       // Phi(-|x|)
    const float phi_cdf = x >= 0.f ? 1.0f - q : q;
    y = x * phi_cdf;
    dy = fmaf(x * 0.3989422804014327f, e, phi_cdf);
}

__device__ __forceinline__ float torch_lerp(float a, float b, float w) {
    // torch.lerp: w < 0.5 ? a + w (b - a) : b - (b - a)(1 - w)
    const float diff = b - a;
    return (w < 0.5f) ? fmaf(w, diff, a) : (b - diff * (1.0f - w));
}

// ---------------------------------------------------------------------------------- philox
// Counter-based noise stream, restated on the CPU in oracle/philox.py (bit-exact integers).
constexpr uint32_t PHILOX_M0 = 0xD2511F53u, PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u, PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t PHILOX_STREAM = 0x5DE5A301u;

__device__ __forceinline__ uint4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3,
                                               uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(PHILOX_M0, c0), lo0 = PHILOX_M0 * c0;
        const uint32_t hi1 = __umulhi(PHILOX_M1, c2), lo1 = PHILOX_M1 * c2;
        c0 = hi1 ^ c1 ^ k0;
        c1 = lo1;
        c2 = hi0 ^ c3 ^ k1;
        c3 = lo0;
        k0 += PHILOX_W0;
        k1 += PHILOX_W1;
    }
    return make_uint4(c0, c1, c2, c3);
}

__device__ __forceinline__ float u01(uint32_t r) {
    // (0, 1]: r * 2^-32 + 2^-33, one FMA (the conversion rounds to nearest; product exact)
    return fmaf(__uint2float_rn(r), 2.3283064365386963e-10f, 1.1641532182693481e-10f);
}

__device__ __forceinline__ void box_muller(uint32_t ra, uint32_t rb, float& n0, float& n1) {
    const float ua = u01(ra), ub = u01(rb);
    // rad = sqrt(-2 ln ua).  log2 via MUFU; -2 ln2 folded into one multiply.
    // (ua is never denormal and the product never negative: the bare MUFU forms skip the range fix-ups
    // that __log2f / sqrtf carry, ~11 instructions per pair)
    float l2, rad;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(l2) : "f"(ua));
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rad) : "f"(-1.3862943611198906f * l2));
    const float theta = fmaf(6.283185307179586f, ub, -3.141592653589793f);
    float sn, cs;
    __sincosf(theta, &sn, &cs);
    n0 = rad * cs;
    n1 = rad * sn;
}

// four standard normals for (trajectory, step, dim chunk).  Deliberately NOT inlined: the rollout
// kernels call it once per 4 dims per step, and one shared copy keeps the hot loop inside the
// instruction cache (the 10 unrolled Philox rounds are ~100 SASS instructions).
static __device__ __noinline__ float4 normal4_call(uint32_t k0, uint32_t k1, uint32_t traj, uint32_t step, uint32_t chunk) {
    const uint4 r = philox4x32_10(traj, step, chunk, PHILOX_STREAM, k0, k1);
    float4 e;
    box_muller(r.x, r.y, e.x, e.y);
    box_muller(r.z, r.w, e.z, e.w);
    return e;
}

__device__ __forceinline__ void normal4(uint64_t seed, uint32_t traj, uint32_t step, uint32_t chunk,
                                        float& e0, float& e1, float& e2, float& e3) {
    const float4 e = normal4_call((uint32_t)seed, (uint32_t)(seed >> 32), traj, step, chunk);
    e0 = e.x; e1 = e.y; e2 = e.z; e3 = e.w;
}

}  // namespace sdes
