// Microbenchmark: issue cost of GELU + bf16 hi/lo split epilogue variants on sm_100a, at the rollout kernel's
// occupancy (512 threads per SM, 4 warps per scheduler).  Each iteration processes 8 values per thread the way
// layer_epilogue4 does (values arrive in 8 registers, bias added, GELU, split into packed bf16 hi / lo words).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gelu_variants gelu_variants.cu ; run: ./gelu_variants
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t pack_bf16x2(float lo_k, float lo_k1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(lo_k1), "f"(lo_k));
    return r;
}
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_bf16x2(a - ha, b - hb);
}
__device__ __forceinline__ void split_bf16_pair2(float2 v, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(v.x, v.y);
    const float2 h = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
    const float2 l = __ffma2_rn(h, make_float2(-1.f, -1.f), v);
    lo = pack_bf16x2(l.x, l.y);
}
// truncation split: hi = top 16 bits (exact bf16 by truncation), lo = bf16(v - hi)
__device__ __forceinline__ void split_trunc_pair2(float2 v, uint32_t& hi, uint32_t& lo) {
    const uint32_t a = __float_as_uint(v.x), b = __float_as_uint(v.y);
    hi = __byte_perm(a, b, 0x7632);
    const float2 h = make_float2(__uint_as_float(a & 0xFFFF0000u), __uint_as_float(b & 0xFFFF0000u));
    const float2 l = __ffma2_rn(h, make_float2(-1.f, -1.f), v);
    lo = pack_bf16x2(l.x, l.y);
}

__device__ __forceinline__ float gelu_as(float x) {  // current product form (A&S 7.1.26)
    const float ax = fabsf(x);
    const float z = ax * 0.8493218002880191f;
    float t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.2727374808792225f, z, 1.0f)));
    float p = fmaf(0.5f * 1.061405429f, t, 0.5f * -1.453152027f);
    p = fmaf(p, t, 0.5f * 1.421413741f);
    p = fmaf(p, t, 0.5f * -0.284496736f);
    p = fmaf(p, t, 0.5f * 0.254829592f);
    p *= t * ax;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-z * z));
    return fmaf(-p, e, fmaxf(x, 0.0f));
}
// logistic form: GELU(x) = x / (1 + 2^(-x P(x^2))), P degree 6 (minimax fit of log2(Phi(x)/Phi(-x))/x)
#define LC0 2.30220913e+00f
#define LC1 1.04834383e-01f
#define LC2 -9.27478302e-05f
#define LC3 -1.60239457e-04f
#define LC4 1.15760618e-05f
#define LC5 -3.93527977e-07f
#define LC6 5.42691260e-09f
__device__ __forceinline__ float gelu_lg(float x) {
    const float u = x * x;
    float p = fmaf(LC6, u, LC5);
    p = fmaf(p, u, LC4);
    p = fmaf(p, u, LC3);
    p = fmaf(p, u, LC2);
    p = fmaf(p, u, LC1);
    p = fmaf(p, u, LC0);
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(-x * p));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}
__device__ __forceinline__ float2 gelu_lg2(float2 x) {
    const float2 u = __fmul2_rn(x, x);
    float2 p = __ffma2_rn(make_float2(LC6, LC6), u, make_float2(LC5, LC5));
    p = __ffma2_rn(p, u, make_float2(LC4, LC4));
    p = __ffma2_rn(p, u, make_float2(LC3, LC3));
    p = __ffma2_rn(p, u, make_float2(LC2, LC2));
    p = __ffma2_rn(p, u, make_float2(LC1, LC1));
    p = __ffma2_rn(p, u, make_float2(-LC0, -LC0));  // negated polynomial: q = -x P
    const float2 q = __fmul2_rn(x, p);
    float2 e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(q.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(q.y));
    const float2 d = __fadd2_rn(e, make_float2(1.f, 1.f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.x) : "f"(d.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r.y) : "f"(d.y));
    return __fmul2_rn(x, r);
}
__device__ __forceinline__ float2 gelu_as2(float2 x) {  // current formula, packed where possible
    const float2 ax = make_float2(fabsf(x.x), fabsf(x.y));
    const float2 z = __fmul2_rn(ax, make_float2(0.8493218002880191f, 0.8493218002880191f));
    const float2 a = __ffma2_rn(z, make_float2(0.2727374808792225f, 0.2727374808792225f), make_float2(1.f, 1.f));
    float2 t, e;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.x) : "f"(a.x));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t.y) : "f"(a.y));
    float2 p = __ffma2_rn(make_float2(0.5f * 1.061405429f, 0.5f * 1.061405429f), t, make_float2(0.5f * -1.453152027f, 0.5f * -1.453152027f));
    p = __ffma2_rn(p, t, make_float2(0.5f * 1.421413741f, 0.5f * 1.421413741f));
    p = __ffma2_rn(p, t, make_float2(0.5f * -0.284496736f, 0.5f * -0.284496736f));
    p = __ffma2_rn(p, t, make_float2(0.5f * 0.254829592f, 0.5f * 0.254829592f));
    p = __fmul2_rn(p, __fmul2_rn(t, ax));
    const float2 mz = __fmul2_rn(z, make_float2(-z.x, -z.y));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.x) : "f"(mz.x));
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e.y) : "f"(mz.y));
    const float2 rl = make_float2(fmaxf(x.x, 0.f), fmaxf(x.y, 0.f));
    return __ffma2_rn(make_float2(-p.x, -p.y), e, rl);
}

template <int MODE>
__global__ void __launch_bounds__(512, 1) k(const float* __restrict__ in, uint32_t* __restrict__ out, int iters, long long* cyc) {
    float v[8];
    const float4* in4 = reinterpret_cast<const float4*>(in) + threadIdx.x * 2;
    float4 b0 = in4[0], b1 = in4[1];
    for (int i = 0; i < 8; ++i) v[i] = (threadIdx.x * 8 + i) * 1e-3f - 2.0f;
    uint32_t acc = 0;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
        uint32_t hi[4], lo[4];
        if (MODE == 0) {
            float a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = gelu_as(v[q] + bb[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) split_bf16_pair(a[2 * q], a[2 * q + 1], hi[q], lo[q]);
        } else if (MODE == 1) {
            float a[8];
#pragma unroll
            for (int q = 0; q < 8; ++q) a[q] = gelu_lg(v[q] + bb[q]);
#pragma unroll
            for (int q = 0; q < 4; ++q) split_bf16_pair(a[2 * q], a[2 * q + 1], hi[q], lo[q]);
        } else if (MODE == 2) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 a = gelu_lg2(__fadd2_rn(make_float2(v[2 * q], v[2 * q + 1]), make_float2(bb[2 * q], bb[2 * q + 1])));
                split_bf16_pair2(a, hi[q], lo[q]);
            }
        } else if (MODE == 3) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 a = gelu_as2(__fadd2_rn(make_float2(v[2 * q], v[2 * q + 1]), make_float2(bb[2 * q], bb[2 * q + 1])));
                split_bf16_pair2(a, hi[q], lo[q]);
            }
        } else if (MODE == 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float2 a = gelu_lg2(__fadd2_rn(make_float2(v[2 * q], v[2 * q + 1]), make_float2(bb[2 * q], bb[2 * q + 1])));
                split_trunc_pair2(a, hi[q], lo[q]);
            }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) acc ^= hi[q] + lo[q];
        // next inputs depend weakly on the iteration so nothing is hoisted
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = __uint_as_float((__float_as_uint(v[q]) ^ (acc & 0x7u)));
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name) {
    float* in; uint32_t* out; long long* cyc; long long h;
    cudaMalloc(&in, 512 * 8 * 4); cudaMemset(in, 0, 512 * 8 * 4);
    cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<MODE><<<148, 512>>>(in, out, iters, cyc);
    k<MODE><<<148, 512>>>(in, out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-44s cycles/8-elem-iter/warp-slot=%.1f  elems/clk/SM=%.2f\n", name, (double)h / iters, 8.0 * 512 * iters / (double)h);
    cudaFree(in); cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0>("A&S scalar + rn split (product r1)");
    run<1>("logistic deg6 scalar + rn split");
    run<2>("logistic deg6 f32x2 + rn split (ffma2)");
    run<3>("A&S f32x2 + rn split (ffma2)");
    run<4>("logistic deg6 f32x2 + trunc split");
    return 0;
}
