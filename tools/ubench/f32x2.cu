// Microbenchmark: issue rate of FFMA vs FFMA2 (fma.rn.f32x2), FMUL2, FADD2, MUFU on sm_100a.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o f32x2 f32x2.cu ; run: ./f32x2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, int iters, long long* cyc) {
    float a[8], b[8];
    for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3f + i; b[i] = 1.0f + i * 1e-4f; }
    const float c = 0.999f, dd = 1e-3f;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, dd);
#pragma unroll
            for (int i = 0; i < 8; ++i) b[i] = fmaf(b[i], c, dd);
        } else if (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(c, c), make_float2(dd, dd));
                a[i] = r.x; a[i + 1] = r.y;
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __ffma2_rn(make_float2(b[i], b[i + 1]), make_float2(c, c), make_float2(dd, dd));
                b[i] = r.x; b[i + 1] = r.y;
            }
        } else if (MODE == 2) {  // FMUL2 + FADD2
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __fmul2_rn(make_float2(a[i], a[i + 1]), make_float2(c, c));
                a[i] = r.x; a[i + 1] = r.y;
            }
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __fadd2_rn(make_float2(b[i], b[i + 1]), make_float2(dd, dd));
                b[i] = r.x; b[i + 1] = r.y;
            }
        } else if (MODE == 3) {  // FFMA with register operands (3-reg form)
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], b[i], b[(i + 1) & 7]);
        } else if (MODE == 4) {  // FFMA2 3-reg form
#pragma unroll
            for (int i = 0; i < 8; i += 2) {
                float2 r = __ffma2_rn(make_float2(a[i], a[i + 1]), make_float2(b[i], b[i + 1]), make_float2(b[(i + 2) & 7], b[(i + 3) & 7]));
                a[i] = r.x; a[i + 1] = r.y;
            }
        } else if (MODE == 5) {  // MUFU ex2
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
        } else if (MODE == 6) {  // FMNMX (alu) + FFMA (fma) mix
#pragma unroll
            for (int i = 0; i < 8; ++i) a[i] = fmaf(a[i], c, dd);
#pragma unroll
            for (int i = 0; i < 8; ++i) b[i] = fminf(b[i], a[i]);
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i] + b[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int ops_per_iter_per_thread, int threads) {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 20000;
    k<MODE><<<148, threads>>>(out, iters, cyc);
    k<MODE><<<148, threads>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // warp-instructions per SM per cycle and fp32 lane-ops per SM per cycle
    double winst = (double)ops_per_iter_per_thread * iters * (threads / 32) / (double)h;
    printf("%-28s threads=%4d cycles=%lld  warp-inst/clk/SM=%.3f\n", name, threads, h, winst);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int threads : {128, 256, 512, 1024}) {
        run<0>("FFMA imm (16/it)", 16, threads);
        run<1>("FFMA2 imm (8/it = 16 fma)", 8, threads);
        run<2>("FMUL2+FADD2 (8/it)", 8, threads);
        run<3>("FFMA 3-reg (8/it)", 8, threads);
        run<4>("FFMA2 3-reg (4/it = 8 fma)", 4, threads);
        run<5>("MUFU.EX2 (8/it)", 8, threads);
        run<6>("FFMA+FMNMX (16/it)", 16, threads);
    }
    return 0;
}
