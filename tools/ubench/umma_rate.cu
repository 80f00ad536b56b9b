// Microbenchmark: issue rate of tcgen05.mma.cta_group::1.kind::f16 (bf16, M=128, N=256, K=16) with both operands in
// shared memory in the K-major NO-SWIZZLE layout the GEMM layer uses, and the latency/throughput of the 16 KB
// cp.async.bulk copies that fill a stage.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../sde_sampler_b200/csrc
#include <cstdio>
#include <cuda_runtime.h>
#include "sdes_tc.cuh"
using namespace sdes;

__device__ __forceinline__ void mma_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

__global__ void __launch_bounds__(128, 1) mma_rate(int n_mma, int N, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tm;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x < 32) { tc::tmem_alloc(&tm, 256); tc::tmem_relinquish(); }
    if (threadIdx.x == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
    tc::fence_proxy_async();
    tc::fence_before(); __syncthreads(); tc::fence_after();
    if (threadIdx.x == 0) {
        const uint32_t idesc = tc::idesc_bf16(128, N);
        const uint32_t a0 = tc::smem_u32(smem), b0 = a0 + 32768, lbo = (uint32_t)N * 16u;
        long long t0 = clock64();
        for (int i = 0; i < n_mma; ++i) {
            const int ks = i & 3;
            const uint64_t da = tc::smem_desc_kmajor(a0 + ks * 4096u, 2048u, 128u);
            const uint64_t db = tc::smem_desc_kmajor(b0 + ks * 2u * lbo, lbo, 128u);
            mma_ss(tm, da, db, idesc, i > 0);
        }
        tc::mma_commit(&bar);
        tc::mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc::fence_before(); __syncthreads();
    if (threadIdx.x < 32) tc::tmem_dealloc(tm, 256);
}

// each CTA streams `n_stage` stages of `stage_kb` KB from global (L2-resident buffer) into a ring of `depth` stages
__global__ void __launch_bounds__(128, 1) bulk_rate(const uint8_t* src, size_t src_bytes, int n_stage, int stage_kb, int depth, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t full[8];
    if (threadIdx.x == 0) { for (int s = 0; s < 8; ++s) tc::mbar_init(&full[s], 1); tc::fence_mbar_init(); }
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t bytes = stage_kb * 1024u;
        long long t0 = clock64();
        for (int i = 0; i < n_stage + depth; ++i) {
            const int s = i % depth;
            if (i >= depth) tc::mbar_wait(&full[s], (uint32_t)(((i - depth) / depth) & 1));   // consume stage (i - depth)
            if (i < n_stage) {
                tc::mbar_arrive_expect_tx(&full[s], bytes);
                const size_t off = ((size_t)blockIdx.x * 7919u * bytes + (size_t)i * bytes) % (src_bytes - bytes);
                for (uint32_t o = 0; o < bytes; o += 16384u) tc::bulk_g2s(smem + (size_t)s * bytes + o, src + (off & ~(size_t)15) + o, 16384u, &full[s]);
            }
        }
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
}

int main() {
    long long* d_out; long long h;
    cudaMalloc(&d_out, 8);
    cudaFuncSetAttribute(mma_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    for (int N : {64, 128, 256}) {
        mma_rate<<<148, 128, 96 * 1024>>>(4096, N, d_out);
        mma_rate<<<148, 128, 96 * 1024>>>(4096, N, d_out);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("tcgen05.mma bf16 M=128 N=%3d K=16, SS no-swizzle, 148 CTAs: %.1f clk per MMA (%s)\n", N, (double)h / 4096, cudaGetErrorString(cudaGetLastError()));
    }
    uint8_t* src; size_t sb = 64u << 20;
    cudaMalloc(&src, sb); cudaMemset(src, 1, sb);
    cudaFuncSetAttribute(bulk_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, 192 * 1024);
    for (int ctas : {1, 148}) for (int stage_kb : {32, 64, 96}) for (int depth : {1, 2, 3, 6}) {
        if (stage_kb * depth > 192) continue;
        bulk_rate<<<ctas, 128, 192 * 1024>>>(src, sb, 64, stage_kb, depth, d_out);
        bulk_rate<<<ctas, 128, 192 * 1024>>>(src, sb, 64, stage_kb, depth, d_out);
        cudaDeviceSynchronize();
        cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
        printf("cp.async.bulk: %3d CTAs, stage %2d KB x depth %d: %.0f clk per stage, %.1f B/clk/SM (%s)\n", ctas, stage_kb, depth, (double)h / 64, 64.0 * stage_kb * 1024 / h, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
