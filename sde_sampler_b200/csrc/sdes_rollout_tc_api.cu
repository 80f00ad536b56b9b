// sdes_rollout_tc_api.cu — host side of the tensor-core rollout engine that is independent of the padded state dimension:
// operand-image size, shared-memory budget, the dispatch onto the per-DPAD translation units of sdes_rollout_mma.cu,
// and the tcgen05 self test.
#include <cuda_bf16.h>

#include "sdes_common.cuh"
#include "sdes_tc.cuh"

namespace sdes {

constexpr int TC_GROUPS = 4;
constexpr int GMM_ACT = 8;

static inline int mma_nout(int dpad) { return (dpad + 15) / 16 * 16; }

// floats of the bf16 operand image: per layer hi[N x K16] then lo[N x K16] (tc::wimg16_offset layout), then the fp32
// biases {b_h[64]} x n_hidden, b_out[NOUT]   (b_in is folded into the time-embedding table)
int64_t mma4_weight_image_floats(const SdesRolloutDesc& d) {
    const int dpad = mma_pad_dim(d.dim), nout = mma_nout(dpad), k0b = (dpad + 15) & ~15;
    const int64_t bf16_elems = 2ll * 64 * k0b + (int64_t)d.n_hidden * 2 * 64 * 64 + 2ll * nout * 64;
    return bf16_elems / 2 + (int64_t)d.n_hidden * 64 + nout;
}

int mma_groups_per_sm() { return TC_GROUPS; }

size_t tc_smem_bytes_host(const KParams& p) {  // mirrors tc_smem_bytes in sdes_rollout_mma.cu
    const int dpad = p.ws.dpad, K = p.d.target_kind == SDES_TARGET_GMM ? p.d.n_components : 0;
    const size_t fl = (size_t)((p.ws.w_mma4_len + 31) & ~31ll) + 2 * (size_t)((K + 3) & ~3) * GMM_ACT + 64 + 2 * (size_t)dpad + 2 * (2 * dpad + 8) +
                      (size_t)TC_GROUPS * dpad * 128;
    return fl * sizeof(float);
}

bool mma_supported(const KParams& p) {
    return p.d.dim <= 64 && p.d.n_hidden <= SDES_MAX_HIDDEN && tc_smem_bytes_host(p) <= 226u * 1024u;
}

cudaError_t launch_rollout_tc_8(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);
cudaError_t launch_rollout_tc_16(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);
cudaError_t launch_rollout_tc_32(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);
cudaError_t launch_rollout_tc_48(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);
cudaError_t launch_rollout_tc_56(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);
cudaError_t launch_rollout_tc_64(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches);

cudaError_t launch_rollout_tc(const KParams& p, int sm_count, cudaStream_t stream, int* n_launches) {
    switch (p.ws.dpad) {
        case 8: return launch_rollout_tc_8(p, sm_count, stream, n_launches);
        case 16: return launch_rollout_tc_16(p, sm_count, stream, n_launches);
        case 32: return launch_rollout_tc_32(p, sm_count, stream, n_launches);
        case 48: return launch_rollout_tc_48(p, sm_count, stream, n_launches);
        case 56: return launch_rollout_tc_56(p, sm_count, stream, n_launches);
        case 64: return launch_rollout_tc_64(p, sm_count, stream, n_launches);
    }
    return cudaErrorInvalidValue;
}

// ------------------------------------------------------------------------------ self test
// D[128,N] = A[128,K] * W[N,K]^T through exactly the code path the rollout uses (A split into bf16 hi/lo and written
// with tcgen05.st into TMEM, W hi/lo images in shared memory, the bf16x3 issue, tcgen05.ld).  sdes_tcgen05_selftest.
__global__ void __launch_bounds__(128, 1) mma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                              float* __restrict__ D, int K, int N) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int K16 = (K + 15) & ~15;
    __nv_bfloat16* w_hi = reinterpret_cast<__nv_bfloat16*>(sm);
    __nv_bfloat16* w_lo = w_hi + N * K16;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < N * K16; e += 128) {
        const int n = e / K16, k = e % K16;
        const float w = k < K ? W[n * K + k] : 0.f;
        const __nv_bfloat16 hi = __float2bfloat16_rn(w);
        w_hi[tc::wimg16_offset(n, k, N)] = hi;
        w_lo[tc::wimg16_offset(n, k, N)] = __float2bfloat16_rn(w - __bfloat162float(hi));
    }
    if (warp == 0) {
        tc::tmem_alloc(&tmem_base_s, 128);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(&bar, 1);
        tc::fence_mbar_init();
    }
    tc::fence_proxy_async();  // generic-proxy smem writes (weights) -> visible to the tensor-core (async) proxy
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tbase = tmem_base_s;
    const uint32_t lane_addr = tbase + ((uint32_t)(warp * 32) << 16);
    const uint32_t col_d = 0, col_hi = 64, col_lo = 96;
    for (int c = 0; c < K16; c += 8) {
        uint32_t hi[4], lo[4];
        for (int q = 0; q < 4; ++q) {
            const int k = c + 2 * q;
            const float2 a = make_float2(k < K ? A[tid * K + k] : 0.f, k + 1 < K ? A[tid * K + k + 1] : 0.f);
            tc::split_bf16_pair2(a, hi[q], lo[q]);
        }
        tc::tmem_st4(lane_addr + col_hi + c / 2, hi);
        tc::tmem_st4(lane_addr + col_lo + c / 2, lo);
    }
    tc::wait_st();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after();
        tc::issue_layer_bf16x3(tbase + col_d, tbase + col_hi, tbase + col_lo, tc::smem_u32(w_hi), tc::smem_u32(w_lo), K16, N);
        tc::mma_commit(&bar);
    }
    tc::mbar_wait(&bar, 0);
    tc::fence_after();
    for (int c = 0; c < N; c += 8) {
        float v[8];
        tc::tmem_ld8(lane_addr + col_d + c, v);
        tc::wait_ld_tie<8>(v);
        for (int q = 0; q < 8; ++q) D[tid * N + c + q] = v[q];
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 128);
}

cudaError_t launch_mma_selftest(const float* A, const float* W, float* D, int K, int N, cudaStream_t stream) {
    const size_t smem = 2 * (size_t)N * ((K + 15) & ~15) * 2;
    cudaError_t e = cudaFuncSetAttribute(mma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mma_selftest_kernel<<<1, 128, smem, stream>>>(A, W, D, K, N);
    return cudaGetLastError();
}

}  // namespace sdes
