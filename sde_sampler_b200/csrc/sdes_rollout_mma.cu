// sdes_rollout_mma.cu — persistent rollout kernel, control MLP on tcgen05 tensor cores.
//
// One CTA per SM, resident for the whole rollout.  It holds the split (hi/lo) weight images of
// every layer in shared memory (loaded once with TMA bulk copies), owns all 512 TMEM columns,
// and runs GROUPS independent groups of 4 warps.  A group pulls 128-trajectory tiles from a
// global counter and carries each tile through all T time steps: thread r owns trajectory r —
// its state x, running cost and network activations stay in registers / TMEM lane r.  Per step
// and layer the group writes the activations (hi, lo) into TMEM as the A operand, one thread
// issues the 3xTF32 MMAs against the shared-memory weights and commits to the group's mbarrier,
// and the threads read the fp32 accumulator row back with tcgen05.ld for the fused epilogue
// (bias, exact-erf GELU, then — after the last layer — target score, control reparametrisation,
// cost increments, Philox noise, Euler-Maruyama update: sdes_step.cuh).  While one group waits
// on its MMAs the other group's epilogue keeps the FP32/MUFU pipes busy.
#include <cuda_bf16.h>

#include "sdes_step.cuh"
#include "sdes_tc.cuh"

namespace sdes {

#ifndef SDES_MMA_GROUPS
#define SDES_MMA_GROUPS 3
#endif
constexpr int MMA_GROUPS = SDES_MMA_GROUPS;
constexpr int MMA_THREADS = MMA_GROUPS * 128;
constexpr int TMEM_COLS = 512;
constexpr int GROUP_COLS = 160;  // D[64] | A_hi tf32 [64] | A_lo bf16x2 [32]

__host__ __device__ inline int mma_nout(int dpad) { return (dpad + 15) / 16 * 16; }

// Workspace image for the tcgen05 engine (floats), in the order the kernel keeps it in smem.  Per layer:
// hi (tf32-truncated fp32), lo = w - hi (fp32), w16 (bf16, K padded to a multiple of 16):
//   L0:  hi[64*K0] lo[64*K0] w16[64*K0b/2]     (N=64, K=K0=dpad, K0b = round16(dpad))
//   Lh:  { hi[64*64] lo[64*64] w16[64*64/2] } x n_hidden
//   Lo:  hi[NOUT*64] lo[NOUT*64] w16[NOUT*64/2] (N=NOUT, K=64)
//   bias: { b_h[64] } x n_hidden, b_out[NOUT]     (b_in is folded into the time-embedding table)
int64_t mma_weight_image_floats(const SdesRolloutDesc& d, int dpad_simt) {
    (void)dpad_simt;
    const int dpad = mma_pad_dim(d.dim), nout = mma_nout(dpad), k0b = (dpad + 15) & ~15;
    return (2ll * 64 * dpad + 32ll * k0b) + (int64_t)d.n_hidden * (2ll * 64 * 64 + 32 * 64) + (2ll * nout * 64 + nout * 32ll) +
           (int64_t)d.n_hidden * 64 + nout;
}

int mma_groups_per_sm(int variant) { return variant == 1 ? 4 : MMA_GROUPS; }

bool mma_supported(const KParams& p) { return p.d.dim <= 64 && p.d.n_hidden <= SDES_MAX_HIDDEN; }

// ------------------------------------------------------------------------------ self test
// D[128,N] = A[128,K] * W[N,K]^T through exactly the code path the rollout uses (A via tcgen05.st
// into TMEM, W image in smem, 3xTF32 issue, tcgen05.ld).  Exposed as sdes_tcgen05_selftest.
__global__ void __launch_bounds__(128, 1) mma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ W,
                                                              float* __restrict__ D, int K, int N, int mode) {
    extern __shared__ __align__(128) float sm[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int K16 = (K + 15) & ~15;
    float* w_hi = sm;
    float* w_lo = sm + N * K;
    __nv_bfloat16* w16 = reinterpret_cast<__nv_bfloat16*>(sm + 2 * N * K);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int e = tid; e < N * K; e += 128) {
        const int n = e / K, k = e % K;
        const float w = W[e];
        const float hi = __uint_as_float(tc::tf32_hi_bits(w));
        w_hi[tc::wimg_offset_floats(n, k, N)] = hi;
        w_lo[tc::wimg_offset_floats(n, k, N)] = w - hi;
    }
    for (int e = tid; e < N * K16; e += 128) {
        const int n = e / K16, k = e % K16;
        w16[tc::wimg16_offset(n, k, N)] = __float2bfloat16_rn(k < K ? W[n * K + k] : 0.f);
    }
    if (warp == 0) {
        tc::tmem_alloc(&tmem_base_s, 256);
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
    const uint32_t col_d = 0, col_hi = 64, col_lo = 128;
    for (int c = 0; c < K16; c += 8) {
        uint32_t hi[8], lo[8], lo16[4];
        float lof[8];
        for (int q = 0; q < 8; ++q) {
            const float a = c + q < K ? A[tid * K + c + q] : 0.f;
            hi[q] = tc::tf32_hi_bits(a);
            lof[q] = a - __uint_as_float(hi[q]);
            lo[q] = __float_as_uint(lof[q]);
        }
        for (int q = 0; q < 4; ++q) lo16[q] = tc::pack_bf16x2(lof[2 * q], lof[2 * q + 1]);
        if (c < K) tc::tmem_st8(lane_addr + col_hi + c, hi);
        if (mode == 0) {
            if (c < K) tc::tmem_st8(lane_addr + col_lo + c, lo);
        } else {
            tc::tmem_st4(lane_addr + col_lo + c / 2, lo16);
        }
    }
    tc::wait_st();
    tc::fence_before();
    __syncthreads();
    if (tid == 0) {
        tc::fence_after();
        if (mode == 0)
            tc::issue_layer_3xtf32(tbase + col_d, tbase + col_hi, tbase + col_lo, tc::smem_u32(w_hi), tc::smem_u32(w_lo), K, N);
        else
            tc::issue_layer_mixed(tbase + col_d, tbase + col_hi, tbase + col_lo, tc::smem_u32(w_hi), tc::smem_u32(w_lo),
                                  tc::smem_u32(w16), K, K16, N);
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
    if (warp == 0) tc::tmem_dealloc(tbase, 256);
}

cudaError_t launch_mma_selftest(const float* A, const float* W, float* D, int K, int N, int mode, cudaStream_t stream) {
    const size_t smem = 2 * (size_t)N * K * sizeof(float) + (size_t)N * ((K + 15) & ~15) * 2;
    cudaError_t e = cudaFuncSetAttribute(mma_selftest_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    mma_selftest_kernel<<<1, 128, smem, stream>>>(A, W, D, K, N, mode);
    return cudaGetLastError();
}


// ---------------------------------------------------------------------------- the kernel
__device__ __forceinline__ void group_bar(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

// activations (fp32) -> TMEM A operand: tf32 hi half (one column per value) and bf16 lo half (two per column).
// ZERO_PAD: also clear the bf16 columns up to the next multiple of 16 values (input layer, K0 % 16 == 8).
template <int N>
__device__ __forceinline__ void store_a_split(uint32_t addr_hi, uint32_t addr_lo16, const float (&a)[N]) {
#pragma unroll
    for (int c = 0; c < N; c += 8) {
        uint32_t hi[8], lo16[4];
        float lo[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            hi[q] = tc::tf32_hi_bits(a[c + q]);
            lo[q] = a[c + q] - __uint_as_float(hi[q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) lo16[q] = tc::pack_bf16x2(lo[2 * q], lo[2 * q + 1]);
        tc::tmem_st8(addr_hi + c, hi);
        tc::tmem_st4(addr_lo16 + c / 2, lo16);
    }
    if (N % 16 == 8) {
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
        tc::tmem_st4(addr_lo16 + N / 2, z);
    }
}

template <int N>
__device__ __forceinline__ void load_acc(uint32_t addr_d, float (&acc)[N]) {
#pragma unroll
    for (int c = 0; c < N; c += 8) tc::tmem_ld8(addr_d + c, &acc[c]);
    tc::wait_ld_tie<N>(acc);
}

// 8 accumulator columns -> + bias -> exact GELU -> tf32 hi/lo split -> A operand of the next layer
__device__ __forceinline__ void gelu_split_store8(uint32_t addr_hi, uint32_t addr_lo, const float (&v)[8],
                                                  const float4 b0, const float4 b1) {
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    uint32_t hi[8], lo16[4];
    float lo[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float a = gelu_fast(v[q] + bb[q]);
        hi[q] = tc::tf32_hi_bits(a);
        lo[q] = a - __uint_as_float(hi[q]);
    }
#pragma unroll
    for (int q = 0; q < 4; ++q) lo16[q] = tc::pack_bf16x2(lo[2 * q], lo[2 * q + 1]);
    tc::tmem_st8(addr_hi, hi);
    tc::tmem_st4(addr_lo, lo16);
}

struct GroupCtx;
__device__ __forceinline__ void layer_epilogue(const GroupCtx& c, const float* __restrict__ bias);

struct GroupCtx {
    int g;               // group index
    int gtid;            // thread index in the group = row in the tile = TMEM lane
    uint32_t t_d, t_hi, t_lo;          // TMEM addresses, lane 0 of the tile (for the issuing thread)
    uint32_t l_d, l_hi, l_lo;          // same, at this thread's warp lane window (for ld / st)
    uint64_t* bar;
    uint32_t phase;
};

// A operand is in TMEM; run one layer and leave the accumulator ready to be read.
__device__ __forceinline__ void run_layer(GroupCtx& c, uint32_t w_hi_saddr, uint32_t w_lo_saddr, uint32_t w16_saddr, int K, int N) {
    tc::wait_st();
    tc::fence_before();
    group_bar(c.g);
    if (c.gtid == 0) {
        tc::fence_after();
        tc::issue_layer_mixed(c.t_d, c.t_hi, c.t_lo, w_hi_saddr, w_lo_saddr, w16_saddr, K, (K + 15) & ~15, N);
        tc::mma_commit(c.bar);
    }
    tc::mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    tc::fence_after();
}

// The fused epilogue between two layers, streamed 8 columns at a time straight from the
// accumulator (TMEM) into the next A operand (TMEM): the 64-wide activation row never sits in
// registers, and one compact loop serves every layer (instruction-cache footprint matters: the
// first version of this kernel spent 42% of its stall samples on instruction fetch).
// `bias` may point to global (time-embedding row, input layer) or shared memory (hidden biases).
__device__ __forceinline__ void layer_epilogue(const GroupCtx& c, const float* __restrict__ bias) {
    float a[8], b[8];
    tc::tmem_ld8(c.l_d, a);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll 1
    for (int ch = 0; ch < 8; ch += 2) {
        // bias loads are issued before the wait so their latency hides behind the TMEM load
        const float4 p0 = b4[2 * ch], p1 = b4[2 * ch + 1], p2 = b4[2 * ch + 2], p3 = b4[2 * ch + 3];
        tc::wait_ld_tie<8>(a);
        tc::tmem_ld8(c.l_d + 8u * (ch + 1), b);
        gelu_split_store8(c.l_hi + 8u * ch, c.l_lo + 4u * ch, a, p0, p1);
        tc::wait_ld_tie<8>(b);
        if (ch + 2 < 8) tc::tmem_ld8(c.l_d + 8u * (ch + 2), a);
        gelu_split_store8(c.l_hi + 8u * (ch + 1), c.l_lo + 4u * (ch + 1), b, p2, p3);
    }
}

template <int DPAD>
__global__ void __launch_bounds__(MMA_THREADS, 1) rollout_mma_kernel(const __grid_constant__ KParams p) {
    constexpr int NOUT = (DPAD + 15) / 16 * 16;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t s_wbar;
    __shared__ uint64_t s_mbar[MMA_GROUPS];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_tile[MMA_GROUPS];

    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, T = d.n_steps, K = d.n_components, nh = d.n_hidden;
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- shared memory carve-up: [weight image + biases | gmm mu | gmm h | gmm c | prior | ref]
    float* s_w = smem;
    const int K2 = (K + 1) & ~1;  // GMM images are padded to an even number of components
    float* s_mu = s_w + p.ws.w_mma_len;
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    float* s_prior = s_c + 64;
    float* s_ref = s_prior + 2 * DPAD + 8;

    if (warp == 0) {
        tc::tmem_alloc(&s_tmem, TMEM_COLS);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(&s_wbar, 1);
        for (int g = 0; g < MMA_GROUPS; ++g) tc::mbar_init(&s_mbar[g], 1);
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (tid == 0) {
        // weights: TMA bulk copies global -> shared, all counted on one mbarrier
        const uint32_t total = (uint32_t)(p.ws.w_mma_len * sizeof(float));
        tc::mbar_arrive_expect_tx(&s_wbar, total);
        const char* src = reinterpret_cast<const char*>(ws + p.ws.w_mma);
        char* dst = reinterpret_cast<char*>(s_w);
        for (uint32_t off = 0; off < total; off += 16384u) {
            const uint32_t n = total - off < 16384u ? total - off : 16384u;
            tc::bulk_g2s(dst + off, src + off, n, &s_wbar);
        }
    }
    for (int e = tid; e < K2 * DPAD; e += blockDim.x) {
        s_mu[e] = ws[p.ws.gmm_mu + e];
        s_h[e] = ws[p.ws.gmm_h + e];
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c[e] = ws[p.ws.gmm_c + e];
    for (int e = tid; e < 2 * DPAD + 8; e += blockDim.x) {
        s_prior[e] = e <= 2 * DPAD ? ws[p.ws.prior + e] : 0.f;
        s_ref[e] = e <= 2 * DPAD ? ws[p.ws.ref + e] : 0.f;
    }
    tc::mbar_wait(&s_wbar, 0);
    __syncthreads();

    // weight image addresses
    const uint32_t w_base = tc::smem_u32(s_w);
    constexpr uint32_t K0B = (DPAD + 15) & ~15;
    constexpr uint32_t LH_BYTES = 16384u + 16384u + 8192u;             // one hidden layer: hi | lo | w16
    const uint32_t l0_hi = w_base, l0_lo = l0_hi + 64u * DPAD * 4u, l0_16 = l0_lo + 64u * DPAD * 4u;
    const uint32_t lh_base = l0_16 + 64u * K0B * 2u;
    const uint32_t lo_hi = lh_base + (uint32_t)nh * LH_BYTES, lo_lo = lo_hi + (uint32_t)NOUT * 256u, lo_16 = lo_lo + (uint32_t)NOUT * 256u;
    const float* s_bias = s_w + (2 * 64 * DPAD + 32 * K0B) + nh * (2 * 64 * 64 + 32 * 64) + (2 * NOUT * 64 + NOUT * 32);  // {b_h[64]} x nh, b_out[NOUT]

    TargetSmem tsm{s_mu, s_h, s_c, reinterpret_cast<const uint32_t*>(ws + p.ws.counter)[1], s_prior, s_ref};
    GroupCtx c;
    c.g = warp >> 2;
    c.gtid = tid & 127;
    const uint32_t tbase = s_tmem + (uint32_t)(c.g * GROUP_COLS);
    c.t_d = tbase;
    c.t_hi = tbase + 64;
    c.t_lo = tbase + 128;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    c.l_d = c.t_d + lane_off;
    c.l_hi = c.t_hi + lane_off;
    c.l_lo = c.t_lo + lane_off;
    c.bar = &s_mbar[c.g];
    c.phase = 0;

    uint32_t* counter = reinterpret_cast<uint32_t*>(const_cast<float*>(ws) + p.ws.counter);
    const bool from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    const bool ret_traj = (d.flags & SDES_F_RETURN_TRAJ) != 0;
    const int64_t B = d.batch;
    const uint32_t n_tiles = (uint32_t)((B + 127) / 128);

    // Work items are (time chunk, tile) pairs handed out chunk-major from one counter: a tile's T steps
    // are cut into n_chunks pieces so that 512 tiles on 296 groups do not quantise into 2 rounds with
    // the second 73% full.  Between chunks the tile's state (x, rnd) parks in the workspace
    // ([tile][j][128] so warps read and write whole 128-byte lines) and a per-tile progress word orders
    // producer and consumer.  An item only ever waits on an item handed out earlier, so there is no
    // deadlock whatever the residency.
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

        float x[DPAD];
        float rnd;
        if (chunk == 0) {
#pragma unroll
            for (int j = 0; j < DPAD; ++j) x[j] = (j < dim) ? __ldg(d.x0 + rrow * dim + j) : 0.f;
            if (ret_traj && valid) {
                const TrajRef o = traj_ref(d, d.xs, 0, rrow);
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) o.p[j * o.stride] = x[j];
            }
            rnd = initial_rnd<DPAD>(d, x, tsm);
        } else {
            if (c.gtid == 0) {
                uint32_t seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + tile) : "memory");
                } while (seen < chunk);
            }
            group_bar(c.g);
#pragma unroll
            for (int j = 0; j < DPAD; ++j) x[j] = __ldcg(st + j * 128);  // L2 reads: another SM wrote them
            rnd = __ldcg(st + DPAD * 128);
        }
        const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)rrow);
        const int i_begin = (int)chunk * chunk_steps;
        const int i_end = (i_begin + chunk_steps < T) ? i_begin + chunk_steps : T;

        for (int i = i_begin; i < i_end; ++i) {
            const float* tab = ws + p.ws.tab + (int64_t)i * TAB_STRIDE;
            // ---- score part of the control (needs x only): before the MLP, kept in registers
            float sc[DPAD];
            score_part<DPAD>(d, x, sc, tsm, ws + p.ws.gate + (int64_t)i * DPAD, tab[TAB_SIGMA], tab[TAB_LERP_W]);
            // ---- control MLP on the tensor cores (models/mlp.py:114-122)
            store_a_split<DPAD>(c.l_hi, c.l_lo, x);
            run_layer(c, l0_hi, l0_lo, l0_16, DPAD, C);
            layer_epilogue(c, ws + p.ws.emb + (int64_t)i * C);  // + (emb_t + b_in), GELU
#pragma unroll 1
            for (int l = 0; l < nh; ++l) {
                run_layer(c, lh_base + (uint32_t)l * LH_BYTES, lh_base + (uint32_t)l * LH_BYTES + 16384u, lh_base + (uint32_t)l * LH_BYTES + 32768u, C, C);
                layer_epilogue(c, s_bias + l * C);
            }
            run_layer(c, lo_hi, lo_lo, lo_16, C, NOUT);
            // ---- network output streamed from TMEM into the control / cost / state update
            {
                const StepCoef sc_ = make_step_coef(d, tab);
                const float* nrow = from_hbm ? d.noise + ((int64_t)i * B + rrow) * dim : nullptr;
                const float* bo = s_bias + nh * C;
                float cost = 0.f, ito = 0.f;
                float na[8], nb[8];
                tc::tmem_ld8(c.l_d, na);
#pragma unroll
                for (int q = 0; q < DPAD / 8; q += 2) {
                    tc::wait_ld_tie<8>(na);
                    if (q + 1 < DPAD / 8) tc::tmem_ld8(c.l_d + 8u * (q + 1), nb);
#pragma unroll
                    for (int r = 0; r < 8; ++r) na[r] += bo[8 * q + r];
                    update4(sc_, &x[8 * q], &na[0], &sc[8 * q], s_prior + 8 * q, s_prior + DPAD + 8 * q, 8 * q, i, traj, nrow, cost, ito);
                    update4(sc_, &x[8 * q + 4], &na[4], &sc[8 * q + 4], s_prior + 8 * q + 4, s_prior + DPAD + 8 * q + 4, 8 * q + 4, i, traj, nrow, cost, ito);
                    if (q + 1 < DPAD / 8) {
                        tc::wait_ld_tie<8>(nb);
                        if (q + 2 < DPAD / 8) tc::tmem_ld8(c.l_d + 8u * (q + 2), na);
#pragma unroll
                        for (int r = 0; r < 8; ++r) nb[r] += bo[8 * (q + 1) + r];
                        update4(sc_, &x[8 * q + 8], &nb[0], &sc[8 * q + 8], s_prior + 8 * q + 8, s_prior + DPAD + 8 * q + 8, 8 * q + 8, i, traj, nrow, cost, ito);
                        update4(sc_, &x[8 * q + 12], &nb[4], &sc[8 * q + 12], s_prior + 8 * q + 12, s_prior + DPAD + 8 * q + 12, 8 * q + 12, i, traj, nrow, cost, ito);
                    }
                }
                finish_step(d, sc_, tab, cost, ito, rnd);
            }
            if (ret_traj && valid) {
                const TrajRef o = traj_ref(d, d.xs, i + 1, rrow);
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) o.p[j * o.stride] = x[j];
            }
        }
        if (i_end == T) {
            rnd += terminal_rnd<DPAD>(d, x, tsm);
            if (valid) {
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) d.x_T[rrow * dim + j] = x[j];
                d.rnd[rrow] = rnd;
            }
        } else {
#pragma unroll
            for (int j = 0; j < DPAD; ++j) __stcg(st + j * 128, x[j]);
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

// =========================================================================== 4-group engine
// Same work decomposition (items = (time chunk, 128-trajectory tile), one trajectory per thread, MLP on tcgen05
// with the A operand in TMEM), but FOUR groups per SM instead of three: one more warp per scheduler to hide the
// epilogue's dependency stalls, which is what bounds this kernel (issue slots ~58 % busy with three).  What makes
// the fourth group fit:
//   * TMEM: both operand halves are bf16 (hi = bf16(v), lo = bf16(v - hi), 16 significant bits — the wide engine's
//     split, parity-tested there): D 64 | A_hi 32 | A_lo 32 = 128 columns per group, 4 x 128 = 512;
//   * registers (128 per thread at 512 threads): the state x lives in shared memory ([j][128] per group, conflict
//     free) and is read where it is needed; only the score part sc[DPAD] stays in registers across the MLP;
//   * shared memory: bf16 hi/lo weights are 4 bytes per element instead of 10 (65 KB instead of 160 KB at d = 50),
//     which pays for the 112 KB of state.
// The layer is three kind::f16 MMAs per K=16 step (tc::issue_layer_bf16x3): 12 per 64x64 layer instead of 20.
constexpr int MMA4_GROUPS = 4;
constexpr int MMA4_THREADS = MMA4_GROUPS * 128;
constexpr int GROUP4_COLS = 128;  // D[64] | A_hi bf16x2 [32] | A_lo bf16x2 [32]

// bytes of the bf16 operand image: per layer hi then lo (wimg16 layout), then the fp32 biases
int64_t mma4_weight_image_floats(const SdesRolloutDesc& d) {
    const int dpad = mma_pad_dim(d.dim), nout = mma_nout(dpad), k0b = (dpad + 15) & ~15;
    const int64_t bf16_elems = 2ll * 64 * k0b + (int64_t)d.n_hidden * 2 * 64 * 64 + 2ll * nout * 64;
    return bf16_elems / 2 + (int64_t)d.n_hidden * 64 + nout;
}

struct XSmem {  // the state of one trajectory in shared memory: element j at p[j * 128]
    float* p;
    __device__ __forceinline__ float operator[](int j) const { return p[j * 128]; }
};

template <int DPAD>
__device__ __forceinline__ void store_a_split16(uint32_t addr_hi, uint32_t addr_lo, const XSmem& x) {
#pragma unroll
    for (int c = 0; c < DPAD; c += 8) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) tc::split_bf16_pair(x[c + 2 * q], x[c + 2 * q + 1], hi[q], lo[q]);
        tc::tmem_st4(addr_hi + c / 2, hi);
        tc::tmem_st4(addr_lo + c / 2, lo);
    }
    if (DPAD % 16 == 8) {
        const uint32_t z[4] = {0u, 0u, 0u, 0u};
        tc::tmem_st4(addr_hi + DPAD / 2, z);
        tc::tmem_st4(addr_lo + DPAD / 2, z);
    }
}

__device__ __forceinline__ void gelu_split16_store8(uint32_t addr_hi, uint32_t addr_lo, const float (&v)[8], const float4 b0, const float4 b1) {
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
    float a[8];
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 8; ++q) a[q] = gelu_fast(v[q] + bb[q]);
#pragma unroll
    for (int q = 0; q < 4; ++q) tc::split_bf16_pair(a[2 * q], a[2 * q + 1], hi[q], lo[q]);
    tc::tmem_st4(addr_hi, hi);
    tc::tmem_st4(addr_lo, lo);
}

__device__ __forceinline__ void group_bar4(int g) { asm volatile("bar.sync %0, 128;" ::"r"(g + 1) : "memory"); }

__device__ __forceinline__ void run_layer4(GroupCtx& c, uint32_t w_hi_saddr, uint32_t w_lo_saddr, int K16, int N) {
    tc::wait_st();
    tc::fence_before();
    group_bar4(c.g);
    if (c.gtid == 0) {
        tc::fence_after();
        tc::issue_layer_bf16x3(c.t_d, c.t_hi, c.t_lo, w_hi_saddr, w_lo_saddr, K16, N);
        tc::mma_commit(c.bar);
    }
    tc::mbar_wait(c.bar, c.phase);
    c.phase ^= 1u;
    tc::fence_after();
}

__device__ __forceinline__ void layer_epilogue4(const GroupCtx& c, const float* __restrict__ bias) {
    float a[8], b[8];
    tc::tmem_ld8(c.l_d, a);
    const float4* b4 = reinterpret_cast<const float4*>(bias);
#pragma unroll 1
    for (int ch = 0; ch < 8; ch += 2) {
        const float4 p0 = b4[2 * ch], p1 = b4[2 * ch + 1], p2 = b4[2 * ch + 2], p3 = b4[2 * ch + 3];
        tc::wait_ld_tie<8>(a);
        tc::tmem_ld8(c.l_d + 8u * (ch + 1), b);
        gelu_split16_store8(c.l_hi + 4u * ch, c.l_lo + 4u * ch, a, p0, p1);
        tc::wait_ld_tie<8>(b);
        if (ch + 2 < 8) tc::tmem_ld8(c.l_d + 8u * (ch + 2), a);
        gelu_split16_store8(c.l_hi + 4u * (ch + 1), c.l_lo + 4u * (ch + 1), b, p2, p3);
    }
}

template <int DPAD>
__global__ void __launch_bounds__(MMA4_THREADS, 1) rollout_mma4_kernel(const __grid_constant__ KParams p) {
    constexpr int NOUT = (DPAD + 15) / 16 * 16;
    constexpr uint32_t K0B = (DPAD + 15) & ~15;
    extern __shared__ __align__(128) float smem[];
    __shared__ uint64_t s_wbar;
    __shared__ uint64_t s_mbar[MMA4_GROUPS];
    __shared__ uint32_t s_tmem;
    __shared__ uint32_t s_tile[MMA4_GROUPS];

    const SdesRolloutDesc& d = p.d;
    const float* ws = reinterpret_cast<const float*>(d.workspace);
    const int dim = d.dim, T = d.n_steps, K = d.n_components, nh = d.n_hidden;
    const int tid = threadIdx.x, warp = tid >> 5;

    // ---- shared memory: [bf16 weight image + biases | gmm mu | gmm h | gmm c | prior | ref | state x]
    float* s_w = smem;
    const int K2 = (K + 1) & ~1;
    float* s_mu = s_w + ((p.ws.w_mma4_len + 31) & ~31ll);
    float* s_h = s_mu + K2 * DPAD;
    float* s_c = s_h + K2 * DPAD;
    float* s_prior = s_c + 64;
    float* s_ref = s_prior + 2 * DPAD + 8;
    float* s_x = s_ref + 2 * DPAD + 8;

    if (warp == 0) {
        tc::tmem_alloc(&s_tmem, TMEM_COLS);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        tc::mbar_init(&s_wbar, 1);
        for (int g = 0; g < MMA4_GROUPS; ++g) tc::mbar_init(&s_mbar[g], 1);
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    if (tid == 0) {
        const uint32_t total = (uint32_t)(p.ws.w_mma4_len * sizeof(float));
        tc::mbar_arrive_expect_tx(&s_wbar, total);
        const char* src = reinterpret_cast<const char*>(ws + p.ws.w_mma4);
        char* dst = reinterpret_cast<char*>(s_w);
        for (uint32_t off = 0; off < total; off += 16384u) {
            const uint32_t n = total - off < 16384u ? total - off : 16384u;
            tc::bulk_g2s(dst + off, src + off, n, &s_wbar);
        }
    }
    for (int e = tid; e < K2 * DPAD; e += blockDim.x) {
        s_mu[e] = ws[p.ws.gmm_mu + e];
        s_h[e] = ws[p.ws.gmm_h + e];
    }
    for (int e = tid; e < 64; e += blockDim.x) s_c[e] = ws[p.ws.gmm_c + e];
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

    TargetSmem tsm{s_mu, s_h, s_c, reinterpret_cast<const uint32_t*>(ws + p.ws.counter)[1], s_prior, s_ref};
    GroupCtx c;
    c.g = warp >> 2;
    c.gtid = tid & 127;
    const uint32_t tbase = s_tmem + (uint32_t)(c.g * GROUP4_COLS);
    c.t_d = tbase;
    c.t_hi = tbase + 64;
    c.t_lo = tbase + 96;
    const uint32_t lane_off = (uint32_t)((warp & 3) * 32) << 16;
    c.l_d = c.t_d + lane_off;
    c.l_hi = c.t_hi + lane_off;
    c.l_lo = c.t_lo + lane_off;
    c.bar = &s_mbar[c.g];
    c.phase = 0;
    const XSmem xs{s_x + c.g * (DPAD * 128) + c.gtid};

    uint32_t* counter = reinterpret_cast<uint32_t*>(const_cast<float*>(ws) + p.ws.counter);
    const bool from_hbm = (d.flags & SDES_F_NOISE_FROM_HBM) != 0;
    const bool ret_traj = (d.flags & SDES_F_RETURN_TRAJ) != 0;
    const int64_t B = d.batch;
    const uint32_t n_tiles = (uint32_t)((B + 127) / 128);
    const int n_chunks = p.n_chunks, chunk_steps = p.chunk_steps;
    const uint32_t n_items = n_tiles * (uint32_t)n_chunks;
    float* state = const_cast<float*>(ws) + p.ws.state;
    uint32_t* progress = reinterpret_cast<uint32_t*>(const_cast<float*>(ws) + p.ws.progress);
    for (;;) {
        if (c.gtid == 0) s_tile[c.g] = atomicAdd(counter, 1u);
        group_bar4(c.g);
        const uint32_t item = s_tile[c.g];
        if (item >= n_items) break;
        const uint32_t chunk = item / n_tiles, tile = item - chunk * n_tiles;
        const int64_t row = (int64_t)tile * 128 + c.gtid;
        const bool valid = row < B;
        const int64_t rrow = valid ? row : (B - 1);
        float* st = state + (int64_t)tile * (DPAD + 1) * 128 + c.gtid;

        float rnd;
        if (chunk == 0) {
            const TrajRef o0 = traj_ref(d, d.xs, 0, rrow);
#pragma unroll
            for (int j = 0; j < DPAD; ++j) {
                const float v = (j < dim) ? __ldg(d.x0 + rrow * dim + j) : 0.f;
                xs.p[j * 128] = v;
                if (ret_traj && valid && j < dim) o0.p[j * o0.stride] = v;
            }
            rnd = initial_rnd<DPAD>(d, xs, tsm);
        } else {
            if (c.gtid == 0) {
                uint32_t seen;
                do {
                    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(progress + tile) : "memory");
                } while (seen < chunk);
            }
            group_bar4(c.g);
#pragma unroll
            for (int j = 0; j < DPAD; ++j) xs.p[j * 128] = __ldcg(st + j * 128);
            rnd = __ldcg(st + DPAD * 128);
        }
        const uint32_t traj = (uint32_t)(d.traj_offset + (uint64_t)rrow);
        const int i_begin = (int)chunk * chunk_steps;
        const int i_end = (i_begin + chunk_steps < T) ? i_begin + chunk_steps : T;

        for (int i = i_begin; i < i_end; ++i) {
            const float* tab = ws + p.ws.tab + (int64_t)i * TAB_STRIDE;
            float sc[DPAD];
            score_part<DPAD>(d, xs, sc, tsm, ws + p.ws.gate + (int64_t)i * DPAD, tab[TAB_SIGMA], tab[TAB_LERP_W]);
            store_a_split16<DPAD>(c.l_hi, c.l_lo, xs);
            run_layer4(c, l0_hi, l0_lo, (int)K0B, C);
            layer_epilogue4(c, ws + p.ws.emb + (int64_t)i * C);
#pragma unroll 1
            for (int l = 0; l < nh; ++l) {
                run_layer4(c, lh_base + (uint32_t)l * 2u * LH_HALF, lh_base + (uint32_t)l * 2u * LH_HALF + LH_HALF, C, C);
                layer_epilogue4(c, s_bias + l * C);
            }
            run_layer4(c, lo_hi, lo_lo, C, NOUT);
            {
                const StepCoef sc_ = make_step_coef(d, tab);
                const float* nrow = from_hbm ? d.noise + ((int64_t)i * B + rrow) * dim : nullptr;
                const float* bo = s_bias + nh * C;
                const TrajRef xo_ref = traj_ref(d, d.xs, i + 1, rrow);
                float* xo = (ret_traj && valid) ? xo_ref.p : nullptr;
                const int xo_st = xo_ref.stride;
                float cost = 0.f, ito = 0.f;
                float na[8], nb[8];
                tc::tmem_ld8(c.l_d, na);
#pragma unroll
                for (int q = 0; q < DPAD / 8; q += 2) {
                    tc::wait_ld_tie<8>(na);
                    if (q + 1 < DPAD / 8) tc::tmem_ld8(c.l_d + 8u * (q + 1), nb);
                    {
                        float xv[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) { na[r] += bo[8 * q + r]; xv[r] = xs[8 * q + r]; }
                        update4(sc_, &xv[0], &na[0], &sc[8 * q], s_prior + 8 * q, s_prior + DPAD + 8 * q, 8 * q, i, traj, nrow, cost, ito);
                        update4(sc_, &xv[4], &na[4], &sc[8 * q + 4], s_prior + 8 * q + 4, s_prior + DPAD + 8 * q + 4, 8 * q + 4, i, traj, nrow, cost, ito);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            xs.p[(8 * q + r) * 128] = xv[r];
                            if (xo != nullptr && 8 * q + r < dim) xo[(8 * q + r) * xo_st] = xv[r];
                        }
                    }
                    if (q + 1 < DPAD / 8) {
                        tc::wait_ld_tie<8>(nb);
                        if (q + 2 < DPAD / 8) tc::tmem_ld8(c.l_d + 8u * (q + 2), na);
                        float xv[8];
#pragma unroll
                        for (int r = 0; r < 8; ++r) { nb[r] += bo[8 * (q + 1) + r]; xv[r] = xs[8 * (q + 1) + r]; }
                        update4(sc_, &xv[0], &nb[0], &sc[8 * q + 8], s_prior + 8 * q + 8, s_prior + DPAD + 8 * q + 8, 8 * q + 8, i, traj, nrow, cost, ito);
                        update4(sc_, &xv[4], &nb[4], &sc[8 * q + 12], s_prior + 8 * q + 12, s_prior + DPAD + 8 * q + 12, 8 * q + 12, i, traj, nrow, cost, ito);
#pragma unroll
                        for (int r = 0; r < 8; ++r) {
                            xs.p[(8 * (q + 1) + r) * 128] = xv[r];
                            if (xo != nullptr && 8 * (q + 1) + r < dim) xo[(8 * (q + 1) + r) * xo_st] = xv[r];
                        }
                    }
                }
                finish_step(d, sc_, tab, cost, ito, rnd);
            }
        }
        if (i_end == T) {
            rnd += terminal_rnd<DPAD>(d, xs, tsm);
            if (valid) {
#pragma unroll
                for (int j = 0; j < DPAD; ++j)
                    if (j < dim) d.x_T[rrow * dim + j] = xs[j];
                d.rnd[rrow] = rnd;
            }
        } else {
#pragma unroll
            for (int j = 0; j < DPAD; ++j) __stcg(st + j * 128, xs[j]);
            __stcg(st + DPAD * 128, rnd);
            __threadfence();
            group_bar4(c.g);
            if (c.gtid == 0) asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(progress + tile), "r"(chunk + 1u) : "memory");
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(s_tmem, TMEM_COLS);
}

size_t mma4_smem_bytes(const KParams& p) {
    const int dpad = p.ws.dpad, K = p.d.n_components;
    const size_t fl = (size_t)((p.ws.w_mma4_len + 31) & ~31ll) + 2 * (size_t)((K + 1) & ~1) * dpad + 64 + 2 * (2 * dpad + 8) +
                      (size_t)MMA4_GROUPS * dpad * 128;
    return fl * sizeof(float);
}

bool mma4_supported(const KParams& p) { return mma_supported(p) && mma4_smem_bytes(p) <= 226u * 1024u; }

template <int DPAD>
static cudaError_t launch_mma4_t(const KParams& p, int sm_count, cudaStream_t stream) {
    const size_t smem = mma4_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(rollout_mma4_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    // work items = tiles x time chunks: a group that finds no first-chunk tile left starts on a second chunk and waits
    // for its predecessor, so every SM is used even when there are fewer tiles than resident groups
    const int64_t items = ((p.d.batch + 127) / 128) * (int64_t)(p.n_chunks > 0 ? p.n_chunks : 1);
    int grid = (int)((items + MMA4_GROUPS - 1) / MMA4_GROUPS);
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    rollout_mma4_kernel<DPAD><<<grid, MMA4_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_rollout_mma4(const KParams& p, int sm_count, cudaStream_t stream) {
    switch (p.ws.dpad) {
        case 8: return launch_mma4_t<8>(p, sm_count, stream);
        case 16: return launch_mma4_t<16>(p, sm_count, stream);
        case 32: return launch_mma4_t<32>(p, sm_count, stream);
        case 48: return launch_mma4_t<48>(p, sm_count, stream);
        case 56: return launch_mma4_t<56>(p, sm_count, stream);
        case 64: return launch_mma4_t<64>(p, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

size_t mma_smem_bytes(const KParams& p) {
    const int dpad = p.ws.dpad, K = p.d.n_components;
    const size_t fl = (size_t)p.ws.w_mma_len + 2 * (size_t)((K + 1) & ~1) * dpad + 64 + 2 * (2 * dpad + 8);
    return fl * sizeof(float);
}

template <int DPAD>
static cudaError_t launch_mma_t(const KParams& p, int sm_count, cudaStream_t stream) {
    const size_t smem = mma_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(rollout_mma_kernel<DPAD>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    const int64_t items = ((p.d.batch + 127) / 128) * (int64_t)(p.n_chunks > 0 ? p.n_chunks : 1);
    int grid = (int)((items + MMA_GROUPS - 1) / MMA_GROUPS);
    if (grid > sm_count) grid = sm_count;
    if (grid < 1) grid = 1;
    rollout_mma_kernel<DPAD><<<grid, MMA_THREADS, smem, stream>>>(p);
    return cudaGetLastError();
}

cudaError_t launch_rollout_mma(const KParams& p, int sm_count, cudaStream_t stream) {
    switch (p.ws.dpad) {
        case 8: return launch_mma_t<8>(p, sm_count, stream);
        case 16: return launch_mma_t<16>(p, sm_count, stream);
        case 32: return launch_mma_t<32>(p, sm_count, stream);
        case 48: return launch_mma_t<48>(p, sm_count, stream);
        case 56: return launch_mma_t<56>(p, sm_count, stream);
        case 64: return launch_mma_t<64>(p, sm_count, stream);
    }
    return cudaErrorInvalidValue;
}

}  // namespace sdes
