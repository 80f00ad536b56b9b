// sdes_linear.cuh — the tcgen05 GEMM layer of the wide engine and of the lv-gradient path: operand images,
// the warp-specialised `linear_mma_kernel` with its fused epilogue, the CUDA-core cross-check kernel and the
// weight-imaging kernel.  See sdes_wide.cu for the design notes (operand images, split precision).
#pragma once

#include <cuda_bf16.h>

#include <cstdlib>

#include "sdes_common.cuh"
#include "sdes_tc.cuh"

namespace sdes {
namespace wide {

constexpr int KC = 64;                       // K elements per image block / pipeline stage
constexpr uint32_t A_HALF = 128u * KC * 2u;  // bytes of one bf16 half (hi or lo) of an activation block
constexpr uint32_t A_BLOCK = 2u * A_HALF;    // hi | lo
// warp 0: TMA producer, warp 1: MMA issuer, warps 2..: epilogue — EPI warps, EPI/4 per TMEM lane quadrant, each on its
// share of the columns.  A single warp per scheduler runs the ~80 dependent instructions of an 8-column group at
// ~10 cycles each, so wide layers (one CTA per SM) use 12 epilogue warps; skinny layers keep 4 and co-schedule 3 CTAs.
constexpr int EPI_WIDE = 12, EPI_SKINNY = 4;

static inline int round_up(int v, int a) { return (v + a - 1) / a * a; }
static inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }

// --------------------------------------------------------------------- weight operand plan
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

// What the lv-gradient pass (sdes_grad.cu) needs to know about a wide-engine workspace written in keep mode
struct WideGradView {
    int d, Hp, P, pc, T, nh, m_tiles;
    int64_t B, Bp;
    int64_t tab, emb, gate, ximg, ximg_slot, qgate, grad_base;
    int64_t vec_prior, vec_ref, gmm_h, xst, logp, sc_keep;  // kl gradient: planar parameter vectors, final state / log-density, kept scores
    bool keep_score;
    Lin mlp_in, mlp_h[SDES_MAX_HIDDEN], mlp_out;
};

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

static __global__ void __launch_bounds__(256) weight_image_kernel(const ImgArgs a) {
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

static __global__ void pad_bias_kernel(const float* __restrict__ src, int n, float* __restrict__ dst, int n_pad) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_pad; e += gridDim.x * blockDim.x) dst[e] = e < n ? src[e] : 0.f;
}

// ------------------------------------------------------------------------ the GEMM layer
struct LinArgs {
    const uint8_t* a_img; int64_t a_mt_stride;    // A operand: [m_tile] rows of k_chunks blocks
    const uint8_t* w_img;                          // B operand: [n_tile][k_chunk] blocks
    int k_chunks, n_tiles, tile_n;
    const float* bias;                             // n_pad floats or NULL
    int bias_mt_div; int64_t bias_mt_stride;       // > 0: row tile mt uses bias + (mt / div) * stride (per-time-step bias)
    int act;                                       // ACT_*
    const uint8_t* mask_img; int64_t mask_mt_stride;  // ReLU backward: keep where the stored activation > 0
    const uint8_t* mul_img; int64_t mul_mt_stride;    // multiply by a stored hi+lo image (GELU backward: gelu'(h))
    const float* resid;                            // out = resid + acc (additive coupling / gradient accumulation)
    float* out_f32; int ld_f32;                    // row-major fp32 output (same leading dimension as resid)
    uint8_t* out_img; int64_t out_mt_stride;       // next layer's A operand
    uint8_t* aux_img;                              // ACT_GELU_GRAD: image of gelu'(h), strides as out_img
};
enum { ACT_NONE = 0, ACT_RELU = 1, ACT_GELU = 2, ACT_GELU_GRAD = 3 };

// 8 consecutive output columns of one row: the fused epilogue of every layer, split into the global loads it needs
// (issued one group ahead so their latency hides behind the previous group's arithmetic and stores — with four
// epilogue warps per CTA nothing else would hide it) and the arithmetic + stores.
struct EpiPre {
    float4 b0, b1, r0, r1;
    uint4 m, mh, ml;
};

__device__ __forceinline__ void epi_prefetch(const LinArgs& a, int mt, int r, int col, EpiPre& p) {
    const int64_t row = (int64_t)mt * 128 + r;
    if (a.bias != nullptr) {
        const float* bp = a.bias + (a.bias_mt_div > 0 ? (int64_t)(mt / a.bias_mt_div) * a.bias_mt_stride : 0) + col;
        p.b0 = __ldg(reinterpret_cast<const float4*>(bp));
        p.b1 = __ldg(reinterpret_cast<const float4*>(bp + 4));
    }
    if (a.resid != nullptr) {
        const float4* rp = reinterpret_cast<const float4*>(a.resid + row * a.ld_f32 + col);
        p.r0 = rp[0];
        p.r1 = rp[1];
    }
    const int64_t goff = img_group_offset(r, col);
    if (a.mask_img != nullptr) p.m = *reinterpret_cast<const uint4*>(a.mask_img + (int64_t)mt * a.mask_mt_stride + goff);
    if (a.mul_img != nullptr) {
        const uint8_t* mp = a.mul_img + (int64_t)mt * a.mul_mt_stride + goff;
        p.mh = *reinterpret_cast<const uint4*>(mp);
        p.ml = *reinterpret_cast<const uint4*>(mp + A_HALF);
    }
}

__device__ __forceinline__ void epi_finish(const LinArgs& a, int mt, int r, int col, float (&v)[8], const EpiPre& p) {
    const int64_t row = (int64_t)mt * 128 + r;
    if (a.bias != nullptr) {
        v[0] += p.b0.x; v[1] += p.b0.y; v[2] += p.b0.z; v[3] += p.b0.w; v[4] += p.b1.x; v[5] += p.b1.y; v[6] += p.b1.z; v[7] += p.b1.w;
    }
    if (a.resid != nullptr) {
        v[0] += p.r0.x; v[1] += p.r0.y; v[2] += p.r0.z; v[3] += p.r0.w; v[4] += p.r1.x; v[5] += p.r1.y; v[6] += p.r1.z; v[7] += p.r1.w;
    }
    if (a.act == ACT_RELU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = fmaxf(v[q], 0.f);
    } else if (a.act == ACT_GELU) {
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = gelu_fast(v[q]);
    }
    const int64_t goff = img_group_offset(r, col);
    if (a.act == ACT_GELU_GRAD) {
        float gp[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) gelu_and_grad(v[q], v[q], gp[q]);
        uint4 hi, lo;
        split_pair(gp[0], gp[1], hi.x, lo.x);
        split_pair(gp[2], gp[3], hi.y, lo.y);
        split_pair(gp[4], gp[5], hi.z, lo.z);
        split_pair(gp[6], gp[7], hi.w, lo.w);
        uint8_t* o = a.aux_img + (int64_t)mt * a.out_mt_stride + goff;
        *reinterpret_cast<uint4*>(o) = hi;
        *reinterpret_cast<uint4*>(o + A_HALF) = lo;
    }
    if (a.mul_img != nullptr) {
        float mh[8], ml[8];
        unpack8(p.mh, mh);
        unpack8(p.ml, ml);
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] *= mh[q] + ml[q];
    }
    if (a.mask_img != nullptr) {
        const uint32_t w[4] = {p.m.x, p.m.y, p.m.z, p.m.w};
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

__device__ __forceinline__ void epilogue8(const LinArgs& a, int mt, int r, int col, float (&v)[8]) {
    EpiPre p;
    epi_prefetch(a, mt, r, col, p);
    epi_finish(a, mt, r, col, v, p);
}

// the accumulator tile of one thread's row, N columns starting at global column col0: TMEM loads and the global
// loads of group g+1 are in flight while group g is finished
__device__ __forceinline__ void epilogue_row(const LinArgs& a, int mt, int r, int col0, int N, uint32_t taddr) {
    float v[8], w[8];
    EpiPre pa, pb;
    if (N <= 0) return;
    tc::tmem_ld8(taddr, v);
    epi_prefetch(a, mt, r, col0, pa);
    for (int c0 = 0; c0 < N; c0 += 16) {
        const bool second = c0 + 8 < N;
        tc::wait_ld_tie<8>(v);
        if (second) {
            tc::tmem_ld8(taddr + (uint32_t)c0 + 8u, w);
            epi_prefetch(a, mt, r, col0 + c0 + 8, pb);
        }
        epi_finish(a, mt, r, col0 + c0, v, pa);
        if (second) {
            tc::wait_ld_tie<8>(w);
            if (c0 + 16 < N) {
                tc::tmem_ld8(taddr + (uint32_t)c0 + 16u, v);
                epi_prefetch(a, mt, r, col0 + c0 + 16, pa);
            }
            epi_finish(a, mt, r, col0 + c0 + 8, w, pb);
        }
    }
}

// columns [lo, lo + n) of an N-column tile served by epilogue warp `e` of its lane quadrant (EPI_WARPS / 4 warps share a row)
template <int EPI>
__device__ __forceinline__ void epi_col_range(int N, int e_in_quad, int& lo, int& n) {
    constexpr int PER = EPI / 4;
    const int groups = N / 8, g_lo = groups * e_in_quad / PER, g_hi = groups * (e_in_quad + 1) / PER;
    lo = g_lo * 8;
    n = (g_hi - g_lo) * 8;
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

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tc::smem_u32(bar)) : "memory");
}

// D[128, tile_n] = A[128, K] W[tile_n, K]^T per (n_tile, m_tile).  PERSISTENT: at most one CTA per SM, each walks
// the tile list (n fastest, so the CTAs that share an A tile run together and A is read from HBM once); the TMA ring
// runs ahead across tile boundaries and the accumulator is double-buffered in TMEM, so the epilogue of tile i
// overlaps the MMAs of tile i+1 and per-CTA setup (TMEM allocation, barriers) is paid once per launch instead of
// once per tile (the gradient path runs 8 192 row tiles of K = 64 per launch).
template <int EPI>
static __global__ void __launch_bounds__(64 + 32 * EPI, EPI == EPI_SKINNY ? 3 : 1) linear_mma_kernel(const __grid_constant__ LinArgs a, const int stages, const int m_tiles) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t s_full[4], s_empty[4], s_acc_full[2], s_acc_empty[2];
    __shared__ uint32_t s_tmem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t b_half = (uint32_t)a.tile_n * 128u;           // bytes of one bf16 half of a weight block
    const uint32_t stage_bytes = A_BLOCK + 2u * b_half;
    uint32_t ncols = 32;
    while ((int)ncols < a.tile_n) ncols <<= 1;
    const int n_total = a.n_tiles * m_tiles;
    const int n_my = (int)blockIdx.x < n_total ? (n_total - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x : 0;

    if (warp == 1) {
        tc::tmem_alloc(&s_tmem, 2u * ncols);
        tc::tmem_relinquish();
    }
    if (tid == 0) {
        for (int s = 0; s < 4; ++s) {
            tc::mbar_init(&s_full[s], 1);
            tc::mbar_init(&s_empty[s], 1);
        }
        for (int s = 0; s < 2; ++s) {
            tc::mbar_init(&s_acc_full[s], 1);
            tc::mbar_init(&s_acc_empty[s], 32 * EPI);
        }
        tc::fence_mbar_init();
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = s_tmem;

    if (warp == 0) {
        if (lane == 0) {  // ---- producer: one elected thread drives the TMA bulk copies
            int iter = 0;
            for (int i = 0; i < n_my; ++i) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x, nt = tile % a.n_tiles, mt = tile / a.n_tiles;
                const uint8_t* a_src = a.a_img + (int64_t)mt * a.a_mt_stride;
                const uint8_t* w_src = a.w_img + (int64_t)nt * a.k_chunks * (2ll * b_half);
                for (int kc = 0; kc < a.k_chunks; ++kc, ++iter) {
                    const int s = iter % stages, it = iter / stages;
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
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ---- MMA issuer
            const uint32_t idesc = tc::idesc_bf16(128, a.tile_n);
            const uint32_t lbo_b = (uint32_t)a.tile_n * 16u;
            int iter = 0;
            for (int i = 0; i < n_my; ++i) {
                const int buf = i & 1, use = i >> 1;
                if (use > 0) {  // the epilogue must have drained this accumulator buffer
                    tc::mbar_wait(&s_acc_empty[buf], (uint32_t)((use - 1) & 1));
                    tc::fence_after();
                }
                const uint32_t tmem_d = tmem_base + (uint32_t)buf * ncols;
                for (int kc = 0; kc < a.k_chunks; ++kc, ++iter) {
                    const int s = iter % stages, it = iter / stages;
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
                tc::mma_commit(&s_acc_full[buf]);
            }
        }
    } else {  // ---- epilogue warps: TMEM lane quadrant = warp % 4
        const int q = warp & 3, r = q * 32 + lane;
        int c_lo, c_n;
        epi_col_range<EPI>(a.tile_n, (warp - 2) >> 2, c_lo, c_n);
        for (int i = 0; i < n_my; ++i) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x, nt = tile % a.n_tiles, mt = tile / a.n_tiles;
            const int buf = i & 1, use = i >> 1;
            tc::mbar_wait(&s_acc_full[buf], (uint32_t)(use & 1));
            tc::fence_after();
            const uint32_t taddr = tmem_base + (uint32_t)buf * ncols + ((uint32_t)(q * 32) << 16);
            epilogue_row(a, mt, r, nt * a.tile_n + c_lo, c_n, taddr + (uint32_t)c_lo);
            tc::fence_before();
            mbar_arrive(&s_acc_empty[buf]);  // every epilogue thread has read its part of the accumulator
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) tc::tmem_dealloc(tmem_base, 2u * ncols);
}

// The same layer on the CUDA cores, reading the same operand images: the cross-check engine (SDES_F_MLP_SIMT).
static __global__ void __launch_bounds__(128) linear_simt_kernel(const LinArgs a) {
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

static cudaError_t launch_linear(const LinArgs& a, int m_tiles, bool simt, cudaStream_t stream, int64_t& launches) {
    ++launches;
    if (simt) {
        linear_simt_kernel<<<dim3(a.n_tiles, m_tiles), 128, 0, stream>>>(a);
        return cudaGetLastError();
    }
    static bool attr_set = false;  // one process per GPU (the host side is single-threaded like the reference)
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(linear_mma_kernel<EPI_WIDE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(linear_mma_kernel<EPI_SKINNY>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) return e;
        attr_set = true;
    }
    const uint32_t stage_bytes = A_BLOCK + 2u * (uint32_t)a.tile_n * 128u;
    // Skinny layers (K <= 128, N <= 64: the control MLP over millions of rows) are bound by the epilogue's CUDA-core
    // work, not by the tensor pipe: give them three co-resident CTAs per SM (12 epilogue warps) with a one-stage ring
    // each; wide layers keep the SM to themselves and use the shared memory for a deep ring.
    const int ctas_per_sm = (a.k_chunks <= 2 && stage_bytes <= 64u * 1024u) ? 3 : 1;
    int stages = (int)(200u * 1024u / (uint32_t)ctas_per_sm / stage_bytes);
    if (stages > 4) stages = 4;
    if (stages < 1) stages = 1;
    static int sms = 0;
    if (sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    }
    const int64_t total = (int64_t)a.n_tiles * m_tiles;
    const int grid = (int)(total < (int64_t)sms * ctas_per_sm ? total : (int64_t)sms * ctas_per_sm);
    if (ctas_per_sm > 1) linear_mma_kernel<EPI_SKINNY><<<grid, 64 + 32 * EPI_SKINNY, (size_t)stages * stage_bytes, stream>>>(a, stages, m_tiles);
    else linear_mma_kernel<EPI_WIDE><<<grid, 64 + 32 * EPI_WIDE, (size_t)stages * stage_bytes, stream>>>(a, stages, m_tiles);
    return cudaGetLastError();
}

}  // namespace wide
}  // namespace sdes
