// sdes_tc.cuh — tcgen05 / TMEM / mbarrier / bulk-copy primitives (inline PTX, sm_100a) and the
// split-precision dense layer used by the control MLP.
//
// One GROUP = 4 consecutive warps = 128 trajectories = one M=128 MMA tile; thread r of the group
// owns TMEM lane r (row r).  Per layer
//     D[128, N] (TMEM, fp32)  =  A[128, K] (TMEM)  x  W[N, K]^T (shared memory, K-major)
// is issued by one thread as K/16 k-steps of three kind::f16 (bf16) MMAs
//     A_lo*W_hi + A_hi*W_lo + A_hi*W_hi          (hi = bf16(v), lo = bf16(v - hi): 16 significant bits)
// which recovers ~2^-16 relative accuracy of sum|a||w| from the 8-bit-mantissa tensor-core format — the
// reference computes in true fp32 (TF32 is off by default in PyTorch), so a single bf16 / tf32 pass
// (~4e-3 / 5e-4 per layer) would not hold the stated 2e-4 tolerance over 100 steps.
#pragma once

#include "sdes_common.cuh"

namespace sdes {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ------------------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
        "@P1 bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}

// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(smem_dst)),
                 "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// TMA 1-D bulk copy shared -> global (bulk async-group completion).  The caller makes its generic-proxy writes of the
// source visible first (fence_proxy_async + a barrier), and waits for the read side before reusing the buffer.
__device__ __forceinline__ void bulk_s2g(void* gmem_dst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gmem_dst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---------------------------------------------------------------------------------- TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)), "r"(ncols)
                 : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// tcgen05.ld fills its destination registers asynchronously; they may only be read after
// tcgen05.wait::ld.  The compiler does not know that, so tie the registers to the wait: every
// value passes through an asm statement that is ordered after the wait (volatile asms keep
// their order), which keeps consumers from being scheduled above it.
__device__ __forceinline__ void tie8(float* v) {
    asm volatile("" : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7])::"memory");
}
template <int N>
__device__ __forceinline__ void wait_ld_tie(float (&v)[N]) {
    wait_ld();
#pragma unroll
    for (int c = 0; c < N; c += 8) tie8(&v[c]);
}
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// 8 consecutive 32-bit columns of this thread's lane
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float* v) {
    uint32_t r0, r1, r2, r3, r4, r5, r6, r7;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3), "=r"(r4), "=r"(r5), "=r"(r6), "=r"(r7)
                 : "r"(taddr));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1); v[2] = __uint_as_float(r2); v[3] = __uint_as_float(r3);
    v[4] = __uint_as_float(r4); v[5] = __uint_as_float(r5); v[6] = __uint_as_float(r6); v[7] = __uint_as_float(r7);
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}

// ----------------------------------------------------------------------------------- MMA
// Instruction descriptor (cute/arch/mma_sm100_desc.hpp InstrDescriptor): c_format[4,6)=1 (F32),
// a_format[7,10)=1 (BF16), b_format[10,13)=1 (BF16), a_major[15]=0 (K), b_major[16]=0 (K),
// n_dim[17,23)=N>>3, m_dim[24,29)=M>>4.
//
// Shared-memory matrix descriptor, K-major, no swizzle ("interleave"): 8-row x 16-byte core
// matrices; LBO = byte distance between the two 16-byte K-chunks of one MMA (K=16 bf16),
// SBO = byte distance between consecutive 8-row groups.  version=1 (Blackwell) at bit 46.
__device__ __forceinline__ uint64_t smem_desc_kmajor(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
           (1ull << 46);
}
// completion of all previously issued MMAs of this thread -> one arrive on an mbarrier
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__host__ __device__ constexpr uint32_t idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// bf16 weight image (N rows x K cols, K multiple of 16), K-major, no swizzle: core matrix = 8 rows x 16 bytes
// = 8 rows x 8 bf16.  offset in bf16 elements: (k/8) * (N*8) + (n/8) * 64 + (n%8) * 8 + (k%8)
//   => LBO = N*16 bytes, SBO = 128 bytes; one MMA (K=16) spans two K-chunks, k-step s starts at byte s*2*LBO.
__host__ __device__ inline int64_t wimg16_offset(int n, int k, int N) {
    return (int64_t)(k / 8) * (N * 8) + (n / 8) * 64 + (n % 8) * 8 + (k % 8);
}
// two fp32 -> packed bf16x2 (round to nearest even); `lo_k` lands in the low half (even k)
__device__ __forceinline__ uint32_t pack_bf16x2(float lo_k, float lo_k1) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(lo_k1), "f"(lo_k));
    return r;
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, const uint32_t* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]),
                 "r"(v[3])
                 : "memory");
}

// v = hi + lo with hi = bf16(v), lo = bf16(v - hi), two values at a time (a lands in the low half)
__device__ __forceinline__ void split_bf16_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(a, b);
    const float ha = __uint_as_float(hi << 16), hb = __uint_as_float(hi & 0xFFFF0000u);
    lo = pack_bf16x2(a - ha, b - hb);
}

// the same on a register pair, the subtraction as one packed FFMA2 (v - hi = fma(hi, -1, v), exact)
__device__ __forceinline__ void split_bf16_pair2(float2 v, uint32_t& hi, uint32_t& lo) {
    hi = pack_bf16x2(v.x, v.y);
    const float2 h = make_float2(__uint_as_float(hi << 16), __uint_as_float(hi & 0xFFFF0000u));
    const float2 l = __ffma2_rn(h, make_float2(-1.f, -1.f), v);
    lo = pack_bf16x2(l.x, l.y);
}

// bf16x3 layer (4-group engine): A = A_hi + A_lo and W = W_hi + W_lo, both halves bf16 (16 significant bits per
// operand), D = A_lo*W_hi + A_hi*W_lo + A_hi*W_hi in fp32 — the wide engine's scheme with A in TMEM.  A halves are
// packed two values per 32-bit column (8 columns per K=16 step).  Called by ONE thread.
__device__ __forceinline__ void issue_layer_bf16x3(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo,
                                                   uint32_t w_hi_saddr, uint32_t w_lo_saddr, int K16, int N) {
    const uint32_t id16 = idesc_bf16(128, N);
    const uint32_t lbo = (uint32_t)N * 16u;
    for (int s = 0; s < (K16 >> 4); ++s) {
        const uint64_t bh = smem_desc_kmajor(w_hi_saddr + (uint32_t)s * 2u * lbo, lbo, 128u);
        const uint64_t bl = smem_desc_kmajor(w_lo_saddr + (uint32_t)s * 2u * lbo, lbo, 128u);
        mma_f16_ts(tmem_d, tmem_a_lo + 8u * s, bh, id16, s > 0 ? 1u : 0u);
        mma_f16_ts(tmem_d, tmem_a_hi + 8u * s, bl, id16, 1u);
        mma_f16_ts(tmem_d, tmem_a_hi + 8u * s, bh, id16, 1u);
    }
}

// The same issued by a CONVERGED warp: every lane runs the warp-uniform descriptor arithmetic, one elected lane issues.
// From a divergent `if (lane == 0)` the compiler wraps each tcgen05.mma in an elect / branch loop (~10 instructions and a
// branch per MMA); converged, the MMAs are straight-line predicated instructions.
__device__ __forceinline__ void mma_f16_ts_elect(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
    asm volatile(
        "{\n\t"
        ".reg .pred q;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t"
        "}" ::"r"(smem_u32(bar))
        : "memory");
}
// The three MMAs of one k-step behind ONE elect; the weight descriptors as (low word = address / LBO, constant high word):
// no 64-bit shift / mask arithmetic per MMA (the issuing warp is one of the group's four epilogue warps).
__device__ __forceinline__ void mma3_f16_ts_elect(uint32_t tmem_d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi_lo32, uint32_t b_lo_lo32,
                                                  uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p, q, t;\n\t"
        ".reg .b64 dbh, dbl;\n\t"
        "mov.b64 dbh, {%3, %7};\n\t"
        "mov.b64 dbl, {%4, %7};\n\t"
        "setp.ne.b32 p, %6, 0;\n\t"
        "setp.eq.u32 t, %0, %0;\n\t"
        "elect.sync _|q, 0xffffffff;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%2], dbh, %5, p;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], dbl, %5, t;\n\t"
        "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], dbh, %5, t;\n\t"
        "}" ::"r"(tmem_d), "r"(a_hi), "r"(a_lo), "r"(b_hi_lo32), "r"(b_lo_lo32), "r"(idesc), "r"(accumulate), "n"((128u >> 4) | (1u << 14))
        : "memory");
}
__device__ __forceinline__ void issue_layer_bf16x3_warp(uint32_t tmem_d, uint32_t tmem_a_hi, uint32_t tmem_a_lo,
                                                        uint32_t w_hi_saddr, uint32_t w_lo_saddr, int K16, int N) {
    const uint32_t id16 = idesc_bf16(128, N);
    const uint32_t lbo = (uint32_t)N * 16u;
    const uint32_t bh = ((w_hi_saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16), bl = ((w_lo_saddr & 0x3FFFFu) >> 4) | ((lbo >> 4) << 16);
    const uint32_t step = (2u * lbo) >> 4;  // one k-step = two K-chunks of the weight image
    for (int s = 0; s < (K16 >> 4); ++s)
        mma3_f16_ts_elect(tmem_d, tmem_a_hi + 8u * s, tmem_a_lo + 8u * s, bh + (uint32_t)s * step, bl + (uint32_t)s * step, id16, s > 0 ? 1u : 0u);
}

}  // namespace tc
}  // namespace sdes
