// tcgen05 / TMEM / mbarrier primitives (inline PTX, sm_100a) shared by the tensor-core decoders.
//
// Operand images in shared memory (SWIZZLE_NONE "interleave" canonical layout, 16-byte units):
//   activation tile  [F/4 chunks][128 samples][4 x tf32]   chunk stride 2048 B
//   weight image     [IN/4 chunks][OUTP rows ][4 x tf32]   chunk stride OUTP*16 B   (W[out][in], OUTP = OUT padded to 8)
// One image serves several operand roles (see cute/atom/mma_traits_sm100.hpp, make_umma_desc):
//   K-major  operand (rows = M or N index, chunks = K):  LBO = chunk stride, SBO = 128 B, start += 2 chunks per K-step
//   MN-major operand (chunks = M or N index, rows = K):  SBO = chunk stride, LBO = 128 B, start += 128 B  per K-step
// so  Y = X W^T   uses tile K-major x weights K-major,
//     dX = dY W   uses tile K-major x weights MN-major,
//     dW = G^T X  uses tile MN-major x tile MN-major (K = the 128 samples of the tile).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define TC_TILE_M 128
#define TC_CHUNK_BYTES 2048  // 128 samples x 16 B

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier -------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// try_wait with a suspend-time hint: the waiting warp is parked by the hardware (and woken when the phase completes) instead
// of re-issuing the probe in a tight loop -- measured: with the plain form the three waiting warps of a scheduler left the
// one working warp (MMA issue, semantic head) about a quarter of the issue slots
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "LAB_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
        "@P1 bra DONE;\n\t"
        "bra LAB_WAIT;\n\t"
        "DONE:\n\t"
        "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680u) : "memory");
}

// ---- single-lane election in warp-uniform control flow ------------------------------------------------------------
// tcgen05.mma / commit live on the uniform datapath.  Issued under a divergent `if (threadIdx.x == 0)` the compiler wraps
// every one of them in a per-lane waterfall loop (~100 cycles per MMA, measured with tools/mma_bench.py); issued under
// `warp_id_uniform() == w && elect_one()` they are straight-line uniform instructions.
__device__ __forceinline__ int warp_id_uniform() { return __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}" : "=r"(pred));
    return pred != 0;
}

// ---- bulk asynchronous copies (TMA, 1-D): global <-> shared, 16-byte aligned, size a multiple of 16 ---------------------
// load: completes on an mbarrier armed with the byte count; store: tracked by the thread's bulk async-group
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulk_wait_read_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }

// ---- proxies / fences -----------------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM allocation (one full warp) ----------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------
// shared-memory matrix descriptor, SWIZZLE_NONE, version 1 (Blackwell)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // version
    return d;
}
// instruction descriptor, kind::tf32, fp32 accumulate
__host__ __device__ constexpr uint32_t idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem];  issued by ONE thread
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// arrive on `bar` when all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- TMEM -> registers: thread t of warp w reads lane 32*(w%4)+t, 16 consecutive columns -----
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// registers -> TMEM: the inverse of tmem_ld16 (thread t of warp w writes lane 32*(w%4)+t, 16 consecutive columns)
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
           "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
           "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
           "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_st16_nowait(uint32_t taddr, const float (&v)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};"
        :: "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
           "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
           "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
           "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15]))
        : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// 2^x on the SFU, one instruction (flushes denormals; 2^-inf = 0)
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ---- MMA chains over the canonical images -------------------------------------------------------
// Y[128 x N] = X_tile[128 x K] * W[N x K]^T          (A K-major, B K-major)
__device__ __forceinline__ void mma_fwd(uint32_t tmem_d, uint32_t a_tile, uint32_t w_img, int N, int NP, int K, bool accumulate) {
    const uint32_t id = idesc_tf32(128, N, 0, 0);
    for (int k = 0; k < K / 8; ++k) {
        const uint64_t ad = smem_desc(a_tile + k * 2 * TC_CHUNK_BYTES, TC_CHUNK_BYTES, 128);
        const uint64_t bd = smem_desc(w_img + k * 2 * NP * 16, NP * 16, 128);
        umma_tf32(tmem_d, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}
// dX[128 x N] = dY_tile[128 x K] * W[K x N]           (A K-major, B = the same weight image read MN-major)
__device__ __forceinline__ void mma_bwd_data(uint32_t tmem_d, uint32_t g_tile, uint32_t w_img, int N, int KP, int K, bool accumulate) {
    const uint32_t id = idesc_tf32(128, N, 0, 1);
    for (int k = 0; k < K / 8; ++k) {
        const uint64_t ad = smem_desc(g_tile + k * 2 * TC_CHUNK_BYTES, TC_CHUNK_BYTES, 128);
        const uint64_t bd = smem_desc(w_img + k * 128, 128, KP * 16);
        umma_tf32(tmem_d, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}
// dW[128(pad) x N] (+)= G_tile^T[feat x 128 samples] * X_tile[128 samples x N]   (A MN-major, B MN-major, K = samples)
__device__ __forceinline__ void mma_bwd_weight(uint32_t tmem_d, uint32_t g_tile, uint32_t x_tile, int N, bool accumulate) {
    const uint32_t id = idesc_tf32(128, N, 1, 1);
    for (int k = 0; k < TC_TILE_M / 8; ++k) {
        const uint64_t ad = smem_desc(g_tile + k * 128, 128, TC_CHUNK_BYTES);
        const uint64_t bd = smem_desc(x_tile + k * 128, 128, TC_CHUNK_BYTES);
        umma_tf32(tmem_d, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}

// ---- kind::f16 (fp16 operands, fp32 accumulate): 16-byte unit = 8 halfs, K = 16 per instruction ---------------
//   activation tile  [F/8 chunks][128 samples][8 x half]   chunk stride 2048 B
//   weight image     [IN/8 chunks][OUTP rows ][8 x half]   chunk stride OUTP*16 B
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
    return (1u << 4) | (0u << 7) | (0u << 10) | ((uint32_t)a_mn_major << 15) | ((uint32_t)b_mn_major << 16) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// Y[128 x N] = X_tile[128 x K] * W[N x K]^T          (A K-major, B K-major); K multiple of 16
__device__ __forceinline__ void mma16_fwd(uint32_t tmem_d, uint32_t a_tile, uint32_t w_img, int N, int NP, int K, bool accumulate) {
    const uint32_t id = idesc_f16(128, N, 0, 0);
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t ad = smem_desc(a_tile + k * 2 * TC_CHUNK_BYTES, TC_CHUNK_BYTES, 128);
        const uint64_t bd = smem_desc(w_img + k * 2 * NP * 16, NP * 16, 128);
        umma_f16(tmem_d, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}
// dX[128 x N] = dY_tile[128 x K] * W[K x N]           (A K-major, B = the same weight image read MN-major)
__device__ __forceinline__ void mma16_bwd_data(uint32_t tmem_d, uint32_t g_tile, uint32_t w_img, int N, int KP, int K, bool accumulate) {
    const uint32_t id = idesc_f16(128, N, 0, 1);
    for (int k = 0; k < K / 16; ++k) {
        const uint64_t ad = smem_desc(g_tile + k * 2 * TC_CHUNK_BYTES, TC_CHUNK_BYTES, 128);
        const uint64_t bd = smem_desc(w_img + k * 256, 128, KP * 16);
        umma_f16(tmem_d, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}
// dW[128(pad) x N] (+)= G_tile^T * X_tile   (A MN-major, B MN-major, K = the 128 samples)
__device__ __forceinline__ void mma16_bwd_weight(uint32_t tmem_d, uint32_t g_tile, uint32_t x_tile, int N, bool accumulate) {
    const uint32_t id = idesc_f16(128, N, 1, 1);
    for (int k = 0; k < TC_TILE_M / 16; ++k) {
        const uint64_t ad = smem_desc(g_tile + k * 256, 128, TC_CHUNK_BYTES);
        const uint64_t bd = smem_desc(x_tile + k * 256, 128, TC_CHUNK_BYTES);
        umma_f16(tmem_d, ad, bd, id, (accumulate || k > 0) ? 1u : 0u);
    }
}
