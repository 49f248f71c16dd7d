// Parity probe for the tcgen05 building blocks (descriptors, TMEM, commit/mbarrier): one CTA, one tile.
// mode 0: D[128,N] = A[128,K] * B[N,K]^T          (K-major x K-major)        -- forward layer
// mode 1: D[128,N] = A[128,K] * B[K,N]            (K-major x MN-major)       -- backward data
// mode 2: D[FA(pad 128),N] = A[128,FA]^T * B[128,N]  (MN-major x MN-major)   -- backward weights (K = 128 samples)
#include "common.cuh"
#include "tc_common.cuh"

__global__ void __launch_bounds__(128) tc_gemm_test_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B,
                                                           float* __restrict__ D, int N, int K, int FA, int reps) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    float* a_img = reinterpret_cast<float*>(smem_raw);                       // 64 KB: up to 32 chunks
    float* b_img = reinterpret_cast<float*>(smem_raw + 32 * TC_CHUNK_BYTES); // 64 KB
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 512 * 2; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    __syncthreads();
    const int KP = (K + 7) & ~7, NP = (N + 7) & ~7;
    if (mode == 0 || mode == 1) {
        for (int f = 0; f < K; ++f) a_img[((f >> 2) * 128 + tid) * 4 + (f & 3)] = A[tid * K + f];
        if (mode == 0) {  // weight image of W = B [N][K]: [(k/4)][n (NP rows)][4]
            for (int i = tid; i < N * K; i += 128) { const int n = i / K, k = i - n * K; b_img[((k >> 2) * NP + n) * 4 + (k & 3)] = B[i]; }
        } else {          // weight image of W = B [K(out)][N(in)]: [(in/4)][out (KP rows)][4]
            for (int i = tid; i < K * N; i += 128) { const int o = i / N, in = i - o * N; b_img[((in >> 2) * KP + o) * 4 + (in & 3)] = B[i]; }
        }
    } else {
        for (int f = 0; f < FA; ++f) a_img[((f >> 2) * 128 + tid) * 4 + (f & 3)] = A[tid * FA + f];
        for (int f = 0; f < N; ++f) b_img[((f >> 2) * 128 + tid) * 4 + (f & 3)] = B[tid * N + f];
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {   // reps > 1 exercises accumulate = 1 (result scales by reps)
        if (tid == 0) {
            if (mode == 0) mma_fwd(tmem, smem_u32(a_img), smem_u32(b_img), N, NP, K, rep > 0);
            else if (mode == 1) mma_bwd_data(tmem, smem_u32(a_img), smem_u32(b_img), N, KP, K, rep > 0);
            else mma_bwd_weight(tmem, smem_u32(a_img), smem_u32(b_img), N, rep > 0);
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
    }
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

#include <cuda_fp16.h>
__global__ void __launch_bounds__(128) tc_gemm_test16_kernel(int mode, const float* __restrict__ A, const float* __restrict__ B,
                                                             float* __restrict__ D, int N, int K, int FA, int reps, int lbo_alt) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __half* a_img = reinterpret_cast<__half*>(smem_raw);
    __half* b_img = reinterpret_cast<__half*>(smem_raw + 32 * TC_CHUNK_BYTES);
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 32 * 512 * 2; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    __syncthreads();
    const int KP = (K + 7) & ~7, NP = (N + 7) & ~7;
    if (mode == 0 || mode == 1) {
        for (int f = 0; f < K; ++f) a_img[((f >> 3) * 128 + tid) * 8 + (f & 7)] = __float2half(A[tid * K + f]);
        if (mode == 0) {
            for (int i = tid; i < N * K; i += 128) { const int n = i / K, k = i - n * K; b_img[((k >> 3) * NP + n) * 8 + (k & 7)] = __float2half(B[i]); }
        } else {
            for (int i = tid; i < K * N; i += 128) { const int o = i / N, in = i - o * N; b_img[((in >> 3) * KP + o) * 8 + (in & 7)] = __float2half(B[i]); }
        }
    } else {
        for (int f = 0; f < FA; ++f) a_img[((f >> 3) * 128 + tid) * 8 + (f & 7)] = __float2half(A[tid * FA + f]);
        for (int f = 0; f < N; ++f) b_img[((f >> 3) * 128 + tid) * 8 + (f & 7)] = __float2half(B[tid * N + f]);
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base;
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            if (mode == 0) mma16_fwd(tmem, smem_u32(a_img), smem_u32(b_img), N, NP, K, rep > 0);
            else if (mode == 1) mma16_bwd_data(tmem, smem_u32(a_img), smem_u32(b_img), N, KP, K, rep > 0);
            else mma16_bwd_weight(tmem, smem_u32(a_img), smem_u32(b_img), N, rep > 0);
            umma_commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
    }
    for (int c0 = 0; c0 < N; c0 += 16) {
        float v[16];
        tmem_ld16(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
        for (int i = 0; i < 16; ++i) D[tid * N + c0 + i] = v[i];
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}

// MMA issue / completion timing probe: thread 0 issues `chains` operand chains (K/16 MMAs each) back to back, commits and
// waits, `reps` times; cycles[0] = total clock64 ticks, cycles[1] = ticks spent issuing.
__global__ void __launch_bounds__(128) tc_mma_bench_kernel(int mode, int N, int K, int chains, int reps, long long* __restrict__ cycles) {
    extern __shared__ __align__(128) uint8_t smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    const int tid = threadIdx.x, warp = warp_id_uniform();
    for (int i = tid; i < 32 * 512 * 2; i += 128) reinterpret_cast<float*>(smem_raw)[i] = 0.f;
    __syncthreads();
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(&tmem_base, 256);
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = tmem_base, a = smem_u32(smem_raw), b = smem_u32(smem_raw + 32 * TC_CHUNK_BYTES);
    const int KP = (K + 7) & ~7, NP = (N + 7) & ~7;
    uint32_t parity = 0;
    long long t_issue = 0;
    const long long t0 = clock64();
    for (int rep = 0; rep < reps; ++rep) {
        if (warp == 0 && elect_one()) {
            const long long ti = clock64();
            for (int c = 0; c < chains; ++c) {
                if (mode == 0) mma16_fwd(tmem, a, b, N, NP, K, true);
                else if (mode == 1) mma16_bwd_data(tmem, a, b, N, KP, K, true);
                else mma16_bwd_weight(tmem, a, b, N, true);
            }
            umma_commit(&bar);
            t_issue += clock64() - ti;
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        tc_fence_after();
    }
    const long long t1 = clock64();
    if (tid == 0) { cycles[0] = t1 - t0; cycles[1] = t_issue; }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 256);
}
extern "C" int pag_tc_mma_bench(int mode, int N, int K, int chains, int reps, int64_t* cycles, void* stream) {
    if (N % 16 || N > 256 || N < 16 || K % 16 || K > 256) return PAG_ERR_ARG;
    const size_t bytes = 64 * TC_CHUNK_BYTES;
    cudaError_t e = cudaFuncSetAttribute(tc_mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    tc_mma_bench_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(mode, N, K, chains, reps, reinterpret_cast<long long*>(cycles));
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

extern "C" int pag_tc_gemm_test16(int mode, const float* A, const float* B, float* D, int N, int K, int FA, int reps, void* stream) {
    if (N % 16 || N > 256 || N < 16) return PAG_ERR_ARG;
    const size_t bytes = 64 * TC_CHUNK_BYTES;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_test16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    tc_gemm_test16_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(mode, A, B, D, N, K, FA, reps, 0);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

extern "C" int pag_tc_gemm_test(int mode, const float* A, const float* B, float* D, int N, int K, int FA, int reps, void* stream) {
    if (N % 16 || N > 256 || N < 16) return PAG_ERR_ARG;
    const size_t bytes = 64 * TC_CHUNK_BYTES;
    cudaError_t e = cudaFuncSetAttribute(tc_gemm_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (e != cudaSuccess) return (int)e;
    tc_gemm_test_kernel<<<1, 128, bytes, (cudaStream_t)stream>>>(mode, A, B, D, N, K, FA, reps);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
