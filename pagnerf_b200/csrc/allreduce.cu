// Gradient all-reduce over NVLink 5 / NVSwitch peer memory (ray-sharded data parallelism).
//
// The reference is single-GPU; its multi-GPU counterpart would be DistributedDataParallel's NCCL all-reduce of the two 50 MB grid
// gradient tables.  Here the tables live in SYMMETRIC memory (the same virtual layout on every rank, mapped into every peer and
// into one NVSwitch multicast address), the scatter kernels accumulate into the local copy, and the exchange is ONE kernel per
// table, in place:
//   multicast path (NVLS): rank r owns the r-th slice; `multimem.ld_reduce.add.v4.f32` returns the sum over all ranks of 16 bytes
//     in one instruction -- the addition happens inside the switch --, the mean is broadcast back with `multimem.st`: every byte
//     crosses each GPU's link once in each direction (the two-shot optimum), no staging buffers, no SM-side reduction tree;
//   peer path (no multicast): the same slice ownership with explicit 16-byte loads from every peer and stores to every peer.
// Cross-GPU ordering (all scatters finished before the first load, all stores landed before the first consumer) is provided by
// the caller with symmetric-memory signal-pad barriers on the same stream.  256-thread CTAs with ~20 registers and no shared
// memory: they fit beside the persistent decoder CTAs, so the exchange of the first table overlaps the rest of the backward.
#include "common.cuh"

struct PeerPtrs { float* p[16]; };

__device__ __forceinline__ float4 multimem_ld_reduce_add(const float* mc) {
    float4 v;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(mc) : "memory");
    return v;
}
__device__ __forceinline__ void multimem_st(float* mc, float4 v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// AR_UNROLL independent 16-byte switch reductions in flight per thread: an NVLink round trip is a few microseconds, so the
// bytes in flight (grid x 256 threads x AR_UNROLL x 16 B ~ 10 MB), not the instruction rate, set the bandwidth
#define AR_UNROLL 8
__global__ void __launch_bounds__(256) allreduce_mc_kernel(float* __restrict__ mc, int64_t q0, int64_t q1, float mult) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; q + (AR_UNROLL - 1) * stride < q1; q += AR_UNROLL * stride) {
        float4 v[AR_UNROLL];
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u) v[u] = multimem_ld_reduce_add(mc + 4 * (q + u * stride));
#pragma unroll
        for (int u = 0; u < AR_UNROLL; ++u) {
            v[u].x *= mult; v[u].y *= mult; v[u].z *= mult; v[u].w *= mult;
            multimem_st(mc + 4 * (q + u * stride), v[u]);
        }
    }
    for (; q < q1; q += stride) {
        float4 v = multimem_ld_reduce_add(mc + 4 * q);
        v.x *= mult; v.y *= mult; v.z *= mult; v.w *= mult;
        multimem_st(mc + 4 * q, v);
    }
}

__device__ __forceinline__ float4 ld_peer(const float* p) {      // straight from the owner's memory, never a stale L1 line
    float4 v;
    asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");      // L1 is clean at kernel start, peer lines bypass the local L2
    return v;
}
template <int WORLD>
__global__ void __launch_bounds__(256) allreduce_p2p_kernel(PeerPtrs peers, int world, int64_t q0, int64_t q1, float mult) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    constexpr int U = WORLD ? (WORLD <= 2 ? 8 : (WORLD <= 4 ? 4 : 2)) : 1;      // U x world independent 16-byte peer loads in flight
    const int W = WORLD ? WORLD : world;
    int64_t q = q0 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; q + (U - 1) * stride < q1; q += U * stride) {
        float4 acc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) acc[u] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < (WORLD ? WORLD : 16); ++r) {
            if (r < W) {
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const float4 v = ld_peer(peers.p[r] + 4 * (q + u * stride));
                    acc[u].x += v.x; acc[u].y += v.y; acc[u].z += v.z; acc[u].w += v.w;
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            acc[u].x *= mult; acc[u].y *= mult; acc[u].z *= mult; acc[u].w *= mult;
#pragma unroll
            for (int r = 0; r < (WORLD ? WORLD : 16); ++r)
                if (r < W) *reinterpret_cast<float4*>(peers.p[r] + 4 * (q + u * stride)) = acc[u];
        }
    }
    for (; q < q1; q += stride) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r = 0; r < W; ++r) {
            const float4 v = ld_peer(peers.p[r] + 4 * q);
            acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
        }
        acc.x *= mult; acc.y *= mult; acc.z *= mult; acc.w *= mult;
        for (int r = 0; r < W; ++r) *reinterpret_cast<float4*>(peers.p[r] + 4 * q) = acc;
    }
}

// Cross-GPU barrier through flags in the symmetric buffer itself: thread t of ONE CTA publishes this rank's epoch into peer t's
// flag slot [channel][rank] (system-scope release store over NVLink) and spins (system-scope acquire loads of LOCAL memory)
// until peer t's epoch has arrived in its own slot [channel][t].  The epoch counter lives in device memory and is advanced by
// the kernel, so a captured CUDA graph replays correctly.  ~2 us of NVLink latency + one launch; everything the stream ran before
// is ordered before the release (kernel boundary + __threadfence_system), everything after sees the peers' data.
#define SYMM_FLAG_STRIDE 16      // ranks per channel row
__global__ void symm_barrier_kernel(PeerPtrs peers, int64_t flag_off, int* __restrict__ epoch, int rank, int world, int channel) {
    __shared__ int e_s;
    if (threadIdx.x == 0) { e_s = epoch[channel] + 1; epoch[channel] = e_s; }
    __threadfence_system();
    __syncthreads();
    const int t = threadIdx.x, e = e_s;
    if (t < world) {
        int* remote = reinterpret_cast<int*>(peers.p[t] + flag_off) + channel * SYMM_FLAG_STRIDE + rank;
        asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(remote), "r"(e) : "memory");
        const int* local = reinterpret_cast<const int*>(peers.p[rank] + flag_off) + channel * SYMM_FLAG_STRIDE + t;
        int v;
        do {
            asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(local) : "memory");
        } while (v - e < 0);
    }
    __syncthreads();
    __threadfence_system();
}

extern "C" {

// Barrier across the ranks of a symmetric buffer on `stream` (see symm_barrier_kernel).  peers: host array of the `world` ranks'
// buffer base pointers; the flags are int32 [n_channels][16] at element offset flag_offset of every rank's buffer (zeroed once
// before first use); epoch: device int32[n_channels], zeroed once, private to this rank.
int pag_symm_barrier(float* const* peers, int64_t flag_offset, int* epoch, int rank, int world, int channel, void* stream) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world || channel < 0 || !peers || !epoch) return PAG_ERR_ARG;
    PeerPtrs pp;
    for (int r = 0; r < world; ++r) pp.p[r] = peers[r];
    symm_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pp, flag_offset, epoch, rank, world, channel);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// In-place scaled all-reduce (x <- mult * sum over ranks of x) of elements [offset, offset + n) of a symmetric float buffer
// (offset and n multiples of 4).  multicast (nullable): the buffer's NVSwitch multicast address; peers: host array of `world`
// device pointers to the ranks' copies (used when multicast is NULL; world <= 16).  The caller brackets the call with cross-rank
// barriers on the same stream.  max_ctas bounds the grid (0 = 2 per SM).
int pag_allreduce_symm(float* multicast, float* const* peers, int rank, int world, int64_t offset, int64_t n, float mult, int max_ctas,
                       void* stream) {
    if (world < 1 || world > 16 || rank < 0 || rank >= world || (offset & 3) || (n & 3) || n < 0) return PAG_ERR_ARG;
    if (n == 0) return PAG_OK;
    const int64_t nq = n / 4, per = (nq + world - 1) / world;
    const int64_t q0 = offset / 4 + rank * per, q1 = offset / 4 + ((rank + 1) * per < nq ? (rank + 1) * per : nq);
    if (q1 <= q0) return PAG_OK;
    int64_t blocks = (q1 - q0 + 255) / 256;
    const int cap = max_ctas > 0 ? max_ctas : 4 * 148;
    const int grid = (int)(blocks < cap ? blocks : cap);
    if (multicast) {
        allreduce_mc_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(multicast, q0, q1, mult);
    } else {
        if (!peers) return PAG_ERR_ARG;
        PeerPtrs pp;
        for (int r = 0; r < world; ++r) pp.p[r] = peers[r];
        cudaStream_t st = (cudaStream_t)stream;
        if (world == 2) allreduce_p2p_kernel<2><<<grid, 256, 0, st>>>(pp, world, q0, q1, mult);
        else if (world == 4) allreduce_p2p_kernel<4><<<grid, 256, 0, st>>>(pp, world, q0, q1, mult);
        else if (world == 8) allreduce_p2p_kernel<8><<<grid, 256, 0, st>>>(pp, world, q0, q1, mult);
        else allreduce_p2p_kernel<0><<<grid, 256, 0, st>>>(pp, world, q0, q1, mult);
    }
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
