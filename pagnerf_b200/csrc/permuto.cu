// Multi-resolution permutohedral-lattice encoding, forward + backward (values and positions).
//
// Replaces permutohedral_encoding.PermutoEncoding as the reference uses it
// (grids/permuto_grid.py:57-62 build, :71 call; pc_nerf/panoptic_delta_nef.py:170,219):
// pos_dim 3, F = 2 features per level, L levels, table f32[L, capacity, 2].
//
// B200 design: one thread per sample walks ALL levels with the 4L float2 gathers of a sample
// issued from one thread (unrolled by 4 levels -> 16 independent 8-byte gathers in flight); the
// whole 50 MB table stays L2 resident (126 MB L2) so the gathers are L2-sector bound, not HBM
// bound; each thread writes its own contiguous 8L-byte output row, or -- in the fused trace -- the fp16
// operand image of the tensor-core decoders (coalesced 16-byte stores, see permuto_fwd_kernel).  Backward re-derives the
// lattice (cheaper than storing 32 B/level/sample) and scatters with red.global.add.v2.f32;
// the coarse levels, where a warp hits a handful of vertices, are first aggregated inside the
// warp (match-any leader reduction) so contention at L2 drops by up to 32x; rows with an exactly-zero
// gradient issue no atomics at all.
// Lattice integers (rem0, rank, key, idx) are bit-exact against oracle/permuto.py.
#include "common.cuh"
#include <string.h>
#include <cuda_fp16.h>

#define PERMUTO_HASH_MUL 2531011u

#define PERMUTO_MAX_LEVELS 64
// per-level parameters into shared memory: lv[2l] = (sf.x, sf.y, sf.z, anneal), lv[2l+1] = (shift.x, shift.y, shift.z, 0)
__device__ __forceinline__ void permuto_stage_levels(float4* lv, int L, const float* __restrict__ sf, const float* __restrict__ sh,
                                                     const float* __restrict__ anneal) {
    for (int l = threadIdx.x; l < L; l += blockDim.x) {
        lv[2 * l] = make_float4(__ldg(sf + 3 * l), __ldg(sf + 3 * l + 1), __ldg(sf + 3 * l + 2), anneal ? __ldg(anneal + l) : 1.f);
        lv[2 * l + 1] = make_float4(__ldg(sh + 3 * l), __ldg(sh + 3 * l + 1), __ldg(sh + 3 * l + 2), 0.f);
    }
    __syncthreads();
}

struct PermutoVertex {
    uint32_t idx[4];
    float bary[4];
    int rank[4];
};

// one level of the lattice for one point; `cap_mask` != 0 means capacity is a power of two
// sfv / shv: the level's scale factors / random shift (xyz).  The encode kernels keep all levels' parameters in shared memory
// (two broadcast LDS.128 per level instead of seven uniform LDG: the uniform loads were 12 % of the kernel's L1 wavefronts).
__device__ __forceinline__ void permuto_lattice(float p0, float p1, float p2, const float4 sfv, const float4 shv, uint32_t cap,
                                                uint32_t cap_mask, PermutoVertex& v) {
    const float cf0 = __fmul_rn(__fadd_rn(p0, shv.x), sfv.x);
    const float cf1 = __fmul_rn(__fadd_rn(p1, shv.y), sfv.y);
    const float cf2 = __fmul_rn(__fadd_rn(p2, shv.z), sfv.z);
    float e[4];
    float sm = 0.f;
    e[3] = __fmaf_rn(-3.f, cf2, sm); sm = __fadd_rn(sm, cf2);
    e[2] = __fmaf_rn(-2.f, cf1, sm); sm = __fadd_rn(sm, cf1);
    e[1] = __fmaf_rn(-1.f, cf0, sm); sm = __fadd_rn(sm, cf0);
    e[0] = sm;
    int rem0[4];
    int sum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float q = e[i] * 0.25f;
        const float up = ceilf(q) * 4.f, down = floorf(q) * 4.f;
        rem0[i] = (__fsub_rn(up, e[i]) < __fsub_rn(e[i], down)) ? (int)up : (int)down;
        sum += rem0[i];
    }
    sum /= 4;
    float dlt[4];
    int rank[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 4; ++i) dlt[i] = __fsub_rn(e[i], (float)rem0[i]);
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = i + 1; j < 4; ++j) {
            if (dlt[i] < dlt[j]) rank[i]++; else rank[j]++;
        }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rank[i] += sum;
        if (rank[i] < 0) { rank[i] += 4; rem0[i] += 4; }
        else if (rank[i] > 3) { rank[i] -= 4; rem0[i] -= 4; }
    }
    // barycentric weights: D[k] = delta of the coordinate whose rank is k (ranks are a permutation)
    float D[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float delta = __fsub_rn(e[i], (float)rem0[i]) * 0.25f;
#pragma unroll
        for (int k = 0; k < 4; ++k) D[k] = (rank[i] == k) ? delta : D[k];
        v.rank[i] = rank[i];
    }
    v.bary[0] = (D[3] + 1.0f) - D[0];
    v.bary[1] = D[2] - D[3];
    v.bary[2] = D[1] - D[2];
    v.bary[3] = D[0] - D[1];
    // hash of vertex r: k = ((key0 * m + key1) * m + key2) * m (mod 2^32), key_i = rem0[i] + r - 4 [rank_i > 3 - r].
    // The hash is linear mod 2^32, so k_r = sum_i rem0[i] m^(3-i) + r (m + m^2 + m^3) - 4 sum_i [rank_i > 3 - r] m^(3-i):
    // three multiplies for the base instead of three per vertex (bit-exact: same wrap-around arithmetic).
    constexpr uint32_t M1 = PERMUTO_HASH_MUL, M2 = M1 * M1, M3 = M2 * M1, MS = M1 + M2 + M3;
    const uint32_t base = (uint32_t)rem0[0] * M3 + (uint32_t)rem0[1] * M2 + (uint32_t)rem0[2] * M1;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
        uint32_t k = base + (uint32_t)r * MS;
        k -= (rank[0] > 3 - r) ? 4u * M3 : 0u;
        k -= (rank[1] > 3 - r) ? 4u * M2 : 0u;
        k -= (rank[2] > 3 - r) ? 4u * M1 : 0u;
        v.idx[r] = cap_mask ? (k & cap_mask) : (k % cap);
    }
}

// IMG16: features go out as the decoders' fp16 operand image instead of f32[M, 2L]: tile t = m / 128 holds
// [2L/8 chunks][128 rows][8 halfs] (12 KB for L = 24), so that a decoder CTA fetches its whole input tile with one bulk copy
// and feeds it to tcgen05.mma unchanged.  Four levels (8 features) are packed per 16-byte store and consecutive lanes write
// consecutive rows: fully coalesced (the f32 row-major layout makes every float2 store its own sector).  Rows between M and
// the end of the last tile are zero-filled (the tensor core reads whole tiles).  Matches the reference's autocast, where the
// encoder output is half (grids/permuto_grid.py:65 custom_fwd(cast_inputs=torch.half)).
template <bool IMG16>
__global__ void __launch_bounds__(128) permuto_fwd_kernel(
    const float* __restrict__ pos, int64_t M, const float* __restrict__ table, uint32_t cap, uint32_t cap_mask, int L,
    const float* __restrict__ sf, const float* __restrict__ sh, const float* __restrict__ anneal,
    float* __restrict__ out, const int64_t* __restrict__ m_dev, int pos_half) {
    __shared__ float4 lv[2 * PERMUTO_MAX_LEVELS];
    permuto_stage_levels(lv, L, sf, sh, anneal);
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m_dev) M = min(M, __ldg(m_dev));   // packed-sample count produced on the device by the marcher
    uint4* img = nullptr;
    if (IMG16) {
        const int64_t Mpad = (M + 127) & ~(int64_t)127;
        if (m >= Mpad) return;
        img = reinterpret_cast<uint4*>(out) + ((m >> 7) * (L >> 2)) * 128 + (m & 127);    // chunk c at img[c * 128]
        if (m >= M) {
            for (int c = 0; c < (L >> 2); ++c) img[c * 128] = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
    } else if (m >= M) return;
    float p0 = pos[3 * m], p1 = pos[3 * m + 1], p2 = pos[3 * m + 2];
    if (pos_half) {   // autocast: custom_fwd(cast_inputs=torch.half) then .float() (grids/permuto_grid.py:65,71)
        p0 = __half2float(__float2half_rn(p0)); p1 = __half2float(__float2half_rn(p1)); p2 = __half2float(__float2half_rn(p2));
    }
    float2* orow = reinterpret_cast<float2*>(out + m * (int64_t)(2 * L));
    const bool row16 = !IMG16 && !(L & 1);   // f32 rows of an even level count are 16-byte aligned: two levels per store (every lane's
    float2 prev = make_float2(0.f, 0.f);     // store is its own L1 wavefront at this 8L-byte stride -- half as many of them)
    uint32_t pk[4];
#pragma unroll 4
    for (int l = 0; l < L; ++l) {
        PermutoVertex v;
        const float4 sfv = lv[2 * l];
        permuto_lattice(p0, p1, p2, sfv, lv[2 * l + 1], cap, cap_mask, v);
        const float* tl = table + (size_t)l * cap * 2;
        const float2 a = ldg2(tl + 2 * (size_t)v.idx[0]);
        const float2 b = ldg2(tl + 2 * (size_t)v.idx[1]);
        const float2 c = ldg2(tl + 2 * (size_t)v.idx[2]);
        const float2 d = ldg2(tl + 2 * (size_t)v.idx[3]);
        const float w = sfv.w;
        float2 acc;
        acc.x = (a.x * v.bary[0] + b.x * v.bary[1] + c.x * v.bary[2] + d.x * v.bary[3]) * w;
        acc.y = (a.y * v.bary[0] + b.y * v.bary[1] + c.y * v.bary[2] + d.y * v.bary[3]) * w;
        if (IMG16) {
            const __half2 h = __floats2half2_rn(acc.x, acc.y);
            pk[l & 3] = *reinterpret_cast<const uint32_t*>(&h);
            if ((l & 3) == 3) img[(l >> 2) * 128] = make_uint4(pk[0], pk[1], pk[2], pk[3]);
        } else if (row16) {
            if (l & 1) reinterpret_cast<float4*>(orow)[l >> 1] = make_float4(prev.x, prev.y, acc.x, acc.y);
            else prev = acc;
        } else {
            orow[l] = acc;
        }
    }
}

// debug / parity: dump lattice integers  idx[L,M,4] u32, rank[L,M,4] i32
__global__ void permuto_indices_kernel(const float* __restrict__ pos, int64_t M, uint32_t cap, uint32_t cap_mask, int L,
                                       const float* __restrict__ sf, const float* __restrict__ sh,
                                       uint32_t* __restrict__ idx, int* __restrict__ rank, float* __restrict__ bary) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float p0 = pos[3 * m], p1 = pos[3 * m + 1], p2 = pos[3 * m + 2];
    for (int l = 0; l < L; ++l) {
        PermutoVertex v;
        permuto_lattice(p0, p1, p2, make_float4(__ldg(sf + 3 * l), __ldg(sf + 3 * l + 1), __ldg(sf + 3 * l + 2), 1.f),
                        make_float4(__ldg(sh + 3 * l), __ldg(sh + 3 * l + 1), __ldg(sh + 3 * l + 2), 0.f), cap, cap_mask, v);
        for (int r = 0; r < 4; ++r) {
            idx[((size_t)l * M + m) * 4 + r] = v.idx[r];
            rank[((size_t)l * M + m) * 4 + r] = v.rank[r];
            bary[((size_t)l * M + m) * 4 + r] = v.bary[r];
        }
    }
}

// GIMG: the upstream gradient arrives as the decoders' fp16 tile image (see permuto_fwd_kernel), still multiplied by the
// power-of-two loss scale the tensor-core backward used; *img_scale is that scale (coalesced 16-byte reads, half the bytes).
template <bool POS_GRAD, bool GIMG>
__global__ void __launch_bounds__(128) permuto_bwd_kernel(
    const float* __restrict__ pos, int64_t M, const float* __restrict__ table, uint32_t cap, uint32_t cap_mask, int L,
    const float* __restrict__ sf, const float* __restrict__ sh, const float* __restrict__ anneal,
    const float* __restrict__ gout, float* __restrict__ gtable, float* __restrict__ gpos, int n_agg_levels,
    const int64_t* __restrict__ m_dev, int pos_half, const float* __restrict__ img_scale, int l0, int l1) {
    __shared__ float4 lv[2 * PERMUTO_MAX_LEVELS];
    permuto_stage_levels(lv, L, sf, sh, anneal);
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m_dev) M = min(M, __ldg(m_dev));
    if (M <= 0 || (m & ~31ll) >= M) return;   // whole warp beyond the packed samples
    const bool valid = m < M;
    // aggregated levels need the full warp converged: clamp instead of returning early
    const int64_t mm = valid ? m : (M - 1);
    float p0 = pos[3 * mm], p1 = pos[3 * mm + 1], p2 = pos[3 * mm + 2];
    if (pos_half) {
        p0 = __half2float(__float2half_rn(p0)); p1 = __half2float(__float2half_rn(p1)); p2 = __half2float(__float2half_rn(p2));
    }
    const float2* grow = reinterpret_cast<const float2*>(gout + mm * (int64_t)(2 * L));
    const uint4* gimg = reinterpret_cast<const uint4*>(gout) + ((mm >> 7) * (L >> 2)) * 128 + (mm & 127);   // chunk c at gimg[c * 128]
    const float inv_scale = (GIMG && img_scale) ? 1.f / __ldg(img_scale) : 1.f;
    // Samples the integrator gave zero weight (sigma clamped to 0, or behind an opaque surface with an underflowed
    // transmittance) arrive with an exactly-zero gradient row: adding zeros is skipped -- a whole warp of them costs one
    // pass over its rows, a single one costs no atomics.
    bool rownz = false;
    if (valid) {
        if (GIMG) {
            for (int c = 0; c < (L >> 2); ++c) { const uint4 u = __ldg(gimg + c * 128); rownz |= ((u.x | u.y | u.z | u.w) & 0x7fff7fffu) != 0u; }
        } else {
            for (int l = 0; l < L; ++l) { const float2 g = __ldg(grow + l); rownz |= (g.x != 0.f) | (g.y != 0.f); }
        }
    }
    // [l0, l1): the levels of this launch (the caller may split the table into level ranges to pipeline the scatter with
    // the all-reduce of the finished ranges); position gradients accumulate over the ranges (first range writes)
    if (!__any_sync(0xffffffffu, rownz)) {
        if (POS_GRAD && valid && l0 == 0) { gpos[3 * m] = 0.f; gpos[3 * m + 1] = 0.f; gpos[3 * m + 2] = 0.f; }
        return;
    }
    float gp0 = 0.f, gp1 = 0.f, gp2 = 0.f;
    uint4 gq = make_uint4(0u, 0u, 0u, 0u);
    for (int l = l0; l < l1; ++l) {
        PermutoVertex v;
        const float4 sfv = lv[2 * l];
        permuto_lattice(p0, p1, p2, sfv, lv[2 * l + 1], cap, cap_mask, v);
        const float w = sfv.w;
        float2 g;
        if (GIMG) {
            if ((l & 3) == 0 || l == l0) gq = __ldg(gimg + (l >> 2) * 128);
            const uint32_t u = (l & 3) == 0 ? gq.x : ((l & 3) == 1 ? gq.y : ((l & 3) == 2 ? gq.z : gq.w));
            g = __half22float2(*reinterpret_cast<const __half2*>(&u));
            g.x *= inv_scale; g.y *= inv_scale;
        } else {
            g = __ldg(grow + l);
        }
        g.x = valid ? g.x * w : 0.f;
        g.y = valid ? g.y * w : 0.f;
        const bool nz = rownz && ((g.x != 0.f) | (g.y != 0.f));
        float* gl = gtable + (size_t)l * cap * 2;
        if (l < n_agg_levels) {
            const unsigned todo0 = __ballot_sync(0xffffffffu, nz);
#pragma unroll
            for (int r = 0; r < 4; ++r)
                scatter_aggregated(gl, v.idx[r], g.x * v.bary[r], g.y * v.bary[r], 0xffffffffu, todo0);
        } else if (nz) {
#pragma unroll
            for (int r = 0; r < 4; ++r) red_add_f32x2(gl + 2 * (size_t)v.idx[r], g.x * v.bary[r], g.y * v.bary[r]);
        }
        if (POS_GRAD) {
            const float* tl = table + (size_t)l * cap * 2;
            float s[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const float2 t = ldg2(tl + 2 * (size_t)v.idx[r]);
                s[r] = g.x * t.x + g.y * t.y;
            }
            // dL/dD[k]; D[k] = delta of the coordinate with rank k
            const float dD[4] = {s[3] - s[0], s[2] - s[3], s[1] - s[2], s[0] - s[1]};
            float de[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float t = 0.f;
#pragma unroll
                for (int k = 0; k < 4; ++k) t = (v.rank[i] == k) ? dD[k] : t;
                de[i] = 0.25f * t;
            }
            // e0 = c0+c1+c2, e1 = c2+c1-c0, e2 = c2-2c1, e3 = -3c2
            gp0 += (de[0] - de[1]) * sfv.x;
            gp1 += (de[0] + de[1] - 2.f * de[2]) * sfv.y;
            gp2 += (de[0] + de[1] + de[2] - 3.f * de[3]) * sfv.z;
        }
    }
    if (POS_GRAD && valid) {
        if (l0 == 0) { gpos[3 * m] = gp0; gpos[3 * m + 1] = gp1; gpos[3 * m + 2] = gp2; }
        else { gpos[3 * m] += gp0; gpos[3 * m + 1] += gp1; gpos[3 * m + 2] += gp2; }
    }
}

// uniformly random 8-byte gathers (see pag_gather_probe): 16 independent loads in flight per thread per iteration -- what the
// encoder keeps in flight with four lattice levels unrolled
__global__ void __launch_bounds__(128) gather_probe_kernel(const float* __restrict__ table, uint32_t entries, int64_t threads,
                                                           int loads_per_thread, float* __restrict__ sink) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= threads) return;
    uint32_t state = (uint32_t)t * 2654435761u + 12345u;
    const bool pow2 = (entries & (entries - 1)) == 0;
    float acc = 0.f;
    for (int i = 0; i < loads_per_thread; i += 16) {
        float2 v[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            const uint32_t h = pag_lowbias32(state + k);
            const uint32_t idx = pow2 ? (h & (entries - 1)) : (h % entries);
            v[k] = ldg2(table + 2 * (size_t)idx);
        }
        state += 16;
#pragma unroll
        for (int k = 0; k < 16; ++k) acc += v[k].x + v[k].y;
    }
    sink[t] = acc;
}

// L2 read-bandwidth probe: every CTA streams the whole buffer `iters` times with coalesced 16-byte loads, each CTA starting at
// its own offset so that the L2 slices are loaded evenly; the buffer (<= 64 MB) stays resident in the 126 MB L2 after the first pass
__global__ void __launch_bounds__(256) l2_stream_probe_kernel(const uint4* __restrict__ buf, int64_t n16, int iters, float* __restrict__ sink) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    uint32_t acc = 0;
    for (int it = 0; it < iters; ++it) {
        const int64_t start = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x + (int64_t)it * 4099 * blockDim.x) % n16;
        int64_t i = start;
        for (int64_t k = 0; k < n16; k += stride * 4) {
            uint4 v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int64_t j = i + u * stride;
                if (j >= n16) j -= n16;
                if (j >= n16) j %= n16;
                asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w) : "l"(buf + j));
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) acc ^= v[u].x ^ v[u].y ^ v[u].z ^ v[u].w;
            i += 4 * stride;
            if (i >= n16) i %= n16;
        }
    }
    sink[(int64_t)blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc & 0x007fffffu);
}

extern "C" {

// forward: out[M, 2L] (level-major, feature-minor)
int pag_permuto_fwd(const float* pos, int64_t M, const float* table, int64_t capacity, int L, int F,
                    const float* scale_factor, const float* shift, const float* anneal, float* out, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (capacity <= 0 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    if (cap == 1) return PAG_ERR_ARG;
    permuto_fwd_kernel<false><<<pag_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(pos, M, table, cap, mask, L, scale_factor,
                                                                          shift, anneal, out, nullptr, 0);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// same, with the sample count read on the device (m_dev[0] <= M_max, no host sync) and optional fp16 rounding of pos
int pag_permuto_fwd_dyn(const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table,
                        int64_t capacity, int L, int F, const float* scale_factor, const float* shift, const float* anneal,
                        float* out, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (capacity <= 1 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M_max == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    permuto_fwd_kernel<false><<<pag_grid(M_max, 128), 128, 0, (cudaStream_t)stream>>>(pos, M_max, table, cap, mask, L, scale_factor,
                                                                                     shift, anneal, out, m_dev, pos_half);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// fp16 operand-image output (see permuto_fwd_kernel<true>): img16 holds ceil(M_max / 128) tiles of 2L/8 * 2048 bytes; L % 4 == 0
int pag_permuto_fwd_img16_dyn(const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table,
                              int64_t capacity, int L, int F, const float* scale_factor, const float* shift, const float* anneal,
                              void* img16, void* stream) {
    if (F != 2 || (L & 3)) return PAG_ERR_UNSUPPORTED;
    if (capacity <= 1 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M_max == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    const int64_t Mpad = (M_max + 127) & ~(int64_t)127;
    permuto_fwd_kernel<true><<<pag_grid(Mpad, 128), 128, 0, (cudaStream_t)stream>>>(
        pos, M_max, table, cap, mask, L, scale_factor, shift, anneal, reinterpret_cast<float*>(img16), m_dev, pos_half);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// backward: gtable[L,cap,2] is ACCUMULATED into (caller zeroes it); gpos[M,3] written iff non-null.
int pag_permuto_bwd(const float* pos, int64_t M, const float* table, int64_t capacity, int L, int F,
                    const float* scale_factor, const float* shift, const float* anneal, const float* grad_out,
                    float* grad_table, float* grad_pos, int n_agg_levels, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (capacity <= 1 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    if (grad_pos)
        permuto_bwd_kernel<true, false><<<pag_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(
            pos, M, table, cap, mask, L, scale_factor, shift, anneal, grad_out, grad_table, grad_pos, n_agg_levels, nullptr, 0, nullptr, 0, L);
    else
        permuto_bwd_kernel<false, false><<<pag_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(
            pos, M, table, cap, mask, L, scale_factor, shift, anneal, grad_out, grad_table, grad_pos, n_agg_levels, nullptr, 0, nullptr, 0, L);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_permuto_bwd_dyn(const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table,
                        int64_t capacity, int L, int F, const float* scale_factor, const float* shift, const float* anneal,
                        const float* grad_out, float* grad_table, float* grad_pos, int n_agg_levels, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (capacity <= 1 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M_max == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    if (grad_pos)
        permuto_bwd_kernel<true, false><<<pag_grid(M_max, 128), 128, 0, (cudaStream_t)stream>>>(
            pos, M_max, table, cap, mask, L, scale_factor, shift, anneal, grad_out, grad_table, grad_pos, n_agg_levels, m_dev, pos_half, nullptr, 0, L);
    else
        permuto_bwd_kernel<false, false><<<pag_grid(M_max, 128), 128, 0, (cudaStream_t)stream>>>(
            pos, M_max, table, cap, mask, L, scale_factor, shift, anneal, grad_out, grad_table, grad_pos, n_agg_levels, m_dev, pos_half, nullptr, 0, L);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// upstream gradient as the decoders' fp16 tile image, scaled by *img_scale (device float, nullable = 1): see permuto_bwd_kernel
int pag_permuto_bwd_img16_dyn(const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table,
                              int64_t capacity, int L, int F, const float* scale_factor, const float* shift, const float* anneal,
                              const void* grad_img16, const float* img_scale, float* grad_table, float* grad_pos, int n_agg_levels,
                              int level_begin, int level_end, void* stream) {
    if (F != 2 || (L & 3)) return PAG_ERR_UNSUPPORTED;
    if (level_begin < 0 || level_end > L || level_begin >= level_end) return PAG_ERR_ARG;
    if (capacity <= 1 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M_max == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    const float* g = reinterpret_cast<const float*>(grad_img16);
    if (grad_pos)
        permuto_bwd_kernel<true, true><<<pag_grid(M_max, 128), 128, 0, (cudaStream_t)stream>>>(
            pos, M_max, table, cap, mask, L, scale_factor, shift, anneal, g, grad_table, grad_pos, n_agg_levels, m_dev, pos_half, img_scale, level_begin, level_end);
    else
        permuto_bwd_kernel<false, true><<<pag_grid(M_max, 128), 128, 0, (cudaStream_t)stream>>>(
            pos, M_max, table, cap, mask, L, scale_factor, shift, anneal, g, grad_table, grad_pos, n_agg_levels, m_dev, pos_half, img_scale, level_begin, level_end);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_permuto_indices(const float* pos, int64_t M, int64_t capacity, int L, const float* scale_factor,
                        const float* shift, uint32_t* idx, int32_t* rank, float* bary, void* stream) {
    if (capacity <= 1 || capacity > 0xFFFFFFFFll || L <= 0) return PAG_ERR_ARG;
    if (L > PERMUTO_MAX_LEVELS) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    const uint32_t cap = (uint32_t)capacity;
    const uint32_t mask = ((cap & (cap - 1)) == 0) ? (cap - 1) : 0;
    permuto_indices_kernel<<<pag_grid(M, 128), 128, 0, (cudaStream_t)stream>>>(pos, M, cap, mask, L, scale_factor, shift,
                                                                              idx, rank, bary);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// ---- achievable-gather-bandwidth probe (the denominator of the encoder's roofline fraction) ----------------
// `threads` threads each issue `loads_per_thread` (multiple of 16) uniformly random 8-byte loads (16 independent loads in
// flight per iteration, like the 4 simplex vertices of four unrolled lattice levels) from a table of `entries` float2 -- the same access shape as
// the encoder's vertex reads without any of its arithmetic.  sink receives one float per thread so the loads stay live.
int pag_gather_probe(const float* table, int64_t entries, int64_t threads, int loads_per_thread, float* sink, void* stream) {
    if (entries <= 0 || entries > 0xFFFFFFFFll || threads <= 0 || loads_per_thread <= 0 || (loads_per_thread & 15)) return PAG_ERR_ARG;
    gather_probe_kernel<<<pag_grid(threads, 128), 128, 0, (cudaStream_t)stream>>>(table, (uint32_t)entries, threads, loads_per_thread, sink);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// ---- L2 persisting window (north_star: "per-level tables ... in L2-persisting windows") ------------------------------------
// Marks [base, base + bytes) as persisting for every kernel launched on `stream` afterwards (captured into CUDA-graph kernel
// nodes as well); bytes == 0 clears the window.  The persisting carve-out of the L2 is raised to the device maximum on first
// use.  hit_ratio: fraction of the window's lines that get the persisting property (set < 1 when the window is larger than the
// carve-out, so that the persisting lines do not thrash each other).
int pag_set_l2_window(const void* base, int64_t bytes, float hit_ratio, void* stream) {
    static bool carved = false;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return (int)e;
    int max_persist = 0, max_window = 0;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persist <= 0 || max_window <= 0) return PAG_ERR_UNSUPPORTED;
    if (!carved) {
        e = cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)max_persist);
        if (e != cudaSuccess) return (int)e;
        carved = true;
    }
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    if (bytes > 0) {
        v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        v.accessPolicyWindow.num_bytes = (size_t)(bytes < (int64_t)max_window ? bytes : (int64_t)max_window);
        v.accessPolicyWindow.hitRatio = hit_ratio;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    } else {
        v.accessPolicyWindow.num_bytes = 0;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyNormal;
        v.accessPolicyWindow.missProp = cudaAccessPropertyNormal;
    }
    e = cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &v);
    return e == cudaSuccess ? PAG_OK : (int)e;
}
// max persisting carve-out / max window size of the current device (bytes)
int pag_l2_limits(int64_t* max_persisting, int64_t* max_window) {
    int dev = 0, a = 0, b = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return PAG_ERR_ARG;
    cudaDeviceGetAttribute(&a, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&b, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    if (max_persisting) *max_persisting = a;
    if (max_window) *max_window = b;
    return PAG_OK;
}

// ---- L2 read-bandwidth probe (the denominator of the encoders' "l2" roofline) -----------------------------------------------
// All CTAs together read `bytes` (multiple of 16, <= L2 size) `iters` times: bytes * iters / time = L2 -> SM read bandwidth once
// the buffer is L2 resident.  sink: 148 * 8 * 256 floats.
int pag_l2_stream_probe(const void* buf, int64_t bytes, int iters, float* sink, void* stream) {
    if (bytes <= 0 || (bytes & 15) || iters <= 0) return PAG_ERR_ARG;
    l2_stream_probe_kernel<<<148 * 8, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint4*>(buf), bytes / 16, iters, sink);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
