// Occupancy-octree query, ray marching ('ray' and 'voxel') and packed-ray bookkeeping.
//
// Replaces (behind the same grid.raymarch() surface, reference grids/occtree.py:85-91):
//   kaolin.ops.spc.unbatched_query, kaolin.render.spc.unbatched_raytrace,
//   wisp OctreeAS.raymarch (sample_from_depth_intervals / linspace+rand+mask compaction),
//   kaolin.render.spc.mark_pack_boundaries.
// Design (B200): one warp per ray for 'ray' mode (lanes = steps, ballot compaction), one thread
// per ray DFS for 'voxel' mode; both are count -> scan -> emit so that no per-level host sync
// exists (upstream syncs once per octree level).  The octree bytes + prefix (<= 1.5 MB at
// level 7) stay L1/L2 resident; outputs are written once, packed.
// Integer outputs are bit-exact against oracle/spc.py + oracle/raymarch.py.
#include "common.cuh"

#define PAG_MAX_LEVEL 10

// ---------------------------------------------------------------------------------------------
// point query
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int octree_query_point(const uint8_t* __restrict__ octree,
                                                  const int* __restrict__ prefix,
                                                  float x, float y, float z, int level) {
    const float res = (float)(1 << level);
    const float qx = floorf(__fmul_rn(res, __fadd_rn(__fmul_rn(x, 0.5f), 0.5f)));
    const float qy = floorf(__fmul_rn(res, __fadd_rn(__fmul_rn(y, 0.5f), 0.5f)));
    const float qz = floorf(__fmul_rn(res, __fadd_rn(__fmul_rn(z, 0.5f), 0.5f)));
    if (!(qx >= 0.f && qx < res && qy >= 0.f && qy < res && qz >= 0.f && qz < res)) return -1;
    if (isinf(x) || isinf(y) || isinf(z)) return -1;
    const int ix = (int)qx, iy = (int)qy, iz = (int)qz;
    int node = 0;
    for (int l = 0; l < level; ++l) {
        const int s = level - 1 - l;
        const uint32_t j = (((ix >> s) & 1) << 2) | (((iy >> s) & 1) << 1) | ((iz >> s) & 1);
        const uint32_t byte = __ldg(octree + node);
        if (!((byte >> j) & 1u)) return -1;
        node = __ldg(prefix + node) + __popc(byte & ((2u << j) - 1u));
    }
    return node;
}

__global__ void octree_query_kernel(const uint8_t* __restrict__ octree, const int* __restrict__ prefix,
                                    const float* __restrict__ coords, int64_t P, int level,
                                    int* __restrict__ pidx) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    pidx[i] = octree_query_point(octree, prefix, coords[3 * i], coords[3 * i + 1], coords[3 * i + 2], level);
}

// ---------------------------------------------------------------------------------------------
// single-CTA exclusive scan (N <= ~1M per-ray counts); writes total to out[N]
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) exclusive_scan_kernel(const int* __restrict__ in, int64_t N,
                                                               int64_t* __restrict__ out) {
    __shared__ int64_t warp_sums[32];
    __shared__ int64_t carry_s;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int64_t per = (N + 1023) / 1024;
    const int64_t lo = min((int64_t)tid * per, N), hi = min(lo + per, N);
    // up to 16 elements per thread live in registers: the loads are independent (one memory latency instead of `per`)
    constexpr int REG = 16;
    int v[REG];
    const bool in_regs = per <= REG;
    int64_t s = 0;
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < REG; ++k) v[k] = (lo + k < hi) ? __ldg(in + lo + k) : 0;
#pragma unroll
        for (int k = 0; k < REG; ++k) s += v[k];
    } else {
        for (int64_t i = lo; i < hi; ++i) s += in[i];
    }
    int64_t incl = s;
    for (int o = 1; o < 32; o <<= 1) {
        int64_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    if (wid == 0) {
        int64_t w = warp_sums[lane];
        int64_t wi = w;
        for (int o = 1; o < 32; o <<= 1) {
            int64_t t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        warp_sums[lane] = wi - w;  // exclusive warp offsets
        if (lane == 31) carry_s = wi;
    }
    __syncthreads();
    int64_t run = warp_sums[wid] + incl - s;
    if (in_regs) {
#pragma unroll
        for (int k = 0; k < REG; ++k)
            if (lo + k < hi) { out[lo + k] = run; run += v[k]; }
    } else {
        for (int64_t i = lo; i < hi; ++i) { out[i] = run; run += in[i]; }
    }
    if (tid == 0) out[N] = carry_s;
}

// ---------------------------------------------------------------------------------------------
// live-sample compaction of a packed sample list (training hot path)
// A sample whose density came out exactly 0 (ReLU-clamped) has integration weight T * (1 - exp(-0 * delta)) = 0 and
// leaves the transmittance unchanged: it contributes an exact 0 to every composited output and receives an exact 0
// gradient (ReLU' = 0, everything else is scaled by the weight).  Dropping those samples from the packed list before
// the colour / panoptic decoders, the delta-grid encode and the whole backward changes no result bit.
// Per ray (one warp): count the live samples, scan the counts over rays, then copy the survivors in order.
// ---------------------------------------------------------------------------------------------
__global__ void compact_count_kernel(const float* __restrict__ sigma, const int64_t* __restrict__ offsets, int64_t N,
                                     int32_t* __restrict__ counts) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const int64_t lo = offsets[r], hi = offsets[r + 1];
    int cnt = 0;
    for (int64_t s0 = lo; s0 < hi; s0 += 32) {
        const int64_t s = s0 + lane;
        const bool live = s < hi && __ldg(sigma + s) > 0.f;
        cnt += __popc(__ballot_sync(0xffffffffu, live));
    }
    if (lane == 0) counts[r] = cnt;
}
__global__ void compact_emit_kernel(const float* __restrict__ sigma, const int64_t* __restrict__ offsets,
                                    const int64_t* __restrict__ offsets_c, int64_t N, const float* __restrict__ samples,
                                    const float* __restrict__ depths, const float* __restrict__ deltas,
                                    const float* __restrict__ feats, int F, int64_t* __restrict__ ridx_c,
                                    float* __restrict__ samples_c, float* __restrict__ depths_c, float* __restrict__ deltas_c,
                                    float* __restrict__ feats_c) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= N) return;
    const int64_t lo = offsets[r], hi = offsets[r + 1];
    int64_t base = offsets_c[r];
    const int nq = F >> 2;                       // float4 quads per feature row (F % 4 == 0)
    for (int64_t s0 = lo; s0 < hi; s0 += 32) {
        const int64_t s = s0 + lane;
        const bool live = s < hi && __ldg(sigma + s) > 0.f;
        const unsigned b = __ballot_sync(0xffffffffu, live);
        if (live) {
            const int64_t dst = base + __popc(b & ((1u << lane) - 1u));
            ridx_c[dst] = r;
            samples_c[3 * dst] = samples[3 * s]; samples_c[3 * dst + 1] = samples[3 * s + 1]; samples_c[3 * dst + 2] = samples[3 * s + 2];
            depths_c[dst] = depths[s];
            deltas_c[dst] = deltas[s];
        }
        // feature rows: the warp copies one surviving row per step (two when a row needs <= 16 lanes), coalesced
        unsigned todo = b;
        int k = 0;
        if (nq <= 16) {
            const int half = lane >> 4, ql = lane & 15;
            while (todo) {
                const int b0 = __ffs(todo) - 1;
                todo &= todo - 1;
                int b1 = -1;
                if (todo) { b1 = __ffs(todo) - 1; todo &= todo - 1; }
                const int bit = half ? b1 : b0;
                if (bit >= 0 && ql < nq)
                    reinterpret_cast<float4*>(feats_c + (base + k + half) * F)[ql] = __ldg(reinterpret_cast<const float4*>(feats + (s0 + bit) * F) + ql);
                k += 2;
            }
        } else {
            while (todo) {
                const int bit = __ffs(todo) - 1;
                todo &= todo - 1;
                for (int ql = lane; ql < nq; ql += 32)
                    reinterpret_cast<float4*>(feats_c + (base + k) * F)[ql] = __ldg(reinterpret_cast<const float4*>(feats + (s0 + bit) * F) + ql);
                ++k;
            }
        }
        base += __popc(b);
    }
}

// ---------------------------------------------------------------------------------------------
// 'ray' mode
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float ray_step_depth(int64_t ray, int i, int S, const float* __restrict__ lin,
                                                const float* __restrict__ jitter, uint32_t seed,
                                                float near, float range) {
    const uint64_t flat = (uint64_t)ray * (uint64_t)S + (uint64_t)i;
    const float u = jitter ? __ldg(jitter + flat) : pag_jitter(seed, flat);
    float t = __fadd_rn(__ldg(lin + i), __fdiv_rn(u, (float)S));
    t = __fmul_rn(t, range);
    return __fadd_rn(t, near);
}

__global__ void march_ray_count_kernel(const float* __restrict__ org, const float* __restrict__ dir, int64_t N,
                                       int S, const float* __restrict__ lin, const float* __restrict__ jitter,
                                       uint32_t seed, float near, float range,
                                       const uint8_t* __restrict__ octree, const int* __restrict__ prefix,
                                       int level, int* __restrict__ pidx_tmp, int* __restrict__ counts,
                                       const uint32_t* __restrict__ seed_dev) {
    if (seed_dev) seed = __ldg(seed_dev);   // CUDA-graph replays advance the jitter stream through device memory
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= N) return;
    const float ox = org[3 * ray], oy = org[3 * ray + 1], oz = org[3 * ray + 2];
    const float dx = dir[3 * ray], dy = dir[3 * ray + 1], dz = dir[3 * ray + 2];
    int cnt = 0;
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        int p = -1;
        if (i < S) {
            const float t = ray_step_depth(ray, i, S, lin, jitter, seed, near, range);
            const float x = __fadd_rn(ox, __fmul_rn(dx, t));
            const float y = __fadd_rn(oy, __fmul_rn(dy, t));
            const float z = __fadd_rn(oz, __fmul_rn(dz, t));
            p = octree_query_point(octree, prefix, x, y, z, level);
            pidx_tmp[ray * S + i] = p;
        }
        cnt += __popc(__ballot_sync(0xffffffffu, p >= 0));
    }
    if (lane == 0) counts[ray] = cnt;
}

__global__ void march_ray_emit_kernel(const float* __restrict__ org, const float* __restrict__ dir, int64_t N,
                                      int S, const float* __restrict__ lin, const float* __restrict__ jitter,
                                      uint32_t seed, float near, float range,
                                      const int* __restrict__ pidx_tmp, const int64_t* __restrict__ offsets,
                                      int64_t* __restrict__ ridx, int64_t* __restrict__ pidx,
                                      float* __restrict__ samples, float* __restrict__ depths,
                                      float* __restrict__ deltas, uint8_t* __restrict__ boundary,
                                      const uint32_t* __restrict__ seed_dev) {
    if (seed_dev) seed = __ldg(seed_dev);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= N) return;
    const int64_t first = offsets[ray];
    if (offsets[ray + 1] == first) return;
    const float ox = org[3 * ray], oy = org[3 * ray + 1], oz = org[3 * ray + 2];
    const float dx = dir[3 * ray], dy = dir[3 * ray + 1], dz = dir[3 * ray + 2];
    int64_t off = first;
    for (int base = 0; base < S; base += 32) {
        const int i = base + lane;
        const int p = (i < S) ? pidx_tmp[ray * S + i] : -1;
        const unsigned bal = __ballot_sync(0xffffffffu, p >= 0);
        if (p >= 0) {
            const int64_t dst = off + __popc(bal & ((1u << lane) - 1u));
            const float t = ray_step_depth(ray, i, S, lin, jitter, seed, near, range);
            const float tp = (i == 0) ? near : ray_step_depth(ray, i - 1, S, lin, jitter, seed, near, range);
            ridx[dst] = ray;
            if (pidx) pidx[dst] = p;
            samples[3 * dst + 0] = __fadd_rn(ox, __fmul_rn(dx, t));
            samples[3 * dst + 1] = __fadd_rn(oy, __fmul_rn(dy, t));
            samples[3 * dst + 2] = __fadd_rn(oz, __fmul_rn(dz, t));
            depths[dst] = t;
            deltas[dst] = __fsub_rn(t, tp);
            if (boundary) boundary[dst] = (dst == first) ? 1 : 0;
        }
        off += __popc(bal);
    }
}

// ---------------------------------------------------------------------------------------------
// 'ray' mode without point indices (training hot path): occupancy bit field of the marching level
// The fused trace never uses pidx; it only needs to know WHICH of the S steps of a ray fall into an occupied cell.  The
// level-7 occupancy is 2 M cells = 256 KB as a bit field (cell (ix,iy,iz) -> bit (ix*res + iy)*res + iz): one cached word
// per step instead of a 7-level descent through octree bytes + prefix sums, and the count pass hands 128-bit step masks
// (16 B per ray) to the emit pass instead of N*S point indices (8 MB each way).  The kept set is the same by construction:
// a cell's bit is the result of octree_query_point on that cell, and the step -> cell arithmetic is shared.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t cell_of_point(float x, float y, float z, int level) {
    const float res = (float)(1 << level);
    const float qx = floorf(__fmul_rn(res, __fadd_rn(__fmul_rn(x, 0.5f), 0.5f)));
    const float qy = floorf(__fmul_rn(res, __fadd_rn(__fmul_rn(y, 0.5f), 0.5f)));
    const float qz = floorf(__fmul_rn(res, __fadd_rn(__fmul_rn(z, 0.5f), 0.5f)));
    if (!(qx >= 0.f && qx < res && qy >= 0.f && qy < res && qz >= 0.f && qz < res)) return -1;
    if (isinf(x) || isinf(y) || isinf(z)) return -1;
    return (((int64_t)qx << level) | (int64_t)qy) << level | (int64_t)qz;
}
__global__ void octree_level_bits_kernel(const uint8_t* __restrict__ octree, const int* __restrict__ prefix, int level,
                                         int64_t ncells, uint32_t* __restrict__ bits) {
    const int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;      // ncells is a multiple of 32 for level >= 2
    bool occ = false;
    if (c < ncells) {
        const int m = (1 << level) - 1;
        const int ix = (int)(c >> (2 * level)) & m, iy = (int)(c >> level) & m, iz = (int)c & m;
        int node = 0;
        occ = true;
        for (int l = 0; l < level && occ; ++l) {
            const int sft = level - 1 - l;
            const uint32_t j = (((ix >> sft) & 1) << 2) | (((iy >> sft) & 1) << 1) | ((iz >> sft) & 1);
            const uint32_t byte = __ldg(octree + node);
            if (!((byte >> j) & 1u)) occ = false;
            else node = __ldg(prefix + node) + __popc(byte & ((2u << j) - 1u));
        }
    }
    const unsigned b = __ballot_sync(0xffffffffu, occ);
    if ((threadIdx.x & 31) == 0 && c < ncells) bits[c >> 5] = b;
}
__global__ void march_ray_bits_count_kernel(const float* __restrict__ org, const float* __restrict__ dir, int64_t N, int S,
                                            const float* __restrict__ lin, const float* __restrict__ jitter, uint32_t seed,
                                            float near, float range, const uint32_t* __restrict__ bits, int level,
                                            uint32_t* __restrict__ masks, int* __restrict__ counts,
                                            const uint32_t* __restrict__ seed_dev) {
    if (seed_dev) seed = __ldg(seed_dev);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= N) return;
    const float ox = org[3 * ray], oy = org[3 * ray + 1], oz = org[3 * ray + 2];
    const float dx = dir[3 * ray], dy = dir[3 * ray + 1], dz = dir[3 * ray + 2];
    const int nw = (S + 31) >> 5;
    int cnt = 0;
    for (int wi = 0; wi < nw; ++wi) {
        const int i = 32 * wi + lane;
        bool keep = false;
        if (i < S) {
            const float t = ray_step_depth(ray, i, S, lin, jitter, seed, near, range);
            const int64_t c = cell_of_point(__fadd_rn(ox, __fmul_rn(dx, t)), __fadd_rn(oy, __fmul_rn(dy, t)),
                                            __fadd_rn(oz, __fmul_rn(dz, t)), level);
            keep = c >= 0 && ((__ldg(bits + (c >> 5)) >> (c & 31)) & 1u);
        }
        const unsigned b = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) masks[ray * nw + wi] = b;
        cnt += __popc(b);
    }
    if (lane == 0) counts[ray] = cnt;
}
__global__ void march_ray_bits_emit_kernel(const float* __restrict__ org, const float* __restrict__ dir, int64_t N, int S,
                                           const float* __restrict__ lin, const float* __restrict__ jitter, uint32_t seed,
                                           float near, float range, const uint32_t* __restrict__ masks,
                                           const int64_t* __restrict__ offsets, int64_t* __restrict__ ridx,
                                           float* __restrict__ samples, float* __restrict__ depths, float* __restrict__ deltas,
                                           const uint32_t* __restrict__ seed_dev) {
    if (seed_dev) seed = __ldg(seed_dev);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= N) return;
    int64_t off = offsets[ray];
    if (offsets[ray + 1] == off) return;
    const float ox = org[3 * ray], oy = org[3 * ray + 1], oz = org[3 * ray + 2];
    const float dx = dir[3 * ray], dy = dir[3 * ray + 1], dz = dir[3 * ray + 2];
    const int nw = (S + 31) >> 5;
    for (int wi = 0; wi < nw; ++wi) {
        const unsigned b = __ldg(masks + ray * nw + wi);
        if ((b >> lane) & 1u) {
            const int i = 32 * wi + lane;
            const int64_t dst = off + __popc(b & ((1u << lane) - 1u));
            const float t = ray_step_depth(ray, i, S, lin, jitter, seed, near, range);
            const float tp = (i == 0) ? near : ray_step_depth(ray, i - 1, S, lin, jitter, seed, near, range);
            ridx[dst] = ray;
            samples[3 * dst + 0] = __fadd_rn(ox, __fmul_rn(dx, t));
            samples[3 * dst + 1] = __fadd_rn(oy, __fmul_rn(dy, t));
            samples[3 * dst + 2] = __fadd_rn(oz, __fmul_rn(dz, t));
            depths[dst] = t;
            deltas[dst] = __fsub_rn(t, tp);
        }
        off += __popc(b);
    }
}

// ---------------------------------------------------------------------------------------------
// 'voxel' mode: per-ray DFS in kaolin's nugget order (children j = code ^ i)
// ---------------------------------------------------------------------------------------------
struct RayCtx { float ox, oy, oz, dx, dy, dz, ix, iy, iz, sx, sy, sz; };

__device__ __forceinline__ float ray_aabb(const RayCtx& r, float vcx, float vcy, float vcz, float rad, float* exit_t) {
    const float ox = __fsub_rn(r.ox, vcx), oy = __fsub_rn(r.oy, vcy), oz = __fsub_rn(r.oz, vcz);
    const float cmax = fmaxf(fmaxf(fabsf(ox), fabsf(oy)), fabsf(oz));
    const float d0 = __fmul_rn(__fmaf_rn(rad, r.sx, -ox), r.ix);
    const float d1 = __fmul_rn(__fmaf_rn(rad, r.sy, -oy), r.iy);
    const float d2 = __fmul_rn(__fmaf_rn(rad, r.sz, -oz), r.iz);
    const bool t0 = (d0 >= 0.f) && (fabsf(__fmaf_rn(r.dy, d0, oy)) < rad) && (fabsf(__fmaf_rn(r.dz, d0, oz)) < rad);
    const bool t1 = (d1 >= 0.f) && (fabsf(__fmaf_rn(r.dx, d1, ox)) < rad) && (fabsf(__fmaf_rn(r.dz, d1, oz)) < rad);
    const bool t2 = (d2 >= 0.f) && (fabsf(__fmaf_rn(r.dx, d2, ox)) < rad) && (fabsf(__fmaf_rn(r.dy, d2, oy)) < rad);
    float entry = t0 ? d0 : (t1 ? d1 : (t2 ? d2 : 0.f));
    if (cmax < rad) entry = -1.f;
    if (exit_t) {
        const float e0 = __fmul_rn(__fmaf_rn(-rad, r.sx, -ox), r.ix);
        const float e1 = __fmul_rn(__fmaf_rn(-rad, r.sy, -oy), r.iy);
        const float e2 = __fmul_rn(__fmaf_rn(-rad, r.sz, -oz), r.iz);
        *exit_t = fminf(fminf(e0, e1), e2);
    }
    return entry;
}

// MODE 0: count nuggets per ray; 1: emit (ridx, pidx, depth) at the packed offsets of a previous count + scan; 2: ONE pass for the
// sync-free fused trace -- depth pairs go to the ray's own staging row depth[ray * stage_cap + n] and the count is written too
// (no second traversal; the filter / emit kernels read the staged rows)
// SPLIT (staged modes, level >= 2): 64 threads per ray.  Thread (ray, i1, i2) walks only the i1-th child of the root and the i2-th
// child of that node in the ray's own visiting order (children j = code ^ i) and everything below; the 64 partial traversals in
// (i1, i2) order ARE the ray's DFS order, so per-thread counts + a prefix over the ray's 64 slots give every nugget its place in
// the ray's staging row.  The sequential per-ray DFS is latency bound (dependent octree-byte / prefix loads, two warps per SM at
// 16 k rays): 64-way splitting turns 0.42 ms into two short passes.
template <int MODE, bool SPLIT>
__global__ void raytrace_kernel(const uint8_t* __restrict__ octree, const int* __restrict__ prefix,
                                const float* __restrict__ org, const float* __restrict__ dir, int64_t N, int level,
                                int* __restrict__ counts, const int64_t* __restrict__ offsets,
                                int64_t* __restrict__ ridx, int64_t* __restrict__ pidx, float* __restrict__ depth, int64_t stage_cap,
                                int* __restrict__ counts64) {
    constexpr bool EMIT = MODE != 0;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t ray = SPLIT ? (tid >> 6) : tid;
    const int sub = SPLIT ? (int)(tid & 63) : 0;
    if (ray >= N) return;
    RayCtx r;
    r.ox = org[3 * ray]; r.oy = org[3 * ray + 1]; r.oz = org[3 * ray + 2];
    r.dx = dir[3 * ray]; r.dy = dir[3 * ray + 1]; r.dz = dir[3 * ray + 2];
    r.ix = __fdiv_rn(1.0f, r.dx); r.iy = __fdiv_rn(1.0f, r.dy); r.iz = __fdiv_rn(1.0f, r.dz);
    r.sx = (__float_as_uint(r.dx) >> 31) ? 1.f : -1.f;
    r.sy = (__float_as_uint(r.dy) >> 31) ? 1.f : -1.f;
    r.sz = (__float_as_uint(r.dz) >> 31) ? 1.f : -1.f;
    const float ax = __fmaf_rn(0.5f, r.ox, 0.5f), ay = __fmaf_rn(0.5f, r.oy, 0.5f), az = __fmaf_rn(0.5f, r.oz, 0.5f);

    int64_t out = MODE == 1 ? offsets[ray] : (MODE == 2 ? ray * stage_cap : 0);
    if (SPLIT && MODE == 2) {      // this thread's place in the ray's row: nuggets of the slots visited before it
        const int* c = counts64 + ray * 64;
        for (int k = 0; k < sub; ++k) out += __ldg(c + k);
    }
    int n = 0;
    int node[PAG_MAX_LEVEL + 1], px[PAG_MAX_LEVEL + 1], py[PAG_MAX_LEVEL + 1], pz[PAG_MAX_LEVEL + 1];
    uint8_t byte[PAG_MAX_LEVEL + 1], code[PAG_MAX_LEVEL + 1], iter[PAG_MAX_LEVEL + 1], lim[PAG_MAX_LEVEL + 1];
    int l = -1;
    {   // root
        float ex;
        const float entry = ray_aabb(r, 0.f, 0.f, 0.f, 1.f, &ex);
        if (level == 0) {
            if (entry > 0.f) {
                if (EMIT) { if (MODE == 1) { ridx[out] = ray; pidx[out] = 0; } depth[2 * out] = entry; depth[2 * out + 1] = ex; ++out; }
                ++n;
            }
        } else if (entry != 0.f) {
            l = 0; node[0] = 0; px[0] = py[0] = pz[0] = 0; byte[0] = __ldg(octree);
            iter[0] = SPLIT ? (uint8_t)(sub >> 3) : 0; lim[0] = SPLIT ? (uint8_t)((sub >> 3) + 1) : 8;
            code[0] = (uint8_t)(((ax > 0.5f) ? 4 : 0) | ((ay > 0.5f) ? 2 : 0) | ((az > 0.5f) ? 1 : 0));
        }
    }
    while (l >= 0) {
        if (iter[l] == lim[l]) { --l; continue; }
        const uint32_t i = iter[l]++;
        const uint32_t j = code[l] ^ i;
        const uint32_t b = byte[l];
        if (!((b >> j) & 1u)) continue;
        const int child = __ldg(prefix + node[l]) + __popc(b & ((2u << j) - 1u));
        const int cl = l + 1;
        const int cx = 2 * px[l] + ((j >> 2) & 1), cy = 2 * py[l] + ((j >> 1) & 1), cz = 2 * pz[l] + (j & 1);
        const float rad = 1.0f / (float)(1 << cl);
        const float vcx = __fmaf_rn(rad, (float)(2 * cx + 1), -1.0f);
        const float vcy = __fmaf_rn(rad, (float)(2 * cy + 1), -1.0f);
        const float vcz = __fmaf_rn(rad, (float)(2 * cz + 1), -1.0f);
        if (cl == level) {
            float ex;
            const float entry = ray_aabb(r, vcx, vcy, vcz, rad, &ex);
            if (entry > 0.f) {
                if (EMIT) { if (MODE == 1) { ridx[out] = ray; pidx[out] = child; } depth[2 * out] = entry; depth[2 * out + 1] = ex; ++out; }
                ++n;
            }
        } else {
            const float entry = ray_aabb(r, vcx, vcy, vcz, rad, nullptr);
            if (entry != 0.f) {
                l = cl; node[l] = child; px[l] = cx; py[l] = cy; pz[l] = cz;
                byte[l] = __ldg(octree + child);
                iter[l] = (SPLIT && l == 1) ? (uint8_t)(sub & 7) : 0; lim[l] = (SPLIT && l == 1) ? (uint8_t)((sub & 7) + 1) : 8;
                const float bx = __fmul_rn(rad, (float)cx + 0.5f), by = __fmul_rn(rad, (float)cy + 0.5f),
                            bz = __fmul_rn(rad, (float)cz + 0.5f);
                code[l] = (uint8_t)(((__fsub_rn(ax, bx) > 0.f) ? 4 : 0) | ((__fsub_rn(ay, by) > 0.f) ? 2 : 0) |
                                    ((__fsub_rn(az, bz) > 0.f) ? 1 : 0));
            }
        }
    }
    if (SPLIT) { if (MODE == 0) counts64[tid] = n; }
    else if (MODE != 1) counts[ray] = n;
}

// per-ray totals of the 64 slot counts
__global__ void raytrace_totals_kernel(const int* __restrict__ counts64, int64_t N, int* __restrict__ counts) {
    const int64_t ray = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (ray >= N) return;
    const int4* c = reinterpret_cast<const int4*>(counts64 + ray * 64);
    int s = 0;
#pragma unroll
    for (int k = 0; k < 16; ++k) { const int4 v = __ldg(c + k); s += v.x + v.y + v.z + v.w; }
    counts[ray] = s;
}

__device__ __forceinline__ float voxel_sample_depth(int64_t k, int s, int S, float t0, float t1,
                                                    const float* __restrict__ jitter, uint32_t seed) {
    const uint64_t flat = (uint64_t)k * (uint64_t)S + (uint64_t)s;
    const float u = jitter ? __ldg(jitter + flat) : pag_jitter(seed, flat);
    const float step = __fdiv_rn(__fadd_rn((float)s, u), (float)S);
    return __fadd_rn(t0, __fmul_rn(__fsub_rn(t1, t0), step));
}

__global__ void voxel_samples_kernel(const float* __restrict__ org, const float* __restrict__ dir,
                                     const int64_t* __restrict__ ridx, const float* __restrict__ depth, int64_t K,
                                     int S, const float* __restrict__ jitter, uint32_t seed,
                                     float* __restrict__ samples, float* __restrict__ depths,
                                     float* __restrict__ deltas, uint8_t* __restrict__ boundary) {
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= K * S) return;
    const int64_t k = t / S;
    const int s = (int)(t - k * S);
    const int64_t ray = ridx[k];
    const float t0 = depth[2 * k], t1 = depth[2 * k + 1];
    const float ds = voxel_sample_depth(k, s, S, t0, t1, jitter, seed);
    const float prev = (s == 0) ? t0 : voxel_sample_depth(k, s - 1, S, t0, t1, jitter, seed);
    depths[t] = ds;
    deltas[t] = __fsub_rn(ds, prev);
    samples[3 * t + 0] = __fadd_rn(org[3 * ray + 0], __fmul_rn(dir[3 * ray + 0], ds));
    samples[3 * t + 1] = __fadd_rn(org[3 * ray + 1], __fmul_rn(dir[3 * ray + 1], ds));
    samples[3 * t + 2] = __fadd_rn(org[3 * ray + 2], __fmul_rn(dir[3 * ray + 2], ds));
    boundary[t] = (s == 0 && (k == 0 || ridx[k - 1] != ray)) ? 1 : 0;
}

// max-travel filter (tracers/panoptic_packed_rf_tracer.py:88-99): keep[k] = d0[k] - d0[first nugget of ray] < max_travel
__global__ void max_travel_mask_kernel(const int64_t* __restrict__ ridx, const float* __restrict__ depths, int64_t K,
                                       int S, const int64_t* __restrict__ ray_first, float max_travel,
                                       uint8_t* __restrict__ keep) {
    const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const int64_t f = ray_first[ridx[k]];
    keep[k] = (__fsub_rn(depths[k * S], depths[f * S]) < max_travel) ? 1 : 0;
}

// ---- sync-free 'voxel' marching (fused training trace) -------------------------------------------------------------------
// The nuggets of the raytrace pass stay on the device (worst-case buffers, count = nug_off[N]); the max-travel filter of
// tracers/panoptic_packed_rf_tracer.py:88-108 is folded into the per-ray count of the kept nuggets, and the emit pass writes the
// S samples of every kept nugget straight into the packed (ray-sorted) sample list the encoders / decoders read.  The jitter
// index stays k * S + s with k the UNFILTERED nugget index, so positions are bit-identical to pag_voxel_samples + filter.
// warp per ray: rel[k] = rank of nugget k among the ray's kept nuggets (-1 = dropped); counts[ray] = kept * S
// stage_cap > 0: depth / rel are per-ray staging rows (entry j of ray r at r * stage_cap + j) instead of packed arrays
__global__ void voxel_filter_kernel(const float* __restrict__ depth, const int64_t* __restrict__ nug_off, int64_t N, int S,
                                    uint32_t seed, const uint32_t* __restrict__ seed_dev, float max_travel, int apply_filter,
                                    int* __restrict__ rel, int* __restrict__ counts, int64_t stage_cap) {
    if (seed_dev) seed = __ldg(seed_dev);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= N) return;
    const int64_t k0 = nug_off[ray], k1 = nug_off[ray + 1];
    const int64_t shift = stage_cap > 0 ? ray * stage_cap - k0 : 0;      // storage index = k + shift
    int kept = 0;
    if (k1 > k0) {
        const float first = voxel_sample_depth(k0, 0, S, depth[2 * (k0 + shift)], depth[2 * (k0 + shift) + 1], nullptr, seed);
        for (int64_t kb = k0; kb < k1; kb += 32) {
            const int64_t k = kb + lane;
            bool keep = false;
            if (k < k1) {
                const float d0 = voxel_sample_depth(k, 0, S, depth[2 * (k + shift)], depth[2 * (k + shift) + 1], nullptr, seed);
                keep = !apply_filter || (__fsub_rn(d0, first) < max_travel);
            }
            const unsigned b = __ballot_sync(0xffffffffu, keep);
            if (k < k1) rel[k + shift] = keep ? kept + __popc(b & ((1u << lane) - 1u)) : -1;
            kept += __popc(b);
        }
    }
    if (lane == 0) counts[ray] = kept * S;
}

// warp per ray over the staged rows: lanes = (nugget, sample) pairs of the ray; the ray's kept samples are contiguous in the packed
// output, so the stores coalesce
__global__ void voxel_emit_staged_kernel(const float* __restrict__ org, const float* __restrict__ dir, const float* __restrict__ depth,
                                         const int* __restrict__ rel, const int64_t* __restrict__ nug_off, const int64_t* __restrict__ offsets,
                                         int64_t N, int64_t stage_cap, int S, uint32_t seed, const uint32_t* __restrict__ seed_dev,
                                         int64_t* __restrict__ ridx, float* __restrict__ samples, float* __restrict__ depths,
                                         float* __restrict__ deltas) {
    if (seed_dev) seed = __ldg(seed_dev);
    const int64_t ray = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (ray >= N) return;
    const int64_t k0 = nug_off[ray];
    const int n = (int)(nug_off[ray + 1] - k0);
    if (n == 0 || offsets[ray + 1] == offsets[ray]) return;
    const float ox = org[3 * ray], oy = org[3 * ray + 1], oz = org[3 * ray + 2];
    const float dx = dir[3 * ray], dy = dir[3 * ray + 1], dz = dir[3 * ray + 2];
    const int64_t base = offsets[ray], row = ray * stage_cap;
    for (int i = lane; i < n * S; i += 32) {
        const int j = i / S, s = i - j * S;
        const int r = rel[row + j];
        if (r < 0) continue;
        const float t0 = depth[2 * (row + j)], t1 = depth[2 * (row + j) + 1];
        const float ds = voxel_sample_depth(k0 + j, s, S, t0, t1, nullptr, seed);
        const float prev = (s == 0) ? t0 : voxel_sample_depth(k0 + j, s - 1, S, t0, t1, nullptr, seed);
        const int64_t dst = base + (int64_t)r * S + s;
        ridx[dst] = ray;
        depths[dst] = ds;
        deltas[dst] = __fsub_rn(ds, prev);
        samples[3 * dst + 0] = __fadd_rn(ox, __fmul_rn(dx, ds));
        samples[3 * dst + 1] = __fadd_rn(oy, __fmul_rn(dy, ds));
        samples[3 * dst + 2] = __fadd_rn(oz, __fmul_rn(dz, ds));
    }
}

// thread per (unfiltered nugget, sample): offsets = per-ray packed SAMPLE offsets of the kept nuggets
__global__ void voxel_emit_kernel(const float* __restrict__ org, const float* __restrict__ dir, const int64_t* __restrict__ nug_ridx,
                                  const float* __restrict__ depth, const int* __restrict__ rel, const int64_t* __restrict__ nug_off,
                                  const int64_t* __restrict__ offsets, int64_t N, int64_t K_max, int S, uint32_t seed,
                                  const uint32_t* __restrict__ seed_dev, int64_t* __restrict__ ridx, float* __restrict__ samples,
                                  float* __restrict__ depths, float* __restrict__ deltas) {
    if (seed_dev) seed = __ldg(seed_dev);
    const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t K = min(K_max, nug_off[N]);
    const int64_t k = t / S;
    if (k >= K) return;
    const int r = rel[k];
    if (r < 0) return;
    const int s = (int)(t - k * S);
    const int64_t ray = nug_ridx[k];
    const float t0 = depth[2 * k], t1 = depth[2 * k + 1];
    const float ds = voxel_sample_depth(k, s, S, t0, t1, nullptr, seed);
    const float prev = (s == 0) ? t0 : voxel_sample_depth(k, s - 1, S, t0, t1, nullptr, seed);
    const int64_t dst = offsets[ray] + (int64_t)r * S + s;
    ridx[dst] = ray;
    depths[dst] = ds;
    deltas[dst] = __fsub_rn(ds, prev);
    samples[3 * dst + 0] = __fadd_rn(org[3 * ray + 0], __fmul_rn(dir[3 * ray + 0], ds));
    samples[3 * dst + 1] = __fadd_rn(org[3 * ray + 1], __fmul_rn(dir[3 * ray + 1], ds));
    samples[3 * dst + 2] = __fadd_rn(org[3 * ray + 2], __fmul_rn(dir[3 * ray + 2], ds));
}

// ---------------------------------------------------------------------------------------------
// octree rebuild from a dense leaf-occupancy mask (prune(): pc_nerf/panoptic_delta_nef.py:63-104 -> kaolin
// unbatched_points_to_octree + wisp OctreeAS.init).  All nodes of all levels live in ONE flat index space
// f = base_l + morton (base_l = (8^l - 1) / 7): a bottom-up OR pass per level gives every node its child byte and existence flag,
// ONE exclusive scan over the flags gives every existing node its point index (points are listed level by level in Morton order,
// and the byte of a non-leaf node sits at the same index in `octree`), a second scan over the popcounts gives `prefix`.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int64_t level_base(int l) { return ((1ll << (3 * l)) - 1) / 7; }

__global__ void octree_build_leaf_kernel(const uint8_t* __restrict__ mask, int L, int* __restrict__ exists) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (3 * L))) return;
    exists[level_base(L) + i] = mask[i] ? 1 : 0;
}
__global__ void octree_build_level_kernel(int l, int* __restrict__ exists, uint8_t* __restrict__ bytes) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (1ll << (3 * l))) return;
    const int* ch = exists + level_base(l + 1) + 8 * i;
    uint32_t b = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) b |= ch[j] ? (1u << j) : 0u;
    bytes[level_base(l) + i] = (uint8_t)b;
    exists[level_base(l) + i] = b ? 1 : 0;
}
__global__ void octree_build_emit_kernel(int L, const int* __restrict__ exists, const uint8_t* __restrict__ bytes,
                                         const int64_t* __restrict__ pos, uint8_t* __restrict__ octree, int16_t* __restrict__ points,
                                         int* __restrict__ popc) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= level_base(L + 1) || !exists[f]) return;
    int l = 0;
    while (l < L && f >= level_base(l + 1)) ++l;
    const int64_t code = f - level_base(l), p = pos[f];
    int x = 0, y = 0, z = 0;
    for (int b = 0; b < l; ++b) {
        x |= (int)((code >> (3 * b + 2)) & 1) << b;
        y |= (int)((code >> (3 * b + 1)) & 1) << b;
        z |= (int)((code >> (3 * b)) & 1) << b;
    }
    points[3 * p] = (int16_t)x; points[3 * p + 1] = (int16_t)y; points[3 * p + 2] = (int16_t)z;
    if (l < L) { octree[p] = bytes[f]; popc[p] = __popc((uint32_t)bytes[f]); }
}
__global__ void octree_build_finish_kernel(int L, const int64_t* __restrict__ pos, const int64_t* __restrict__ prefix64, int* __restrict__ prefix,
                                           int* __restrict__ pyramid) {
    const int64_t n_nodes = pos[level_base(L)];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i <= n_nodes) prefix[i] = (int)prefix64[i];
    if (i == 0) {
        int acc = 0;
        for (int l = 0; l <= L; ++l) {
            const int cnt = (int)(pos[level_base(l + 1)] - pos[level_base(l)]);
            pyramid[l] = cnt;
            pyramid[(L + 2) + l] = acc;
            acc += cnt;
        }
        pyramid[L + 1] = 0;
        pyramid[(L + 2) + L + 1] = acc;
    }
}

__global__ void mark_pack_boundaries_kernel(const int64_t* __restrict__ ids, int64_t M, uint8_t* __restrict__ b) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= M) return;
    b[i] = (i == 0 || ids[i] != ids[i - 1]) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------------------------
extern "C" {

int pag_octree_query(const uint8_t* octree, const int32_t* prefix, const float* coords, int64_t P, int level,
                     int32_t* pidx, void* stream) {
    if (level < 0 || level > PAG_MAX_LEVEL) return PAG_ERR_ARG;
    if (P == 0) return PAG_OK;
    octree_query_kernel<<<pag_grid(P, 256), 256, 0, (cudaStream_t)stream>>>(octree, prefix, coords, P, level, pidx);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// offsets i64[N+1] of the packed list (offsets[N] = M); offsets_c i64[N+1] receives the compacted list's (offsets_c[N] = live count)
int pag_compact_count(const float* sigma, const int64_t* offsets, int64_t N, int32_t* counts, int64_t* offsets_c, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        compact_count_kernel<<<pag_grid(N * 32, 256), 256, 0, st>>>(sigma, offsets, N, counts);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, offsets_c);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_compact_emit(const float* sigma, const int64_t* offsets, const int64_t* offsets_c, int64_t N, const float* samples,
                     const float* depths, const float* deltas, const float* feats, int F, int64_t* ridx_c, float* samples_c,
                     float* depths_c, float* deltas_c, float* feats_c, void* stream) {
    if (F <= 0 || (F & 3)) return PAG_ERR_ARG;
    if (N == 0) return PAG_OK;
    compact_emit_kernel<<<pag_grid(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(sigma, offsets, offsets_c, N, samples, depths, deltas,
                                                                                feats, F, ridx_c, samples_c, depths_c, deltas_c, feats_c);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_exclusive_scan_i32(const int32_t* in, int64_t N, int64_t* out /*[N+1]*/, void* stream) {
    exclusive_scan_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(in, N, out);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// pass 1: pidx_tmp[N*S], counts[N]; then offsets[N+1] = exclusive scan (offsets[N] = M)
int pag_march_ray_count(const float* origins, const float* dirs, int64_t N, int S, const float* linspace,
                        const float* jitter, uint32_t seed, float dist_min, float dist_range,
                        const uint8_t* octree, const int32_t* prefix, int level,
                        int32_t* pidx_tmp, int32_t* counts, int64_t* offsets, const uint32_t* seed_dev, void* stream) {
    if (level < 0 || level > PAG_MAX_LEVEL || S <= 0) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        march_ray_count_kernel<<<pag_grid(N * 32, 256), 256, 0, st>>>(origins, dirs, N, S, linspace, jitter, seed,
                                                                     dist_min, dist_range, octree, prefix, level,
                                                                     pidx_tmp, counts, seed_dev);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_march_ray_emit(const float* origins, const float* dirs, int64_t N, int S, const float* linspace,
                       const float* jitter, uint32_t seed, float dist_min, float dist_range,
                       const int32_t* pidx_tmp, const int64_t* offsets,
                       int64_t* ridx, int64_t* pidx, float* samples, float* depths, float* deltas,
                       uint8_t* boundary, const uint32_t* seed_dev, void* stream) {
    if (N == 0) return PAG_OK;
    march_ray_emit_kernel<<<pag_grid(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        origins, dirs, N, S, linspace, jitter, seed, dist_min, dist_range, pidx_tmp, offsets, ridx, pidx, samples,
        depths, deltas, boundary, seed_dev);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// occupancy bit field of `level`: bits u32[8^level / 32], bit (ix*res + iy)*res + iz of the word array
int pag_octree_level_bits(const uint8_t* octree, const int32_t* prefix, int level, uint32_t* bits, void* stream) {
    if (level < 2 || level > PAG_MAX_LEVEL) return PAG_ERR_ARG;
    const int64_t ncells = (int64_t)1 << (3 * level);
    octree_level_bits_kernel<<<pag_grid(ncells, 256), 256, 0, (cudaStream_t)stream>>>(octree, prefix, level, ncells, bits);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
// 'ray' march against the bit field, no point indices: masks u32[N * ceil(S/32)], counts i32[N], offsets i64[N+1]
int pag_march_ray_bits_count(const float* origins, const float* dirs, int64_t N, int S, const float* linspace,
                             const float* jitter, uint32_t seed, float dist_min, float dist_range, const uint32_t* bits, int level,
                             uint32_t* masks, int32_t* counts, int64_t* offsets, const uint32_t* seed_dev, void* stream) {
    if (level < 2 || level > PAG_MAX_LEVEL || S <= 0) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        march_ray_bits_count_kernel<<<pag_grid(N * 32, 256), 256, 0, st>>>(origins, dirs, N, S, linspace, jitter, seed, dist_min,
                                                                          dist_range, bits, level, masks, counts, seed_dev);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_march_ray_bits_emit(const float* origins, const float* dirs, int64_t N, int S, const float* linspace,
                            const float* jitter, uint32_t seed, float dist_min, float dist_range, const uint32_t* masks,
                            const int64_t* offsets, int64_t* ridx, float* samples, float* depths, float* deltas,
                            const uint32_t* seed_dev, void* stream) {
    if (N == 0) return PAG_OK;
    march_ray_bits_emit_kernel<<<pag_grid(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        origins, dirs, N, S, linspace, jitter, seed, dist_min, dist_range, masks, offsets, ridx, samples, depths, deltas, seed_dev);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_raytrace_count(const uint8_t* octree, const int32_t* prefix, const float* origins, const float* dirs,
                       int64_t N, int level, int32_t* counts, int64_t* offsets, void* stream) {
    if (level < 0 || level > PAG_MAX_LEVEL) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        raytrace_kernel<0, false><<<pag_grid(N, 128), 128, 0, st>>>(octree, prefix, origins, dirs, N, level, counts,
                                                                   nullptr, nullptr, nullptr, nullptr, 0, nullptr);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_raytrace_emit(const uint8_t* octree, const int32_t* prefix, const float* origins, const float* dirs,
                      int64_t N, int level, const int64_t* offsets, int64_t* ridx, int64_t* pidx, float* depth,
                      void* stream) {
    if (level < 0 || level > PAG_MAX_LEVEL) return PAG_ERR_ARG;
    if (N == 0) return PAG_OK;
    raytrace_kernel<1, false><<<pag_grid(N, 128), 128, 0, (cudaStream_t)stream>>>(octree, prefix, origins, dirs, N, level,
                                                                                 nullptr, offsets, ridx, pidx, depth, 0, nullptr);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_voxel_samples(const float* origins, const float* dirs, const int64_t* ridx, const float* depth, int64_t K,
                      int S, const float* jitter, uint32_t seed, float* samples, float* depths, float* deltas,
                      uint8_t* boundary, void* stream) {
    if (S <= 0) return PAG_ERR_ARG;
    if (K == 0) return PAG_OK;
    voxel_samples_kernel<<<pag_grid(K * S, 256), 256, 0, (cudaStream_t)stream>>>(origins, dirs, ridx, depth, K, S,
                                                                                jitter, seed, samples, depths, deltas,
                                                                                boundary);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_max_travel_mask(const int64_t* ridx, const float* depths, int64_t K, int S, const int64_t* ray_first,
                        float max_travel, uint8_t* keep, void* stream) {
    if (K == 0) return PAG_OK;
    max_travel_mask_kernel<<<pag_grid(K, 256), 256, 0, (cudaStream_t)stream>>>(ridx, depths, K, S, ray_first,
                                                                              max_travel, keep);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// Sync-free voxel marching, after pag_raytrace_count / pag_raytrace_emit (nuggets in worst-case device buffers):
// max-travel filter + kept count per ray + scan -> offsets[N+1] = packed SAMPLE offsets (offsets[N] = M on the device)
int pag_voxel_filter_count(const float* nug_depth, const int64_t* nug_offsets, int64_t N, int S, uint32_t seed, const uint32_t* seed_dev,
                           float max_travel, int apply_filter, int32_t* rel, int32_t* counts, int64_t* offsets, void* stream) {
    if (S <= 0) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        voxel_filter_kernel<<<pag_grid(N * 32, 256), 256, 0, st>>>(nug_depth, nug_offsets, N, S, seed, seed_dev, max_travel, apply_filter, rel, counts, 0);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
// ... and the samples of the kept nuggets into the packed list (ridx i64, samples f32[.,3], depths, deltas; K_max = nugget capacity)
int pag_voxel_emit_dyn(const float* origins, const float* dirs, const int64_t* nug_ridx, const float* nug_depth, const int32_t* rel,
                       const int64_t* nug_offsets, const int64_t* offsets, int64_t N, int64_t K_max, int S, uint32_t seed,
                       const uint32_t* seed_dev, int64_t* ridx, float* samples, float* depths, float* deltas, void* stream) {
    if (S <= 0) return PAG_ERR_ARG;
    if (N == 0 || K_max == 0) return PAG_OK;
    voxel_emit_kernel<<<pag_grid(K_max * S, 256), 256, 0, (cudaStream_t)stream>>>(origins, dirs, nug_ridx, nug_depth, rel, nug_offsets, offsets,
                                                                                 N, K_max, S, seed, seed_dev, ridx, samples, depths, deltas);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// Octree (kaolin SPC layout: octree bytes, points, prefix, pyramid) from the dense leaf-occupancy mask u8[8^level] in Morton order
// (= the order of grid.dense_points).  Workspaces / outputs sized for F = (8^(level+1) - 1) / 7 nodes: exists i32[F], bytes u8[F],
// pos i64[F+1], popc i32[F], prefix64 i64[F+1]; outputs octree u8[F], points i16[F,3], prefix i32[F+1], pyramid i32[2, level+2]
// (device).  Valid lengths: n_nodes = pyramid[1][level], n_points = pyramid[1][level+1].  An empty mask yields n_points = 0.
int pag_octree_from_mask(const uint8_t* mask, int level, int32_t* exists, uint8_t* bytes, int64_t* pos, int32_t* popc, int64_t* prefix64,
                         uint8_t* octree, int16_t* points, int32_t* prefix, int32_t* pyramid, void* stream) {
    if (level < 1 || level > PAG_MAX_LEVEL) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t F = ((1ll << (3 * (level + 1))) - 1) / 7, leaves = 1ll << (3 * level), inner = ((1ll << (3 * level)) - 1) / 7;
    octree_build_leaf_kernel<<<pag_grid(leaves, 256), 256, 0, st>>>(mask, level, exists);
    PAG_LAUNCH_CHECK();
    for (int l = level - 1; l >= 0; --l) {
        octree_build_level_kernel<<<pag_grid(1ll << (3 * l), 256), 256, 0, st>>>(l, exists, bytes);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(exists, F, pos);
    PAG_LAUNCH_CHECK();
    cudaError_t e = cudaMemsetAsync(popc, 0, sizeof(int32_t) * inner, st);
    if (e != cudaSuccess) return (int)e;
    octree_build_emit_kernel<<<pag_grid(F, 256), 256, 0, st>>>(level, exists, bytes, pos, octree, points, popc);
    PAG_LAUNCH_CHECK();
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(popc, inner, prefix64);      // entries beyond n_nodes are zero: harmless
    PAG_LAUNCH_CHECK();
    octree_build_finish_kernel<<<pag_grid(inner + 1, 256), 256, 0, st>>>(level, pos, prefix64, prefix, pyramid);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// One-traversal variant for the fused trace: nuggets of ray r are staged in row r of stage_depth f32[N, stage_cap, 2] (stage_cap >= the
// DDA worst case 3 * 2^level - 2), counts / nug_offsets as pag_raytrace_count; pag_voxel_filter_count_staged / pag_voxel_emit_staged
// then read the rows (rel i32[N, stage_cap]).  Same nuggets, same order, same samples as the two-pass chain, one DFS per ray instead of two.
int pag_raytrace_stage(const uint8_t* octree, const int32_t* prefix, const float* origins, const float* dirs, int64_t N, int level,
                       int64_t stage_cap, int32_t* counts, int64_t* nug_offsets, float* stage_depth, int32_t* slot_counts, void* stream) {
    if (level < 0 || level > PAG_MAX_LEVEL || stage_cap < 1) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0 && level >= 2 && slot_counts) {
        // 64 threads per ray (see raytrace_kernel<., true>): count per slot, per-ray totals, then the same walk again writing each
        // slot's nuggets at its prefix inside the ray's row -- two short, wide passes instead of one long, narrow one
        raytrace_kernel<0, true><<<pag_grid(N * 64, 128), 128, 0, st>>>(octree, prefix, origins, dirs, N, level, nullptr, nullptr, nullptr,
                                                                       nullptr, nullptr, 0, slot_counts);
        PAG_LAUNCH_CHECK();
        raytrace_totals_kernel<<<pag_grid(N, 128), 128, 0, st>>>(slot_counts, N, counts);
        PAG_LAUNCH_CHECK();
        raytrace_kernel<2, true><<<pag_grid(N * 64, 128), 128, 0, st>>>(octree, prefix, origins, dirs, N, level, nullptr, nullptr, nullptr,
                                                                       nullptr, stage_depth, stage_cap, slot_counts);
        PAG_LAUNCH_CHECK();
    } else if (N > 0) {
        raytrace_kernel<2, false><<<pag_grid(N, 64), 64, 0, st>>>(octree, prefix, origins, dirs, N, level, counts, nullptr, nullptr, nullptr,
                                                                 stage_depth, stage_cap, nullptr);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, nug_offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_voxel_filter_count_staged(const float* stage_depth, const int64_t* nug_offsets, int64_t N, int64_t stage_cap, int S, uint32_t seed,
                                  const uint32_t* seed_dev, float max_travel, int apply_filter, int32_t* rel, int32_t* counts,
                                  int64_t* offsets, void* stream) {
    if (S <= 0 || stage_cap < 1) return PAG_ERR_ARG;
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        voxel_filter_kernel<<<pag_grid(N * 32, 256), 256, 0, st>>>(stage_depth, nug_offsets, N, S, seed, seed_dev, max_travel, apply_filter, rel,
                                                                  counts, stage_cap);
        PAG_LAUNCH_CHECK();
    }
    exclusive_scan_kernel<<<1, 1024, 0, st>>>(counts, N, offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_voxel_emit_staged(const float* origins, const float* dirs, const float* stage_depth, const int32_t* rel, const int64_t* nug_offsets,
                          const int64_t* offsets, int64_t N, int64_t stage_cap, int S, uint32_t seed, const uint32_t* seed_dev,
                          int64_t* ridx, float* samples, float* depths, float* deltas, void* stream) {
    if (S <= 0 || stage_cap < 1) return PAG_ERR_ARG;
    if (N == 0) return PAG_OK;
    voxel_emit_staged_kernel<<<pag_grid(N * 32, 256), 256, 0, (cudaStream_t)stream>>>(origins, dirs, stage_depth, rel, nug_offsets, offsets, N,
                                                                                     stage_cap, S, seed, seed_dev, ridx, samples, depths, deltas);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_mark_pack_boundaries(const int64_t* ids, int64_t M, uint8_t* boundary, void* stream) {
    if (M == 0) return PAG_OK;
    mark_pack_boundaries_kernel<<<pag_grid(M, 256), 256, 0, (cudaStream_t)stream>>>(ids, M, boundary);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
