// Instant-ngp multiresolution hash grids, forward + backward, two index conventions:
//
//  * "tcnn"     -- tiny-cuda-nn GridEncoding as configured by the reference at
//                  grids/hash_grid_tinycudann.py:24-34 and called at :41 (coords in [-1,1] wrap
//                  through (uint32)(int) like upstream; dense levels when res^3 <= table size).
//  * "hashnerf" -- the reference's own torch grid, grids/hash_grid_torch.py:13-108 (always hashed,
//                  clamped cell, weights from the UNclamped point, corner order i,j,k x-major).
//
// Same B200 mapping as permuto.cu: one thread per sample walks all levels (8 float2 gathers per
// level in flight), table L2-resident, red.global.add.v2.f32 scatter with warp aggregation on
// the coarse levels.  Vertex indices are bit-exact against oracle/hashgrid.py (and, for the
// hashnerf flavour, against the reference file itself via tests/golden/hash_torch.npz).
#include "common.cuh"
#include <cuda_fp16.h>

#define P1 2654435761u
#define P2 805459861u

// corner numbering used internally: bit d of c set -> +1 along dim d
struct Cell {
    uint32_t idx[8];
    float w[3];
    float dwdx[3];  // d w / d x per dim
};

__device__ __forceinline__ void tcnn_cell(float x, float y, float z, float scale, uint32_t res, uint32_t size,
                                          Cell& c) {
    const float in[3] = {x, y, z};
    uint32_t pg[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float p = __fmaf_rn(scale, in[d], 0.5f);
        const float f = floorf(p);
        c.w[d] = __fsub_rn(p, f);
        c.dwdx[d] = scale;
        pg[d] = (uint32_t)(int)f;
    }
    const bool hashed = (uint64_t)res * res * res > (uint64_t)size;
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t a = pg[0] + (k & 1), b = pg[1] + ((k >> 1) & 1), e = pg[2] + ((k >> 2) & 1);
        const uint32_t h = hashed ? (a ^ (b * P1) ^ (e * P2)) : (a + b * res + e * (res * res));
        c.idx[k] = h % size;
    }
}

__device__ __forceinline__ void hashnerf_cell(float x, float y, float z, float res, uint32_t mask, Cell& c) {
    const float in[3] = {x, y, z};
    const float gs = __fdiv_rn(2.0f, res);
    int bl[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        const float xc = fminf(fmaxf(in[d], -1.0f), 1.0f);
        bl[d] = (int)floorf(__fdiv_rn(__fsub_rn(xc, -1.0f), gs));
        const float vmin = __fadd_rn(__fmul_rn((float)bl[d], gs), -1.0f);
        const float vmax = __fadd_rn(vmin, __fmul_rn(1.0f, gs));
        const float den = __fsub_rn(vmax, vmin);
        c.w[d] = __fdiv_rn(__fsub_rn(in[d], vmin), den);
        c.dwdx[d] = __fdiv_rn(1.0f, den);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const uint32_t a = (uint32_t)(bl[0] + (k & 1)), b = (uint32_t)(bl[1] + ((k >> 1) & 1)),
                       e = (uint32_t)(bl[2] + ((k >> 2) & 1));
        c.idx[k] = (a ^ (b * P1) ^ (e * P2)) & mask;
    }
}

__device__ __forceinline__ float corner_weight(const Cell& c, int k) {
    float w = 1.f;
#pragma unroll
    for (int d = 0; d < 3; ++d) w *= ((k >> d) & 1) ? c.w[d] : (1.f - c.w[d]);
    return w;
}

struct LevelDesc {  // per level, passed as parallel device arrays
    const float* fparam;     // tcnn: scale ; hashnerf: resolution
    const uint32_t* res;     // tcnn only
    const uint32_t* offset;  // entry offset of the level in the table
    const uint32_t* size;    // tcnn: entries in level ; hashnerf: mask+1
};

template <int FLAVOUR>  // 0 = tcnn, 1 = hashnerf
__device__ __forceinline__ void make_cell(float x, float y, float z, const LevelDesc& d, int l, Cell& c) {
    if (FLAVOUR == 0) tcnn_cell(x, y, z, __ldg(d.fparam + l), __ldg(d.res + l), __ldg(d.size + l), c);
    else hashnerf_cell(x, y, z, __ldg(d.fparam + l), __ldg(d.size + l) - 1u, c);
}

// IMG16: features go out as the tensor-core decoders' fp16 operand image (see permuto_fwd_kernel in permuto.cu): tile m / 128
// holds [nXc chunks][128 rows][8 halfs], nXc = ceil(2L / 16) * 2 chunks of four levels each; levels >= L and rows between M and
// the end of the last tile are zero (the tensor core reads whole, 16-feature-padded tiles).
template <int FLAVOUR, bool IMG16>
__global__ void __launch_bounds__(128) hash_fwd_kernel(const float* __restrict__ pos, int64_t M,
                                                       const float* __restrict__ table, int L, LevelDesc d,
                                                       float* __restrict__ out, int round_half,
                                                       const int64_t* __restrict__ m_dev, int pos_half) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m_dev) M = min(M, __ldg(m_dev));   // packed-sample count produced on the device by the marcher (fused trace)
    uint4* img = nullptr;
    const int nXc = ((2 * L + 15) & ~15) >> 3;
    if (IMG16) {
        const int64_t Mpad = (M + 127) & ~(int64_t)127;
        if (m >= Mpad) return;
        img = reinterpret_cast<uint4*>(out) + ((m >> 7) * nXc) * 128 + (m & 127);    // chunk c at img[c * 128]
        if (m >= M) {
            for (int c = 0; c < nXc; ++c) img[c * 128] = make_uint4(0u, 0u, 0u, 0u);
            return;
        }
    } else if (m >= M) return;
    uint32_t pk[4] = {0u, 0u, 0u, 0u};
    float x = pos[3 * m], y = pos[3 * m + 1], z = pos[3 * m + 2];
    if (pos_half) {   // autocast: custom_fwd(cast_inputs=torch.half), grids/hash_grid_tinycudann.py:36
        x = __half2float(__float2half_rn(x)); y = __half2float(__float2half_rn(y)); z = __half2float(__float2half_rn(z));
    }
    float2* orow = reinterpret_cast<float2*>(out + m * (int64_t)(2 * L));
    const bool row16 = !IMG16 && !(L & 1);   // 16-byte aligned f32 rows: two levels per store (see permuto_fwd_kernel)
    float2 prev = make_float2(0.f, 0.f);
#pragma unroll 2
    for (int l = 0; l < L; ++l) {
        Cell c;
        make_cell<FLAVOUR>(x, y, z, d, l, c);
        const float* tl = table + 2 * (size_t)__ldg(d.offset + l);
        float2 v[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) v[k] = ldg2(tl + 2 * (size_t)c.idx[k]);
        float2 acc = make_float2(0.f, 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float w = corner_weight(c, k);
            acc.x = fmaf(w, v[k].x, acc.x);
            acc.y = fmaf(w, v[k].y, acc.y);
        }
        if (IMG16) {
            const __half2 h = __floats2half2_rn(acc.x, acc.y);
            pk[l & 3] = *reinterpret_cast<const uint32_t*>(&h);
            if ((l & 3) == 3) { img[(l >> 2) * 128] = make_uint4(pk[0], pk[1], pk[2], pk[3]); pk[0] = pk[1] = pk[2] = pk[3] = 0u; }
        } else {
            if (round_half) {  // tcnn returns __half, the wrapper casts back to float (:41)
                acc.x = __half2float(__float2half_rn(acc.x));
                acc.y = __half2float(__float2half_rn(acc.y));
            }
            if (!row16) orow[l] = acc;
            else if (l & 1) reinterpret_cast<float4*>(orow)[l >> 1] = make_float4(prev.x, prev.y, acc.x, acc.y);
            else prev = acc;
        }
    }
    if (IMG16) {      // partially filled chunk (L % 4 != 0) and the all-zero padding chunks up to nXc
        int c = L >> 2;
        if (L & 3) { img[c * 128] = make_uint4(pk[0], pk[1], pk[2], pk[3]); ++c; }
        for (; c < nXc; ++c) img[c * 128] = make_uint4(0u, 0u, 0u, 0u);
    }
}

template <int FLAVOUR>
__global__ void hash_indices_kernel(const float* __restrict__ pos, int64_t M, int L, LevelDesc d,
                                    uint32_t* __restrict__ idx) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m >= M) return;
    const float x = pos[3 * m], y = pos[3 * m + 1], z = pos[3 * m + 2];
    for (int l = 0; l < L; ++l) {
        Cell c;
        make_cell<FLAVOUR>(x, y, z, d, l, c);
        for (int k = 0; k < 8; ++k) idx[((size_t)l * M + m) * 8 + k] = c.idx[k];
    }
}

// GIMG: the upstream gradient arrives as the decoders' fp16 tile image, still multiplied by the power-of-two loss scale
// *img_scale of the tensor-core backward (see permuto_bwd_kernel)
template <int FLAVOUR, bool POS_GRAD, bool GIMG>
__global__ void __launch_bounds__(128) hash_bwd_kernel(const float* __restrict__ pos, int64_t M,
                                                       const float* __restrict__ table, int L, LevelDesc d,
                                                       const float* __restrict__ gout, float* __restrict__ gtable,
                                                       float* __restrict__ gpos, int n_agg_levels,
                                                       const int64_t* __restrict__ m_dev, int pos_half,
                                                       const float* __restrict__ img_scale) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m_dev) M = min(M, __ldg(m_dev));
    if (M <= 0 || (m & ~31ll) >= M) return;   // whole warp beyond the packed samples
    const bool valid = m < M;
    const int64_t mm = valid ? m : (M - 1);
    float x = pos[3 * mm], y = pos[3 * mm + 1], z = pos[3 * mm + 2];
    if (pos_half) {
        x = __half2float(__float2half_rn(x)); y = __half2float(__float2half_rn(y)); z = __half2float(__float2half_rn(z));
    }
    const float2* grow = reinterpret_cast<const float2*>(gout + mm * (int64_t)(2 * L));
    const int nXc = ((2 * L + 15) & ~15) >> 3;
    const uint4* gimg = reinterpret_cast<const uint4*>(gout) + ((mm >> 7) * nXc) * 128 + (mm & 127);   // chunk c at gimg[c * 128]
    const float inv_scale = (GIMG && img_scale) ? 1.f / __ldg(img_scale) : 1.f;
    uint4 gq = make_uint4(0u, 0u, 0u, 0u);
    float gp[3] = {0.f, 0.f, 0.f};
    for (int l = 0; l < L; ++l) {
        Cell c;
        make_cell<FLAVOUR>(x, y, z, d, l, c);
        float2 g;
        if (GIMG) {
            if ((l & 3) == 0) gq = __ldg(gimg + (l >> 2) * 128);
            const uint32_t u = (l & 3) == 0 ? gq.x : ((l & 3) == 1 ? gq.y : ((l & 3) == 2 ? gq.z : gq.w));
            g = __half22float2(*reinterpret_cast<const __half2*>(&u));
            g.x *= inv_scale; g.y *= inv_scale;
        } else {
            g = __ldg(grow + l);
        }
        if (!valid) { g.x = 0.f; g.y = 0.f; }
        const size_t off = 2 * (size_t)__ldg(d.offset + l);
        float* gl = gtable + off;
        if (l < n_agg_levels) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float w = corner_weight(c, k);
                scatter_aggregated(gl, c.idx[k], g.x * w, g.y * w, 0xffffffffu, 0xffffffffu);
            }
        } else if (valid) {
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float w = corner_weight(c, k);
                red_add_f32x2(gl + 2 * (size_t)c.idx[k], g.x * w, g.y * w);
            }
        }
        if (POS_GRAD) {
            const float* tl = table + off;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
                const float2 t = ldg2(tl + 2 * (size_t)c.idx[k]);
                const float s = g.x * t.x + g.y * t.y;
#pragma unroll
                for (int dd = 0; dd < 3; ++dd) {
                    float w = ((k >> dd) & 1) ? 1.f : -1.f;
#pragma unroll
                    for (int o = 0; o < 3; ++o)
                        if (o != dd) w *= ((k >> o) & 1) ? c.w[o] : (1.f - c.w[o]);
                    gp[dd] = fmaf(s * w, c.dwdx[dd], gp[dd]);
                }
            }
        }
    }
    if (POS_GRAD && valid) { gpos[3 * m] = gp[0]; gpos[3 * m + 1] = gp[1]; gpos[3 * m + 2] = gp[2]; }
}

template <int FLAVOUR>
static int launch_bwd(const float* pos, int64_t M, const float* table, int L, LevelDesc d, const float* gout,
                      float* gtable, float* gpos, int n_agg, cudaStream_t st, const int64_t* m_dev = nullptr, int pos_half = 0,
                      bool gimg = false, const float* img_scale = nullptr) {
    const dim3 g = pag_grid(M, 128);
    if (gimg) {
        if (gpos) hash_bwd_kernel<FLAVOUR, true, true><<<g, 128, 0, st>>>(pos, M, table, L, d, gout, gtable, gpos, n_agg, m_dev, pos_half, img_scale);
        else hash_bwd_kernel<FLAVOUR, false, true><<<g, 128, 0, st>>>(pos, M, table, L, d, gout, gtable, gpos, n_agg, m_dev, pos_half, img_scale);
    } else {
        if (gpos) hash_bwd_kernel<FLAVOUR, true, false><<<g, 128, 0, st>>>(pos, M, table, L, d, gout, gtable, gpos, n_agg, m_dev, pos_half, nullptr);
        else hash_bwd_kernel<FLAVOUR, false, false><<<g, 128, 0, st>>>(pos, M, table, L, d, gout, gtable, gpos, n_agg, m_dev, pos_half, nullptr);
    }
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

extern "C" {

// flavour 0 = tcnn (fparam = per-level scale), 1 = hashnerf (fparam = per-level resolution, res unused,
// size = 2^log2_T).  table f32[total_entries, 2]; out f32[M, 2L].
int pag_hash_fwd(int flavour, const float* pos, int64_t M, const float* table, int L, int F, const float* fparam,
                 const uint32_t* res, const uint32_t* offset, const uint32_t* size, float* out, int round_half,
                 void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    if (flavour == 0) hash_fwd_kernel<0, false><<<pag_grid(M, 128), 128, 0, st>>>(pos, M, table, L, d, out, round_half, nullptr, 0);
    else hash_fwd_kernel<1, false><<<pag_grid(M, 128), 128, 0, st>>>(pos, M, table, L, d, out, round_half, nullptr, 0);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_hash_bwd(int flavour, const float* pos, int64_t M, const float* table, int L, int F, const float* fparam,
                 const uint32_t* res, const uint32_t* offset, const uint32_t* size, const float* grad_out,
                 float* grad_table, float* grad_pos, int n_agg_levels, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    return flavour == 0 ? launch_bwd<0>(pos, M, table, L, d, grad_out, grad_table, grad_pos, n_agg_levels, st)
                        : launch_bwd<1>(pos, M, table, L, d, grad_out, grad_table, grad_pos, n_agg_levels, st);
}

// variants for the sync-free fused trace: sample count read on the device (m_dev[0] <= M_max), optional fp16 rounding of pos
int pag_hash_fwd_dyn(int flavour, const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table, int L,
                     int F, const float* fparam, const uint32_t* res, const uint32_t* offset, const uint32_t* size, float* out,
                     int round_half, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M_max == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    if (flavour == 0) hash_fwd_kernel<0, false><<<pag_grid(M_max, 128), 128, 0, st>>>(pos, M_max, table, L, d, out, round_half, m_dev, pos_half);
    else hash_fwd_kernel<1, false><<<pag_grid(M_max, 128), 128, 0, st>>>(pos, M_max, table, L, d, out, round_half, m_dev, pos_half);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_hash_bwd_dyn(int flavour, const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table, int L,
                     int F, const float* fparam, const uint32_t* res, const uint32_t* offset, const uint32_t* size,
                     const float* grad_out, float* grad_table, float* grad_pos, int n_agg_levels, void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M_max == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    return flavour == 0 ? launch_bwd<0>(pos, M_max, table, L, d, grad_out, grad_table, grad_pos, n_agg_levels, st, m_dev, pos_half)
                        : launch_bwd<1>(pos, M_max, table, L, d, grad_out, grad_table, grad_pos, n_agg_levels, st, m_dev, pos_half);
}

// fp16 operand-image interchange with the tensor-core decoders (fused trace): img16 holds ceil(M_max / 128) tiles of
// ceil(2L / 16) * 2 chunks of 2048 bytes; grad_img16 likewise, scaled by *img_scale (device float, nullable = 1)
int pag_hash_fwd_img16_dyn(int flavour, const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table, int L,
                           int F, const float* fparam, const uint32_t* res, const uint32_t* offset, const uint32_t* size, void* img16,
                           void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M_max == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t Mpad = (M_max + 127) & ~(int64_t)127;
    float* out = reinterpret_cast<float*>(img16);
    if (flavour == 0) hash_fwd_kernel<0, true><<<pag_grid(Mpad, 128), 128, 0, st>>>(pos, M_max, table, L, d, out, 0, m_dev, pos_half);
    else hash_fwd_kernel<1, true><<<pag_grid(Mpad, 128), 128, 0, st>>>(pos, M_max, table, L, d, out, 0, m_dev, pos_half);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_hash_bwd_img16_dyn(int flavour, const float* pos, int64_t M_max, const int64_t* m_dev, int pos_half, const float* table, int L,
                           int F, const float* fparam, const uint32_t* res, const uint32_t* offset, const uint32_t* size,
                           const void* grad_img16, const float* img_scale, float* grad_table, float* grad_pos, int n_agg_levels,
                           void* stream) {
    if (F != 2) return PAG_ERR_UNSUPPORTED;
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M_max == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    const float* g = reinterpret_cast<const float*>(grad_img16);
    return flavour == 0 ? launch_bwd<0>(pos, M_max, table, L, d, g, grad_table, grad_pos, n_agg_levels, st, m_dev, pos_half, true, img_scale)
                        : launch_bwd<1>(pos, M_max, table, L, d, g, grad_table, grad_pos, n_agg_levels, st, m_dev, pos_half, true, img_scale);
}

int pag_hash_indices(int flavour, const float* pos, int64_t M, int L, const float* fparam, const uint32_t* res,
                     const uint32_t* offset, const uint32_t* size, uint32_t* idx, void* stream) {
    if (L <= 0 || (flavour != 0 && flavour != 1)) return PAG_ERR_ARG;
    if (M == 0) return PAG_OK;
    LevelDesc d{fparam, res, offset, size};
    cudaStream_t st = (cudaStream_t)stream;
    if (flavour == 0) hash_indices_kernel<0><<<pag_grid(M, 128), 128, 0, st>>>(pos, M, L, d, idx);
    else hash_indices_kernel<1><<<pag_grid(M, 128), 128, 0, st>>>(pos, M, L, d, idx);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
