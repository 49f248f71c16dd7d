// Packed alpha compositing of RGB, depth, semantic probabilities and instance embeddings (+ backward).
//
// Replaces, behind PanopticPackedRFTracer.trace (reference tracers/panoptic_packed_rf_tracer.py:134-205):
//   kaolin.render.spc.exponential_integration (x2: attached + detached tau), sum_reduce (x5),
//   the [M,C] `feats * w` temporaries, and the index_put scatters to dense [N,C] buffers.
// Conventions preserved exactly:
//   tau = sigma*delta;  w_i = exp(-sum_{j<i} tau_j) * (1 - exp(-tau_i));  alpha = sum w
//   rgb   = (1-alpha) + alpha * sum w c   (white)   |   alpha * sum w c   (black)   -- alpha ON TOP
//   depth = sum w t  (no alpha);  semantics/inst = alpha_p * sum w_p f  with w_p, alpha_p DETACHED
//   rays without samples: rgb = background, everything else 0, hit = false.
//
// B200 design: one warp per ray over the packed samples (offsets[R+1] from the marcher's scan).
// Scalars (tau, depth, rgb) use lanes = samples with a warp-segmented exclusive scan carried
// across 32-sample chunks; the wide channels (C up to 256 per pass, 200 for instances) use
// lanes = channels so the [M,C] rows are read exactly once, fully coalesced (800 B/sample).
// Everything is HBM-stream bound: algorithmic bytes per sample fwd = 4+4+4+12+4C_sem+4C_inst (+1).
#include "common.cuh"
#include <cuda_fp16.h>

__device__ __forceinline__ float warp_excl_scan(float v, float& total) {
    const int lane = threadIdx.x & 31;
    float incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    total = __shfl_sync(0xffffffffu, incl, 31);
    return incl - v;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// offsets[r] = first packed index whose ridx >= r  (ridx ascending); offsets[R] = M
__global__ void ray_offsets_kernel(const int64_t* __restrict__ ridx, int64_t M, int64_t R, int64_t* __restrict__ offsets) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > M) return;
    const int64_t lo = (i == 0) ? 0 : (ridx[i - 1] + 1);
    const int64_t hi = (i == M) ? R : ridx[i];
    for (int64_t r = lo; r <= hi && r <= R; ++r) offsets[r] = i;
}

__global__ void __launch_bounds__(256) composite_fwd_kernel(
    const float* __restrict__ sigma, const float* __restrict__ deltas, const float* __restrict__ depths,
    const float* __restrict__ rgb, const float* __restrict__ sem, int Cs, const float* __restrict__ inst, int Ci,
    const int64_t* __restrict__ offsets, int64_t R, int bg_white,
    float* __restrict__ w_out, float* __restrict__ T_out, float* __restrict__ alpha_out, uint8_t* __restrict__ hit_out,
    float* __restrict__ rgb_out, float* __restrict__ rgbsum_out, float* __restrict__ depth_out,
    float* __restrict__ sem_out, float* __restrict__ inst_out) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    float carry = 0.f, a_acc = 0.f, d_acc = 0.f, c0 = 0.f, c1 = 0.f, c2 = 0.f;
    for (int64_t base = s0; base < s1; base += 32) {
        const int64_t i = base + lane;
        const bool ok = i < s1;
        const float tau = ok ? sigma[i] * deltas[i] : 0.f;
        float tot;
        const float ex = warp_excl_scan(tau, tot) + carry;
        carry += tot;
        const float T = expf(-ex);
        const float w = T * (1.f - expf(-tau));
        if (ok) {
            w_out[i] = w; T_out[i] = T;
            a_acc += w;
            if (depths) d_acc += w * depths[i];
            if (rgb) { c0 += w * rgb[3 * i]; c1 += w * rgb[3 * i + 1]; c2 += w * rgb[3 * i + 2]; }
        }
    }
    const float alpha = warp_sum(a_acc);
    if (lane == 0) {
        alpha_out[r] = alpha;
        if (hit_out) hit_out[r] = (alpha > 0.f) ? 1 : 0;
    }
    if (depths) { d_acc = warp_sum(d_acc); if (lane == 0) depth_out[r] = d_acc; }
    if (rgb) {
        c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2);
        if (lane == 0) {
            rgbsum_out[3 * r] = c0; rgbsum_out[3 * r + 1] = c1; rgbsum_out[3 * r + 2] = c2;
            const float bgc = bg_white ? (1.f - alpha) : 0.f;
            const bool empty = (s1 == s0);
            rgb_out[3 * r + 0] = empty ? (bg_white ? 1.f : 0.f) : bgc + alpha * c0;
            rgb_out[3 * r + 1] = empty ? (bg_white ? 1.f : 0.f) : bgc + alpha * c1;
            rgb_out[3 * r + 2] = empty ? (bg_white ? 1.f : 0.f) : bgc + alpha * c2;
        }
    }
    __syncwarp();
    // wide channels: lanes = channels, rows read once, coalesced
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        const float* __restrict__ f = which ? inst : sem;
        float* __restrict__ o = which ? inst_out : sem_out;
        const int C = which ? Ci : Cs;
        if (!f) continue;
        for (int cb = 0; cb < C; cb += 256) {
            float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            for (int64_t i = s0; i < s1; ++i) {
                const float w = w_out[i];
                const float* row = f + i * (int64_t)C + cb + lane;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (cb + lane + 32 * k < C) acc[k] = fmaf(w, __ldg(row + 32 * k), acc[k]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k)
                if (cb + lane + 32 * k < C) o[r * (int64_t)C + cb + lane + 32 * k] = alpha * acc[k];
        }
    }
}

__global__ void __launch_bounds__(256) composite_bwd_kernel(
    const float* __restrict__ sigma, const float* __restrict__ deltas, const float* __restrict__ depths,
    const float* __restrict__ rgb, const int64_t* __restrict__ offsets, int64_t R, int bg_white,
    const float* __restrict__ w_in, const float* __restrict__ T_in, const float* __restrict__ alpha_in,
    const float* __restrict__ rgbsum_in,
    const float* __restrict__ g_alpha, const float* __restrict__ g_rgb, const float* __restrict__ g_depth,
    const float* __restrict__ g_sem, int Cs, const float* __restrict__ g_inst, int Ci,
    float* __restrict__ g_sigma, float* __restrict__ g_rgb_s, float* __restrict__ g_sem_s, float* __restrict__ g_inst_s) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    if (s1 == s0) return;
    const float alpha = alpha_in[r];
    float gr0 = 0.f, gr1 = 0.f, gr2 = 0.f, ga = g_alpha ? g_alpha[r] : 0.f;
    if (g_rgb) {
        gr0 = g_rgb[3 * r]; gr1 = g_rgb[3 * r + 1]; gr2 = g_rgb[3 * r + 2];
        const float sub = bg_white ? 1.f : 0.f;  // d color / d alpha = S_c - 1 (white) | S_c (black)
        ga += gr0 * (rgbsum_in[3 * r] - sub) + gr1 * (rgbsum_in[3 * r + 1] - sub) + gr2 * (rgbsum_in[3 * r + 2] - sub);
    }
    const float gd = g_depth ? g_depth[r] : 0.f;
    // reverse walk: suffix_j>i (dw_j * w_j)
    float carry = 0.f;
    const int64_t n = s1 - s0;
    const int64_t nchunks = (n + 31) / 32;
    for (int64_t ch = nchunks - 1; ch >= 0; --ch) {
        const int64_t i = s0 + ch * 32 + lane;
        const bool ok = i < s1;
        float w = 0.f, T = 0.f, dw = 0.f, dl = 0.f;
        if (ok) {
            w = w_in[i]; T = T_in[i]; dl = deltas[i];
            dw = ga;
            if (depths) dw += gd * depths[i];
            if (g_rgb) {
                const float x0 = rgb[3 * i], x1 = rgb[3 * i + 1], x2 = rgb[3 * i + 2];
                dw += alpha * (gr0 * x0 + gr1 * x1 + gr2 * x2);
                g_rgb_s[3 * i] = w * alpha * gr0; g_rgb_s[3 * i + 1] = w * alpha * gr1; g_rgb_s[3 * i + 2] = w * alpha * gr2;
            }
        }
        const float v = dw * w;
        // reverse exclusive scan inside the chunk: suffix over lanes > lane
        float incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += t;
        }
        const float suffix = incl - v + carry;
        carry += __shfl_sync(0xffffffffu, incl, 0);
        if (ok) {
            const float dtau = dw * (T - w) - suffix;   // T*exp(-tau) = T - w
            g_sigma[i] = dtau * dl;
        }
    }
    // panoptic channels: w, alpha detached -> only d f_i = alpha * w_i * g_out
#pragma unroll 1
    for (int which = 0; which < 2; ++which) {
        const float* __restrict__ g = which ? g_inst : g_sem;
        float* __restrict__ o = which ? g_inst_s : g_sem_s;
        const int C = which ? Ci : Cs;
        if (!g) continue;
        for (int cb = 0; cb < C; cb += 256) {
            float gv[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) gv[k] = (cb + lane + 32 * k < C) ? alpha * g[r * (int64_t)C + cb + lane + 32 * k] : 0.f;
            for (int64_t i = s0; i < s1; ++i) {
                const float w = w_in[i];
                float* row = o + i * (int64_t)C + cb + lane;
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    if (cb + lane + 32 * k < C) row[32 * k] = w * gv[k];
            }
        }
    }
}

// kaolin-compatible segmented sum  [M,C] -> [R,C]  and its backward (broadcast)
__global__ void sum_reduce_kernel(const float* __restrict__ x, int C, const int64_t* __restrict__ offsets, int64_t R,
                                  float* __restrict__ out) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    for (int cb = 0; cb < C; cb += 32) {
        const int c = cb + lane;
        float acc = 0.f;
        if (c < C)
            for (int64_t i = s0; i < s1; ++i) acc += x[i * (int64_t)C + c];
        if (c < C) out[r * (int64_t)C + c] = acc;
    }
}
__global__ void sum_reduce_bwd_kernel(const float* __restrict__ g, int C, const int64_t* __restrict__ offsets, int64_t R,
                                      float* __restrict__ gx) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    for (int cb = 0; cb < C; cb += 32) {
        const int c = cb + lane;
        if (c >= C) continue;
        const float v = g[r * (int64_t)C + c];
        for (int64_t i = s0; i < s1; ++i) gx[i * (int64_t)C + c] = v;
    }
}

// kaolin-compatible exponential integration weights: w = exp(-excl_cumsum(tau)) * (1-exp(-tau)); also T
__global__ void expint_fwd_kernel(const float* __restrict__ tau, const int64_t* __restrict__ offsets, int64_t R,
                                  float* __restrict__ w_out, float* __restrict__ T_out) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    float carry = 0.f;
    for (int64_t base = s0; base < s1; base += 32) {
        const int64_t i = base + lane;
        const bool ok = i < s1;
        const float t = ok ? tau[i] : 0.f;
        float tot;
        const float ex = warp_excl_scan(t, tot) + carry;
        carry += tot;
        const float T = expf(-ex);
        if (ok) { w_out[i] = T * (1.f - expf(-t)); T_out[i] = T; }
    }
}
__global__ void expint_bwd_kernel(const float* __restrict__ gw, const float* __restrict__ w_in,
                                  const float* __restrict__ T_in, const int64_t* __restrict__ offsets, int64_t R,
                                  float* __restrict__ gtau) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    float carry = 0.f;
    const int64_t nchunks = (s1 - s0 + 31) / 32;
    for (int64_t ch = nchunks - 1; ch >= 0; --ch) {
        const int64_t i = s0 + ch * 32 + lane;
        const bool ok = i < s1;
        const float w = ok ? w_in[i] : 0.f, T = ok ? T_in[i] : 0.f, dw = ok ? gw[i] : 0.f;
        const float v = dw * w;
        float incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const float t = __shfl_down_sync(0xffffffffu, incl, o);
            if (lane + o < 32) incl += t;
        }
        const float suffix = incl - v + carry;
        carry += __shfl_sync(0xffffffffu, incl, 0);
        if (ok) gtau[i] = dw * (T - w) - suffix;
    }
}

// ---- loss-scale for the fp16 tensor-core backward: scale = 2^floor(log2(target / max|g|)) (one launch + finalize) ----
// One launch: vectorised abs-max over a (and b), atomicMax into scratch[0]; the LAST block to finish (ticket in scratch[1]) turns
// the maximum into the power-of-two scale and resets both words, so the scratch stays zero between calls (no memset launch).
__global__ void __launch_bounds__(256) absmax_scale_kernel(const float* __restrict__ a, int64_t na, const float* __restrict__ b, int64_t nb,
                                                           const int64_t* __restrict__ m_dev, int wa, int wb, float target,
                                                           unsigned* __restrict__ scratch, float* __restrict__ out) {
    if (m_dev) { const int64_t mv = __ldg(m_dev); na = min(na, mv * wa); nb = min(nb, mv * wb); }
    float mx = 0.f;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nth = (int64_t)gridDim.x * blockDim.x;
#pragma unroll
    for (int t = 0; t < 2; ++t) {
        const float* p = t ? b : a;
        const int64_t n = t ? nb : na;
        if (!p || n <= 0) continue;
        if (!(reinterpret_cast<uintptr_t>(p) & 15)) {
            const int64_t n4 = n >> 2;
            const float4* p4 = reinterpret_cast<const float4*>(p);
            int64_t i = tid;
            for (; i + 3 * nth < n4; i += 4 * nth) {      // four independent 16-byte loads in flight
                const float4 v0 = __ldg(p4 + i), v1 = __ldg(p4 + i + nth), v2 = __ldg(p4 + i + 2 * nth), v3 = __ldg(p4 + i + 3 * nth);
                mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(v0.x), fabsf(v0.y)), fmaxf(fabsf(v0.z), fabsf(v0.w))),
                                     fmaxf(fmaxf(fabsf(v1.x), fabsf(v1.y)), fmaxf(fabsf(v1.z), fabsf(v1.w)))));
                mx = fmaxf(mx, fmaxf(fmaxf(fmaxf(fabsf(v2.x), fabsf(v2.y)), fmaxf(fabsf(v2.z), fabsf(v2.w))),
                                     fmaxf(fmaxf(fabsf(v3.x), fabsf(v3.y)), fmaxf(fabsf(v3.z), fabsf(v3.w)))));
            }
            for (; i < n4; i += nth) {
                const float4 v = __ldg(p4 + i);
                mx = fmaxf(mx, fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w))));     // NaN is dropped by fmaxf
            }
            for (int64_t j = (n4 << 2) + tid; j < n; j += nth) mx = fmaxf(mx, fabsf(p[j]));
        } else {
            for (int64_t j = tid; j < n; j += nth) mx = fmaxf(mx, fabsf(p[j]));
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    __shared__ float wmax[8];
    __shared__ bool last;
    if ((threadIdx.x & 31) == 0) wmax[threadIdx.x >> 5] = mx;
    __syncthreads();
    if (threadIdx.x == 0) {
        float m = wmax[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) m = fmaxf(m, wmax[w]);
        if (m > 0.f) atomicMax(scratch, __float_as_uint(m));
        __threadfence();
        last = atomicAdd(scratch + 1, 1u) == gridDim.x - 1;
    }
    __syncthreads();
    if (last && threadIdx.x == 0) {
        __threadfence();
        const float amax = fmaxf(__uint_as_float(atomicExch(scratch, 0u)), 1e-30f);
        float s = exp2f(floorf(log2f(target / amax)));
        s = fminf(fmaxf(s, 5.9604645e-08f), 1.1529215e18f);   // [2^-24, 2^60]
        *out = s;
        scratch[1] = 0u;
    }
}

// fp16 transport of a gradient table through the all-reduce: out16 = half(x * scale), x = float(in16) * mult / scale
__global__ void __launch_bounds__(256) pack_f16_kernel(const float4* __restrict__ x, int64_t n4, const float* __restrict__ scale, uint2* __restrict__ out) {
    const float s = __ldg(scale);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const float4 v = __ldg(x + i);
        const __half2 a = __floats2half2_rn(v.x * s, v.y * s), b = __floats2half2_rn(v.z * s, v.w * s);
        out[i] = make_uint2(*reinterpret_cast<const uint32_t*>(&a), *reinterpret_cast<const uint32_t*>(&b));
    }
}
__global__ void __launch_bounds__(256) unpack_f16_kernel(const uint2* __restrict__ in, int64_t n4, const float* __restrict__ scale, float mult,
                                                         float4* __restrict__ x) {
    const float s = mult / __ldg(scale);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
        const uint2 u = __ldg(in + i);
        const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&u.x)), b = __half22float2(*reinterpret_cast<const __half2*>(&u.y));
        x[i] = make_float4(a.x * s, a.y * s, b.x * s, b.y * s);
    }
}

extern "C" {

// Gradient tables on the wire as halfs (multi-GPU all-reduce): out16[n] = half(x[n] * scale[0]) and back,
// x[n] = float(in16[n]) * mult / scale[0]; n a multiple of 4, 16-byte aligned buffers; scale on the device (pag_grad_scale).
int pag_pack_f16(const float* x, int64_t n, const float* scale, void* out16, void* stream) {
    if (n < 0 || (n & 3)) return PAG_ERR_ARG;
    if (n == 0) return PAG_OK;
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    pack_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float4*>(x), n / 4, scale, reinterpret_cast<uint2*>(out16));
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_unpack_f16(const void* in16, int64_t n, const float* scale, float mult, float* x, void* stream) {
    if (n < 0 || (n & 3)) return PAG_ERR_ARG;
    if (n == 0) return PAG_OK;
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    unpack_f16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const uint2*>(in16), n / 4, scale, mult, reinterpret_cast<float4*>(x));
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// out_scale[0] = power-of-two loss scale 2^floor(log2(target / max|.|)) for the gradients a[na] (and b[nb], nullable).
// scratch: TWO uint32 on the device, zero on entry; the kernel leaves them zero again (one launch, no memset).
// With m_dev the element counts are min(na, m_dev[0]*wa) / min(nb, m_dev[0]*wb) (packed-sample tensors of width wa / wb).
int pag_grad_scale(const float* a, int64_t na, int wa, const float* b, int64_t nb, int wb, const int64_t* m_dev,
                   float target, uint32_t* scratch, float* out_scale, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    if (!b) nb = 0;
    const int64_t n = na + nb;
    int grid = (int)((n / 4 + 255) / 256);
    if (grid > 148 * 8) grid = 148 * 8;
    if (grid < 1) grid = 1;
    absmax_scale_kernel<<<grid, 256, 0, st>>>(a, na, b, nb, m_dev, wa, wb, target, scratch, out_scale);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// debugging aid: cudaStreamCaptureStatus of `stream` (0 none, 1 active, 2 invalidated), or -(cudaError) on failure
int pag_capture_status(void* stream) {
    cudaStreamCaptureStatus st;
    cudaError_t e = cudaStreamIsCapturing((cudaStream_t)stream, &st);
    if (e != cudaSuccess) return -(int)e;
    return (int)st;
}

int pag_ray_offsets(const int64_t* ridx, int64_t M, int64_t R, int64_t* offsets /*[R+1]*/, void* stream) {
    ray_offsets_kernel<<<pag_grid(M + 1, 256), 256, 0, (cudaStream_t)stream>>>(ridx, M, R, offsets);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_composite_fwd(const float* sigma, const float* deltas, const float* depths, const float* rgb,
                      const float* sem, int Cs, const float* inst, int Ci, const int64_t* offsets, int64_t R,
                      int bg_white, float* w, float* T, float* alpha, uint8_t* hit, float* rgb_out, float* rgbsum_out,
                      float* depth_out, float* sem_out, float* inst_out, void* stream) {
    if (R == 0) return PAG_OK;
    if ((rgb && (!rgb_out || !rgbsum_out)) || (depths && !depth_out) || (sem && !sem_out) || (inst && !inst_out))
        return PAG_ERR_ARG;
    composite_fwd_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        sigma, deltas, depths, rgb, sem, Cs, inst, Ci, offsets, R, bg_white, w, T, alpha, hit, rgb_out, rgbsum_out,
        depth_out, sem_out, inst_out);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// g_sigma[M] always written for packed samples; g_rgb_s[M,3] iff g_rgb; g_sem_s / g_inst_s iff their grads.
int pag_composite_bwd(const float* sigma, const float* deltas, const float* depths, const float* rgb,
                      const int64_t* offsets, int64_t R, int bg_white, const float* w, const float* T,
                      const float* alpha, const float* rgbsum, const float* g_alpha, const float* g_rgb,
                      const float* g_depth, const float* g_sem, int Cs, const float* g_inst, int Ci, float* g_sigma,
                      float* g_rgb_s, float* g_sem_s, float* g_inst_s, void* stream) {
    if (R == 0) return PAG_OK;
    if ((g_rgb && (!rgb || !g_rgb_s || !rgbsum)) || (g_depth && !depths) || (g_sem && !g_sem_s) || (g_inst && !g_inst_s))
        return PAG_ERR_ARG;
    composite_bwd_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(
        sigma, deltas, g_depth ? depths : nullptr, rgb, offsets, R, bg_white, w, T, alpha, rgbsum, g_alpha, g_rgb,
        g_depth, g_sem, Cs, g_inst, Ci, g_sigma, g_rgb_s, g_sem_s, g_inst_s);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

int pag_sum_reduce_fwd(const float* x, int64_t C, const int64_t* offsets, int64_t R, float* out, void* stream) {
    if (R == 0 || C == 0) return PAG_OK;
    sum_reduce_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(x, (int)C, offsets, R, out);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_sum_reduce_bwd(const float* g, int64_t C, const int64_t* offsets, int64_t R, float* gx, void* stream) {
    if (R == 0 || C == 0) return PAG_OK;
    sum_reduce_bwd_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(g, (int)C, offsets, R, gx);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_expint_fwd(const float* tau, const int64_t* offsets, int64_t R, float* w, float* T, void* stream) {
    if (R == 0) return PAG_OK;
    expint_fwd_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(tau, offsets, R, w, T);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_expint_bwd(const float* gw, const float* w, const float* T, const int64_t* offsets, int64_t R, float* gtau,
                   void* stream) {
    if (R == 0) return PAG_OK;
    expint_bwd_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(gw, w, T, offsets, R, gtau);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
