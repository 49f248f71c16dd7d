// Fused shallow MLP decoders of the panoptic field, forward + backward, FP32 FMA path.
//
// Replaces the 4 wisp BasicDecoders + view PositionalEmbedder + activations that
// PanopticNeF / PanopticDeltaNeF.rgb_semantics run as 9 cuBLAS GEMMs and ~15 elementwise kernels
// (reference pc_nerf/panoptic_nef.py:114-164,309-361; pc_nerf/panoptic_delta_nef.py:184-257):
//   "dc"  kernel: density  IN->64->16 (ch0 -> relu -> sigma)  +  color [16 | PE(-d) 27] ->64->64->3 sigmoid
//   "pan" kernel: panop = feats(.detach) + delta_feats;  semantics IN->64->Cs (softmax)
//                                                        instance  IN->64->64->Ci (softmax, temperature)
// This file is the exact-FP32 path (1e-4 parity, validation / no-autocast mode).  One thread owns
// one sample: its activations sit in registers, every layer streams weight rows from shared
// memory as broadcast LDS.128 (weights are staged once per persistent CTA, rows padded to 16 B),
// layer outputs bounce through a private shared-memory column.  Backward recomputes the hidden
// activations (cheaper than 1.5 KB/sample of saved state), forms dX with the same row-major
// weights (no transpose), and accumulates dW with a cooperative register-tiled (4x4) product
// over the CTA's 128-sample tile straight out of the staging columns.
#include "decoder_common.cuh"
#include <cuda_fp16.h>

// cooperative copy of W[rows][cols] (global) into smem [pad4(rows)][colsp], zero padded
__device__ __forceinline__ void stage_weights(float* dst, const float* __restrict__ W, int rows, int cols, int colsp) {
    const int rp = pad4(rows);
    for (int i = threadIdx.x; i < rp * colsp; i += blockDim.x) {
        const int r = i / colsp, c = i - r * colsp;
        dst[i] = (r < rows && c < cols) ? __ldg(W + (size_t)r * cols + c) : 0.f;
    }
}
__device__ __forceinline__ void stage_bias(float* dst, const float* __restrict__ b, int n) {
    for (int i = threadIdx.x; i < pad4(n); i += blockDim.x) dst[i] = (i < n) ? __ldg(b + i) : 0.f;
}

// y[j] = act(b[j] + sum_k W[j][k] x[k]) for j < pad4(OUT); written to column `dst` (stride LD)
template <int INP, int LD>
__device__ __forceinline__ void fwd_layer(const float (&x)[INP], const float* __restrict__ Ws,
                                          const float* __restrict__ bs, int OUT, float* __restrict__ dst, bool relu) {
    for (int j0 = 0; j0 < OUT; j0 += 4) {
        float a0 = bs[j0], a1 = bs[j0 + 1], a2 = bs[j0 + 2], a3 = bs[j0 + 3];
        const float4* r0 = reinterpret_cast<const float4*>(Ws + (size_t)j0 * INP);
        const float4* r1 = r0 + INP / 4;
        const float4* r2 = r1 + INP / 4;
        const float4* r3 = r2 + INP / 4;
#pragma unroll
        for (int k = 0; k < INP / 4; ++k) {
            const float4 w0 = r0[k], w1 = r1[k], w2 = r2[k], w3 = r3[k];
            a0 = fmaf(w0.x, x[4 * k], a0); a0 = fmaf(w0.y, x[4 * k + 1], a0); a0 = fmaf(w0.z, x[4 * k + 2], a0); a0 = fmaf(w0.w, x[4 * k + 3], a0);
            a1 = fmaf(w1.x, x[4 * k], a1); a1 = fmaf(w1.y, x[4 * k + 1], a1); a1 = fmaf(w1.z, x[4 * k + 2], a1); a1 = fmaf(w1.w, x[4 * k + 3], a1);
            a2 = fmaf(w2.x, x[4 * k], a2); a2 = fmaf(w2.y, x[4 * k + 1], a2); a2 = fmaf(w2.z, x[4 * k + 2], a2); a2 = fmaf(w2.w, x[4 * k + 3], a2);
            a3 = fmaf(w3.x, x[4 * k], a3); a3 = fmaf(w3.y, x[4 * k + 1], a3); a3 = fmaf(w3.z, x[4 * k + 2], a3); a3 = fmaf(w3.w, x[4 * k + 3], a3);
        }
        if (relu) { a0 = fmaxf(a0, 0.f); a1 = fmaxf(a1, 0.f); a2 = fmaxf(a2, 0.f); a3 = fmaxf(a3, 0.f); }
        dst[(j0 + 0) * LD] = a0; dst[(j0 + 1) * LD] = a1; dst[(j0 + 2) * LD] = a2; dst[(j0 + 3) * LD] = a3;
    }
}

template <int N, int LD>
__device__ __forceinline__ void load_col(float (&x)[N], const float* __restrict__ col) {
#pragma unroll
    for (int k = 0; k < N; ++k) x[k] = col[k * LD];
}

// dx[k] += sum_{j<OUT} W[j][k] * g[j]   (g from column `gcol`)
template <int INP, int LD>
__device__ __forceinline__ void bwd_layer(float (&dx)[INP], const float* __restrict__ Ws, int OUT,
                                          const float* __restrict__ gcol) {
    for (int j = 0; j < OUT; ++j) {
        const float g = gcol[j * LD];
        const float4* r = reinterpret_cast<const float4*>(Ws + (size_t)j * INP);
#pragma unroll
        for (int k = 0; k < INP / 4; ++k) {
            const float4 w = r[k];
            dx[4 * k] = fmaf(w.x, g, dx[4 * k]); dx[4 * k + 1] = fmaf(w.y, g, dx[4 * k + 1]);
            dx[4 * k + 2] = fmaf(w.z, g, dx[4 * k + 2]); dx[4 * k + 3] = fmaf(w.w, g, dx[4 * k + 3]);
        }
    }
}

// dW[j][k] += sum_s G[j][s] A[k][s]  (j<RJ, k<RK), db[j] += sum_s G[j][s];  G/A are staging rows [.][LD]
template <int LD, int NTHREADS>
__device__ void accum_dw(const float* __restrict__ Gs, int RJ, const float* __restrict__ As, int RK,
                         float* __restrict__ dW, int ldw, float* __restrict__ db) {
    const int NT = LD - 4;
    const int sj = (RJ + 3) >> 2, sk = (RK + 3) >> 2;
    for (int t = threadIdx.x; t < sj * sk; t += NTHREADS) {
        const int jq = t / sk, kq = t - jq * sk;
        const float4* g[4];
        const float4* a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            g[i] = reinterpret_cast<const float4*>(Gs + (size_t)min(jq + sj * i, RJ - 1) * LD);
            a[i] = reinterpret_cast<const float4*>(As + (size_t)min(kq + sk * i, RK - 1) * LD);
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
        for (int s = 0; s < NT / 4; ++s) {
            float4 gv[4], av[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) { gv[i] = g[i][s]; av[i] = a[i][s]; }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    acc[i][j] = fmaf(gv[i].x, av[j].x, acc[i][j]);
                    acc[i][j] = fmaf(gv[i].y, av[j].y, acc[i][j]);
                    acc[i][j] = fmaf(gv[i].z, av[j].z, acc[i][j]);
                    acc[i][j] = fmaf(gv[i].w, av[j].w, acc[i][j]);
                }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int jj = jq + sj * i, kk = kq + sk * j;
                if (jj < RJ && kk < RK) red_add_f32(dW + (size_t)jj * ldw + kk, acc[i][j]);
            }
    }
    if (db) {
        for (int j = threadIdx.x; j < RJ; j += NTHREADS) {
            const float4* g = reinterpret_cast<const float4*>(Gs + (size_t)j * LD);
            float s = 0.f;
            for (int i = 0; i < NT / 4; ++i) { const float4 v = g[i]; s += (v.x + v.y) + (v.z + v.w); }
            red_add_f32(db + j, s);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// density + color, forward
// ---------------------------------------------------------------------------------------------
template <int INP>
struct DcSmem {
    static constexpr int W_FLOATS = H * INP + H + DOUT * H + DOUT + H * CINP + H + H * H + H + 4 * H + 4;
};

template <int INP>
__device__ __forceinline__ void dc_stage(float* sm, const DcParams& p, int IN, float*& Wd1, float*& bd1, float*& Wd2,
                                         float*& bd2, float*& Wc1, float*& bc1, float*& Wc2, float*& bc2, float*& Wc3,
                                         float*& bc3) {
    Wd1 = sm; sm += H * INP; bd1 = sm; sm += H; Wd2 = sm; sm += DOUT * H; bd2 = sm; sm += DOUT;
    Wc1 = sm; sm += H * CINP; bc1 = sm; sm += H; Wc2 = sm; sm += H * H; bc2 = sm; sm += H;
    Wc3 = sm; sm += 4 * H; bc3 = sm; sm += 4;
    stage_weights(Wd1, p.Wd1, H, IN, INP); stage_bias(bd1, p.bd1, H);
    stage_weights(Wd2, p.Wd2, DOUT, H, H); stage_bias(bd2, p.bd2, DOUT);
    stage_weights(Wc1, p.Wc1, H, CIN, CINP); stage_bias(bc1, p.bc1, H);
    stage_weights(Wc2, p.Wc2, H, H, H); stage_bias(bc2, p.bc2, H);
    stage_weights(Wc3, p.Wc3, 3, H, H); stage_bias(bc3, p.bc3, 3);
}

template <int INP>
__device__ __forceinline__ void load_feats(float (&x)[INP], const float* __restrict__ a, const float* __restrict__ b,
                                           const float* __restrict__ lodw, int IN, int64_t m) {
#pragma unroll
    for (int k = 0; k < INP; ++k) {
        float v = 0.f;
        if (k < IN) {
            v = a[m * IN + k];
            if (b) v += b[m * IN + k];
            if (lodw) v *= __ldg(lodw + k);
        }
        x[k] = v;
    }
}

template <int INP, int NT>
__global__ void __launch_bounds__(NT) dc_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ lodw,
                                                    const float* __restrict__ ray_d, int S, int64_t M, int IN,
                                                    DcParams p, int want_rgb, float* __restrict__ sigma,
                                                    float* __restrict__ rgb) {
    constexpr int LD = NT + 4;
    extern __shared__ __align__(16) float smem[];
    float *Wd1, *bd1, *Wd2, *bd2, *Wc1, *bc1, *Wc2, *bc2, *Wc3, *bc3;
    dc_stage<INP>(smem, p, IN, Wd1, bd1, Wd2, bd2, Wc1, bc1, Wc2, bc2, Wc3, bc3);
    float* act = smem + DcSmem<INP>::W_FLOATS;  // [H][LD]
    __syncthreads();
    float* col = act + threadIdx.x;
    const int64_t ntiles = (M + NT - 1) / NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m = tile * NT + threadIdx.x;
        const bool valid = m < M;
        const int64_t mm = valid ? m : M - 1;
        {
            float x[INP];
            load_feats<INP>(x, feats, nullptr, lodw, IN, mm);
            fwd_layer<INP, LD>(x, Wd1, bd1, H, col, true);
        }
        float h[H];
        load_col<H, LD>(h, col);
        fwd_layer<H, LD>(h, Wd2, bd2, DOUT, col, false);
        if (valid) sigma[m] = fmaxf(col[0], 0.f);
        if (want_rgb) {
            float cin[CINP], pe[PE_DIM];
            const int64_t r = mm / S;
            view_embed(ray_d[3 * r], ray_d[3 * r + 1], ray_d[3 * r + 2], pe);
#pragma unroll
            for (int k = 0; k < DOUT; ++k) cin[k] = col[k * LD];
#pragma unroll
            for (int k = 0; k < PE_DIM; ++k) cin[DOUT + k] = pe[k];
            cin[CIN] = 0.f;
            fwd_layer<CINP, LD>(cin, Wc1, bc1, H, col, true);
            load_col<H, LD>(h, col);
            fwd_layer<H, LD>(h, Wc2, bc2, H, col, true);
            load_col<H, LD>(h, col);
#pragma unroll
            for (int j = 0; j < 3; ++j) {
                float a = bc3[j];
#pragma unroll
                for (int k = 0; k < H; ++k) a = fmaf(Wc3[j * H + k], h[k], a);
                if (valid) rgb[3 * m + j] = 1.f / (1.f + expf(-a));
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// density + color, backward
// ---------------------------------------------------------------------------------------------
template <int INP, int NT>
__global__ void __launch_bounds__(NT) dc_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ lodw,
                                                    const float* __restrict__ ray_d, int S, int64_t M, int IN,
                                                    DcParams p, const float* __restrict__ g_sigma,
                                                    const float* __restrict__ g_rgb, float* __restrict__ g_feats,
                                                    float* __restrict__ g_dir) {
    constexpr int LD = NT + 4;
    extern __shared__ __align__(16) float smem[];
    float *Wd1, *bd1, *Wd2, *bd2, *Wc1, *bc1, *Wc2, *bc2, *Wc3, *bc3;
    dc_stage<INP>(smem, p, IN, Wd1, bd1, Wd2, bd2, Wc1, bc1, Wc2, bc2, Wc3, bc3);
    float* Xs = smem + DcSmem<INP>::W_FLOATS;   // [INP][LD]
    float* Ad = Xs + INP * LD;                  // [H][LD]   density hidden / later its gradient
    float* Ci = Ad + H * LD;                    // [CINP][LD] color input (y16 | pe | 0) / later dY16 in rows 0..15
    float* A1 = Ci + CINP * LD;                 // [H][LD]
    float* A2 = A1 + H * LD;                    // [H][LD]
    float* Go = A2 + H * LD;                    // [4][LD]
    __syncthreads();
    const int tid = threadIdx.x;
    const bool do_rgb = g_rgb != nullptr;
    const int64_t ntiles = (M + NT - 1) / NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m = tile * NT + tid;
        const bool valid = m < M;
        const int64_t mm = valid ? m : M - 1;
        float y0;  // pre-relu density channel
        {   // ---- recompute forward, fill staging
            float x[INP];
            load_feats<INP>(x, feats, nullptr, lodw, IN, mm);
#pragma unroll
            for (int k = 0; k < INP; ++k) Xs[k * LD + tid] = x[k];
            fwd_layer<INP, LD>(x, Wd1, bd1, H, Ad + tid, true);
            float h[H];
            load_col<H, LD>(h, Ad + tid);
            fwd_layer<H, LD>(h, Wd2, bd2, DOUT, Ci + tid, false);
            y0 = Ci[tid];
            if (do_rgb) {
                float cin[CINP], pe[PE_DIM];
                const int64_t r = mm / S;
                view_embed(ray_d[3 * r], ray_d[3 * r + 1], ray_d[3 * r + 2], pe);
#pragma unroll
                for (int k = 0; k < DOUT; ++k) cin[k] = Ci[k * LD + tid];
#pragma unroll
                for (int k = 0; k < PE_DIM; ++k) cin[DOUT + k] = pe[k];
                cin[CIN] = 0.f;
#pragma unroll
                for (int k = DOUT; k < CINP; ++k) Ci[k * LD + tid] = cin[k];
                fwd_layer<CINP, LD>(cin, Wc1, bc1, H, A1 + tid, true);
                load_col<H, LD>(h, A1 + tid);
                fwd_layer<H, LD>(h, Wc2, bc2, H, A2 + tid, true);
                load_col<H, LD>(h, A2 + tid);
#pragma unroll
                for (int j = 0; j < 3; ++j) {
                    float a = bc3[j];
#pragma unroll
                    for (int k = 0; k < H; ++k) a = fmaf(Wc3[j * H + k], h[k], a);
                    const float c = 1.f / (1.f + expf(-a));
                    Go[j * LD + tid] = valid ? g_rgb[3 * m + j] * c * (1.f - c) : 0.f;
                }
                Go[3 * LD + tid] = 0.f;
            }
        }
        float dy[DOUT];
#pragma unroll
        for (int k = 0; k < DOUT; ++k) dy[k] = 0.f;
        if (do_rgb) {
            __syncthreads();
            accum_dw<LD, NT>(Go, 3, A2, H, p.gWc3, H, p.gbc3);
            __syncthreads();
            {
                float g[H];
#pragma unroll
                for (int k = 0; k < H; ++k) g[k] = 0.f;
                bwd_layer<H, LD>(g, Wc3, 3, Go + tid);
#pragma unroll
                for (int k = 0; k < H; ++k) A2[k * LD + tid] = (A2[k * LD + tid] > 0.f) ? g[k] : 0.f;
            }
            __syncthreads();
            accum_dw<LD, NT>(A2, H, A1, H, p.gWc2, H, p.gbc2);
            __syncthreads();
            {
                float g[H];
#pragma unroll
                for (int k = 0; k < H; ++k) g[k] = 0.f;
                bwd_layer<H, LD>(g, Wc2, H, A2 + tid);
#pragma unroll
                for (int k = 0; k < H; ++k) A1[k * LD + tid] = (A1[k * LD + tid] > 0.f) ? g[k] : 0.f;
            }
            __syncthreads();
            accum_dw<LD, NT>(A1, H, Ci, CIN, p.gWc1, CIN, p.gbc1);
            __syncthreads();
            {
                float g[CINP];
#pragma unroll
                for (int k = 0; k < CINP; ++k) g[k] = 0.f;
                bwd_layer<CINP, LD>(g, Wc1, H, A1 + tid);
#pragma unroll
                for (int k = 0; k < DOUT; ++k) dy[k] = g[k];
                if (g_dir && valid) {
                    const int64_t r = m / S;
                    const float v[3] = {-ray_d[3 * r], -ray_d[3 * r + 1], -ray_d[3 * r + 2]};
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float gv = g[DOUT + c];
#pragma unroll
                        for (int f = 0; f < PE_F; ++f) {
                            const float b = (float)(1 << f), a = v[c] * b;
                            gv += b * (cosf(a) * g[DOUT + 3 + 3 * f + c] - sinf(a) * g[DOUT + 3 + 3 * PE_F + 3 * f + c]);
                        }
                        g_dir[3 * m + c] = -gv;
                    }
                }
            }
        } else if (g_dir && valid) {
            g_dir[3 * m] = 0.f; g_dir[3 * m + 1] = 0.f; g_dir[3 * m + 2] = 0.f;
        }
        if (g_sigma && valid && y0 > 0.f) dy[0] += g_sigma[m];
        if (!valid) {
#pragma unroll
            for (int k = 0; k < DOUT; ++k) dy[k] = 0.f;
        }
#pragma unroll
        for (int k = 0; k < DOUT; ++k) Ci[k * LD + tid] = dy[k];
        __syncthreads();
        accum_dw<LD, NT>(Ci, DOUT, Ad, H, p.gWd2, H, p.gbd2);
        __syncthreads();
        {
            float g[H];
#pragma unroll
            for (int k = 0; k < H; ++k) g[k] = 0.f;
            bwd_layer<H, LD>(g, Wd2, DOUT, Ci + tid);
#pragma unroll
            for (int k = 0; k < H; ++k) Ad[k * LD + tid] = (Ad[k * LD + tid] > 0.f) ? g[k] : 0.f;
        }
        __syncthreads();
        accum_dw<LD, NT>(Ad, H, Xs, IN, p.gWd1, IN, p.gbd1);
        if (g_feats) {
            float g[INP];
#pragma unroll
            for (int k = 0; k < INP; ++k) g[k] = 0.f;
            bwd_layer<INP, LD>(g, Wd1, H, Ad + tid);
            if (valid) {
#pragma unroll
                for (int k = 0; k < INP; ++k)
                    if (k < IN) g_feats[m * IN + k] = lodw ? g[k] * __ldg(lodw + k) : g[k];
            }
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------------------------------------
// semantics + instance, forward
// ---------------------------------------------------------------------------------------------
struct PanLayout {
    int Csp, Cip;
    int oWs1, obs1, oWs2, obs2, oWi1, obi1, oWi2, obi2, oWi3, obi3, total;
};
__host__ __device__ inline PanLayout pan_layout(int INP, int Cs, int Ci) {
    PanLayout l;
    l.Csp = pad4(Cs > 0 ? Cs : 1); l.Cip = pad4(Ci > 0 ? Ci : 1);
    int o = 0;
    l.oWs1 = o; o += H * INP; l.obs1 = o; o += H; l.oWs2 = o; o += l.Csp * H; l.obs2 = o; o += l.Csp;
    l.oWi1 = o; o += H * INP; l.obi1 = o; o += H; l.oWi2 = o; o += H * H; l.obi2 = o; o += H;
    l.oWi3 = o; o += l.Cip * H; l.obi3 = o; o += l.Cip;
    l.total = o;
    return l;
}
__device__ __forceinline__ void pan_stage(float* sm, const PanLayout& l, const PanParams& p, int IN, int INP, int Cs, int Ci) {
    if (Cs > 0) {
        stage_weights(sm + l.oWs1, p.Ws1, H, IN, INP); stage_bias(sm + l.obs1, p.bs1, H);
        stage_weights(sm + l.oWs2, p.Ws2, Cs, H, H); stage_bias(sm + l.obs2, p.bs2, Cs);
    }
    if (Ci > 0) {
        stage_weights(sm + l.oWi1, p.Wi1, H, IN, INP); stage_bias(sm + l.obi1, p.bi1, H);
        stage_weights(sm + l.oWi2, p.Wi2, H, H, H); stage_bias(sm + l.obi2, p.bi2, H);
        stage_weights(sm + l.oWi3, p.Wi3, Ci, H, H); stage_bias(sm + l.obi3, p.bi3, Ci);
    }
}

// logits -> (optional) softmax, written to row `out` (global); two passes over the thread's own row
__device__ __forceinline__ void finish_row(float* __restrict__ out, int C, bool softmax, float inv_temp, float mx) {
    if (!softmax) {
        if (inv_temp != 1.f) for (int j = 0; j < C; ++j) out[j] *= inv_temp;
        return;
    }
    float s = 0.f;
    for (int j = 0; j < C; ++j) { const float e = expf((out[j] - mx) * inv_temp); out[j] = e; s += e; }
    const float inv = 1.f / s;
    for (int j = 0; j < C; ++j) out[j] *= inv;
}

// last layer: logits[j] = b[j] + W[j].h  streamed 4 rows at a time to the global row; returns the row max
template <int LD>
__device__ __forceinline__ float logits_to_row(const float (&h)[H], const float* __restrict__ Ws,
                                               const float* __restrict__ bs, int C, float* __restrict__ out, bool valid) {
    float mx = -INFINITY;
    for (int j0 = 0; j0 < C; j0 += 4) {
        float a[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) a[i] = bs[j0 + i];
        const float4* r = reinterpret_cast<const float4*>(Ws + (size_t)j0 * H);
#pragma unroll
        for (int k = 0; k < H / 4; ++k) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float4 w = r[i * (H / 4) + k];
                a[i] = fmaf(w.x, h[4 * k], a[i]); a[i] = fmaf(w.y, h[4 * k + 1], a[i]);
                a[i] = fmaf(w.z, h[4 * k + 2], a[i]); a[i] = fmaf(w.w, h[4 * k + 3], a[i]);
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
            if (j0 + i < C) { mx = fmaxf(mx, a[i]); if (valid) out[j0 + i] = a[i]; }
    }
    return mx;
}

template <int INP, int NT>
__global__ void __launch_bounds__(NT) pan_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ dfeats,
                                                     const float* __restrict__ lodw, int64_t M, int IN, PanParams p,
                                                     int Cs, int Ci, int sem_softmax, int inst_softmax, float inst_inv_temp,
                                                     float* __restrict__ sem, float* __restrict__ inst) {
    constexpr int LD = NT + 4;
    extern __shared__ __align__(16) float smem[];
    const PanLayout l = pan_layout(INP, Cs, Ci);
    pan_stage(smem, l, p, IN, INP, Cs, Ci);
    float* act = smem + l.total;  // [H][LD]
    __syncthreads();
    float* col = act + threadIdx.x;
    const int64_t ntiles = (M + NT - 1) / NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m = tile * NT + threadIdx.x;
        const bool valid = m < M;
        const int64_t mm = valid ? m : M - 1;
        float x[INP];
        load_feats<INP>(x, feats, dfeats, lodw, IN, mm);
        float h[H];
        if (Cs > 0) {
            fwd_layer<INP, LD>(x, smem + l.oWs1, smem + l.obs1, H, col, true);
            load_col<H, LD>(h, col);
            float* row = sem + mm * Cs;
            const float mx = logits_to_row<LD>(h, smem + l.oWs2, smem + l.obs2, Cs, row, valid);
            if (valid) finish_row(row, Cs, sem_softmax, 1.f, mx);
        }
        if (Ci > 0) {
            fwd_layer<INP, LD>(x, smem + l.oWi1, smem + l.obi1, H, col, true);
            load_col<H, LD>(h, col);
            fwd_layer<H, LD>(h, smem + l.oWi2, smem + l.obi2, H, col, true);
            load_col<H, LD>(h, col);
            float* row = inst + mm * Ci;
            const float mx = logits_to_row<LD>(h, smem + l.oWi3, smem + l.obi3, Ci, row, valid);
            if (valid) finish_row(row, Ci, inst_softmax, inst_inv_temp, mx);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// semantics + instance, backward
// ---------------------------------------------------------------------------------------------
#define GO_ROWS 40

// d logits for rows [c0, c0+n) of one head into Go (softmax backward from the saved probabilities)
template <int LD>
__device__ __forceinline__ void dlogits_chunk(float* __restrict__ Go, int tid, const float* __restrict__ prob,
                                              const float* __restrict__ g, int c0, int n, int C, bool softmax,
                                              float inv_temp, float dot, bool valid) {
    for (int jj = 0; jj < GO_ROWS; ++jj) {
        float d = 0.f;
        const int j = c0 + jj;
        if (valid && jj < n && j < C) {
            d = softmax ? prob[j] * (g[j] - dot) : g[j];
            d *= inv_temp;
        }
        if (jj < pad4(n)) Go[jj * LD + tid] = d;
    }
}

template <int INP, int NT>
__global__ void __launch_bounds__(NT) pan_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ dfeats,
                                                     const float* __restrict__ lodw, int64_t M, int IN, PanParams p,
                                                     int Cs, int Ci, int sem_softmax, int inst_softmax, float inst_inv_temp,
                                                     const float* __restrict__ sem, const float* __restrict__ inst,
                                                     const float* __restrict__ g_sem, const float* __restrict__ g_inst,
                                                     float* __restrict__ g_panop) {
    constexpr int LD = NT + 4;
    extern __shared__ __align__(16) float smem[];
    const PanLayout l = pan_layout(INP, Cs, Ci);
    pan_stage(smem, l, p, IN, INP, Cs, Ci);
    float* Xs = smem + l.total;     // [INP][LD]
    float* A1 = Xs + INP * LD;      // [H][LD]
    float* A2 = A1 + H * LD;        // [H][LD]
    float* Go = A2 + H * LD;        // [GO_ROWS][LD]
    __syncthreads();
    const int tid = threadIdx.x;
    const bool do_sem = (Cs > 0) && g_sem, do_inst = (Ci > 0) && g_inst;
    const int64_t ntiles = (M + NT - 1) / NT;
    for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int64_t m = tile * NT + tid;
        const bool valid = m < M;
        const int64_t mm = valid ? m : M - 1;
        float x[INP], dx[INP];
        load_feats<INP>(x, feats, dfeats, lodw, IN, mm);
#pragma unroll
        for (int k = 0; k < INP; ++k) { Xs[k * LD + tid] = x[k]; dx[k] = 0.f; }
        if (do_sem) {
            fwd_layer<INP, LD>(x, smem + l.oWs1, smem + l.obs1, H, A1 + tid, true);
            const float* pr = sem + mm * Cs;
            const float* gr = g_sem + mm * Cs;
            float dot = 0.f;
            if (sem_softmax) for (int j = 0; j < Cs; ++j) dot = fmaf(pr[j], gr[j], dot);
            for (int c0 = 0; c0 < Cs; c0 += GO_ROWS) {
                const int n = min(GO_ROWS, Cs - c0);
                dlogits_chunk<LD>(Go, tid, pr, gr, c0, n, Cs, sem_softmax, 1.f, dot, valid);
                __syncthreads();
                accum_dw<LD, NT>(Go, n, A1, H, p.gWs2 + (size_t)c0 * H, H, p.gbs2 + c0);
                // hidden gradient accumulates in A2 (free during the semantic head)
                {
                    float g[H];
                    if (c0 == 0) {
#pragma unroll
                        for (int k = 0; k < H; ++k) g[k] = 0.f;
                    } else load_col<H, LD>(g, A2 + tid);
                    bwd_layer<H, LD>(g, smem + l.oWs2 + (size_t)c0 * H, n, Go + tid);
#pragma unroll
                    for (int k = 0; k < H; ++k) A2[k * LD + tid] = g[k];
                }
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < H; ++k) A1[k * LD + tid] = (A1[k * LD + tid] > 0.f) ? A2[k * LD + tid] : 0.f;
            __syncthreads();
            accum_dw<LD, NT>(A1, H, Xs, IN, p.gWs1, IN, p.gbs1);
            bwd_layer<INP, LD>(dx, smem + l.oWs1, H, A1 + tid);
            __syncthreads();
        }
        if (do_inst) {
            float dh2[H];
            {
                fwd_layer<INP, LD>(x, smem + l.oWi1, smem + l.obi1, H, A1 + tid, true);
                float h[H];
                load_col<H, LD>(h, A1 + tid);
                fwd_layer<H, LD>(h, smem + l.oWi2, smem + l.obi2, H, A2 + tid, true);
            }
#pragma unroll
            for (int k = 0; k < H; ++k) dh2[k] = 0.f;
            const float* pr = inst + mm * Ci;
            const float* gr = g_inst + mm * Ci;
            float dot = 0.f;
            if (inst_softmax) for (int j = 0; j < Ci; ++j) dot = fmaf(pr[j], gr[j], dot);
            for (int c0 = 0; c0 < Ci; c0 += GO_ROWS) {
                const int n = min(GO_ROWS, Ci - c0);
                dlogits_chunk<LD>(Go, tid, pr, gr, c0, n, Ci, inst_softmax, inst_inv_temp, dot, valid);
                __syncthreads();
                accum_dw<LD, NT>(Go, n, A2, H, p.gWi3 + (size_t)c0 * H, H, p.gbi3 + c0);
                bwd_layer<H, LD>(dh2, smem + l.oWi3 + (size_t)c0 * H, n, Go + tid);
                __syncthreads();
            }
#pragma unroll
            for (int k = 0; k < H; ++k) A2[k * LD + tid] = (A2[k * LD + tid] > 0.f) ? dh2[k] : 0.f;
            __syncthreads();
            accum_dw<LD, NT>(A2, H, A1, H, p.gWi2, H, p.gbi2);
            __syncthreads();
            {
                float g[H];
#pragma unroll
                for (int k = 0; k < H; ++k) g[k] = 0.f;
                bwd_layer<H, LD>(g, smem + l.oWi2, H, A2 + tid);
#pragma unroll
                for (int k = 0; k < H; ++k) A1[k * LD + tid] = (A1[k * LD + tid] > 0.f) ? g[k] : 0.f;
            }
            __syncthreads();
            accum_dw<LD, NT>(A1, H, Xs, IN, p.gWi1, IN, p.gbi1);
            bwd_layer<INP, LD>(dx, smem + l.oWi1, H, A1 + tid);
            __syncthreads();
        }
        if (g_panop && valid) {
#pragma unroll
            for (int k = 0; k < INP; ++k)
                if (k < IN) g_panop[m * IN + k] = lodw ? dx[k] * __ldg(lodw + k) : dx[k];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int g_num_sms = 0;
static int num_sms() {
    if (!g_num_sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (g_num_sms <= 0) g_num_sms = 148;
    }
    return g_num_sms;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
    if (bytes > 227 * 1024) return PAG_ERR_UNSUPPORTED;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    return e == cudaSuccess ? PAG_OK : (int)e;
}

#define FWD_NT 256
#define BWD_NT 128

template <int INP>
static int dc_fwd_launch(const float* feats, const float* lodw, const float* ray_d, int S, int64_t M, int IN,
                         const DcParams& p, int want_rgb, float* sigma, float* rgb, cudaStream_t st) {
    const size_t bytes = (DcSmem<INP>::W_FLOATS + (size_t)H * (FWD_NT + 4)) * sizeof(float);
    int rc = set_smem(dc_fwd_kernel<INP, FWD_NT>, bytes);
    if (rc) return rc;
    const int64_t tiles = (M + FWD_NT - 1) / FWD_NT;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    dc_fwd_kernel<INP, FWD_NT><<<grid, FWD_NT, bytes, st>>>(feats, lodw, ray_d, S, M, IN, p, want_rgb, sigma, rgb);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
template <int INP>
static int dc_bwd_launch(const float* feats, const float* lodw, const float* ray_d, int S, int64_t M, int IN,
                         const DcParams& p, const float* g_sigma, const float* g_rgb, float* g_feats, float* g_dir,
                         cudaStream_t st) {
    const size_t rows = INP + H + CINP + H + H + 4;
    const size_t bytes = (DcSmem<INP>::W_FLOATS + rows * (BWD_NT + 4)) * sizeof(float);
    int rc = set_smem(dc_bwd_kernel<INP, BWD_NT>, bytes);
    if (rc) return rc;
    const int64_t tiles = (M + BWD_NT - 1) / BWD_NT;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    dc_bwd_kernel<INP, BWD_NT><<<grid, BWD_NT, bytes, st>>>(feats, lodw, ray_d, S, M, IN, p, g_sigma, g_rgb, g_feats, g_dir);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
template <int INP>
static int pan_fwd_launch(const float* feats, const float* dfeats, const float* lodw, int64_t M, int IN,
                          const PanParams& p, int Cs, int Ci, int ss, int is, float it, float* sem, float* inst,
                          cudaStream_t st) {
    const PanLayout l = pan_layout(INP, Cs, Ci);
    const size_t bytes = (l.total + (size_t)H * (FWD_NT + 4)) * sizeof(float);
    int rc = set_smem(pan_fwd_kernel<INP, FWD_NT>, bytes);
    if (rc) return rc;
    const int64_t tiles = (M + FWD_NT - 1) / FWD_NT;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    pan_fwd_kernel<INP, FWD_NT><<<grid, FWD_NT, bytes, st>>>(feats, dfeats, lodw, M, IN, p, Cs, Ci, ss, is, it, sem, inst);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
template <int INP>
static int pan_bwd_launch(const float* feats, const float* dfeats, const float* lodw, int64_t M, int IN,
                          const PanParams& p, int Cs, int Ci, int ss, int is, float it, const float* sem,
                          const float* inst, const float* g_sem, const float* g_inst, float* g_panop, cudaStream_t st) {
    const PanLayout l = pan_layout(INP, Cs, Ci);
    const size_t rows = INP + H + H + GO_ROWS;
    const size_t bytes = (l.total + rows * (BWD_NT + 4)) * sizeof(float);
    int rc = set_smem(pan_bwd_kernel<INP, BWD_NT>, bytes);
    if (rc) return rc;
    const int64_t tiles = (M + BWD_NT - 1) / BWD_NT;
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    pan_bwd_kernel<INP, BWD_NT><<<grid, BWD_NT, bytes, st>>>(feats, dfeats, lodw, M, IN, p, Cs, Ci, ss, is, it, sem, inst,
                                                            g_sem, g_inst, g_panop);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

// ---------------------------------------------------------------------------------------------
// linear scalar head: y[m] = b + sum_k (feats[m,k] + dfeats[m,k]) * lodw[k] * w[k]
// PanopticDDensityNeF's delta-density decoder (pc_nerf/panoptic_dd_nef.py:41-58) is a BasicDecoder with activation 'none'
// (Identity): feat -> 64 -> 1 without a nonlinearity is ONE linear map, w = W2 W1, b = W2 b1 + b2 (collapsed on the host,
// where autograd carries the gradient back to W1 / b1 / W2 / b2).  Thread per sample, float4 rows.
// ---------------------------------------------------------------------------------------------
// pre (nullable): added before the optional ReLU; post (nullable): per-sample factor applied after it -- the DD field's
// tau_p = relu(y0.detach() + delta_density) * delta in one pass.  m_dev (nullable): device-side sample count.
__global__ void linear_head_fwd_kernel(const float* __restrict__ feats, const float* __restrict__ dfeats,
                                       const float* __restrict__ lodw, int64_t M, int IN, const float* __restrict__ w,
                                       const float* __restrict__ b, float* __restrict__ y, const float* __restrict__ pre,
                                       int relu, const float* __restrict__ post, const int64_t* __restrict__ m_dev) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m_dev) M = min(M, __ldg(m_dev));
    if (m >= M) return;
    float acc = __ldg(b);
    for (int k = 0; k < IN; ++k) {
        float x = feats[m * IN + k];
        if (dfeats) x += dfeats[m * IN + k];
        if (lodw) x *= __ldg(lodw + k);
        acc = fmaf(x, __ldg(w + k), acc);
    }
    if (pre) acc += pre[m];
    if (relu) acc = fmaxf(acc, 0.f);
    if (post) acc *= post[m];
    y[m] = acc;
}
// g_x[m,k] (+)= g[m] * lodw[k] * w[k] (same for feats and dfeats); g_w[k] += sum_m g[m] * x[m,k]; g_b += sum_m g[m].
// gate (nullable): the forward output; g is zeroed where gate <= 0 (ReLU) ; post as in the forward.
// Persistent grid, weight-gradient partials of all (<= 64) input features in registers, one reduction + IN atomics per CTA.
#define LHR_MAXIN 64
__global__ void __launch_bounds__(256) linear_head_bwd_kernel(const float* __restrict__ feats, const float* __restrict__ dfeats,
                                                              const float* __restrict__ lodw, int64_t M, int IN,
                                                              const float* __restrict__ w, const float* __restrict__ g,
                                                              float* __restrict__ g_x, float* __restrict__ g_w, float* __restrict__ g_b,
                                                              const float* __restrict__ gate, const float* __restrict__ post,
                                                              int accumulate_x, const int64_t* __restrict__ m_dev) {
    __shared__ float red_s[8][LHR_MAXIN + 1];
    if (m_dev) M = min(M, __ldg(m_dev));
    float acc[LHR_MAXIN];
#pragma unroll
    for (int k = 0; k < LHR_MAXIN; ++k) acc[k] = 0.f;
    float accb = 0.f;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        float gm = g[m];
        if (post) gm *= post[m];
        if (gate && !(gate[m] > 0.f)) gm = 0.f;
        accb += gm;
#pragma unroll
        for (int k = 0; k < LHR_MAXIN; ++k) {
            if (k < IN) {
                const float lw = lodw ? __ldg(lodw + k) : 1.f;
                float x = feats[m * IN + k];
                if (dfeats) x += dfeats[m * IN + k];
                if (g_x) {
                    const float v = gm * lw * __ldg(w + k);
                    if (accumulate_x) g_x[m * IN + k] += v; else g_x[m * IN + k] = v;
                }
                acc[k] = fmaf(gm * x, lw, acc[k]);
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < LHR_MAXIN; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red_s[wid][k] = v;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accb += __shfl_xor_sync(0xffffffffu, accb, o);
    if (lane == 0) red_s[wid][LHR_MAXIN] = accb;
    __syncthreads();
    for (int i = threadIdx.x; i <= LHR_MAXIN; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += red_s[k][i];
        if (v != 0.f) {
            if (i < IN) red_add_f32(g_w + i, v);
            else if (i == LHR_MAXIN) red_add_f32(g_b, v);
        }
    }
}
// the same head on the fp16 operand images of the fused trace (tile t = m / 128: [IN/8 chunks][128 rows][8 halfs], see
// pag_permuto_fwd_img16_dyn): coalesced 16-byte accesses; x = half(feats) + half(dfeats) exactly as the heads kernels form it.
// The backward adds its input gradient, multiplied by the loss scale *img_scale the heads' dX image carries, into that image.
__device__ __forceinline__ void unpack8(const uint4& u, float (&x)[8]) {
    const __half2* h = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 f = __half22float2(h[i]); x[2 * i] = f.x; x[2 * i + 1] = f.y; }
}
__device__ __forceinline__ uint4 hadd8(const uint4& a, const uint4& b) {
    uint4 r;
    const __half2 *x = reinterpret_cast<const __half2*>(&a), *y = reinterpret_cast<const __half2*>(&b);
    __half2* o = reinterpret_cast<__half2*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = __hadd2(x[i], y[i]);
    return r;
}
__global__ void linear_head_fwd_img_kernel(const uint4* __restrict__ fimg, const uint4* __restrict__ dimg,
                                           const float* __restrict__ lodw, int64_t M, int IN, const float* __restrict__ w,
                                           const float* __restrict__ b, float* __restrict__ y, const float* __restrict__ pre,
                                           int relu, const float* __restrict__ post, const int64_t* __restrict__ m_dev) {
    const int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (m_dev) M = min(M, __ldg(m_dev));
    if (m >= M) return;
    const int C = IN >> 3;
    const int64_t base = (m >> 7) * C * 128 + (m & 127);
    float acc = __ldg(b);
    for (int c = 0; c < C; ++c) {
        uint4 u = __ldg(fimg + base + c * 128);
        if (dimg) u = hadd8(u, __ldg(dimg + base + c * 128));
        float x[8];
        unpack8(u, x);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(x[k] * (lodw ? __ldg(lodw + 8 * c + k) : 1.f), __ldg(w + 8 * c + k), acc);
    }
    if (pre) acc += pre[m];
    if (relu) acc = fmaxf(acc, 0.f);
    if (post) acc *= post[m];
    y[m] = acc;
}
// Persistent grid: every thread walks samples with a grid stride and keeps the weight-gradient partials of all (<= 64) input
// features in registers; one shuffle / shared-memory reduction and IN atomics per CTA at the end (a per-warp, per-feature
// reduction with an atomic each made 590 k atomics on 48 addresses: 0.52 ms for 392 k samples).
#define LH_MAXC 8
__global__ void __launch_bounds__(256) linear_head_bwd_img_kernel(const uint4* __restrict__ fimg, const uint4* __restrict__ dimg,
                                                                  const float* __restrict__ lodw, int64_t M, int IN,
                                                                  const float* __restrict__ w, const float* __restrict__ g,
                                                                  uint4* __restrict__ gx_img, const float* __restrict__ img_scale,
                                                                  float* __restrict__ g_w, float* __restrict__ g_b,
                                                                  const float* __restrict__ gate, const float* __restrict__ post,
                                                                  const int64_t* __restrict__ m_dev) {
    __shared__ float red_s[8][8 * LH_MAXC + 1];
    if (m_dev) M = min(M, __ldg(m_dev));
    const int C = IN >> 3;
    const float scale = img_scale ? __ldg(img_scale) : 1.f;
    float acc[LH_MAXC][8];
#pragma unroll
    for (int c = 0; c < LH_MAXC; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[c][k] = 0.f;
    float accb = 0.f;
    for (int64_t m = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; m < M; m += (int64_t)gridDim.x * blockDim.x) {
        float gm = g[m];
        if (post) gm *= post[m];
        if (gate && !(gate[m] > 0.f)) gm = 0.f;
        if (gm == 0.f) continue;
        const int64_t base = (m >> 7) * C * 128 + (m & 127);
        const float gs = gm * scale;
        accb += gm;
#pragma unroll
        for (int c = 0; c < LH_MAXC; ++c) {
            if (c < C) {
                uint4 u = __ldg(fimg + base + c * 128);
                if (dimg) u = hadd8(u, __ldg(dimg + base + c * 128));
                float x[8];
                unpack8(u, x);
                if (gx_img) {      // accumulate into the heads' dX image (scaled)
                    uint4 gi = gx_img[base + c * 128];
                    __half2* gh = reinterpret_cast<__half2*>(&gi);
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        const float l0 = lodw ? __ldg(lodw + 8 * c + 2 * k) : 1.f, l1 = lodw ? __ldg(lodw + 8 * c + 2 * k + 1) : 1.f;
                        const float2 f = __half22float2(gh[k]);
                        gh[k] = __floats2half2_rn(fmaf(gs * l0, __ldg(w + 8 * c + 2 * k), f.x), fmaf(gs * l1, __ldg(w + 8 * c + 2 * k + 1), f.y));
                    }
                    gx_img[base + c * 128] = gi;
                }
#pragma unroll
                for (int k = 0; k < 8; ++k) acc[c][k] = fmaf(gm * x[k], lodw ? __ldg(lodw + 8 * c + k) : 1.f, acc[c][k]);
            }
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int c = 0; c < LH_MAXC; ++c)
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            float v = acc[c][k];
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) red_s[wid][8 * c + k] = v;
        }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) accb += __shfl_xor_sync(0xffffffffu, accb, o);
    if (lane == 0) red_s[wid][8 * LH_MAXC] = accb;
    __syncthreads();
    for (int i = threadIdx.x; i <= 8 * LH_MAXC; i += blockDim.x) {
        float v = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) v += red_s[k][i];
        if (v != 0.f) {
            if (i < IN) red_add_f32(g_w + i, v);
            else if (i == 8 * LH_MAXC) red_add_f32(g_b, v);
        }
    }
}

// DD tracer backward glue: the panoptic outputs are out[ray] = alpha_p * sum_s w_p[s] f[s] with alpha_p = sum_s w_p[s] and the
// weights NOT detached (tracers/panoptic_dd_packed_rf_tracer.py:128-162).  Given the per-sample <f_s, g_ray> from the fused heads
// backward: d L / d w_p[s] = alpha_p * (gw_sem + gw_inst)[s] + d L / d alpha_p[ray],
//           d L / d alpha_p[ray] = sum_c g[ray][c] * out[ray][c] / alpha_p[ray].     One warp per ray.
__global__ void dd_weight_grads_kernel(const float* __restrict__ g_sem, const float* __restrict__ out_sem, int Cs,
                                       const float* __restrict__ g_inst, const float* __restrict__ out_inst, int Ci,
                                       const float* __restrict__ alpha_p, const float* __restrict__ gw_sem,
                                       const float* __restrict__ gw_inst, const int64_t* __restrict__ offsets, int64_t R,
                                       float* __restrict__ gw) {
    const int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (r >= R) return;
    const int64_t s0 = offsets[r], s1 = offsets[r + 1];
    if (s0 == s1) return;
    float acc = 0.f;
    if (g_sem) for (int c = lane; c < Cs; c += 32) acc = fmaf(g_sem[r * Cs + c], out_sem[r * Cs + c], acc);
    if (g_inst) for (int c = lane; c < Ci; c += 32) acc = fmaf(g_inst[r * Ci + c], out_inst[r * Ci + c], acc);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    const float a = alpha_p[r];
    const float ga = a != 0.f ? acc / a : 0.f;
    for (int64_t i = s0 + lane; i < s1; i += 32)
        gw[i] = a * ((gw_sem ? gw_sem[i] : 0.f) + (gw_inst ? gw_inst[i] : 0.f)) + ga;
}

extern "C" {

// weights: 10 pointers in the order Wd1,bd1,Wd2,bd2,Wc1,bc1,Wc2,bc2,Wc3,bc3 (torch Linear [out][in])
int pag_decode_dc_fwd(const float* feats, const float* lodw, const float* ray_d, int S, int64_t M, int IN,
                      const float* const* weights, int hidden, int view_dim, int want_rgb, float* sigma, float* rgb,
                      void* stream) {
    if (hidden != H || view_dim != PE_DIM) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    DcParams p{};
    p.Wd1 = weights[0]; p.bd1 = weights[1]; p.Wd2 = weights[2]; p.bd2 = weights[3]; p.Wc1 = weights[4];
    p.bc1 = weights[5]; p.Wc2 = weights[6]; p.bc2 = weights[7]; p.Wc3 = weights[8]; p.bc3 = weights[9];
    cudaStream_t st = (cudaStream_t)stream;
    switch (pad4(IN)) {
        case 12: return dc_fwd_launch<12>(feats, lodw, ray_d, S, M, IN, p, want_rgb, sigma, rgb, st);
        case 28: return dc_fwd_launch<28>(feats, lodw, ray_d, S, M, IN, p, want_rgb, sigma, rgb, st);
        case 32: return dc_fwd_launch<32>(feats, lodw, ray_d, S, M, IN, p, want_rgb, sigma, rgb, st);
        case 48: return dc_fwd_launch<48>(feats, lodw, ray_d, S, M, IN, p, want_rgb, sigma, rgb, st);
        default: return PAG_ERR_UNSUPPORTED;
    }
}

// grads: 10 pointers (same order) ACCUMULATED into; g_sigma / g_rgb / g_feats / g_dir nullable
int pag_decode_dc_bwd(const float* feats, const float* lodw, const float* ray_d, int S, int64_t M, int IN,
                      const float* const* weights, float* const* grads, int hidden, int view_dim,
                      const float* g_sigma, const float* g_rgb, float* g_feats, float* g_dir, void* stream) {
    if (hidden != H || view_dim != PE_DIM) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    DcParams p{};
    p.Wd1 = weights[0]; p.bd1 = weights[1]; p.Wd2 = weights[2]; p.bd2 = weights[3]; p.Wc1 = weights[4];
    p.bc1 = weights[5]; p.Wc2 = weights[6]; p.bc2 = weights[7]; p.Wc3 = weights[8]; p.bc3 = weights[9];
    p.gWd1 = grads[0]; p.gbd1 = grads[1]; p.gWd2 = grads[2]; p.gbd2 = grads[3]; p.gWc1 = grads[4];
    p.gbc1 = grads[5]; p.gWc2 = grads[6]; p.gbc2 = grads[7]; p.gWc3 = grads[8]; p.gbc3 = grads[9];
    cudaStream_t st = (cudaStream_t)stream;
    switch (pad4(IN)) {
        case 12: return dc_bwd_launch<12>(feats, lodw, ray_d, S, M, IN, p, g_sigma, g_rgb, g_feats, g_dir, st);
        case 28: return dc_bwd_launch<28>(feats, lodw, ray_d, S, M, IN, p, g_sigma, g_rgb, g_feats, g_dir, st);
        case 32: return dc_bwd_launch<32>(feats, lodw, ray_d, S, M, IN, p, g_sigma, g_rgb, g_feats, g_dir, st);
        case 48: return dc_bwd_launch<48>(feats, lodw, ray_d, S, M, IN, p, g_sigma, g_rgb, g_feats, g_dir, st);
        default: return PAG_ERR_UNSUPPORTED;
    }
}

// weights: Ws1,bs1,Ws2,bs2,Wi1,bi1,Wi2,bi2,Wi3,bi3; Cs or Ci may be 0 (head skipped)
int pag_decode_pan_fwd(const float* feats, const float* dfeats, const float* lodw, int64_t M, int IN,
                       const float* const* weights, int hidden, int Cs, int Ci, int sem_softmax, int inst_softmax,
                       float inst_temperature, float* sem, float* inst, void* stream) {
    if (hidden != H || Cs < 0 || Ci < 0) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    PanParams p{};
    p.Ws1 = weights[0]; p.bs1 = weights[1]; p.Ws2 = weights[2]; p.bs2 = weights[3]; p.Wi1 = weights[4];
    p.bi1 = weights[5]; p.Wi2 = weights[6]; p.bi2 = weights[7]; p.Wi3 = weights[8]; p.bi3 = weights[9];
    const float it = inst_temperature > 0.f ? 1.f / inst_temperature : 1.f;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pad4(IN)) {
        case 12: return pan_fwd_launch<12>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, st);
        case 28: return pan_fwd_launch<28>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, st);
        case 32: return pan_fwd_launch<32>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, st);
        case 48: return pan_fwd_launch<48>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, st);
        default: return PAG_ERR_UNSUPPORTED;
    }
}

int pag_decode_pan_bwd(const float* feats, const float* dfeats, const float* lodw, int64_t M, int IN,
                       const float* const* weights, float* const* grads, int hidden, int Cs, int Ci, int sem_softmax,
                       int inst_softmax, float inst_temperature, const float* sem, const float* inst,
                       const float* g_sem, const float* g_inst, float* g_panop, void* stream) {
    if (hidden != H || Cs < 0 || Ci < 0) return PAG_ERR_UNSUPPORTED;
    if (M == 0) return PAG_OK;
    PanParams p{};
    p.Ws1 = weights[0]; p.bs1 = weights[1]; p.Ws2 = weights[2]; p.bs2 = weights[3]; p.Wi1 = weights[4];
    p.bi1 = weights[5]; p.Wi2 = weights[6]; p.bi2 = weights[7]; p.Wi3 = weights[8]; p.bi3 = weights[9];
    p.gWs1 = grads[0]; p.gbs1 = grads[1]; p.gWs2 = grads[2]; p.gbs2 = grads[3]; p.gWi1 = grads[4];
    p.gbi1 = grads[5]; p.gWi2 = grads[6]; p.gbi2 = grads[7]; p.gWi3 = grads[8]; p.gbi3 = grads[9];
    const float it = inst_temperature > 0.f ? 1.f / inst_temperature : 1.f;
    cudaStream_t st = (cudaStream_t)stream;
    switch (pad4(IN)) {
        case 12: return pan_bwd_launch<12>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, g_sem, g_inst, g_panop, st);
        case 28: return pan_bwd_launch<28>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, g_sem, g_inst, g_panop, st);
        case 32: return pan_bwd_launch<32>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, g_sem, g_inst, g_panop, st);
        case 48: return pan_bwd_launch<48>(feats, dfeats, lodw, M, IN, p, Cs, Ci, sem_softmax, inst_softmax, it, sem, inst, g_sem, g_inst, g_panop, st);
        default: return PAG_ERR_UNSUPPORTED;
    }
}

// linear scalar head on (feats + dfeats) * lodw: y f32[M]; w f32[IN], b f32[1].  g_x nullable; g_w / g_b accumulate (zeroed by the caller).
int pag_linear_head_fwd(const float* feats, const float* dfeats, const float* lodw, int64_t M, int IN, const float* w,
                        const float* b, float* y, void* stream) {
    if (IN <= 0) return PAG_ERR_ARG;
    if (M == 0) return PAG_OK;
    linear_head_fwd_kernel<<<pag_grid(M, 256), 256, 0, (cudaStream_t)stream>>>(feats, dfeats, lodw, M, IN, w, b, y, nullptr, 0, nullptr, nullptr);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_linear_head_bwd(const float* feats, const float* dfeats, const float* lodw, int64_t M, int IN, const float* w,
                        const float* g, float* g_x, float* g_w, float* g_b, void* stream) {
    if (IN <= 0) return PAG_ERR_ARG;
    if (M == 0) return PAG_OK;
    if (IN > LHR_MAXIN) return PAG_ERR_UNSUPPORTED;
    linear_head_bwd_kernel<<<pag_grid(M, 256) < 592 ? pag_grid(M, 256) : 592, 256, 0, (cudaStream_t)stream>>>(
        feats, dfeats, lodw, M, IN, w, g, g_x, g_w, g_b, nullptr, nullptr, 0, nullptr);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
// fused-trace variants: y = post * relu?(pre + head(x)) with the sample count on the device; the backward gates g with the
// forward output (ReLU), multiplies it by post, and can accumulate into g_x (the heads' dX is already there)
int pag_linear_head_fwd_dyn(const float* feats, const float* dfeats, const float* lodw, int64_t M_max, const int64_t* m_dev, int IN,
                            const float* w, const float* b, const float* pre, int relu, const float* post, float* y, int x_img16,
                            void* stream) {
    if (IN <= 0 || (x_img16 && (IN & 7))) return PAG_ERR_ARG;
    if (M_max == 0) return PAG_OK;
    if (x_img16) {
        linear_head_fwd_img_kernel<<<pag_grid(M_max, 256), 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4*>(feats), reinterpret_cast<const uint4*>(dfeats), lodw, M_max, IN, w, b, y, pre, relu, post, m_dev);
        PAG_LAUNCH_CHECK();
        return PAG_OK;
    }
    linear_head_fwd_kernel<<<pag_grid(M_max, 256), 256, 0, (cudaStream_t)stream>>>(feats, dfeats, lodw, M_max, IN, w, b, y, pre, relu, post, m_dev);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
int pag_linear_head_bwd_dyn(const float* feats, const float* dfeats, const float* lodw, int64_t M_max, const int64_t* m_dev, int IN,
                            const float* w, const float* g, const float* gate, const float* post, float* g_x, int accumulate_x,
                            float* g_w, float* g_b, int x_img16, const float* img_scale, void* stream) {
    if (IN <= 0 || (x_img16 && ((IN & 7) || !accumulate_x))) return PAG_ERR_ARG;
    if (M_max == 0) return PAG_OK;
    if (x_img16) {
        const int gridp = pag_grid(M_max, 256) < 592 ? pag_grid(M_max, 256) : 592;     // persistent: ~2 resident CTAs per SM, 2 waves
        linear_head_bwd_img_kernel<<<gridp, 256, 0, (cudaStream_t)stream>>>(
            reinterpret_cast<const uint4*>(feats), reinterpret_cast<const uint4*>(dfeats), lodw, M_max, IN, w, g,
            reinterpret_cast<uint4*>(g_x), img_scale, g_w, g_b, gate, post, m_dev);
        PAG_LAUNCH_CHECK();
        return PAG_OK;
    }
    if (IN > LHR_MAXIN) return PAG_ERR_UNSUPPORTED;
    linear_head_bwd_kernel<<<pag_grid(M_max, 256) < 592 ? pag_grid(M_max, 256) : 592, 256, 0, (cudaStream_t)stream>>>(
        feats, dfeats, lodw, M_max, IN, w, g, g_x, g_w, g_b, gate, post, accumulate_x, m_dev);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
// see dd_weight_grads_kernel; gw f32[M] out
int pag_dd_weight_grads(const float* g_sem, const float* out_sem, int Cs, const float* g_inst, const float* out_inst, int Ci,
                        const float* alpha_p, const float* gw_sem, const float* gw_inst, const int64_t* offsets, int64_t R,
                        float* gw, void* stream) {
    if (R == 0) return PAG_OK;
    dd_weight_grads_kernel<<<pag_grid(R * 32, 256), 256, 0, (cudaStream_t)stream>>>(g_sem, out_sem, Cs, g_inst, out_inst, Ci, alpha_p,
                                                                                    gw_sem, gw_inst, offsets, R, gw);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
