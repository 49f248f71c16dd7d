// Camera-pose transform of the base rays + its backward (bundle adjustment inside the training step).
//
// Replaces BAPipeline.transform_rays (pc_nerf/ba_pipeline.py:85-92): kaolin's `Camera.extrinsics` in the
// 'matrix_6dof_rotation' backend (:44) holds 9 parameters per camera -- a 6-D rotation representation (Zhou et al. 2019:
// two 3-vectors a1, a2, orthonormalised by Gram-Schmidt into the rows b1, b2, b3 = b1 x b2 of the view rotation R) and the
// view translation t --, `inv_transform_rays` maps camera-space rays to world space, o_w = R^T (o_c - t), d_w = R^T d_c, and
// the reference renormalises d_w (:89).  Gradients flow back to the 9 parameters (pc_nerf/trainer.py:297, grad mask for anchor
// frames ba_pipeline.py:53-62).
//
// B200 mapping: rays arrive grouped by camera (base_rays.reshape(len(cameras), -1, 3), :87), B rays each.  Forward: one thread
// per ray, R rebuilt from the 9 parameters in registers (36 B per camera, L1 resident).  Backward: one CTA per camera reduces
// the 12 sums  G = sum_r [(o_c - t) (x) g_o + d_c (x) g_v],  S = sum_r g_o  over its rays in registers / shuffles, then ONE
// thread runs the Gram-Schmidt chain rule and adds 9 floats into the camera's gradient row -- no per-ray atomics.
#include "common.cuh"

struct Pose { float b1[3], b2[3], b3[3], t[3], a2[3], n1, n2, dot; };

__device__ __forceinline__ void pose_load(const float* __restrict__ p, Pose& q) {
    const float a1[3] = {__ldg(p), __ldg(p + 1), __ldg(p + 2)};
    q.a2[0] = __ldg(p + 3); q.a2[1] = __ldg(p + 4); q.a2[2] = __ldg(p + 5);
    q.t[0] = __ldg(p + 6); q.t[1] = __ldg(p + 7); q.t[2] = __ldg(p + 8);
    q.n1 = sqrtf(a1[0] * a1[0] + a1[1] * a1[1] + a1[2] * a1[2]);
    const float i1 = 1.f / q.n1;
#pragma unroll
    for (int i = 0; i < 3; ++i) q.b1[i] = a1[i] * i1;
    q.dot = q.b1[0] * q.a2[0] + q.b1[1] * q.a2[1] + q.b1[2] * q.a2[2];
    float u[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) u[i] = q.a2[i] - q.dot * q.b1[i];
    q.n2 = sqrtf(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
    const float i2 = 1.f / q.n2;
#pragma unroll
    for (int i = 0; i < 3; ++i) q.b2[i] = u[i] * i2;
    q.b3[0] = q.b1[1] * q.b2[2] - q.b1[2] * q.b2[1];
    q.b3[1] = q.b1[2] * q.b2[0] - q.b1[0] * q.b2[2];
    q.b3[2] = q.b1[0] * q.b2[1] - q.b1[1] * q.b2[0];
}

__global__ void __launch_bounds__(256) pose_fwd_kernel(const float* __restrict__ params, const int64_t* __restrict__ cam_idx,
                                                       const float* __restrict__ base_o, const float* __restrict__ base_d,
                                                       int64_t C, int64_t B, float* __restrict__ out_o, float* __restrict__ out_d) {
    const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= C * B) return;
    const int64_t c = r / B;
    Pose q;
    pose_load(params + 9 * (cam_idx ? cam_idx[c] : c), q);
    const float oc[3] = {base_o[3 * r] - q.t[0], base_o[3 * r + 1] - q.t[1], base_o[3 * r + 2] - q.t[2]};
    const float dc[3] = {base_d[3 * r], base_d[3 * r + 1], base_d[3 * r + 2]};
    float v[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        out_o[3 * r + j] = q.b1[j] * oc[0] + q.b2[j] * oc[1] + q.b3[j] * oc[2];
        v[j] = q.b1[j] * dc[0] + q.b2[j] * dc[1] + q.b3[j] * dc[2];
    }
    const float inv = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);      // rays_dir / norm (ba_pipeline.py:89)
#pragma unroll
    for (int j = 0; j < 3; ++j) out_d[3 * r + j] = v[j] * inv;
}

__global__ void __launch_bounds__(256) pose_bwd_kernel(const float* __restrict__ params, const int64_t* __restrict__ cam_idx,
                                                       const float* __restrict__ base_o, const float* __restrict__ base_d,
                                                       const float* __restrict__ g_o, const float* __restrict__ g_d, int64_t B,
                                                       float* __restrict__ g_params) {
    const int64_t c = blockIdx.x;
    const int64_t row = cam_idx ? cam_idx[c] : c;
    Pose q;
    pose_load(params + 9 * row, q);
    float acc[12];      // G[3][3] (row i = dL/db_i), S[3] = sum g_o
#pragma unroll
    for (int k = 0; k < 12; ++k) acc[k] = 0.f;
    for (int64_t i = threadIdx.x; i < B; i += blockDim.x) {
        const int64_t r = c * B + i;
        const float oc[3] = {base_o[3 * r] - q.t[0], base_o[3 * r + 1] - q.t[1], base_o[3 * r + 2] - q.t[2]};
        const float dc[3] = {base_d[3 * r], base_d[3 * r + 1], base_d[3 * r + 2]};
        float go[3] = {0.f, 0.f, 0.f}, gv[3] = {0.f, 0.f, 0.f};
        if (g_o) { go[0] = g_o[3 * r]; go[1] = g_o[3 * r + 1]; go[2] = g_o[3 * r + 2]; }
        if (g_d) {
            float v[3];
#pragma unroll
            for (int j = 0; j < 3; ++j) v[j] = q.b1[j] * dc[0] + q.b2[j] * dc[1] + q.b3[j] * dc[2];
            const float inv = 1.f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
            const float dw[3] = {v[0] * inv, v[1] * inv, v[2] * inv};
            const float gd[3] = {g_d[3 * r], g_d[3 * r + 1], g_d[3 * r + 2]};
            const float pr = dw[0] * gd[0] + dw[1] * gd[1] + dw[2] * gd[2];
#pragma unroll
            for (int j = 0; j < 3; ++j) gv[j] = (gd[j] - dw[j] * pr) * inv;
        }
#pragma unroll
        for (int a = 0; a < 3; ++a)
#pragma unroll
            for (int j = 0; j < 3; ++j) acc[3 * a + j] += oc[a] * go[j] + dc[a] * gv[j];
#pragma unroll
        for (int j = 0; j < 3; ++j) acc[9 + j] += go[j];
    }
    __shared__ float red[8][12];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 12; ++k) {
        float v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[warp][k] = v;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float G[12];
        for (int k = 0; k < 12; ++k) { float v = 0.f; for (int w = 0; w < 8; ++w) v += red[w][k]; G[k] = v; }
        float* g1 = G, *g2 = G + 3, *g3 = G + 6, *S = G + 9;
        // dL/dt_i = -b_i . S
        const float gt[3] = {-(q.b1[0] * S[0] + q.b1[1] * S[1] + q.b1[2] * S[2]), -(q.b2[0] * S[0] + q.b2[1] * S[1] + q.b2[2] * S[2]),
                             -(q.b3[0] * S[0] + q.b3[1] * S[1] + q.b3[2] * S[2])};
        // b3 = b1 x b2:  g_b1 += b2 x g3,  g_b2 += g3 x b1
        float gb1[3] = {g1[0] + q.b2[1] * g3[2] - q.b2[2] * g3[1], g1[1] + q.b2[2] * g3[0] - q.b2[0] * g3[2], g1[2] + q.b2[0] * g3[1] - q.b2[1] * g3[0]};
        float gb2[3] = {g2[0] + g3[1] * q.b1[2] - g3[2] * q.b1[1], g2[1] + g3[2] * q.b1[0] - g3[0] * q.b1[2], g2[2] + g3[0] * q.b1[1] - g3[1] * q.b1[0]};
        // b2 = u / |u|
        const float p2 = q.b2[0] * gb2[0] + q.b2[1] * gb2[1] + q.b2[2] * gb2[2];
        float gu[3];
        for (int j = 0; j < 3; ++j) gu[j] = (gb2[j] - q.b2[j] * p2) / q.n2;
        // u = a2 - (b1 . a2) b1
        const float pu = q.b1[0] * gu[0] + q.b1[1] * gu[1] + q.b1[2] * gu[2];
        float ga2[3];
        for (int j = 0; j < 3; ++j) { ga2[j] = gu[j] - q.b1[j] * pu; gb1[j] += -q.dot * gu[j] - pu * q.a2[j]; }
        // b1 = a1 / |a1|
        const float p1 = q.b1[0] * gb1[0] + q.b1[1] * gb1[1] + q.b1[2] * gb1[2];
        float* gp = g_params + 9 * row;
        for (int j = 0; j < 3; ++j) {
            red_add_f32(gp + j, (gb1[j] - q.b1[j] * p1) / q.n1);
            red_add_f32(gp + 3 + j, ga2[j]);
            red_add_f32(gp + 6 + j, gt[j]);
        }
    }
}

extern "C" {

// params f32[n_cameras, 9] = (a1, a2, t); cam_idx i64[C] (nullable = identity) selects the parameter row of each of the C ray
// groups; base_o / base_d f32[C*B, 3] camera-space rays grouped by camera; out_o / out_d f32[C*B, 3] world-space, d normalised.
int pag_pose_transform_fwd(const float* params, const int64_t* cam_idx, const float* base_o, const float* base_d, int64_t C, int64_t B,
                           float* out_o, float* out_d, void* stream) {
    if (C < 0 || B < 0) return PAG_ERR_ARG;
    if (C * B == 0) return PAG_OK;
    pose_fwd_kernel<<<pag_grid(C * B, 256), 256, 0, (cudaStream_t)stream>>>(params, cam_idx, base_o, base_d, C, B, out_o, out_d);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}
// g_params f32[n_cameras, 9] is ACCUMULATED into (caller zeroes it; a camera may appear in several groups); g_o / g_d nullable
int pag_pose_transform_bwd(const float* params, const int64_t* cam_idx, const float* base_o, const float* base_d, const float* g_o,
                           const float* g_d, int64_t C, int64_t B, float* g_params, void* stream) {
    if (C < 0 || B < 0) return PAG_ERR_ARG;
    if (C * B == 0) return PAG_OK;
    pose_bwd_kernel<<<(unsigned)C, 256, 0, (cudaStream_t)stream>>>(params, cam_idx, base_o, base_d, g_o, g_d, B, g_params);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
