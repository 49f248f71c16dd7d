// Fused multi-tensor Adam step (fp32 parameters / gradients / moments), CUDA-graph capturable.
//
// The reference trains with torch.optim.Adam over six parameter groups (pc_nerf/trainer.py:229-300, configs/bup20/best.yaml:114:
// decoders, sem, inst, delta_grid and grid at lr x 100, camera extrinsics) -- ~20 small tensors plus the two 50 MB tables.  One
// launch walks all of them: a flat grid over 16-byte quads, each CTA finding its tensor by a linear scan of <= 48 prefix sums held
// in kernel parameters; the step count lives in device memory (bias corrections are computed on the device, so a captured graph
// replays correctly); gradients may arrive multiplied by a loss scale (*inv_scale, nullable).  Update rule = torch.optim.Adam
// (amsgrad off, L2 weight decay folded into the gradient): 7 x 4 B of HBM traffic per element, the bound of this kernel.
#include "common.cuh"

#define ADAM_MAX_TENSORS 48

struct AdamTensors {
    float* p[ADAM_MAX_TENSORS];
    const float* g[ADAM_MAX_TENSORS];
    float* m[ADAM_MAX_TENSORS];
    float* v[ADAM_MAX_TENSORS];
    long long start[ADAM_MAX_TENSORS + 1];   // prefix sums of ceil(numel / 4) quads
    long long numel[ADAM_MAX_TENSORS];
    float lr[ADAM_MAX_TENSORS], wd[ADAM_MAX_TENSORS];
    int n;
};

__global__ void adam_tick_kernel(int* __restrict__ step) { *step += 1; }

__global__ void __launch_bounds__(256) adam_kernel(AdamTensors T, float beta1, float beta2, float eps, const int* __restrict__ step,
                                                   const float* __restrict__ inv_scale) {
    const float t = (float)__ldg(step);
    const float bc1 = 1.f - powf(beta1, t), bc2 = 1.f - powf(beta2, t);
    const float rs2 = rsqrtf(bc2);
    const float isc = inv_scale ? __ldg(inv_scale) : 1.f;
    const long long total = T.start[T.n];
    for (long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x; q < total; q += (long long)gridDim.x * blockDim.x) {
        int k = 0;
        while (k + 1 < T.n && q >= T.start[k + 1]) ++k;
        const long long e0 = (q - T.start[k]) * 4;
        const long long n = T.numel[k];
        const float lr = T.lr[k], wd = T.wd[k], ss = lr / bc1;
        float* p = T.p[k]; const float* g = T.g[k]; float* m = T.m[k]; float* v = T.v[k];
        const bool vec = (e0 + 4 <= n) && !((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
                                              reinterpret_cast<uintptr_t>(v)) & 15);
        float pp[4], gg[4], mm[4], vv[4];
        if (vec) {
            *reinterpret_cast<float4*>(pp) = *reinterpret_cast<const float4*>(p + e0);
            *reinterpret_cast<float4*>(gg) = __ldg(reinterpret_cast<const float4*>(g + e0));
            *reinterpret_cast<float4*>(mm) = *reinterpret_cast<const float4*>(m + e0);
            *reinterpret_cast<float4*>(vv) = *reinterpret_cast<const float4*>(v + e0);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (e0 + i < n) { pp[i] = p[e0 + i]; gg[i] = g[e0 + i]; mm[i] = m[e0 + i]; vv[i] = v[e0 + i]; }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float gr = gg[i] * isc + wd * pp[i];
            mm[i] = beta1 * mm[i] + (1.f - beta1) * gr;
            vv[i] = beta2 * vv[i] + (1.f - beta2) * gr * gr;
            pp[i] -= ss * mm[i] / (sqrtf(vv[i]) * rs2 + eps);
        }
        if (vec) {
            *reinterpret_cast<float4*>(p + e0) = *reinterpret_cast<const float4*>(pp);
            *reinterpret_cast<float4*>(m + e0) = *reinterpret_cast<const float4*>(mm);
            *reinterpret_cast<float4*>(v + e0) = *reinterpret_cast<const float4*>(vv);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (e0 + i < n) { p[e0 + i] = pp[i]; m[e0 + i] = mm[i]; v[e0 + i] = vv[i]; }
        }
    }
}

extern "C" {

// One Adam step over n_tensors (<= 48) fp32 tensors: host arrays of device pointers p / g / m / v, element counts, per-tensor
// learning rate and weight decay; step (device int) is incremented first and used for the bias corrections; inv_scale (device
// float, nullable) multiplies the gradients.  Replaces the reference's torch.optim.Adam.step() (pc_nerf/trainer.py:229-300, :590).
int pag_adam_step(float* const* p, const float* const* g, float* const* m, float* const* v, const int64_t* numel, const float* lr,
                  const float* weight_decay, int n_tensors, float beta1, float beta2, float eps, int* step, const float* inv_scale,
                  void* stream) {
    if (n_tensors < 0 || n_tensors > ADAM_MAX_TENSORS) return PAG_ERR_ARG;
    if (n_tensors == 0) return PAG_OK;
    AdamTensors T;
    T.n = n_tensors;
    long long acc = 0;
    for (int i = 0; i < n_tensors; ++i) {
        T.p[i] = p[i]; T.g[i] = g[i]; T.m[i] = m[i]; T.v[i] = v[i];
        T.numel[i] = numel[i]; T.lr[i] = lr[i]; T.wd[i] = weight_decay[i];
        T.start[i] = acc;
        acc += (numel[i] + 3) / 4;
    }
    T.start[n_tensors] = acc;
    cudaStream_t st = (cudaStream_t)stream;
    adam_tick_kernel<<<1, 1, 0, st>>>(step);
    PAG_LAUNCH_CHECK();
    const long long blocks = (acc + 255) / 256;
    const int grid = (int)(blocks < 148 * 16 ? (blocks > 0 ? blocks : 1) : 148 * 16);
    adam_kernel<<<grid, 256, 0, st>>>(T, beta1, beta2, eps, step, inv_scale);
    PAG_LAUNCH_CHECK();
    return PAG_OK;
}

}  // extern "C"
