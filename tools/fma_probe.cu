// FP32 FMA issue-rate probe: scalar FFMA vs packed FFMA2 (fma.rn.f32x2) with a broadcast scalar operand / with a pair operand.
// build + run on the GPU box:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/fma_probe.cu -o /tmp/fma_probe && /tmp/fma_probe
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void ffma2(float2& c, const float2 a, const float2 b) {
    unsigned long long cc = *reinterpret_cast<unsigned long long*>(&c);
    asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(cc) : "l"(*reinterpret_cast<const unsigned long long*>(&a)),
        "l"(*reinterpret_cast<const unsigned long long*>(&b)));
    c = *reinterpret_cast<float2*>(&cc);
}

template <int MODE, int NACC>
__global__ void probe(float* out, int iters, float x, float y) {
    float2 acc[NACC];
    for (int i = 0; i < NACC; ++i) acc[i] = make_float2(threadIdx.x * 1e-3f + i, i * 0.5f);
    float2 a0 = make_float2(x, y), a1 = make_float2(y, x), a2 = make_float2(x + 1.f, y + 1.f), a3 = make_float2(y - 1.f, x - 1.f);
    float w0 = x * 0.5f, w1 = y * 0.25f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; i += 4) {
            if (MODE == 0) {          // scalar FFMA: 8 per group of 4 pairs
                acc[i].x = fmaf(a0.x, w0, acc[i].x); acc[i].y = fmaf(a0.y, w0, acc[i].y);
                acc[i + 1].x = fmaf(a1.x, w0, acc[i + 1].x); acc[i + 1].y = fmaf(a1.y, w0, acc[i + 1].y);
                acc[i + 2].x = fmaf(a2.x, w1, acc[i + 2].x); acc[i + 2].y = fmaf(a2.y, w1, acc[i + 2].y);
                acc[i + 3].x = fmaf(a3.x, w1, acc[i + 3].x); acc[i + 3].y = fmaf(a3.y, w1, acc[i + 3].y);
            } else if (MODE == 1) {   // FFMA2, broadcast scalar operand
                ffma2(acc[i], a0, make_float2(w0, w0)); ffma2(acc[i + 1], a1, make_float2(w0, w0));
                ffma2(acc[i + 2], a2, make_float2(w1, w1)); ffma2(acc[i + 3], a3, make_float2(w1, w1));
            } else {                  // FFMA2, pair operand
                ffma2(acc[i], a0, a1); ffma2(acc[i + 1], a1, a2); ffma2(acc[i + 2], a2, a3); ffma2(acc[i + 3], a3, a0);
            }
        }
    }
    float s = 0.f;
    for (int i = 0; i < NACC; ++i) s += acc[i].x + acc[i].y;
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int NACC>
void run(const char* name, int threads, float* out) {
    const int iters = 20000, grid = 148;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    probe<MODE, NACC><<<grid, threads>>>(out, 100, 1.0001f, 0.9999f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<MODE, NACC><<<grid, threads>>>(out, iters, 1.0001f, 0.9999f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)grid * threads * iters * NACC * 2;
    printf("%-28s acc=%2d threads=%4d: %7.3f ms  %6.1f TFMA/s  (%.1f FMA/clk/SM at 1.965 GHz)\n", name, NACC, threads, ms,
           fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * sizeof(float));
    for (int threads : {128, 256, 512, 1024}) {
        run<0, 16>("FFMA scalar", threads, out);
        run<1, 16>("FFMA2 broadcast scalar", threads, out);
        run<2, 16>("FFMA2 pair operand", threads, out);
    }
    run<0, 52>("FFMA scalar", 256, out);
    run<1, 52>("FFMA2 broadcast scalar", 256, out);
    run<2, 52>("FFMA2 pair operand", 256, out);
    return 0;
}
