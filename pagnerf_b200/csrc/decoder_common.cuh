// Shared definitions of the decoder kernels (FP32 FMA path: decoder.cu, tensor-core path: decoder_tc.cu).
#pragma once
#include "common.cuh"

#define H 64        // hidden width (configs/bup20/best.yaml:70)
#define DOUT 16     // density decoder output width (pc_nerf/panoptic_nef.py:115)
#define PE_F 4      // view_multires (best.yaml:39) -> 3 + 3*2*4 = 27
#define PE_DIM 27
#define CIN 43      // 16 + 27
#define CINP 44     // padded to a multiple of 4

struct DcParams {   // density + color decoders; torch Linear layout W[out][in] row-major
    const float *Wd1, *bd1, *Wd2, *bd2, *Wc1, *bc1, *Wc2, *bc2, *Wc3, *bc3;
    float *gWd1, *gbd1, *gWd2, *gbd2, *gWc1, *gbc1, *gWc2, *gbc2, *gWc3, *gbc3;
};
struct PanParams {  // semantic + instance decoders
    const float *Ws1, *bs1, *Ws2, *bs2, *Wi1, *bi1, *Wi2, *bi2, *Wi3, *bi3;
    float *gWs1, *gbs1, *gWs2, *gbs2, *gWi1, *gbi1, *gWi2, *gbi2, *gWi3, *gbi3;
};

__host__ __device__ inline int pad4(int x) { return (x + 3) & ~3; }

// view-direction positional embedding of v = -d : [v, sin(2^f v), cos(2^f v)], f-major / xyz-minor
__device__ __forceinline__ void view_embed(float dx, float dy, float dz, float* pe /*27*/) {
    const float v[3] = {-dx, -dy, -dz};
#pragma unroll
    for (int c = 0; c < 3; ++c) pe[c] = v[c];
#pragma unroll
    for (int f = 0; f < PE_F; ++f)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float a = v[c] * (float)(1 << f);
            pe[3 + 3 * f + c] = sinf(a);
            pe[3 + 3 * PE_F + 3 * f + c] = cosf(a);
        }
}

