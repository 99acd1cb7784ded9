// Generic channels-last convolution on CUDA cores (fp32 accumulate, fp32 weights).
// Covers what the tcgen05 path does not: strided convs, tiny channel counts (3, 8), transposed 3-D convs,
// and serves as the on-device cross-check of the tensor-core kernel in the parity tests.
#pragma once
#include "common.cuh"

namespace adp {

struct DirectConvParams {
    // input: bf16 hi(+lo) or fp32, [B, Di, Hi, Wi, Cin]
    const bf16* in_hi = nullptr;
    const bf16* in_lo = nullptr;
    const float* in_f32 = nullptr;
    int B = 0, Di = 1, Hi = 0, Wi = 0, Cin = 0;
    int Do = 1, Ho = 0, Wo = 0, Cout = 0;
    int kd = 1, kh = 1, kw = 1;
    int sd = 1, sh = 1, sw = 1;
    int pd = 0, ph = 0, pw = 0;
    int dil = 1;             // dilation along h/w
    int transposed = 0;      // 1: ConvTranspose (gather form o = i*s - p + k)
    int f16 = 0;             // 16-bit planes are IEEE half
    const float* w = nullptr;   // [taps][Cin][Cout]
    const float* scale = nullptr;
    const float* bias = nullptr;
    float prelu = 0.f;
    int act = 0;
    int res_after_act = 0;
    const bf16* res_hi = nullptr;
    const bf16* res_lo = nullptr;
    int res_cs = 0;
    bf16* out_hi = nullptr;
    bf16* out_lo = nullptr;
    float* out_f32 = nullptr;
};

int direct_conv_launch(const DirectConvParams& p, int batch, cudaStream_t stream);

}  // namespace adp
