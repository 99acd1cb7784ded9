// Micro-benchmark: issue rate of the sm_100 mixed-precision FMA (PTX fma.rn.f32.f16 -> SASS FHFMA, fp16 x fp16 + fp32
// with .H0/.H1 operand selectors) against the cvt + FFMA pair it replaces in the fp16 gather-and-blend kernels
// (volume builder, upconv_blend).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fhfma_rate fhfma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

__device__ __forceinline__ float fhfma_lo(uint32_t a, uint32_t w, float c) {
    float d; unsigned short al, ah, wl, wh; (void)al; (void)ah; (void)wl; (void)wh;
    asm("mov.b32 {%0,%1}, %2;" : "=h"(al), "=h"(ah) : "r"(a));
    asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(w));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(al), "h"(wl), "f"(c)); return d;
}
__device__ __forceinline__ float fhfma_hi(uint32_t a, uint32_t w, float c) {
    float d; unsigned short al, ah, wl, wh; (void)al; (void)ah; (void)wl; (void)wh;
    asm("mov.b32 {%0,%1}, %2;" : "=h"(al), "=h"(ah) : "r"(a));
    asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(w));
    asm("fma.rn.f32.f16 %0, %1, %2, %3;" : "=f"(d) : "h"(ah), "h"(wl), "f"(c)); return d;
}

template <int MODE>
__global__ void __launch_bounds__(256) rate_kernel(const uint4* __restrict__ in, float* __restrict__ out, int iters, float wf) {
    __shared__ uint4 sm[256 + 64];
    sm[threadIdx.x] = in[threadIdx.x & 31];
    if (threadIdx.x < 64) sm[256 + threadIdx.x] = in[threadIdx.x & 31];
    __syncthreads();
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const __half2 wh2 = __floats2half2_rn(wf, wf);
    const uint32_t w = *reinterpret_cast<const uint32_t*>(&wh2);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
            const uint4 u = sm[threadIdx.x + ((it + r) & 63)];      // one LDS.128 per 8 channels, as in the real kernels
            const uint32_t q[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (MODE == 0) {            // cvt + FFMA (today's kernels)
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&q[k]));
                    acc[2 * k] = fmaf(wf, f.x, acc[2 * k]);
                    acc[2 * k + 1] = fmaf(wf, f.y, acc[2 * k + 1]);
                } else if (MODE == 1) {     // FHFMA
                    acc[2 * k] = fhfma_lo(q[k], w, acc[2 * k]);
                    acc[2 * k + 1] = fhfma_hi(q[k], w, acc[2 * k + 1]);
                } else {                    // FFMA only (upper bound)
                    acc[2 * k] = fmaf(wf, __uint_as_float(q[k]), acc[2 * k]);
                    acc[2 * k + 1] = fmaf(wf, __uint_as_float(q[k] ^ 0x1000u), acc[2 * k + 1]);
                }
            }
        }
    }
    float s = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += acc[k];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, const uint4* in, float* out) {
    const int iters = 4096, blocks = 148 * 8;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    rate_kernel<MODE><<<blocks, 256>>>(in, out, 16, 0.37f);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    rate_kernel<MODE><<<blocks, 256>>>(in, out, iters, 0.37f);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fmas = (double)blocks * 256 * iters * 64;
    printf("%-12s %.3f ms  %.2f T fma/s  (%.1f fma/clk/SM at 1.8 GHz)\n", name, ms, fmas / ms * 1e-9, fmas / ms * 1e-9 * 1e12 / 148 / 1.8e9 );
}

int main() {
    uint4* in; float* out;
    cudaMalloc(&in, 32 * 16); cudaMemset(in, 0x3c, 32 * 16);
    cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<2>("ffma", in, out);
    run<0>("cvt+ffma", in, out);
    run<1>("fhfma", in, out);
    float h[4]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("check %g (err %s)\n", h[0], cudaGetErrorString(cudaGetLastError()));
    return 0;
}
