// Shared declarations for libadapose_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>

typedef __nv_bfloat16 bf16;

#define ADP_OK 0
#define ADP_ERR_CUDA -1
#define ADP_ERR_ARG -2
#define ADP_ERR_STATE -3

namespace adp {

void set_last_error(const char* fmt, ...);

#define ADP_CUDA(expr)                                                                          \
    do {                                                                                        \
        cudaError_t _e = (expr);                                                                \
        if (_e != cudaSuccess) {                                                                \
            adp::set_last_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
            return ADP_ERR_CUDA;                                                                \
        }                                                                                       \
    } while (0)

#define ADP_CHECK_ARG(cond, msg)                                                  \
    do {                                                                          \
        if (!(cond)) {                                                            \
            adp::set_last_error("%s:%d argument check failed: %s (%s)", __FILE__, __LINE__, #cond, msg); \
            return ADP_ERR_ARG;                                                   \
        }                                                                         \
    } while (0)

#define ADP_TRY(expr)              \
    do {                           \
        int _r = (expr);           \
        if (_r != ADP_OK) return _r; \
    } while (0)

// A channels-last activation: value = hi (+ lo when the split-precision path is on).
// dims: [B, D, H, W, C] (D == 1 for 2-D maps).
struct Act {
    bf16* hi = nullptr;
    bf16* lo = nullptr;   // nullptr -> single bf16
    int B = 0, D = 1, H = 0, W = 0, C = 0;
    int f16 = 0;          // 1: the planes hold IEEE half instead of bf16 (3-D stage; no lo plane)
    size_t numel() const { return (size_t)B * D * H * W * C; }
};

__device__ __forceinline__ float ld_act(const bf16* __restrict__ hi, const bf16* __restrict__ lo, size_t i) {
    float v = __bfloat162float(hi[i]);
    if (lo) v += __bfloat162float(lo[i]);
    return v;
}

__device__ __forceinline__ void st_act(bf16* __restrict__ hi, bf16* __restrict__ lo, size_t i, float v) {
    bf16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    if (lo) lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
}

// 16-bit storage that is either bf16 (f16 == 0) or IEEE half (f16 == 1); the pointer type stays bf16* for brevity
__device__ __forceinline__ float ld_act16(const bf16* __restrict__ hi, const bf16* __restrict__ lo, size_t i, int f16) {
    if (f16) return __half2float(reinterpret_cast<const __half*>(hi)[i]);
    return ld_act(hi, lo, i);
}
__device__ __forceinline__ void st_act16(bf16* __restrict__ hi, bf16* __restrict__ lo, size_t i, float v, int f16) {
    if (f16) reinterpret_cast<__half*>(hi)[i] = __float2half_rn(v);
    else st_act(hi, lo, i, v);
}

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

}  // namespace adp
