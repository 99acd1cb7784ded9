// Shared declarations for libadapose_b200.so (sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
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
    uint8_t* q8 = nullptr;   // optional fp8 (e4m3) twin holding value / 2: the A operand of the fp8 low-order pass ("fp16+fp8lo" layers)
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

// 8 consecutive 16-bit values (one 128-bit access) <-> fp32
__device__ __forceinline__ void ld8_16(const bf16* __restrict__ p, size_t i, int f16, float* v, bool accumulate) {
    const uint4 u = *reinterpret_cast<const uint4*>(p + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float a, b;
        if (f16) {
            a = __half2float(__ushort_as_half((unsigned short)(w[q] & 0xffffu)));
            b = __half2float(__ushort_as_half((unsigned short)(w[q] >> 16)));
        } else {
            a = __uint_as_float(w[q] << 16);
            b = __uint_as_float(w[q] & 0xffff0000u);
        }
        if (accumulate) { v[2 * q] += a; v[2 * q + 1] += b; } else { v[2 * q] = a; v[2 * q + 1] = b; }
    }
}
// stores hi (and lo = v - hi for bf16 split precision)
__device__ __forceinline__ void st8_16(bf16* __restrict__ hi, bf16* __restrict__ lo, size_t i, int f16, const float* v) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (f16) {
            h[q] = (uint32_t)__half_as_ushort(__float2half_rn(v[2 * q])) | ((uint32_t)__half_as_ushort(__float2half_rn(v[2 * q + 1])) << 16);
            l[q] = 0;
        } else {
            const bf16 h0 = __float2bfloat16_rn(v[2 * q]), h1 = __float2bfloat16_rn(v[2 * q + 1]);
            h[q] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
            const bf16 l0 = __float2bfloat16_rn(v[2 * q] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(v[2 * q + 1] - __bfloat162float(h1));
            l[q] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
        }
    }
    *reinterpret_cast<uint4*>(hi + i) = make_uint4(h[0], h[1], h[2], h[3]);
    if (lo && !f16) *reinterpret_cast<uint4*>(lo + i) = make_uint4(l[0], l[1], l[2], l[3]);
}

// general 4x4 inverse (Gauss-Jordan with partial pivoting), fp64: np.linalg.inv(view1_extrinsic) at interface_v5.py:369
__device__ inline bool invert4x4(const double* m, double* inv) {
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = m[4 * r + c]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(a[col][col]);
        for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > best) { best = fabs(a[r][col]); piv = r; }
        if (!(best > 0.0)) return false;
        if (piv != col) for (int c = 0; c < 8; ++c) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        const double d = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= d;
        for (int r = 0; r < 4; ++r) if (r != col) {
            const double f = a[r][col];
            if (f != 0.0) for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    for (int r = 0; r < 4; ++r) for (int c = 0; c < 4; ++c) inv[4 * r + c] = a[r][4 + c];
    return true;
}

// 8 fp32 values -> 8 e4m3 bytes of (value * 0.5), saturating (cvt.rn.satfinite.e4m3x2.f32): the fp8 twin of an fp16 activation
__device__ __forceinline__ uint2 pack8_q8(const float* v) {
    uint32_t w[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
        const uint32_t a = __nv_cvt_float2_to_fp8x2(make_float2(0.5f * v[4 * q], 0.5f * v[4 * q + 1]), __NV_SATFINITE, __NV_E4M3);
        const uint32_t b = __nv_cvt_float2_to_fp8x2(make_float2(0.5f * v[4 * q + 2], 0.5f * v[4 * q + 3]), __NV_SATFINITE, __NV_E4M3);
        w[q] = a | (b << 16);
    }
    return make_uint2(w[0], w[1]);
}

// max that PROPAGATES NaN (fmaxf returns the non-NaN operand): a ReLU / max-pool built on fmaxf would launder the NaNs that an
// fp16 overflow turns into (inf - inf) back into zeros and hide it from the range guard at the end of the backbone.
__device__ __forceinline__ float max_nan(float a, float b) {
    float r;
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// fp32 accumulator += fp16 x fp16 with no conversion instruction: the sm_100 mixed-precision FMA (PTX ISA 8.6 `fma.rn.f32.f16`,
// SASS `FHFMA Rd, Ra.H0|H1, Rb.H0|H1, Rc`; the register-pair moves below fold into the operand selectors).  Measured 114 FMA per
// clock and SM against 53 for the HADD2.F32 (fp16 -> fp32) + FFMA pair it replaces (tools/micro/fhfma_rate.cu).  The product of two
// halves is exact in fp32, so the only rounding beyond the fp32 path is that of the fp16 multiplier itself.
// v[0..7] += half(u's 8 channels) * half(w16's low half)
__device__ __forceinline__ void fhfma8(const uint4& u, uint32_t w16, float* v) {
    const uint32_t q[4] = {u.x, u.y, u.z, u.w};
    unsigned short wl, wh;
    asm("mov.b32 {%0,%1}, %2;" : "=h"(wl), "=h"(wh) : "r"(w16));
    (void)wh;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        unsigned short lo, hi;
        asm("mov.b32 {%0,%1}, %2;" : "=h"(lo), "=h"(hi) : "r"(q[k]));
        asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(v[2 * k]) : "h"(lo), "h"(wl));
        asm("fma.rn.f32.f16 %0, %1, %2, %0;" : "+f"(v[2 * k + 1]) : "h"(hi), "h"(wl));
    }
}

// 16-byte asynchronous global -> shared copy (LDGSTS, through L1) and the wait for all copies of this thread
__device__ __forceinline__ void cp_async16_ca(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) property of a kernel: one engine per GPU in the
// same process must set it on every device it launches on.  `per_dev` is a zero-initialised static table owned by the call
// site (one slot per device ordinal, holding the largest size configured so far on that device).
constexpr int kMaxDevices = 64;
template <typename Fn>
static inline int ensure_dyn_smem(Fn* fn, int bytes, int* per_dev) {
    int dev = 0;
    ADP_CUDA(cudaGetDevice(&dev));
    ADP_CHECK_ARG(dev >= 0 && dev < kMaxDevices, "device ordinal");
    if (bytes > per_dev[dev]) {
        ADP_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
        per_dev[dev] = bytes;
    }
    return ADP_OK;
}

}  // namespace adp
