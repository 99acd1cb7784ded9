// Decode of the sampled pixels of the reference view (ADA/lib/network_v5.py:432-465,486-499):
//   depth logits   = the `prob` 3x3x3 conv (8 -> 1) evaluated ONLY at the P sampled pixels x D depths
//   prob           = softmax over the D depths; depth = sum_d prob_d * depth_d                 (soft-argmax)
//   fused feature  = sum_d prob_d * (feat_ref + warp(feat_src, d))                              (depth-guided fusion)
//   NOCS           = tanh(nocs_head(instance_color(feat_ref at the pixel)))
//   pose feature   = pose_mlp1(cat(fused, nocs_pts_mlp(NOCS)))  -> per-env mean -> pose_mlp2 -> mean -> rotation head
// Three kernels: decode_points (gather-heavy, memory-bound part + per-point MLPs), pose_mlp2, rotation head.
#include "../../include/adapose_b200.h"
#include "common.cuh"
#include "warp.cuh"

namespace adp {

constexpr int PB = 32;          // sampled pixels per block
constexpr int DEC_THREADS = 256;
constexpr int LDS = 260;        // padded row pitch (floats) of the activation ping-pong buffers, multiple of 4

struct DecodeWeights {
    // all fp32, transposed to [K][N] so that threads read consecutive output columns
    const float* ic_w;  const float* ic_b;      // instance_color 32 -> 64
    const float* nh0_w; const float* nh0_b;     // nocs_head 64 -> 128
    const float* nh1_w; const float* nh1_b;     // 128 -> 64
    const float* nh2_w; const float* nh2_b;     // 64 -> 3
    const float* np0_w; const float* np0_b;     // nocs_pts_mlp 3 -> 32
    const float* np1_w; const float* np1_b;     // 32 -> 64
    const float* pm0_w; const float* pm0_b;     // pose_mlp1 96 -> 128
    const float* pm1_w; const float* pm1_b;     // 128 -> 128
    const float* q0_w;  const float* q0_b;      // pose_mlp2 256 -> 256
    const float* q1_w;  const float* q1_b;      // 256 -> 256
    const float* r0_w;  const float* r0_b;      // rotation_estimator 256 -> 256
    const float* r1_w;  const float* r1_b;      // 256 -> 128
    const float* r2_w;  const float* r2_b;      // 128 -> 6
    const float* prob_w;                        // [27][8] depth-logit conv
};

// out[r][n] = act(bias[n] + sum_k in[r][k] * Wt[k][n]) for r < PB; N in {32,64,128,256}; K multiple of 4 (or K == 3)
template <int N>
__device__ __forceinline__ void mlp_layer(const float* __restrict__ in_s, int K, const float* __restrict__ Wt,
                                          const float* __restrict__ bias, float* __restrict__ out_s, int out_col0, bool relu,
                                          const float* __restrict__ row_bias = nullptr) {
    constexpr int G = DEC_THREADS / N;   // row groups
    constexpr int R = PB / G;            // rows per thread
    const int n = threadIdx.x % N, rg = threadIdx.x / N;
    float acc[R];
    const float b0 = row_bias ? row_bias[n] : (bias ? bias[n] : 0.f);
#pragma unroll
    for (int r = 0; r < R; ++r) acc[r] = b0;
    if ((K & 3) == 0) {
        for (int k = 0; k < K; k += 4) {
            const float w0 = __ldg(Wt + (size_t)(k + 0) * N + n), w1 = __ldg(Wt + (size_t)(k + 1) * N + n);
            const float w2 = __ldg(Wt + (size_t)(k + 2) * N + n), w3 = __ldg(Wt + (size_t)(k + 3) * N + n);
#pragma unroll
            for (int r = 0; r < R; ++r) {
                const float4 a = *reinterpret_cast<const float4*>(in_s + (rg * R + r) * LDS + k);
                acc[r] = fmaf(a.x, w0, acc[r]); acc[r] = fmaf(a.y, w1, acc[r]);
                acc[r] = fmaf(a.z, w2, acc[r]); acc[r] = fmaf(a.w, w3, acc[r]);
            }
        }
    } else {
        for (int k = 0; k < K; ++k) {
            const float w0 = __ldg(Wt + (size_t)k * N + n);
#pragma unroll
            for (int r = 0; r < R; ++r) acc[r] = fmaf(in_s[(rg * R + r) * LDS + k], w0, acc[r]);
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) out_s[(rg * R + r) * LDS + out_col0 + n] = relu ? fmaxf(acc[r], 0.f) : acc[r];
}

struct DecodeArgs {
    const float* feat_ref;   // [B,S,S,32] fp32
    const float* feat_src;   // [B,S,S,32]
    const float* Mw;         // [B,12]
    const float* depths;     // [D]
    const bf16* x11;         // [B,D,S,S,8] last U-Net activation (conv0 + deconv11)
    const int* choose;       // [B,P]
    const uint8_t* valid;    // [B] or nullptr
    float* nocs;             // [B,P,3]
    float* depth;            // [B,P]
    float* pf1;              // [B,P,128] pose_mlp1 output
    float* gsum;             // [B,128]  sum over P of pf1 (zeroed by the caller)
    float* dbg_logits;       // [B,P,D] or nullptr
    float* dbg_fused;        // [B,P,32] or nullptr
    int B, S, D, P;
    int x11_f16;             // x11 holds IEEE half instead of bf16
    int x11_s2d;             // x11 is stored space-to-depth(2): [B,D/2,S/2,S/2,64], channel = (pz*4+py*2+px)*8 + c
    int regress;             // 0: stop after NOCS + depth (no pose heads: direct_regression = False)
};

// ------------------------------------------------------------------------------------------------
// decode, stage 1 (the gather-bound part): depth logits at the sampled pixels (the `prob` conv evaluated only there),
// softmax over depth + soft-argmax, depth-guided fused feature, and the reference feature at the pixel.  The per-point
// MLPs then run as 1x1 convolutions on the tcgen05 kernel (engine.py), so this kernel writes their inputs as bf16 hi/lo.
//   xfeat [B,P,32]  (instance_color input)           xcat [B,P,96] columns 0..31 = fused feature (pose_mlp1 input)
// ------------------------------------------------------------------------------------------------
// element offset of the 8 channels of voxel (dz, yy, xx) of x11
__device__ __forceinline__ size_t x11_offset(const DecodeArgs& a, int b, int dz, int yy, int xx) {
    if (a.x11_s2d) {
        const int S2 = a.S >> 1;
        return ((((size_t)b * (a.D >> 1) + (dz >> 1)) * S2 + (yy >> 1)) * S2 + (xx >> 1)) * 64 + (((dz & 1) * 4 + (yy & 1) * 2 + (xx & 1)) * 8);
    }
    return ((((size_t)b * a.D + dz) * a.S + yy) * a.S + xx) * 8;
}

struct GatherArgs {
    DecodeArgs a;
    bf16* xfeat_hi; bf16* xfeat_lo;
    bf16* xcat_hi;  bf16* xcat_lo;
};

__global__ void __launch_bounds__(DEC_THREADS)
decode_gather_kernel(const GatherArgs ga, const float* __restrict__ prob_w) {
    const DecodeArgs& a = ga.a;
    extern __shared__ float smf[];
    float* bufA = smf;                   // [PB][LDS] (only columns 0..31 used)
    float* bufB = smf + PB * LDS;
    float* s_logit = smf + 2 * PB * LDS;
    float* s_probw = s_logit + PB * 24;
    int* s_pix = reinterpret_cast<int*>(s_probw + 216);
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * PB;
    if (a.valid && !a.valid[b]) return;
    const int S = a.S, D = a.D;
    for (int i = tid; i < 216; i += DEC_THREADS) s_probw[i] = prob_w[i];
    if (tid < PB) s_pix[tid] = a.choose[(size_t)b * a.P + p0 + tid];
    __syncthreads();

    // ---- (a) depth logits at the sampled pixels: 3x3x3 conv over the 8-channel volume, zero padding
    for (int e = tid; e < PB * D; e += DEC_THREADS) {
        const int r = e / D, d = e - r * D;
        const int pix = s_pix[r];
        const int y = pix / S, x = pix - y * S;
        float acc = 0.f;
        for (int kz = 0; kz < 3; ++kz) {
            const int dz = d + kz - 1;
            if (dz < 0 || dz >= D) continue;
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                if (yy < 0 || yy >= S) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= S) continue;
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.x11 + x11_offset(a, b, dz, yy, xx)));
                    const float* wt = s_probw + ((kz * 3 + ky) * 3 + kx) * 8;
                    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float lo16, hi16;
                        if (a.x11_f16) {
                            lo16 = __half2float(__ushort_as_half((unsigned short)(u[q] & 0xffffu)));
                            hi16 = __half2float(__ushort_as_half((unsigned short)(u[q] >> 16)));
                        } else {
                            lo16 = __uint_as_float(u[q] << 16);
                            hi16 = __uint_as_float(u[q] & 0xffff0000u);
                        }
                        acc = fmaf(lo16, wt[2 * q], acc);
                        acc = fmaf(hi16, wt[2 * q + 1], acc);
                    }
                }
            }
        }
        s_logit[r * 24 + d] = acc;
        if (a.dbg_logits) a.dbg_logits[((size_t)b * a.P + p0 + r) * D + d] = acc;
    }
    __syncthreads();

    // ---- (b) softmax over depth + expectation: 8 lanes per pixel, shuffle reductions
    {
        const int r = tid >> 3, l = tid & 7;   // 32 pixels x 8 lanes
        float v[3], m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 3; ++j) { v[j] = s_logit[r * 24 + l + 8 * j]; m = fmaxf(m, v[j]); }
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) { v[j] = expf(v[j] - m); s += v[j]; }
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        float dsum = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float pr = v[j] / s;
            s_logit[r * 24 + l + 8 * j] = pr;
            dsum = fmaf(pr, a.depths[l + 8 * j], dsum);
        }
        for (int o = 4; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        if (l == 0) a.depth[(size_t)b * a.P + p0 + r] = dsum;
    }
    __syncthreads();

    // ---- (c) reference features at the pixel -> bufA[:, 0:32]; depth-guided fused features -> bufB[:, 0:32]
    {
        const int r = tid >> 3, cq = tid & 7;   // 4 channels per thread, 128-bit loads
        const int pix = s_pix[r];
        const int y = pix / S, x = pix - y * S;
        const float4 fr = __ldg(reinterpret_cast<const float4*>(a.feat_ref + (((size_t)b * S + y) * S + x) * 32 + cq * 4));
        *reinterpret_cast<float4*>(bufA + r * LDS + cq * 4) = fr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* src = a.feat_src + (size_t)b * S * S * 32 + cq * 4;
        for (int d = 0; d < D; ++d) {
            const float pr = s_logit[r * 24 + d];
            float ix, iy;
            warp_coords(a.Mw + 12 * b, (float)x, (float)y, a.depths[d], S, S, &ix, &iy);
            const Bilin bl = bilin_setup(ix, iy, S, S);
            float4 v = fr;
            if (bl.any) {
                const float wts[4] = {bl.w00, bl.w01, bl.w10, bl.w11};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (wts[k] != 0.f) {
                        const float4 s4 = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(bl.y0 + (k >> 1)) * S + bl.x0 + (k & 1)) * 32));
                        v.x = fmaf(wts[k], s4.x, v.x); v.y = fmaf(wts[k], s4.y, v.y);
                        v.z = fmaf(wts[k], s4.z, v.z); v.w = fmaf(wts[k], s4.w, v.w);
                    }
                }
            }
            acc.x = fmaf(pr, v.x, acc.x); acc.y = fmaf(pr, v.y, acc.y);
            acc.z = fmaf(pr, v.z, acc.z); acc.w = fmaf(pr, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(bufB + r * LDS + cq * 4) = acc;
        if (a.dbg_fused) *reinterpret_cast<float4*>(a.dbg_fused + ((size_t)b * a.P + p0 + r) * 32 + cq * 4) = acc;
    }
    __syncthreads();

    // ---- write the MLP inputs as bf16 hi/lo
    {
        const int r = tid >> 3, c8 = (tid & 7) * 8;        // 32 pixels x 8 groups of 8 channels: groups 0-3 feat, 4-7 fused
        const size_t pt = (size_t)b * a.P + p0 + r;
        if (c8 < 32) st8_16(ga.xfeat_hi, ga.xfeat_lo, pt * 32 + c8, 0, bufA + r * LDS + c8);
        else st8_16(ga.xcat_hi, ga.xcat_lo, pt * 96 + (c8 - 32), 0, bufB + r * LDS + (c8 - 32));
    }
}

// column sums over the P points of an env: out[b, c] = sum_p x[b, p, c]   (hi + lo)
__global__ void colsum_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, const uint8_t* __restrict__ valid,
                              float* __restrict__ out, int P, int C, int rows_per_block) {
    const int b = blockIdx.y;
    if (valid && !valid[b]) return;
    const int r0 = blockIdx.x * rows_per_block;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int r = r0; r < r0 + rows_per_block && r < P; ++r) s += ld_act(hi, lo, ((size_t)b * P + r) * C + c);
        atomicAdd(out + (size_t)b * C + c, s);
    }
}

// per-env bias of pose_mlp2's first layer: gb[b, n] = bias[n] + sum_k W[128 + k][n] * (gsum[b, k] / P)   (network_v5.py:491-493)
__global__ void pose_gbias_kernel(const float* __restrict__ gsum, const float* __restrict__ q0_w, const float* __restrict__ q0_b,
                                  const uint8_t* __restrict__ valid, float* __restrict__ gb, int P) {
    __shared__ float gm[128];
    const int b = blockIdx.x, n = threadIdx.x;     // 256 threads
    if (valid && !valid[b]) return;
    if (n < 128) gm[n] = gsum[(size_t)b * 128 + n] / (float)P;
    __syncthreads();
    float acc = q0_b[n];
    for (int k = 0; k < 128; ++k) acc = fmaf(gm[k], __ldg(q0_w + (size_t)(128 + k) * 256 + n), acc);
    gb[(size_t)b * 256 + n] = acc;
}

__global__ void __launch_bounds__(DEC_THREADS)
decode_points_kernel(const DecodeArgs a, const DecodeWeights w) {
    extern __shared__ float smf[];
    float* bufA = smf;                   // [PB][LDS]
    float* bufB = smf + PB * LDS;        // [PB][LDS]
    float* s_logit = smf + 2 * PB * LDS; // [PB][24] -> probabilities
    float* s_probw = s_logit + PB * 24;  // [27*8]
    int* s_pix = reinterpret_cast<int*>(s_probw + 216);   // [PB]
    const int tid = threadIdx.x;
    const int b = blockIdx.y;
    const int p0 = blockIdx.x * PB;
    if (a.valid && !a.valid[b]) return;
    const int S = a.S, D = a.D;
    for (int i = tid; i < 216; i += DEC_THREADS) s_probw[i] = w.prob_w[i];
    if (tid < PB) s_pix[tid] = a.choose[(size_t)b * a.P + p0 + tid];
    __syncthreads();

    // ---- (a) depth logits at the sampled pixels: 3x3x3 conv over the 8-channel volume, zero padding
    for (int e = tid; e < PB * D; e += DEC_THREADS) {
        const int r = e / D, d = e - r * D;
        const int pix = s_pix[r];
        const int y = pix / S, x = pix - y * S;
        float acc = 0.f;
        for (int kz = 0; kz < 3; ++kz) {
            const int dz = d + kz - 1;
            if (dz < 0 || dz >= D) continue;
            for (int ky = 0; ky < 3; ++ky) {
                const int yy = y + ky - 1;
                if (yy < 0 || yy >= S) continue;
#pragma unroll
                for (int kx = 0; kx < 3; ++kx) {
                    const int xx = x + kx - 1;
                    if (xx < 0 || xx >= S) continue;
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(a.x11 + x11_offset(a, b, dz, yy, xx)));
                    const float* wt = s_probw + ((kz * 3 + ky) * 3 + kx) * 8;
                    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float lo16, hi16;
                        if (a.x11_f16) {
                            lo16 = __half2float(__ushort_as_half((unsigned short)(u[q] & 0xffffu)));
                            hi16 = __half2float(__ushort_as_half((unsigned short)(u[q] >> 16)));
                        } else {
                            lo16 = __uint_as_float(u[q] << 16);
                            hi16 = __uint_as_float(u[q] & 0xffff0000u);
                        }
                        acc = fmaf(lo16, wt[2 * q], acc);
                        acc = fmaf(hi16, wt[2 * q + 1], acc);
                    }
                }
            }
        }
        s_logit[r * 24 + d] = acc;
        if (a.dbg_logits) a.dbg_logits[((size_t)b * a.P + p0 + r) * D + d] = acc;
    }
    __syncthreads();

    // ---- (b) softmax over depth + expectation: 8 lanes per pixel, shuffle reductions
    {
        const int r = tid >> 3, l = tid & 7;   // 32 pixels x 8 lanes
        float v[3], m = -INFINITY;
#pragma unroll
        for (int j = 0; j < 3; ++j) { v[j] = s_logit[r * 24 + l + 8 * j]; m = fmaxf(m, v[j]); }
        for (int o = 4; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) { v[j] = expf(v[j] - m); s += v[j]; }
        for (int o = 4; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        float dsum = 0.f;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float pr = v[j] / s;
            s_logit[r * 24 + l + 8 * j] = pr;
            dsum = fmaf(pr, a.depths[l + 8 * j], dsum);
        }
        for (int o = 4; o > 0; o >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o);
        if (l == 0) a.depth[(size_t)b * a.P + p0 + r] = dsum;
    }
    __syncthreads();

    // ---- (c) reference features at the pixel -> bufA[:, 0:32]; depth-guided fused features -> bufB[:, 0:32]
    {
        const int r = tid >> 3, cq = tid & 7;   // 4 channels per thread, 128-bit loads
        const int pix = s_pix[r];
        const int y = pix / S, x = pix - y * S;
        const float4 fr = __ldg(reinterpret_cast<const float4*>(a.feat_ref + (((size_t)b * S + y) * S + x) * 32 + cq * 4));
        *reinterpret_cast<float4*>(bufA + r * LDS + cq * 4) = fr;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        const float* src = a.feat_src + (size_t)b * S * S * 32 + cq * 4;
        for (int d = 0; d < D; ++d) {
            const float pr = s_logit[r * 24 + d];
            float ix, iy;
            warp_coords(a.Mw + 12 * b, (float)x, (float)y, a.depths[d], S, S, &ix, &iy);
            const Bilin bl = bilin_setup(ix, iy, S, S);
            float4 v = fr;
            if (bl.any) {
                const float wts[4] = {bl.w00, bl.w01, bl.w10, bl.w11};
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (wts[k] != 0.f) {
                        const float4 s4 = __ldg(reinterpret_cast<const float4*>(src + ((size_t)(bl.y0 + (k >> 1)) * S + bl.x0 + (k & 1)) * 32));
                        v.x = fmaf(wts[k], s4.x, v.x); v.y = fmaf(wts[k], s4.y, v.y);
                        v.z = fmaf(wts[k], s4.z, v.z); v.w = fmaf(wts[k], s4.w, v.w);
                    }
                }
            }
            acc.x = fmaf(pr, v.x, acc.x); acc.y = fmaf(pr, v.y, acc.y);
            acc.z = fmaf(pr, v.z, acc.z); acc.w = fmaf(pr, v.w, acc.w);
        }
        *reinterpret_cast<float4*>(bufB + r * LDS + cq * 4) = acc;
        if (a.dbg_fused) *reinterpret_cast<float4*>(a.dbg_fused + ((size_t)b * a.P + p0 + r) * 32 + cq * 4) = acc;
    }
    __syncthreads();

    // ---- (d) NOCS head: 32 -> 64 -> 128 -> 64 -> 3 (tanh).  bufA holds the input; outputs alternate A/B above col 32
    mlp_layer<64>(bufA, 32, w.ic_w, w.ic_b, bufA, 64, true);            // A[:, 64:128]
    __syncthreads();
    mlp_layer<128>(bufA + 64, 64, w.nh0_w, w.nh0_b, bufB, 64, true);    // B[:, 64:192]
    __syncthreads();
    mlp_layer<64>(bufB + 64, 128, w.nh1_w, w.nh1_b, bufA, 128, true);   // A[:, 128:192]
    __syncthreads();
    if (tid < PB * 3) {
        const int r = tid / 3, n = tid - r * 3;
        float acc = w.nh2_b[n];
        for (int k = 0; k < 64; ++k) acc = fmaf(bufA[r * LDS + 128 + k], w.nh2_w[k * 3 + n], acc);
        const float t = tanhf(acc);
        bufA[r * LDS + n] = t;                                           // A[:, 0:3] = NOCS
        a.nocs[((size_t)b * a.P + p0 + r) * 3 + n] = t;
    }
    __syncthreads();
    if (!a.regress) return;
    // ---- (e) nocs_pts_mlp 3 -> 32 -> 64, concat with fused (B[:, 0:32]) -> pose_mlp1 96 -> 128 -> 128
    mlp_layer<32>(bufA, 3, w.np0_w, w.np0_b, bufA, 32, true);           // A[:, 32:64]
    __syncthreads();
    mlp_layer<64>(bufA + 32, 32, w.np1_w, w.np1_b, bufB, 32, true);     // B[:, 32:96]  (B[:, 0:96] = cat(fused, pts))
    __syncthreads();
    mlp_layer<128>(bufB, 96, w.pm0_w, w.pm0_b, bufA, 0, true);          // A[:, 0:128]
    __syncthreads();
    mlp_layer<128>(bufA, 128, w.pm1_w, w.pm1_b, bufB, 0, true);         // B[:, 0:128] = pose feature
    __syncthreads();
    // ---- (f) store pose features, accumulate their per-env sum
    for (int e = tid; e < PB * 128; e += DEC_THREADS) {
        const int r = e >> 7, n = e & 127;
        a.pf1[((size_t)b * a.P + p0 + r) * 128 + n] = bufB[r * LDS + n];
    }
    if (tid < 128) {
        float s = 0.f;
        for (int r = 0; r < PB; ++r) s += bufB[r * LDS + tid];
        atomicAdd(a.gsum + (size_t)b * 128 + tid, s);
    }
}

// pose_mlp2 on cat(pf1, mean_P(pf1)) -> 256 -> 256, summed over the P pixels of the env (AdaptiveAvgPool1d)
__global__ void __launch_bounds__(DEC_THREADS)
pose_mlp2_kernel(const float* __restrict__ pf1, const float* __restrict__ gsum, const uint8_t* __restrict__ valid,
                 float* __restrict__ psum, const DecodeWeights w, int P) {
    extern __shared__ float smf[];
    float* bufA = smf;
    float* bufB = smf + PB * LDS;
    float* gvec = smf + 2 * PB * LDS;     // [256] bias + W0[128:256]^T mean
    float* gmean = gvec + 256;            // [128]
    const int tid = threadIdx.x;
    const int b = blockIdx.y, p0 = blockIdx.x * PB;
    if (valid && !valid[b]) return;
    if (tid < 128) gmean[tid] = gsum[(size_t)b * 128 + tid] / (float)P;
    for (int e = tid; e < PB * 128; e += DEC_THREADS) {
        const int r = e >> 7, n = e & 127;
        bufA[r * LDS + n] = pf1[((size_t)b * P + p0 + r) * 128 + n];
    }
    __syncthreads();
    {
        float acc = w.q0_b[tid];
        for (int k = 0; k < 128; ++k) acc = fmaf(gmean[k], __ldg(w.q0_w + (size_t)(128 + k) * 256 + tid), acc);
        gvec[tid] = acc;
    }
    __syncthreads();
    mlp_layer<256>(bufA, 128, w.q0_w, nullptr, bufB, 0, true, gvec);
    __syncthreads();
    mlp_layer<256>(bufB, 256, w.q1_w, w.q1_b, bufA, 0, true);
    __syncthreads();
    float s = 0.f;
    for (int r = 0; r < PB; ++r) s += bufA[r * LDS + tid];
    atomicAdd(psum + (size_t)b * 256 + tid, s);
}

// rotation head 256 -> 256 -> 128 -> 6 and the 6-D -> SO(3) map (ADA/lib/rotation_utils.py:4-27)
__global__ void __launch_bounds__(256)
rot_head_kernel(const float* __restrict__ psum, const uint8_t* __restrict__ valid, float* __restrict__ Rout, float* __restrict__ r6out,
                const DecodeWeights w, int P) {
    __shared__ float v0[256], v1[256], v2[128], r6[6];
    const int b = blockIdx.x, tid = threadIdx.x;
    if (valid && !valid[b]) {
        if (tid < 9) Rout[9 * b + tid] = (tid % 4 == 0) ? 1.f : 0.f;
        return;
    }
    v0[tid] = psum[(size_t)b * 256 + tid] / (float)P;
    __syncthreads();
    {
        float acc = w.r0_b[tid];
        for (int k = 0; k < 256; ++k) acc = fmaf(v0[k], __ldg(w.r0_w + (size_t)k * 256 + tid), acc);
        v1[tid] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (tid < 128) {
        float acc = w.r1_b[tid];
        for (int k = 0; k < 256; ++k) acc = fmaf(v1[k], __ldg(w.r1_w + (size_t)k * 128 + tid), acc);
        v2[tid] = fmaxf(acc, 0.f);
    }
    __syncthreads();
    if (tid < 6) {
        float acc = w.r2_b[tid];
        for (int k = 0; k < 128; ++k) acc = fmaf(v2[k], w.r2_w[k * 6 + tid], acc);
        r6[tid] = acc;
        if (r6out) r6out[6 * b + tid] = acc;
    }
    __syncthreads();
    if (tid == 0) {
        const float xr[3] = {r6[0], r6[1], r6[2]};
        float y[3] = {r6[3], r6[4], r6[5]};
        float n = fmaxf(sqrtf(y[0] * y[0] + y[1] * y[1] + y[2] * y[2]), 1e-8f);
        y[0] /= n; y[1] /= n; y[2] /= n;
        float z[3] = {xr[1] * y[2] - xr[2] * y[1], xr[2] * y[0] - xr[0] * y[2], xr[0] * y[1] - xr[1] * y[0]};
        n = fmaxf(sqrtf(z[0] * z[0] + z[1] * z[1] + z[2] * z[2]), 1e-8f);
        z[0] /= n; z[1] /= n; z[2] /= n;
        const float x[3] = {y[1] * z[2] - y[2] * z[1], y[2] * z[0] - y[0] * z[2], y[0] * z[1] - y[1] * z[0]};
        float* R = Rout + 9 * b;   // columns [x y z]
        R[0] = x[0]; R[1] = y[0]; R[2] = z[0];
        R[3] = x[1]; R[4] = y[1]; R[5] = z[1];
        R[6] = x[2]; R[7] = y[2]; R[8] = z[2];
    }
}

int decode_run(const DecodeArgs& a, const DecodeWeights& w, float* psum, float* Rout, float* r6out, int regress_pose,
               cudaStream_t stream) {
    ADP_CHECK_ARG(a.D == 24, "24 depth hypotheses expected");
    ADP_CHECK_ARG(a.P % PB == 0, "P must be a multiple of 32");
    if (a.B == 0) return ADP_OK;
    const size_t smem1 = (size_t)(2 * PB * LDS + PB * 24 + 216) * sizeof(float) + PB * sizeof(int);
    const size_t smem2 = (size_t)(2 * PB * LDS + 256 + 128) * sizeof(float);
    static bool attr = false;
    if (!attr) {
        ADP_CUDA(cudaFuncSetAttribute(decode_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        ADP_CUDA(cudaFuncSetAttribute(pose_mlp2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
        attr = true;
    }
    ADP_CUDA(cudaMemsetAsync(a.gsum, 0, (size_t)a.B * 128 * sizeof(float), stream));
    ADP_CUDA(cudaMemsetAsync(psum, 0, (size_t)a.B * 256 * sizeof(float), stream));
    decode_points_kernel<<<dim3(a.P / PB, a.B), DEC_THREADS, smem1, stream>>>(a, w);
    ADP_CUDA(cudaGetLastError());
    if (regress_pose) {
        pose_mlp2_kernel<<<dim3(a.P / PB, a.B), DEC_THREADS, smem2, stream>>>(a.pf1, a.gsum, a.valid, psum, w, a.P);
        ADP_CUDA(cudaGetLastError());
        rot_head_kernel<<<a.B, 256, 0, stream>>>(psum, a.valid, Rout, r6out, w, a.P);
        ADP_CUDA(cudaGetLastError());
    }
    return ADP_OK;
}

int decode_gather_run(const DecodeArgs& a, const float* prob_w, bf16* xfeat_hi, bf16* xfeat_lo, bf16* xcat_hi, bf16* xcat_lo,
                      cudaStream_t stream) {
    ADP_CHECK_ARG(a.D == 24 && a.P % PB == 0, "24 depth hypotheses, P multiple of 32");
    if (a.B == 0) return ADP_OK;
    const size_t smem1 = (size_t)(2 * PB * LDS + PB * 24 + 216) * sizeof(float) + PB * sizeof(int);
    static bool attr = false;
    if (!attr) {
        ADP_CUDA(cudaFuncSetAttribute(decode_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem1));
        attr = true;
    }
    GatherArgs ga;
    ga.a = a; ga.xfeat_hi = xfeat_hi; ga.xfeat_lo = xfeat_lo; ga.xcat_hi = xcat_hi; ga.xcat_lo = xcat_lo;
    decode_gather_kernel<<<dim3(a.P / PB, a.B), DEC_THREADS, smem1, stream>>>(ga, prob_w);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

int colsum_run(const bf16* hi, const bf16* lo, const uint8_t* valid, float* out, int B, int P, int C, cudaStream_t stream) {
    if (B == 0) return ADP_OK;
    ADP_CUDA(cudaMemsetAsync(out, 0, (size_t)B * C * sizeof(float), stream));
    const int rpb = 64;
    colsum_kernel<<<dim3((P + rpb - 1) / rpb, B), C < 256 ? C : 256, 0, stream>>>(hi, lo, valid, out, P, C, rpb);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

int pose_gbias_run(const float* gsum, const float* q0_w, const float* q0_b, const uint8_t* valid, float* gb, int B, int P,
                   cudaStream_t stream) {
    if (B == 0) return ADP_OK;
    pose_gbias_kernel<<<B, 256, 0, stream>>>(gsum, q0_w, q0_b, valid, gb, P);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

int rot_head_run(const float* psum, const uint8_t* valid, float* Rout, float* r6out, const adp_decode_weights* cw, int B, int P,
                 cudaStream_t stream) {
    if (B == 0) return ADP_OK;
    DecodeWeights w{};
    w.r0_w = cw->r0_w; w.r0_b = cw->r0_b; w.r1_w = cw->r1_w; w.r1_b = cw->r1_b; w.r2_w = cw->r2_w; w.r2_b = cw->r2_b;
    rot_head_kernel<<<B, 256, 0, stream>>>(psum, valid, Rout, r6out, w, P);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

int decode_gather_c(const float* feat_ref, const float* feat_src, const float* Mw, const float* depths, const void* x11,
                    const int* choose, const uint8_t* valid, const float* prob_w, float* depth, void* xfeat_hi, void* xfeat_lo,
                    void* xcat_hi, void* xcat_lo, float* dbg_logits, float* dbg_fused, int B, int S, int D, int P, int x11_f16,
                    cudaStream_t stream) {
    DecodeArgs a{};
    a.feat_ref = feat_ref; a.feat_src = feat_src; a.Mw = Mw; a.depths = depths; a.x11 = reinterpret_cast<const bf16*>(x11);
    a.choose = choose; a.valid = valid; a.nocs = nullptr; a.depth = depth; a.pf1 = nullptr; a.gsum = nullptr;
    a.dbg_logits = dbg_logits; a.dbg_fused = dbg_fused; a.B = B; a.S = S; a.D = D; a.P = P; a.x11_f16 = x11_f16 & 1; a.x11_s2d = (x11_f16 >> 1) & 1; a.regress = 0;
    return decode_gather_run(a, prob_w, reinterpret_cast<bf16*>(xfeat_hi), reinterpret_cast<bf16*>(xfeat_lo),
                             reinterpret_cast<bf16*>(xcat_hi), reinterpret_cast<bf16*>(xcat_lo), stream);
}

int decode_run_c(const float* feat_ref, const float* feat_src, const float* Mw, const float* depths, const void* x11,
                 const int* choose, const uint8_t* valid, const adp_decode_weights* cw, float* nocs, float* depth, float* pf1,
                 float* gsum, float* psum, float* R, float* r6, float* dbg_logits, float* dbg_fused, int B, int S, int D, int P,
                 int regress_pose, int x11_f16, cudaStream_t stream) {
    DecodeWeights w;
    w.ic_w = cw->ic_w; w.ic_b = cw->ic_b; w.nh0_w = cw->nh0_w; w.nh0_b = cw->nh0_b; w.nh1_w = cw->nh1_w; w.nh1_b = cw->nh1_b;
    w.nh2_w = cw->nh2_w; w.nh2_b = cw->nh2_b; w.np0_w = cw->np0_w; w.np0_b = cw->np0_b; w.np1_w = cw->np1_w; w.np1_b = cw->np1_b;
    w.pm0_w = cw->pm0_w; w.pm0_b = cw->pm0_b; w.pm1_w = cw->pm1_w; w.pm1_b = cw->pm1_b; w.q0_w = cw->q0_w; w.q0_b = cw->q0_b;
    w.q1_w = cw->q1_w; w.q1_b = cw->q1_b; w.r0_w = cw->r0_w; w.r0_b = cw->r0_b; w.r1_w = cw->r1_w; w.r1_b = cw->r1_b;
    w.r2_w = cw->r2_w; w.r2_b = cw->r2_b; w.prob_w = cw->prob_w;
    DecodeArgs a;
    a.feat_ref = feat_ref; a.feat_src = feat_src; a.Mw = Mw; a.depths = depths; a.x11 = reinterpret_cast<const bf16*>(x11);
    a.choose = choose; a.valid = valid; a.nocs = nocs; a.depth = depth; a.pf1 = pf1; a.gsum = gsum;
    a.dbg_logits = dbg_logits; a.dbg_fused = dbg_fused; a.B = B; a.S = S; a.D = D; a.P = P; a.x11_f16 = x11_f16 & 1; a.x11_s2d = (x11_f16 >> 1) & 1; a.regress = regress_pose;
    return decode_run(a, w, psum, R, r6, regress_pose, stream);
}

}  // namespace adp
