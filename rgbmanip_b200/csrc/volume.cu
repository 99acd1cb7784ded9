// Plane-sweep volume: homography warp of the other view's feature map into the reference view for every
// depth hypothesis, fused with the reference features (ADA/lib/network_v5.py:378-416,429):
//   vol[b, d, y, x, :] = feat_ref[b, y, x, :] + bilinear(feat_src[b], u(d, y, x), v(d, y, x))
// The coordinate convention is reproduced as written: normalise with (W-1)/2, sample with
// align_corners=False, zeros padding.
#include "common.cuh"
#include "warp.cuh"

namespace adp {

// one thread = one voxel x 8 channels (C == 32 -> 4 threads per voxel, 128 B coalesced per corner)
__global__ void __launch_bounds__(256)
build_volume_kernel(const float* __restrict__ f_ref, const float* __restrict__ f_src, const float* __restrict__ Mw,
                    const float* __restrict__ depths, bf16* __restrict__ vol, int B, int D, int H, int W, int f16) {
    constexpr int C = 32;
    const size_t total = (size_t)B * D * H * W * 4;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int cg = (int)(i & 3);
        size_t t = i >> 2;
        const int x = (int)(t % W); t /= W;
        const int y = (int)(t % H); t /= H;
        const int d = (int)(t % D);
        const int b = (int)(t / D);
        float ix, iy;
        warp_coords(Mw + 12 * b, (float)x, (float)y, depths[d], W, H, &ix, &iy);
        const Bilin bl = bilin_setup(ix, iy, W, H);
        const float* ref = f_ref + (((size_t)b * H + y) * W + x) * C + cg * 8;
        float4 a0 = *reinterpret_cast<const float4*>(ref), a1 = *reinterpret_cast<const float4*>(ref + 4);
        float v[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        if (bl.any) {
            const float* src = f_src + (size_t)b * H * W * C + cg * 8;
            const float wts[4] = {bl.w00, bl.w01, bl.w10, bl.w11};
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (wts[k] != 0.f) {
                    const int yy = bl.y0 + (k >> 1), xx = bl.x0 + (k & 1);
                    const float* p = src + ((size_t)yy * W + xx) * C;
                    const float4 s0 = __ldg(reinterpret_cast<const float4*>(p));
                    const float4 s1 = __ldg(reinterpret_cast<const float4*>(p + 4));
                    v[0] = fmaf(wts[k], s0.x, v[0]); v[1] = fmaf(wts[k], s0.y, v[1]);
                    v[2] = fmaf(wts[k], s0.z, v[2]); v[3] = fmaf(wts[k], s0.w, v[3]);
                    v[4] = fmaf(wts[k], s1.x, v[4]); v[5] = fmaf(wts[k], s1.y, v[5]);
                    v[6] = fmaf(wts[k], s1.z, v[6]); v[7] = fmaf(wts[k], s1.w, v[7]);
                }
            }
        }
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            if (f16)
                o[u] = (uint32_t)__half_as_ushort(__float2half_rn(v[2 * u])) | ((uint32_t)__half_as_ushort(__float2half_rn(v[2 * u + 1])) << 16);
            else
                o[u] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * u])) |
                       ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * u + 1])) << 16);
        }
        *reinterpret_cast<uint4*>(vol + ((((size_t)b * D + d) * H + y) * W + x) * C + cg * 8) =
            make_uint4(o[0], o[1], o[2], o[3]);
    }
}

int build_volume(const float* f_ref, const float* f_src, const float* Mw, const float* depths, bf16* vol, int B, int D, int H,
                 int W, int C, int f16, cudaStream_t stream) {
    ADP_CHECK_ARG(C == 32, "feature channels must be 32");
    size_t total = (size_t)B * D * H * W * 4;
    if (total == 0) return ADP_OK;
    size_t blocks = (total + 255) / 256;
    int grid = (int)(blocks < (size_t)148 * 32 ? blocks : (size_t)148 * 32);
    build_volume_kernel<<<grid, 256, 0, stream>>>(f_ref, f_src, Mw, depths, vol, B, D, H, W, f16);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

// M = P_src * inv(P_ref) with P = [K' E[:3,:]; 0 0 0 1] (ADA/interface_v5.py:264-270, network_v5.py:389-392), in fp64.
// Mw[b] = {rot(9) row-major, trans(3)} as fp32.  Kp_* [B,9], E_* [B,16] row-major doubles.
__global__ void warp_matrices_kernel(const double* __restrict__ Kp_ref, const double* __restrict__ E_ref,
                                     const double* __restrict__ Kp_src, const double* __restrict__ E_src, float* __restrict__ Mw,
                                     const uint8_t* __restrict__ valid_ref, const uint8_t* __restrict__ valid_src,
                                     uint8_t* __restrict__ valid_env, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    // an estimate needs both views (interface_v5.py:256-257)
    if (valid_env) valid_env[b] = (valid_ref ? valid_ref[b] : 1) && (valid_src ? valid_src[b] : 1);
    double Pr[16], Ps[16];
    for (int which = 0; which < 2; ++which) {
        const double* K = (which ? Kp_src : Kp_ref) + 9 * b;
        const double* E = (which ? E_src : E_ref) + 16 * b;
        double* P = which ? Ps : Pr;
        for (int r = 0; r < 3; ++r)
            for (int c = 0; c < 4; ++c) P[4 * r + c] = K[3 * r] * E[c] + K[3 * r + 1] * E[4 + c] + K[3 * r + 2] * E[8 + c];
        P[12] = 0; P[13] = 0; P[14] = 0; P[15] = 1;
    }
    // Gauss-Jordan inverse of P_ref
    double a[4][8];
    for (int r = 0; r < 4; ++r)
        for (int c = 0; c < 4; ++c) { a[r][c] = Pr[4 * r + c]; a[r][4 + c] = (r == c) ? 1.0 : 0.0; }
    for (int col = 0; col < 4; ++col) {
        int piv = col;
        double best = fabs(a[col][col]);
        for (int r = col + 1; r < 4; ++r) if (fabs(a[r][col]) > best) { best = fabs(a[r][col]); piv = r; }
        if (piv != col) for (int c = 0; c < 8; ++c) { double t = a[col][c]; a[col][c] = a[piv][c]; a[piv][c] = t; }
        const double d = 1.0 / a[col][col];
        for (int c = 0; c < 8; ++c) a[col][c] *= d;
        for (int r = 0; r < 4; ++r) if (r != col) {
            const double f = a[r][col];
            for (int c = 0; c < 8; ++c) a[r][c] -= f * a[col][c];
        }
    }
    float* out = Mw + 12 * b;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
            double s = 0;
            for (int k = 0; k < 4; ++k) s += Ps[4 * r + k] * a[k][4 + c];
            if (c < 3) out[3 * r + c] = (float)s; else out[9 + r] = (float)s;
        }
    }
}

int warp_matrices(const double* Kp_ref, const double* E_ref, const double* Kp_src, const double* E_src, float* Mw,
                  const uint8_t* valid_ref, const uint8_t* valid_src, uint8_t* valid_env, int B, cudaStream_t stream) {
    if (B == 0) return ADP_OK;
    warp_matrices_kernel<<<cdiv(B, 64), 64, 0, stream>>>(Kp_ref, E_ref, Kp_src, E_src, Mw, valid_ref, valid_src, valid_env, B);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

}  // namespace adp
