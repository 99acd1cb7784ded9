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
                    const float* __restrict__ depths, bf16* __restrict__ vol, int B, int D, int H, int W) {
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
            o[u] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * u])) |
                   ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * u + 1])) << 16);
        }
        *reinterpret_cast<uint4*>(vol + ((((size_t)b * D + d) * H + y) * W + x) * C + cg * 8) =
            make_uint4(o[0], o[1], o[2], o[3]);
    }
}

int build_volume(const float* f_ref, const float* f_src, const float* Mw, const float* depths, bf16* vol, int B, int D, int H,
                 int W, int C, cudaStream_t stream) {
    ADP_CHECK_ARG(C == 32, "feature channels must be 32");
    size_t total = (size_t)B * D * H * W * 4;
    if (total == 0) return ADP_OK;
    size_t blocks = (total + 255) / 256;
    int grid = (int)(blocks < (size_t)148 * 32 ? blocks : (size_t)148 * 32);
    build_volume_kernel<<<grid, 256, 0, stream>>>(f_ref, f_src, Mw, depths, vol, B, D, H, W);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

}  // namespace adp
