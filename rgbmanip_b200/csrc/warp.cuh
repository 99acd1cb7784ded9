// Homography sampling helpers shared by the volume builder and the decode kernel
// (ADA/lib/network_v5.py:389-413: normalise with (W-1)/2, grid_sample align_corners=False, zeros padding).
#pragma once
#include "common.cuh"

namespace adp {

// sample coordinates of reference pixel (x, y) at depth `dep` in the source map; M = [rot(9) | trans(3)]
__device__ __forceinline__ void warp_coords(const float* __restrict__ M, float x, float y, float dep, int W, int H, float* ix,
                                            float* iy) {
    const float rx = M[0] * x + M[1] * y + M[2];
    const float ry = M[3] * x + M[4] * y + M[5];
    const float rz = M[6] * x + M[7] * y + M[8];
    const float X = rx * dep + M[9], Y = ry * dep + M[10], Z = rz * dep + M[11];
    const float px = X / Z, py = Y / Z;
    const float gx = px / ((float)(W - 1) / 2.f) - 1.f;
    const float gy = py / ((float)(H - 1) / 2.f) - 1.f;
    *ix = ((gx + 1.f) * (float)W - 1.f) / 2.f;
    *iy = ((gy + 1.f) * (float)H - 1.f) / 2.f;
}

// The same map for the volume builder's inner loop (one call per voxel, the kernel is instruction bound): r = rot * (x, y, 1)
// hoisted by the caller, one approximate reciprocal (1 ulp) shared by both coordinates and multiplications by the reciprocal
// of the constant divisors.  Agrees with warp_coords to a few fp32 ulps of the coordinate (~5e-5 px at 224), the size of the
// rounding noise the reference's own expression carries ((g + 1) - 1 round trip at magnitude 1).
__device__ __forceinline__ void warp_coords_rcp(const float* __restrict__ M, float rx, float ry, float rz, float dep, int W, int H,
                                                float* ix, float* iy) {
    const float X = rx * dep + M[9], Y = ry * dep + M[10], Z = rz * dep + M[11];
    float iz;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(iz) : "f"(Z));
    const float gx = (X * iz) * (2.f / (float)(W - 1)) - 1.f;
    const float gy = (Y * iz) * (2.f / (float)(H - 1)) - 1.f;
    *ix = ((gx + 1.f) * (float)W - 1.f) * 0.5f;
    *iy = ((gy + 1.f) * (float)H - 1.f) * 0.5f;
}

struct Bilin {
    int x0, y0;
    float w00, w01, w10, w11;   // (y0,x0) (y0,x0+1) (y0+1,x0) (y0+1,x0+1); 0 where the corner is outside
    bool any;
};

__device__ __forceinline__ Bilin bilin_setup(float ix, float iy, int W, int H) {
    Bilin b;
    b.any = false;
    b.x0 = b.y0 = 0;
    b.w00 = b.w01 = b.w10 = b.w11 = 0.f;
    if (!(isfinite(ix) && isfinite(iy)) || ix < -1.f || iy < -1.f || ix > (float)W || iy > (float)H) return b;
    const float fx = floorf(ix), fy = floorf(iy);
    b.x0 = (int)fx; b.y0 = (int)fy;
    const float ex = ix - fx, ey = iy - fy;
    const bool xin0 = b.x0 >= 0 && b.x0 < W, xin1 = b.x0 + 1 >= 0 && b.x0 + 1 < W;
    const bool yin0 = b.y0 >= 0 && b.y0 < H, yin1 = b.y0 + 1 >= 0 && b.y0 + 1 < H;
    b.w00 = (xin0 && yin0) ? (1.f - ex) * (1.f - ey) : 0.f;
    b.w01 = (xin1 && yin0) ? ex * (1.f - ey) : 0.f;
    b.w10 = (xin0 && yin1) ? (1.f - ex) * ey : 0.f;
    b.w11 = (xin1 && yin1) ? ex * ey : 0.f;
    b.any = (xin0 || xin1) && (yin0 || yin1);
    return b;
}

}  // namespace adp
