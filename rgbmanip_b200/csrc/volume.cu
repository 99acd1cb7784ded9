// Plane-sweep volume: homography warp of the other view's feature map into the reference view for every
// depth hypothesis, fused with the reference features (ADA/lib/network_v5.py:378-416,429):
//   vol[b, d, y, x, :] = feat_ref[b, y, x, :] + bilinear(feat_src[b], u(d, y, x), v(d, y, x))
// The coordinate convention is reproduced as written: normalise with (W-1)/2, sample with
// align_corners=False, zeros padding.
#include "common.cuh"
#include "warp.cuh"

namespace adp {

// Tiled builder (fp16 features): one block = a 16 x 16 tile of reference pixels for VT_DG consecutive depth planes.
// The plane-sweep gather reads every source pixel ~4 times (once per bilinear footprint that covers it); done per voxel
// from global memory that is 256 B of L2 traffic per 64 B written, and the kernel is L2-bandwidth bound.  Here the
// source footprint of the tile at one depth (the bounding box of its 256 sample cells, found with a block min/max) is
// staged in shared memory once and the four corners are read from there; the reference features stay in registers
// across the depth loop.  A footprint larger than the staging buffer (strong rotation / scale between the views) falls
// back to the direct gather for that (tile, depth).
constexpr int VT_T = 16;
constexpr int VT_DG = 8;
constexpr int VT_MAXPX = 480;                 // staged source pixels (x 64 B = 30 KB; + 16 KB of transpose buffers < 48 KB static)

__device__ __forceinline__ void vt_unpack8(const uint4& u, float* v) {
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
        v[2 * q] = f.x; v[2 * q + 1] = f.y;
    }
}

__device__ __forceinline__ void vt_pack8(const float* v, int f16, uint4* o) {
    uint32_t w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if (f16) {
            const __half2 h = __floats2half2_rn(v[2 * q], v[2 * q + 1]);
            w[q] = *reinterpret_cast<const uint32_t*>(&h);
        } else {
            w[q] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * q])) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(v[2 * q + 1])) << 16);
        }
    }
    *o = make_uint4(w[0], w[1], w[2], w[3]);
}

__global__ void __launch_bounds__(256, 4)
build_volume_tile_kernel(const __half* __restrict__ f_ref, const __half* __restrict__ f_src, const float* __restrict__ Mw,
                         const float* __restrict__ depths, bf16* __restrict__ vol, int D, int H, int W, int tiles_x, int f16) {
    constexpr int C = 32;
    __shared__ uint4 stage[4][VT_MAXPX];        // [16-byte channel chunk][pixel]: consecutive lanes -> consecutive words
    __shared__ uint4 obuf[8][32 * 4];           // per-warp output transpose: 32 voxels x 64 B
    __shared__ int s_box[4];
    __shared__ unsigned char s_empty[VT_DG];
    const int tid = threadIdx.x;
    const int tyi = blockIdx.x / tiles_x, txi = blockIdx.x - tyi * tiles_x;
    const int b = blockIdx.z, d0 = blockIdx.y * VT_DG;
    const int x = txi * VT_T + (tid & (VT_T - 1)), y = tyi * VT_T + (tid >> 4);
    const bool inb = x < W && y < H;
    const float* M = Mw + 12 * b;
    const __half* src = f_src + (size_t)b * H * W * C;
    uint4 ref[4];
    if (inb) {
        const uint4* rp = reinterpret_cast<const uint4*>(f_ref + (((size_t)b * H + y) * W + x) * C);
#pragma unroll
        for (int q = 0; q < 4; ++q) ref[q] = __ldg(rp + q);
    }
    // Cheap emptiness test per depth plane: with positive projective depth at the four tile corners the tile maps to the convex
    // hull of its corner images, so if all four fall outside the source map on the same side no voxel of the tile samples
    // anything and the plane is the reference features bit for bit (one warp: 8 planes x 4 corners; 0.01 px safety margin).
    if (tid < 32) {
        const int dd = tid >> 2, c = tid & 3;
        const int cx = min(txi * VT_T + ((c & 1) ? VT_T - 1 : 0), W - 1), cy = min(tyi * VT_T + ((c & 2) ? VT_T - 1 : 0), H - 1);
        bool bad = true, left = false, right = false, top = false, bottom = false;
        if (d0 + dd < D) {
            const float dep = depths[d0 + dd];
            const float Z = (M[6] * (float)cx + M[7] * (float)cy + M[8]) * dep + M[11];
            float ix, iy;
            warp_coords(M, (float)cx, (float)cy, dep, W, H, &ix, &iy);
            bad = !(Z > 1e-6f) || !isfinite(ix) || !isfinite(iy);
            left = ix < -1.01f; right = ix > (float)W + 0.01f; top = iy < -1.01f; bottom = iy > (float)H + 0.01f;
        }
        unsigned int bits = (bad ? 1u : 0u) | (left ? 0u : 2u) | (right ? 0u : 4u) | (top ? 0u : 8u) | (bottom ? 0u : 16u);   // OR-reducible
        bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
        bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
        // bit 0 clear: every corner is in front of the camera; bit k clear (k = 1..4): every corner is outside on side k
        if (c == 0) s_empty[dd] = (!(bits & 1u) && ((bits & 30u) != 30u)) ? 1 : 0;
    }
    __syncthreads();
    for (int d = d0; d < d0 + VT_DG && d < D; ++d) {
        if (s_empty[d - d0] && f16) {       // uniform over the block
            const int ln = tid & 31;
            uint4* obw = obuf[tid >> 5];
#pragma unroll
            for (int q = 0; q < 4; ++q) obw[ln * 4 + (q ^ ((ln >> 1) & 3))] = ref[q];
            __syncwarp();
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const int vx = 8 * k + (ln >> 2), c = ln & 3;
                const uint4 o = obw[vx * 4 + (c ^ ((vx >> 1) & 3))];
                const int oy = tyi * VT_T + (tid >> 5) * 2 + (vx >> 4), ox = txi * VT_T + (vx & 15);
                if (oy < H && ox < W)
                    *(reinterpret_cast<uint4*>(vol + ((((size_t)b * D + d) * H + oy) * W + ox) * C) + c) = o;
            }
            __syncwarp();
            continue;
        }
        if (tid == 0) { s_box[0] = INT_MAX; s_box[1] = INT_MAX; s_box[2] = INT_MIN; s_box[3] = INT_MIN; }
        __syncthreads();                         // also: the previous depth's readers of `stage` are done
        Bilin bl;
        bl.any = false;
        if (inb) {
            float ix, iy;
            warp_coords(M, (float)x, (float)y, depths[d], W, H, &ix, &iy);
            bl = bilin_setup(ix, iy, W, H);
        }
        {
            const unsigned int act = __ballot_sync(0xffffffffu, bl.any);
            if (bl.any) {
                const int xmn = __reduce_min_sync(act, bl.x0), ymn = __reduce_min_sync(act, bl.y0);
                const int xmx = __reduce_max_sync(act, bl.x0), ymx = __reduce_max_sync(act, bl.y0);
                if ((tid & 31) == __ffs(act) - 1) {
                    atomicMin(&s_box[0], xmn); atomicMin(&s_box[1], ymn);
                    atomicMax(&s_box[2], xmx + 1); atomicMax(&s_box[3], ymx + 1);
                }
            }
        }
        __syncthreads();
        const bool some = s_box[2] >= s_box[0];                           // any thread of the tile samples inside the source map
        const int bx0 = some ? max(s_box[0], 0) : 0, by0 = some ? max(s_box[1], 0) : 0;
        const int bx1 = some ? min(s_box[2], W - 1) : -1, by1 = some ? min(s_box[3], H - 1) : -1;
        const int bw = bx1 - bx0 + 1, bh = by1 - by0 + 1;
        const bool staged = bw > 0 && bh > 0 && bw * bh <= VT_MAXPX;       // uniform over the block
        if (staged) {
            const int npx = bw * bh;
#pragma unroll
            for (int c = 0; c < 4; ++c)
                for (int p = tid; p < npx; p += 256) {
                    const int py = p / bw, px = p - py * bw;
                    stage[c][p] = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)(by0 + py) * W + bx0 + px) * C) + c);
                }
        }
        __syncthreads();
        const float wts[4] = {bl.w00, bl.w01, bl.w10, bl.w11};
        const int lane = tid & 31;
        uint4* ob = obuf[tid >> 5];
#pragma unroll
        for (int h = 0; h < 2; ++h) {            // two halves of 16 channels: keeps the kernel under 64 registers
            float v[16];
            vt_unpack8(ref[2 * h], v);
            vt_unpack8(ref[2 * h + 1], v + 8);
            if (bl.any) {
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (wts[k] != 0.f) {
                        const int yy = bl.y0 + (k >> 1), xx = bl.x0 + (k & 1);
#pragma unroll
                        for (int q = 0; q < 2; ++q) {
                            uint4 u;
                            if (staged) u = stage[2 * h + q][(yy - by0) * bw + (xx - bx0)];
                            else u = __ldg(reinterpret_cast<const uint4*>(src + ((size_t)yy * W + xx) * C) + 2 * h + q);
                            float s8[8];
                            vt_unpack8(u, s8);
#pragma unroll
                            for (int j = 0; j < 8; ++j) v[8 * q + j] = fmaf(wts[k], s8[j], v[8 * q + j]);
                        }
                    }
                }
            }
            // park the 16-bit result in the warp's transpose buffer (chunk index XOR-swizzled: conflict-free both ways)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                uint4 o;
                vt_pack8(v + 8 * q, f16, &o);
                ob[lane * 4 + ((2 * h + q) ^ ((lane >> 1) & 3))] = o;
            }
        }
        __syncwarp();
        // a per-thread 64-byte row makes every store instruction touch 32 separate lines (3.7 TB/s measured, against 6.9 TB/s
        // for warp-contiguous stores: tools/micro/store_pattern.cu); here instruction k writes 8 voxels x 64 B = 512 contiguous bytes
        // (a warp covers two tile rows of 16 voxels)
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int vx = 8 * k + (lane >> 2), c = lane & 3;           // voxel of the warp, chunk
            const uint4 o = ob[vx * 4 + (c ^ ((vx >> 1) & 3))];
            const int oy = tyi * VT_T + (tid >> 5) * 2 + (vx >> 4), ox = txi * VT_T + (vx & 15);
            if (oy < H && ox < W)
                *(reinterpret_cast<uint4*>(vol + ((((size_t)b * D + d) * H + oy) * W + ox) * C) + c) = o;
        }
        __syncwarp();
    }
}

int build_volume(const void* f_ref, const void* f_src, const float* Mw, const float* depths, bf16* vol, int B, int D, int H,
                 int W, int C, int f16, cudaStream_t stream) {
    ADP_CHECK_ARG(C == 32, "feature channels must be 32");
    if ((size_t)B * D * H * W == 0) return ADP_OK;
    const int tiles_x = cdiv(W, VT_T), tiles_y = cdiv(H, VT_T);
    build_volume_tile_kernel<<<dim3(tiles_x * tiles_y, cdiv(D, VT_DG), B), 256, 0, stream>>>(
        reinterpret_cast<const __half*>(f_ref), reinterpret_cast<const __half*>(f_src), Mw, depths, vol, D, H, W, tiles_x, f16);
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
