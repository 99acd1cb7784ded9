// Memory-bound helpers of the PSPNet backbone (channels-last, hi(+lo) bf16 activations):
// 3x3/2 max-pool, pyramid pooling + 1x1 convs, pyramid concat fused with the first x2 bilinear upsample,
// plain x2 bilinear upsample (align_corners=True).  Reference: ADA/lib/pspnet.py:66-107,142-158.
#include "common.cuh"

namespace adp {

// ---------------------------------------------------------------------------------------------
// max-pool 3x3 stride 2 pad 1 (pspnet.py:39,69)
// ---------------------------------------------------------------------------------------------
__global__ void maxpool3x3s2_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, bf16* __restrict__ out_hi,
                                    bf16* __restrict__ out_lo, int B, int Hi, int Wi, int Ho, int Wo, int C) {
    const size_t total = (size_t)B * Ho * Wo * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        size_t t = i / C;
        const int ox = (int)(t % Wo); t /= Wo;
        const int oy = (int)(t % Ho);
        const int b = (int)(t / Ho);
        float m = -INFINITY;
#pragma unroll
        for (int ky = 0; ky < 3; ++ky) {
            const int iy = oy * 2 - 1 + ky;
            if (iy < 0 || iy >= Hi) continue;
#pragma unroll
            for (int kx = 0; kx < 3; ++kx) {
                const int ix = ox * 2 - 1 + kx;
                if (ix < 0 || ix >= Wi) continue;
                m = fmaxf(m, ld_act(in_hi, in_lo, (((size_t)b * Hi + iy) * Wi + ix) * C + c));
            }
        }
        st_act(out_hi, out_lo, i, m);
    }
}

int maxpool3x3s2(const Act& in, const Act& out, int batch, cudaStream_t stream) {
    ADP_CHECK_ARG(in.C == out.C && out.H == (in.H + 1) / 2 && out.W == (in.W + 1) / 2, "maxpool shapes");
    size_t total = (size_t)batch * out.H * out.W * out.C;
    int grid = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
    if (grid == 0) return ADP_OK;
    maxpool3x3s2_kernel<<<grid, 256, 0, stream>>>(in.hi, in.lo, out.hi, out.lo, batch, in.H, in.W, out.H, out.W, in.C);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

// ---------------------------------------------------------------------------------------------
// pyramid pooling: adaptive average pools with bins (1,2,3,6) = 50 cells (pspnet.py:84-87)
// window of cell i along an axis of size S with b bins: [floor(i S / b), ceil((i+1) S / b))
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void psp_cell(int cell, int* bins, int* cy, int* cx) {
    // cells are ordered stage by stage: 1 + 4 + 9 + 36
    if (cell < 1) { *bins = 1; *cy = 0; *cx = 0; }
    else if (cell < 5) { *bins = 2; *cy = (cell - 1) / 2; *cx = (cell - 1) % 2; }
    else if (cell < 14) { *bins = 3; *cy = (cell - 5) / 3; *cx = (cell - 5) % 3; }
    else { *bins = 6; *cy = (cell - 14) / 6; *cx = (cell - 14) % 6; }
}

__global__ void psp_pool_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, float* __restrict__ pooled,
                                int H, int W, int C) {
    const int b = blockIdx.y, cell = blockIdx.x;
    int bins, cy, cx;
    psp_cell(cell, &bins, &cy, &cx);
    const int y0 = (cy * H) / bins, y1 = ((cy + 1) * H + bins - 1) / bins;
    const int x0 = (cx * W) / bins, x1 = ((cx + 1) * W + bins - 1) / bins;
    const float inv = 1.f / (float)((y1 - y0) * (x1 - x0));
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int y = y0; y < y1; ++y)
            for (int x = x0; x < x1; ++x) s += ld_act(in_hi, in_lo, (((size_t)b * H + y) * W + x) * C + c);
        pooled[((size_t)b * 50 + cell) * C + c] = s * inv;
    }
}

// 1x1 conv 512 -> 128 (no bias) + ReLU per stage; w: [4][Cin][128] fp32 (pspnet.py:88-90)
__global__ void psp_conv_kernel(const float* __restrict__ pooled, const float* __restrict__ w, float* __restrict__ priors,
                                int Cin) {
    extern __shared__ float sv[];
    const int b = blockIdx.y, cell = blockIdx.x;
    const int stage = cell < 1 ? 0 : cell < 5 ? 1 : cell < 14 ? 2 : 3;
    for (int c = threadIdx.x; c < Cin; c += blockDim.x) sv[c] = pooled[((size_t)b * 50 + cell) * Cin + c];
    __syncthreads();
    const int n = threadIdx.x;   // 128 threads
    const float* ws = w + (size_t)stage * Cin * 128;
    float acc = 0.f;
    for (int c = 0; c < Cin; ++c) acc = fmaf(sv[c], ws[(size_t)c * 128 + n], acc);
    priors[((size_t)b * 50 + cell) * 128 + n] = fmaxf(acc, 0.f);
}

int psp_priors(const Act& feat, const float* w, float* pooled, float* priors, int batch, cudaStream_t stream) {
    ADP_CHECK_ARG(feat.C <= 1024, "psp channels");
    if (batch == 0) return ADP_OK;
    psp_pool_kernel<<<dim3(50, batch), 256, 0, stream>>>(feat.hi, feat.lo, pooled, feat.H, feat.W, feat.C);
    ADP_CUDA(cudaGetLastError());
    psp_conv_kernel<<<dim3(50, batch), 128, feat.C * sizeof(float), stream>>>(pooled, w, priors, feat.C);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

// ---------------------------------------------------------------------------------------------
// concat(feat, upsampled priors) followed by the x2 bilinear upsample of up_1, in one pass:
// out[b, Y, X, :] for the (2H x 2W) grid, channels [0,C) from feat, [C + 128 s, C + 128 (s+1)) from stage s.
// Both interpolations use align_corners=True (pspnet.py:92-93,105).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lin_coord(int o, int in_size, int out_size, int* i0, int* i1, float* w1) {
    // align_corners=True: src = o * (in-1)/(out-1)
    const float scale = out_size > 1 ? (float)(in_size - 1) / (float)(out_size - 1) : 0.f;
    const float s = scale * (float)o;
    int a = (int)s;
    if (a > in_size - 1) a = in_size - 1;
    *i0 = a;
    *i1 = a + (a < in_size - 1 ? 1 : 0);
    *w1 = s - (float)a;
}

__device__ __forceinline__ float prior_at(const float* __restrict__ pr, int bins, int y, int x, int H, int W, int n) {
    // value of the stage's b x b prior map, bilinearly resized (align_corners=True) to H x W, at (y, x)
    int y0, y1, x0, x1;
    float wy, wx;
    lin_coord(y, bins, H, &y0, &y1, &wy);
    lin_coord(x, bins, W, &x0, &x1, &wx);
    const float v00 = pr[(y0 * bins + x0) * 128 + n], v01 = pr[(y0 * bins + x1) * 128 + n];
    const float v10 = pr[(y1 * bins + x0) * 128 + n], v11 = pr[(y1 * bins + x1) * 128 + n];
    return (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
}

__global__ void psp_concat_up_kernel(const bf16* __restrict__ f_hi, const bf16* __restrict__ f_lo,
                                     const float* __restrict__ priors, bf16* __restrict__ out_hi, bf16* __restrict__ out_lo,
                                     int B, int H, int W, int C) {
    const int Ho = 2 * H, Wo = 2 * W, Ct = C + 512;
    const size_t total = (size_t)B * Ho * Wo * Ct;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % Ct);
        size_t t = i / Ct;
        const int X = (int)(t % Wo); t /= Wo;
        const int Y = (int)(t % Ho);
        const int b = (int)(t / Ho);
        int y0, y1, x0, x1;
        float wy, wx;
        lin_coord(Y, H, Ho, &y0, &y1, &wy);
        lin_coord(X, W, Wo, &x0, &x1, &wx);
        float v00, v01, v10, v11;
        if (c < C) {
            const size_t base = (size_t)b * H * W;
            v00 = ld_act(f_hi, f_lo, (base + (size_t)y0 * W + x0) * C + c);
            v01 = ld_act(f_hi, f_lo, (base + (size_t)y0 * W + x1) * C + c);
            v10 = ld_act(f_hi, f_lo, (base + (size_t)y1 * W + x0) * C + c);
            v11 = ld_act(f_hi, f_lo, (base + (size_t)y1 * W + x1) * C + c);
        } else {
            const int s = (c - C) / 128, n = (c - C) % 128;
            const int bins = s == 0 ? 1 : s == 1 ? 2 : s == 2 ? 3 : 6;
            const int off = s == 0 ? 0 : s == 1 ? 1 : s == 2 ? 5 : 14;
            const float* pr = priors + ((size_t)b * 50 + off) * 128;
            v00 = prior_at(pr, bins, y0, x0, H, W, n);
            v01 = prior_at(pr, bins, y0, x1, H, W, n);
            v10 = prior_at(pr, bins, y1, x0, H, W, n);
            v11 = prior_at(pr, bins, y1, x1, H, W, n);
        }
        const float v = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
        st_act(out_hi, out_lo, i, v);
    }
}

int psp_concat_up(const Act& feat, const float* priors, const Act& out, int batch, cudaStream_t stream) {
    ADP_CHECK_ARG(out.C == feat.C + 512 && out.H == 2 * feat.H && out.W == 2 * feat.W, "psp concat shapes");
    size_t total = (size_t)batch * out.H * out.W * out.C;
    if (total == 0) return ADP_OK;
    int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    psp_concat_up_kernel<<<grid, 256, 0, stream>>>(feat.hi, feat.lo, priors, out.hi, out.lo, batch, feat.H, feat.W, feat.C);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

// ---------------------------------------------------------------------------------------------
// x2 bilinear upsample, align_corners=True (pspnet.py:105)
// ---------------------------------------------------------------------------------------------
__global__ void upsample2x_kernel(const bf16* __restrict__ in_hi, const bf16* __restrict__ in_lo, bf16* __restrict__ out_hi,
                                  bf16* __restrict__ out_lo, int B, int H, int W, int C) {
    const int Ho = 2 * H, Wo = 2 * W;
    const size_t total = (size_t)B * Ho * Wo * C;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int c = (int)(i % C);
        size_t t = i / C;
        const int X = (int)(t % Wo); t /= Wo;
        const int Y = (int)(t % Ho);
        const int b = (int)(t / Ho);
        int y0, y1, x0, x1;
        float wy, wx;
        lin_coord(Y, H, Ho, &y0, &y1, &wy);
        lin_coord(X, W, Wo, &x0, &x1, &wx);
        const size_t base = (size_t)b * H * W;
        const float v00 = ld_act(in_hi, in_lo, (base + (size_t)y0 * W + x0) * C + c);
        const float v01 = ld_act(in_hi, in_lo, (base + (size_t)y0 * W + x1) * C + c);
        const float v10 = ld_act(in_hi, in_lo, (base + (size_t)y1 * W + x0) * C + c);
        const float v11 = ld_act(in_hi, in_lo, (base + (size_t)y1 * W + x1) * C + c);
        st_act(out_hi, out_lo, i, (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11));
    }
}

int upsample2x(const Act& in, const Act& out, int batch, cudaStream_t stream) {
    ADP_CHECK_ARG(out.C == in.C && out.H == 2 * in.H && out.W == 2 * in.W, "upsample shapes");
    size_t total = (size_t)batch * out.H * out.W * out.C;
    if (total == 0) return ADP_OK;
    int grid = (int)((total + 255) / 256 < 148 * 32 ? (total + 255) / 256 : 148 * 32);
    upsample2x_kernel<<<grid, 256, 0, stream>>>(in.hi, in.lo, out.hi, out.lo, batch, in.H, in.W, in.C);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

}  // namespace adp
