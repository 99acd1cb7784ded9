// CUDA-core implicit-GEMM convolution: BM output positions x BN output channels per block, K loop over
// (tap, KCH-channel chunk) staged through shared memory, TM x TN register tile per thread.
#include "direct_conv.cuh"

namespace adp {

template <int BM, int BN, int TM, int TN, int KCH>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
direct_conv_kernel(const DirectConvParams p, int batch) {
    constexpr int NT = (BM / TM) * (BN / TN);
    __shared__ float As[KCH][BM + 1];
    __shared__ float Bs[KCH][BN];
    __shared__ int pos_b[BM], pos_d[BM], pos_h[BM], pos_w[BM];

    const int tid = threadIdx.x;
    const int tx = tid % (BN / TN);      // channel group
    const int ty = tid / (BN / TN);      // position group
    const long long Mtotal = (long long)batch * p.Do * p.Ho * p.Wo;
    const long long m0 = (long long)blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;

    for (int i = tid; i < BM; i += NT) {
        long long m = m0 + i;
        if (m < Mtotal) {
            int ow = (int)(m % p.Wo); long long t = m / p.Wo;
            int oh = (int)(t % p.Ho); t /= p.Ho;
            int od = (int)(t % p.Do);
            pos_b[i] = (int)(t / p.Do); pos_d[i] = od; pos_h[i] = oh; pos_w[i] = ow;
        } else {
            pos_b[i] = -1; pos_d[i] = 0; pos_h[i] = 0; pos_w[i] = 0;
        }
    }
    __syncthreads();

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    const int taps = p.kd * p.kh * p.kw;
    for (int tap = 0; tap < taps; ++tap) {
        const int kx = tap % p.kw, ky = (tap / p.kw) % p.kh, kz = tap / (p.kw * p.kh);
        for (int c0 = 0; c0 < p.Cin; c0 += KCH) {
            // ---- stage A: [KCH][BM] gathered input values
            for (int e = tid; e < BM * KCH; e += NT) {
                const int i = e / KCH, k = e % KCH;
                float v = 0.f;
                const int b = pos_b[i];
                const int c = c0 + k;
                if (b >= 0 && c < p.Cin) {
                    int iz, iy, ix;
                    bool ok = true;
                    if (!p.transposed) {
                        iz = pos_d[i] * p.sd - p.pd + kz;
                        iy = pos_h[i] * p.sh - p.ph + ky * p.dil;
                        ix = pos_w[i] * p.sw - p.pw + kx * p.dil;
                    } else {
                        iz = pos_d[i] + p.pd - kz; iy = pos_h[i] + p.ph - ky; ix = pos_w[i] + p.pw - kx;
                        ok = (iz % p.sd == 0) && (iy % p.sh == 0) && (ix % p.sw == 0) && iz >= 0 && iy >= 0 && ix >= 0;
                        iz /= p.sd; iy /= p.sh; ix /= p.sw;
                    }
                    if (ok && iz >= 0 && iz < p.Di && iy >= 0 && iy < p.Hi && ix >= 0 && ix < p.Wi) {
                        const size_t idx = ((((size_t)b * p.Di + iz) * p.Hi + iy) * p.Wi + ix) * p.Cin + c;
                        v = p.in_f32 ? p.in_f32[idx] : ld_act16(p.in_hi, p.in_lo, idx, p.f16);
                    }
                }
                As[k][i] = v;
            }
            // ---- stage B: [KCH][BN] weights
            for (int e = tid; e < KCH * BN; e += NT) {
                const int k = e / BN, n = e % BN;
                const int c = c0 + k;
                Bs[k][n] = (c < p.Cin && n0 + n < p.Cout) ? p.w[((size_t)tap * p.Cin + c) * p.Cout + n0 + n] : 0.f;
            }
            __syncthreads();
#pragma unroll
            for (int k = 0; k < KCH; ++k) {
                float a[TM], bb[TN];
#pragma unroll
                for (int i = 0; i < TM; ++i) a[i] = As[k][ty * TM + i];
#pragma unroll
                for (int j = 0; j < TN; ++j) bb[j] = Bs[k][tx * TN + j];
#pragma unroll
                for (int i = 0; i < TM; ++i)
#pragma unroll
                    for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
            }
            __syncthreads();
        }
    }

    // ---- epilogue
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int r = ty * TM + i;
        if (pos_b[r] < 0) continue;
        const size_t pix = (size_t)(m0 + r);
#pragma unroll
        for (int j = 0; j < TN; ++j) {
            const int n = n0 + tx * TN + j;
            if (n >= p.Cout) continue;
            float v = acc[i][j];
            const size_t o = pix * p.Cout + n;
            const size_t ro = pix * (p.res_cs ? p.res_cs : p.Cout) + n;
            if (p.scale) v *= p.scale[n];
            if (p.bias) v += p.bias[n];
            if (p.res_hi && !p.res_after_act) v += ld_act16(p.res_hi, p.res_lo, ro, p.f16);
            if (p.act == 1) v = fmaxf(v, 0.f);
            else if (p.act == 2) v = v > 0.f ? v : v * p.prelu;
            if (p.res_hi && p.res_after_act) v += ld_act16(p.res_hi, p.res_lo, ro, p.f16);
            if (p.out_f32) p.out_f32[o] = v;
            if (p.out_hi) st_act16(p.out_hi, p.out_lo, o, v, p.f16);
        }
    }
}

template <int BM, int BN, int TM, int TN, int KCH>
static int launch_cfg(const DirectConvParams& p, int batch, cudaStream_t stream) {
    long long Mtotal = (long long)batch * p.Do * p.Ho * p.Wo;
    if (Mtotal == 0) return ADP_OK;
    dim3 grid((unsigned)((Mtotal + BM - 1) / BM), (unsigned)cdiv(p.Cout, BN));
    direct_conv_kernel<BM, BN, TM, TN, KCH><<<grid, (BM / TM) * (BN / TN), 0, stream>>>(p, batch);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

int direct_conv_launch(const DirectConvParams& p, int batch, cudaStream_t stream) {
    ADP_CHECK_ARG(p.w != nullptr, "weights");
    ADP_CHECK_ARG((p.in_hi != nullptr) != (p.in_f32 != nullptr), "exactly one input kind");
    ADP_CHECK_ARG(p.out_hi || p.out_f32, "no output");
    if (p.Cout >= 64) return launch_cfg<64, 64, 4, 4, 16>(p, batch, stream);
    if (p.Cout >= 32) return launch_cfg<128, 32, 4, 4, 16>(p, batch, stream);
    if (p.Cout >= 16) return launch_cfg<128, 16, 4, 2, 16>(p, batch, stream);
    return launch_cfg<256, 8, 4, 2, 16>(p, batch, stream);
}

}  // namespace adp
