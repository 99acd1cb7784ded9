// Per-environment pose fit, branch A of ADA/interface_v5.py:318-321 (the one every shipped config takes):
//   back-projection (ADA/lib/utils.py:99-112), scale = exact median over all point pairs of
//   |dc| / |dn| with |dn| > 0.01 and |dc| < 0.3 (utils.py:76-96), t = mean(c) - mean(s R n) (utils.py:114-119),
//   box corners + world transform + finite check (interface_v5.py:354-374, utils.py:40-74).
// One CTA per environment; the 1024 points live in shared memory; the median is an exact radix select over the
// recomputed pair ratios (4 x 8-bit passes on the float bit pattern + one pass for the lower middle element).
#include "common.cuh"

namespace adp {

constexpr int FIT_THREADS = 1024;
constexpr int FIT_MAXP = 1024;

__device__ __forceinline__ float block_reduce_sum(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = 0.f;
    for (int i = 0; i < FIT_THREADS / 32; ++i) s += red[i];
    __syncthreads();
    return s;
}
__device__ __forceinline__ float block_reduce_max(float v, float* red) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float s = red[0];
    for (int i = 1; i < FIT_THREADS / 32; ++i) s = fmaxf(s, red[i]);
    __syncthreads();
    return s;
}

// ratio of pair (i, j), or a negative value when the pair is filtered out
__device__ __forceinline__ float pair_ratio(const float* cx, const float* cy, const float* cz, const float* nx, const float* ny,
                                            const float* nz, int i, int j) {
    const float dnx = nx[i] - nx[j], dny = ny[i] - ny[j], dnz = nz[i] - nz[j];
    const float dn = sqrtf(dnx * dnx + dny * dny + dnz * dnz);
    const float dcx = cx[i] - cx[j], dcy = cy[i] - cy[j], dcz = cz[i] - cz[j];
    const float dc = sqrtf(dcx * dcx + dcy * dcy + dcz * dcz);
    return (dn > 0.01f && dc < 0.3f) ? dc / dn : -1.f;
}

// warp-aggregated shared-memory histogram increment (most keys of a pass share a handful of bins)
__device__ __forceinline__ void hist_add(unsigned int* hist, unsigned int bin, bool active) {
    const unsigned int act = __ballot_sync(0xffffffffu, active);
    if (!active) return;
    const unsigned int peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&hist[bin], (unsigned int)__popc(peers));
}

constexpr int FIT_BINS = 2048;   // 11 + 11 + 10 bit radix passes over the positive-float bit pattern

__global__ void __launch_bounds__(FIT_THREADS)
fit_kernel(const float* __restrict__ nocs, const float* __restrict__ depth, const int* __restrict__ choose,
           const double* __restrict__ Kp, const float* __restrict__ R, const double* __restrict__ E, const uint8_t* __restrict__ valid,
           double* __restrict__ bbox, double* __restrict__ scale_out, double* __restrict__ trans_out, float* __restrict__ scratch,
           int P, int S) {
    __shared__ float cx[FIT_MAXP], cy[FIT_MAXP], cz[FIT_MAXP], nx[FIT_MAXP], ny[FIT_MAXP], nz[FIT_MAXP];
    __shared__ unsigned int hist[FIT_BINS];
    __shared__ float red[FIT_THREADS / 32];
    __shared__ unsigned long long s_cnt;
    __shared__ unsigned int s_sel[4];
    const int b = blockIdx.x, tid = threadIdx.x;
    double* out = bbox + (size_t)b * 24;
    const bool ok_in = !valid || valid[b];
    if (!ok_in) {
        if (tid < 24) out[tid] = 10.0 + (double)((tid / 3 >> (2 - tid % 3)) & 1);   // unit cube + 10 (interface_v5.py:232-241)
        if (tid == 0) { if (scale_out) scale_out[b] = nan(""); }
        return;
    }
    const double* k = Kp + 9 * b;
    const double fx = k[0], fy = k[4], pcx = k[2], pcy = k[5];
    for (int i = tid; i < P; i += FIT_THREADS) {
        const int pix = choose[(size_t)b * P + i];
        const double z = (double)depth[(size_t)b * P + i];
        const int y = pix / S, x = pix - y * S;
        cx[i] = (float)(((double)x - pcx) * z / fx);
        cy[i] = (float)(((double)y - pcy) * z / fy);
        cz[i] = (float)z;
        nx[i] = nocs[((size_t)b * P + i) * 3]; ny[i] = nocs[((size_t)b * P + i) * 3 + 1]; nz[i] = nocs[((size_t)b * P + i) * 3 + 2];
    }
    if (tid == 0) s_cnt = 0ull;
    __syncthreads();

    // unordered pairs (i < j): the reference's ordered-pair list holds every ratio twice, which leaves the median unchanged.
    // pass 0: evaluate every pair once, park the ratio (negative = filtered out) in the per-env scratch row, count the valid ones
    const size_t npairs = (size_t)P * (P - 1) / 2;
    float* rat = scratch + (size_t)b * npairs;
    {
        unsigned int c = 0;
        for (int i = 0; i < P - 1; ++i) {
            const size_t row0 = (size_t)i * (2 * P - i - 1) / 2;      // index of pair (i, i+1)
            for (int j = i + 1 + tid; j < P; j += FIT_THREADS) {
                const float r = pair_ratio(cx, cy, cz, nx, ny, nz, i, j);
                rat[row0 + (j - i - 1)] = r;
                c += r >= 0.f;
            }
        }
        for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
        if ((tid & 31) == 0) atomicAdd(&s_cnt, (unsigned long long)c);
    }
    __syncthreads();
    const long long m = (long long)s_cnt;
    double scale = nan("");
    if (m > 0) {
        // rank (0-based) of the upper middle element among the m sorted ratios; the lower one is rank (m-1)/2
        const long long k_hi = m / 2;
        const long long k_lo = (m - 1) / 2;
        unsigned int prefix = 0, pmask = 0;
        long long need = k_hi + 1;      // find the (k_hi+1)-th smallest
        const int shifts[3] = {21, 10, 0};
        const int widths[3] = {11, 11, 10};
        const size_t npad = (npairs + FIT_THREADS - 1) / FIT_THREADS * FIT_THREADS;
        for (int pass = 0; pass < 3; ++pass) {
            const int shift = shifts[pass];
            const unsigned int bmask = (1u << widths[pass]) - 1u;
            for (int i = tid; i < FIT_BINS; i += FIT_THREADS) hist[i] = 0;
            __syncthreads();
            for (size_t q = tid; q < npad; q += FIT_THREADS) {
                const float r = q < npairs ? rat[q] : -1.f;
                const unsigned int key = __float_as_uint(r);
                const bool act = (r >= 0.f) && ((key & pmask) == prefix);
                hist_add(hist, (key >> shift) & bmask, act);
            }
            __syncthreads();
            if (tid < 32) {
                // warp 0 scans the histogram: each lane sums a contiguous slice, then the slices are walked in order
                const int per = FIT_BINS / 32;
                long long part = 0;
                for (int q = 0; q < per; ++q) part += hist[tid * per + q];
                long long incl = part;
                for (int o = 1; o < 32; o <<= 1) {
                    const long long y = __shfl_up_sync(0xffffffffu, incl, o);
                    if (tid >= o) incl += y;
                }
                const long long excl = incl - part;
                const bool mine = (excl < need) && (need <= incl);
                if (mine) {
                    long long acc = excl;
                    int bsel = tid * per;
                    for (; bsel < tid * per + per; ++bsel) {
                        if (acc + (long long)hist[bsel] >= need) break;
                        acc += hist[bsel];
                    }
                    s_sel[0] = (unsigned int)bsel;
                    s_sel[1] = (unsigned int)(need - acc);
                }
            }
            __syncthreads();
            prefix |= s_sel[0] << shift;
            pmask |= bmask << shift;
            need = (long long)s_sel[1];
            __syncthreads();
        }
        const float v_hi = __uint_as_float(prefix);
        // lower middle: equals v_hi unless exactly k_hi elements are smaller than v_hi and k_lo < k_hi
        float v_lo = v_hi;
        if (k_lo < k_hi) {
            unsigned int less = 0;
            float mx = -1.f;
            for (size_t q = tid; q < npairs; q += FIT_THREADS) {
                const float r = rat[q];
                if (r >= 0.f && r < v_hi) { ++less; mx = fmaxf(mx, r); }
            }
            if (tid == 0) s_cnt = 0ull;
            __syncthreads();
            for (int o = 16; o > 0; o >>= 1) less += __shfl_xor_sync(0xffffffffu, less, o);
            if ((tid & 31) == 0) atomicAdd(&s_cnt, (unsigned long long)less);
            mx = block_reduce_max(mx, red);
            __syncthreads();
            if ((long long)s_cnt > k_lo) v_lo = mx;   // rank k_lo falls below v_hi -> it is the largest element < v_hi
        }
        scale = 0.5 * ((double)v_lo + (double)v_hi);
    }

    // ---- translation: mean(cam) - mean(s R nocs);  half extents: max |nocs|
    const float smx = block_reduce_sum([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s += cx[i]; return s; }(), red);
    const float smy = block_reduce_sum([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s += cy[i]; return s; }(), red);
    const float smz = block_reduce_sum([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s += cz[i]; return s; }(), red);
    const float snx = block_reduce_sum([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s += nx[i]; return s; }(), red);
    const float sny = block_reduce_sum([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s += ny[i]; return s; }(), red);
    const float snz = block_reduce_sum([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s += nz[i]; return s; }(), red);
    const float hx = block_reduce_max([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s = fmaxf(s, fabsf(nx[i])); return s; }(), red);
    const float hy = block_reduce_max([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s = fmaxf(s, fabsf(ny[i])); return s; }(), red);
    const float hz = block_reduce_max([&] { float s = 0.f; for (int i = tid; i < P; i += FIT_THREADS) s = fmaxf(s, fabsf(nz[i])); return s; }(), red);

    if (tid == 0) {
        const float* Rb = R + 9 * b;
        const double mc[3] = {(double)smx / P, (double)smy / P, (double)smz / P};
        const double mn[3] = {(double)snx / P, (double)sny / P, (double)snz / P};
        double t[3];
        for (int r = 0; r < 3; ++r)
            t[r] = mc[r] - scale * ((double)Rb[3 * r] * mn[0] + (double)Rb[3 * r + 1] * mn[1] + (double)Rb[3 * r + 2] * mn[2]);
        if (scale_out) scale_out[b] = scale;
        if (trans_out) for (int r = 0; r < 3; ++r) trans_out[3 * b + r] = t[r];
        const double half[3] = {(double)hx * scale, (double)hy * scale, (double)hz * scale};   // size / 2
        double inv[16];
        bool fin = invert4x4(E + 16 * b, inv);
        for (int i = 0; i < 16 && fin; ++i) fin = isfinite(inv[i]);
        double cam[8][3];
        for (int c = 0; c < 8 && fin; ++c) {
            // corner order of get_3d_bbox (utils.py:40-58): (+++, ++-, -++, -+-, +-+, +--, --+, ---)
            const double sx = (c & 2) ? -1.0 : 1.0, sy = (c & 4) ? -1.0 : 1.0, sz = (c & 1) ? -1.0 : 1.0;
            const double p[3] = {sx * half[0], sy * half[1], sz * half[2]};
            for (int r = 0; r < 3; ++r) {
                // the reference stores R and t in a float32 4x4 (interface_v5.py:359-361)
                cam[c][r] = (double)Rb[3 * r] * p[0] + (double)Rb[3 * r + 1] * p[1] + (double)Rb[3 * r + 2] * p[2] + (double)(float)t[r];
                fin = fin && isfinite(cam[c][r]);
            }
        }
        if (fin) {
            for (int c = 0; c < 8; ++c)
                for (int r = 0; r < 3; ++r)
                    out[3 * c + r] = inv[4 * r] * cam[c][0] + inv[4 * r + 1] * cam[c][1] + inv[4 * r + 2] * cam[c][2] + inv[4 * r + 3];
        } else {
            for (int i = 0; i < 24; ++i) out[i] = 10.0 + (double)(((i / 3) >> (2 - i % 3)) & 1);
        }
    }
}

int fit_run(const float* nocs, const float* depth, const int* choose, const double* Kp, const float* R, const double* E,
            const uint8_t* valid, double* bbox, double* scale_out, double* trans_out, float* scratch, int B, int P, int S,
            cudaStream_t stream) {
    ADP_CHECK_ARG(P <= FIT_MAXP, "at most 1024 points per env");
    ADP_CHECK_ARG(scratch != nullptr, "scratch of B * P*(P-1)/2 floats");
    if (B == 0) return ADP_OK;
    fit_kernel<<<B, FIT_THREADS, 0, stream>>>(nocs, depth, choose, Kp, R, E, valid, bbox, scale_out, trans_out, scratch, P, S);
    ADP_CUDA(cudaGetLastError());
    return ADP_OK;
}

}  // namespace adp
