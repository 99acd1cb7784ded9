// Micro-benchmark: per-thread-contiguous 64-byte stores (4 x STG.128 at 64 B lane stride: every instruction writes half
// sectors) against warp-contiguous stores (every instruction writes 512 contiguous bytes).  nvcc -arch=sm_100a -O3.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void strided(uint4* out, size_t nvox) {
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvox; v += (size_t)gridDim.x * blockDim.x) {
        uint4 val = make_uint4((unsigned)v, 1, 2, 3);
#pragma unroll
        for (int q = 0; q < 4; ++q) out[v * 4 + q] = val;
    }
}
__global__ void coalesced(uint4* out, size_t nvox) {
    const int lane = threadIdx.x & 31;
    for (size_t v0 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x - lane); v0 < nvox; v0 += (size_t)gridDim.x * blockDim.x) {
        uint4 val = make_uint4((unsigned)v0, 1, 2, 3);
#pragma unroll
        for (int k = 0; k < 4; ++k) out[(v0 + 8 * k) * 4 + lane] = val;      // 32 lanes x 16 B contiguous
    }
}
__global__ void strided128(uint4* out, size_t npix) {    // 128 B per thread (a 64-channel fp16 pixel row per lane)
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < npix; v += (size_t)gridDim.x * blockDim.x) {
        uint4 val = make_uint4((unsigned)v, 1, 2, 3);
#pragma unroll
        for (int q = 0; q < 8; ++q) out[v * 8 + q] = val;
    }
}
int main() {
    const size_t bytes = (size_t)4 << 30, nvox = bytes / 64;
    uint4* buf;
    cudaMalloc(&buf, bytes);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int which = 0; which < 3; ++which) {
        for (int it = 0; it < 3; ++it) {
            cudaEventRecord(a);
            if (which == 0) strided<<<148 * 16, 256>>>(buf, nvox);
            else if (which == 1) coalesced<<<148 * 16, 256>>>(buf, nvox);
            else strided128<<<148 * 16, 256>>>(buf, nvox / 2);
            cudaEventRecord(b);
            cudaEventSynchronize(b);
            float ms;
            cudaEventElapsedTime(&ms, a, b);
            if (it == 2) printf("%s: %.3f ms  %.1f GB/s\n", which == 0 ? "64B-per-thread" : which == 1 ? "warp-contiguous" : "128B-per-thread", ms, bytes / ms / 1e6);
        }
    }
    printf("err %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
