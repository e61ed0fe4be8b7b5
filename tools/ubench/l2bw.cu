// L2 vs HBM bandwidth on B200: stream-read (and read+write) a buffer of varying size repeatedly
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void rd(const uint4* __restrict__ p, size_t n, uint4* out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    uint4 acc = make_uint4(0, 0, 0, 0);
    for (; i < n; i += st) { uint4 v = __ldcg(p + i); acc.x ^= v.x; acc.y ^= v.y; acc.z ^= v.z; acc.w ^= v.w; }
    if (acc.x == 0x12345678) out[0] = acc;
}
__global__ void rmw(uint4* __restrict__ p, size_t n) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    for (; i < n; i += st) { uint4 v = __ldcg(p + i); v.x += 1; __stcg(p + i, v); }
}
int main() {
    size_t maxb = 1ull << 30;
    uint4 *buf, *out; cudaMalloc(&buf, maxb); cudaMalloc(&out, 64); cudaMemset(buf, 1, maxb);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    size_t sizes[] = {8ull << 20, 16ull << 20, 32ull << 20, 48ull << 20, 64ull << 20, 96ull << 20, 128ull << 20, 256ull << 20, 1ull << 30};
    for (size_t sz : sizes) {
        size_t n = sz / 16; int reps = (int)((8ull << 30) / sz); if (reps < 4) reps = 4;
        for (int mode = 0; mode < 2; mode++) {
            for (int w = 0; w < 2; w++) { if (mode == 0) rd<<<148 * 8, 512>>>(buf, n, out); else rmw<<<148 * 8, 512>>>(buf, n); }
            cudaEventRecord(a);
            for (int r = 0; r < reps; r++) { if (mode == 0) rd<<<148 * 8, 512>>>(buf, n, out); else rmw<<<148 * 8, 512>>>(buf, n); }
            cudaEventRecord(b); cudaEventSynchronize(b);
            float ms; cudaEventElapsedTime(&ms, a, b);
            double gbs = (double)sz * reps * (mode ? 2 : 1) / (ms * 1e-3) / 1e9;
            printf("%6zu MB %s: %8.1f GB/s\n", sz >> 20, mode ? "read+write" : "read      ", gbs);
        }
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
