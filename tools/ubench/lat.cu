// micro-benchmarks: dependent-chain latency of the instructions on the SGM critical path (sm_100a)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
__global__ void k_credux(unsigned* out, long long* t, unsigned seed) {
    unsigned v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __reduce_min_sync(0xffffffffu, v + threadIdx.x) + 1;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) t[0] = t1 - t0;
}
__global__ void k_shfl(unsigned* out, long long* t, unsigned seed) {
    unsigned v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __shfl_up_sync(0xffffffffu, v, 1) + 1;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) t[1] = t1 - t0;
}
__global__ void k_viaddmin(unsigned* out, long long* t, unsigned seed) {
    unsigned v = seed + threadIdx.x, p = seed * 3;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __viaddmin_u16x2(v, p, 0x7fff7fffu);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) t[2] = t1 - t0;
}
__global__ void k_vadd2(unsigned* out, long long* t, unsigned seed) {
    unsigned v = seed + threadIdx.x, p = seed * 3;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __vadd2(v, p);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) t[3] = t1 - t0;
}
// shuffle-xor butterfly min (5 levels)
__global__ void k_bfly(unsigned* out, long long* t, unsigned seed) {
    unsigned v = seed + threadIdx.x;
    long long t0 = clock64();
    for (int i = 0; i < N; i++) {
#pragma unroll
        for (int o = 16; o; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
        v += threadIdx.x + 1;
    }
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) t[4] = t1 - t0;
}
// REDUX via match-free vector variant: __reduce_min_sync on signed (same instr?) 
__global__ void k_credux_s(int* out, long long* t, int seed) {
    int v = seed + threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) v = __reduce_min_sync(0xffffffffu, v ^ (int)threadIdx.x) + 1;
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) t[5] = t1 - t0;
}
int main() {
    unsigned* o; long long* t;
    cudaMalloc(&o, 4096); cudaMallocManaged(&t, 64);
    for (int rep = 0; rep < 2; rep++) {
        k_credux<<<1, 32>>>(o, t, 5); k_shfl<<<1, 32>>>(o, t, 5); k_viaddmin<<<1, 32>>>(o, t, 5); k_vadd2<<<1, 32>>>(o, t, 5);
        k_bfly<<<1, 32>>>(o, t, 5); k_credux_s<<<1, 32>>>((int*)o, t, 5);
        cudaDeviceSynchronize();
    }
    const char* names[] = {"credux(min.u32)+iadd", "shfl.up+iadd", "viaddmin.u16x2", "vadd2", "butterfly-min(5 shfl)+iadd", "credux(min.s32)+xor+iadd"};
    for (int i = 0; i < 6; i++) printf("%-28s %.1f cycles/iter\n", names[i], (double)t[i] / N);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
