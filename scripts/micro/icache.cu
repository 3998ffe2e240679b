// Cold vs warm instruction fetch: a straight-line block of N FFMA (8 independent chains) executed `REPS` times by one warp per CTA.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o icache icache.cu && ./icache
#include <cstdio>
#include <cuda_runtime.h>
template <int N>
__global__ void k(float *out, long long *cyc, int reps, float a) {
    float x[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) x[i] = threadIdx.x + i;
#pragma unroll 1
    for (int r = 0; r < reps; ++r) {
        long long t0 = clock64();
#pragma unroll
        for (int i = 0; i < N; ++i) x[i & 7] = fmaf(x[i & 7], a, 1.0f + (float)(i & 1023));
        long long t1 = clock64();
        if (threadIdx.x == 0) cyc[blockIdx.x * 8 + r] = t1 - t0;
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int N>
void run(int ctas, int threads) {
    float *o; long long *c;
    cudaMalloc(&o, ctas * threads * 4); cudaMalloc(&c, ctas * 8 * 8);
    // evict: run a different big kernel first is not needed for the very first launch; flush L2 with a memset between
    void *big; cudaMalloc(&big, 512 << 20);
    for (int trial = 0; trial < 3; ++trial) {
        if (trial == 2) cudaMemset(big, trial, 512 << 20);  // L2 flushed before the third launch
        cudaMemset(c, 0, ctas * 8 * 8);
        k<N><<<ctas, threads>>>(o, c, 4, 1.0001f);
        cudaDeviceSynchronize();
        long long h[8];
        cudaMemcpy(h, c, 64, cudaMemcpyDeviceToHost);
        printf("N=%6d (%4d KB) ctas %3d thr %3d trial %d%s: rep cycles %lld %lld %lld %lld  -> cold extra %.0f cyc/128B line\n", N, N * 16 / 1024, ctas, threads, trial,
               trial == 2 ? " (L2 flushed)" : "", h[0], h[1], h[2], h[3], (double)(h[0] - h[3]) / (N * 16 / 128.0));
    }
    cudaFree(o); cudaFree(c); cudaFree(big);
}
int main() {
    run<512>(1, 32); run<2048>(1, 32); run<8192>(1, 32); run<16384>(1, 32);
    run<2048>(148, 128); run<8192>(148, 128);
    return 0;
}
