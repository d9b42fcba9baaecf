// FP64 latency / throughput probe (tuning aid): nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_probe fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <int CH>
__global__ void k(double* out, long long* cyc, int iters, double a, double b) {
    double x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fma(x[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int CH>
__global__ void kf(float* out, long long* cyc, int iters, float a, float b) {
    float x[CH];
#pragma unroll
    for (int i = 0; i < CH; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CH; ++i) x[i] = fmaf(x[i], b, a);
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* o; long long* c; cudaMalloc(&o, 1 << 24); cudaMalloc(&c, 8);
    const int iters = 4096;
    long long h;
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    printf("%s, %d SMs\n", p.name, p.multiProcessorCount);
#define RUN(K, CH, blocks, threads, label) K<CH><<<blocks, threads>>>((decltype(K<CH>)*)nullptr == nullptr ? o : o, c, iters, 1.0, 0.999999); 
    k<1><<<1, 32>>>(o, c, iters, 1.0, 0.999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA dependent chain, 1 warp: %.1f cycles per DFMA\n", (double)h / iters);
    k<8><<<1, 32>>>(o, c, iters, 1.0, 0.999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 independent chains, 1 warp: %.2f cycles per warp-DFMA\n", (double)h / iters / 8);
    k<8><<<1, 128>>>(o, c, iters, 1.0, 0.999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 chains x 4 warps (one per SMSP): %.2f cycles per warp-DFMA per SMSP\n", (double)h / iters / 8);
    k<8><<<1, 256>>>(o, c, iters, 1.0, 0.999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 chains x 8 warps: %.2f cycles per warp-DFMA per SMSP\n", (double)h / iters / 16);
    k<8><<<1, 1024>>>(o, c, iters, 1.0, 0.999999); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DFMA 8 chains x 32 warps: %.2f cycles per warp-DFMA per SMSP  -> %.1f lanes/clk/SM\n", (double)h / iters / 64, 32.0 * 4 / ((double)h / iters / 64));
    kf<1><<<1, 32>>>((float*)o, c, iters, 1.0f, 0.999999f); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("FFMA dependent chain, 1 warp: %.1f cycles per FFMA\n", (double)h / iters);
    kf<8><<<1, 1024>>>((float*)o, c, iters, 1.0f, 0.999999f); cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("FFMA 8 chains x 32 warps: %.2f cycles per warp-FFMA per SMSP\n", (double)h / iters / 64);
    return 0;
}
