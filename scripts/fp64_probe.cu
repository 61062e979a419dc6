// FP64 vs FP32 FMA throughput of the device (is double precision full-rate, half-rate or vestigial on this part?)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_dbg/fp64_probe scripts/fp64_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
template <typename T>
__global__ void fma_chain(T *out, int iters) {
    T a[8];
    for (int i = 0; i < 8; ++i) a[i] = (T)(threadIdx.x + i) * (T)1e-3;
    const T b = (T)1.000001, c = (T)1e-7;
    for (int it = 0; it < iters; ++it)
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = a[i] * b + c;
    T s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename T>
double run(const char *name) {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, threads = 256, iters = 4096;
    T *out;
    cudaMalloc(&out, sizeof(T) * blocks * threads);
    fma_chain<T><<<blocks, threads>>>(out, 16);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    fma_chain<T><<<blocks, threads>>>(out, iters);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)blocks * threads * iters * 8;
    printf("%s: %.2f TFLOP/s (FMA = 2 flop), %.1f FMA/clk/SM at 1.9 GHz\n", name, 2 * fma / (ms * 1e-3) / 1e12,
           fma / (ms * 1e-3) / sms / 1.9e9);
    cudaFree(out);
    return ms;
}
int main() {
    run<float>("fp32");
    run<double>("fp64");
    return 0;
}
