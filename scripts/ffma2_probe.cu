// Latency / throughput of the packed fp32 instructions (FFMA2) against scalar FFMA on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_dbg/ffma2_probe scripts/ffma2_probe.cu && scripts/_dbg/ffma2_probe
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ float fma1(float a, float b, float c) { float d; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c)); return d; }

template <int CHAINS, bool PACKED>
__global__ void probe(float *out, long long *cyc, int iters, float seed) {
    u64 p[CHAINS]; float s[2 * CHAINS];
    for (int i = 0; i < CHAINS; ++i) { float a = seed + i + threadIdx.x; p[i] = ((u64)__float_as_uint(a) << 32) | __float_as_uint(a + 1.f); s[2 * i] = a; s[2 * i + 1] = a + 1.f; }
    const float m = 0.999f; const u64 mm = ((u64)__float_as_uint(m) << 32) | __float_as_uint(m);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            if (PACKED) {
#pragma unroll
                for (int i = 0; i < CHAINS; ++i) p[i] = fma2(p[i], mm, mm);
            } else {
#pragma unroll
                for (int i = 0; i < 2 * CHAINS; ++i) s[i] = fma1(s[i], m, m);
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < CHAINS; ++i) { acc += __uint_as_float((unsigned)p[i]) + __uint_as_float((unsigned)(p[i] >> 32)) + s[2 * i] + s[2 * i + 1]; }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS, bool PACKED>
void run(int warps_per_sm, const char *name) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    probe<CHAINS, PACKED><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0f);
    probe<CHAINS, PACKED><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    // complex-pairs of fp32 FMA results produced per SM per cycle
    const double fmas = (double)iters * 8 * CHAINS * 2 * warps_per_sm * 32;
    printf("%-8s chains %d warps/SM %2d: %8.1f cycles per unrolled step (8 x %d instr), %.1f fp32 FMA lanes/clk/SM\n", name, CHAINS, warps_per_sm,
           (double)h / iters, PACKED ? CHAINS : 2 * CHAINS, fmas / (double)h);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<1, true>(1, "FFMA2"); run<1, false>(1, "FFMA");      // dependent-chain latency: cycles / 8 (packed) or per 2 indep chains
    run<1, true>(4, "FFMA2"); run<1, false>(4, "FFMA");
    run<4, true>(4, "FFMA2"); run<4, false>(4, "FFMA");
    run<4, true>(16, "FFMA2"); run<4, false>(16, "FFMA");
    run<8, true>(16, "FFMA2"); run<8, false>(16, "FFMA");
    run<8, true>(32, "FFMA2"); run<8, false>(32, "FFMA");
    return 0;
}
