// Issue rate of scalar FFMA by operand form on sm_100a: (a) two distinct register sources (x = x*m + m), (b) three distinct
// registers per instruction (x = x*a_i + b_i), (c) an immediate multiplier, (d) mixed FFMA + FMNMX / PRMT streams.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/_dbg/ffma_rt_probe scripts/ffma_rt_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
constexpr int CH = 8;
template <int MODE>
__global__ void probe(float *out, long long *cyc, int iters, float seed) {
    float s[CH], a[CH], b[CH];
    for (int i = 0; i < CH; ++i) { s[i] = seed + i + threadIdx.x; a[i] = 0.999f + 1e-4f * i + 1e-6f * threadIdx.x; b[i] = seed * 0.5f + 1e-3f * i + 1e-7f * threadIdx.x; }
    unsigned u[CH];
    for (int i = 0; i < CH; ++i) u[i] = threadIdx.x * 2654435761u + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (MODE == 0) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(s[i]) : "f"(a[0]));
                if (MODE == 1) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a[i]), "f"(b[i]));
                if (MODE == 2) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, %1;" : "+f"(s[i]) : "f"(b[i]));
                if (MODE == 3) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a[i]), "f"(b[i]));
                                 asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[(i + 1) % CH])); }
                if (MODE == 4) { asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(s[i]) : "f"(a[i]), "f"(b[i]));
                                 asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(u[i]) : "r"(u[(i + 1) % CH])); }
                if (MODE == 5) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[i]) : "f"(b[(i + 1) % CH]));
                if (MODE == 6) asm volatile("prmt.b32 %0, %0, %1, 0x4341;" : "+r"(u[i]) : "r"(u[(i + 1) % CH]));
            }
        }
    }
    long long t1 = clock64();
    float acc = 0.f;
    for (int i = 0; i < CH; ++i) acc += s[i] + a[i] + b[i] + __uint_as_float(u[i] & 0x3fffffffu);
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}
template <int MODE>
void run(int warps, const char *name, int per_iter) {
    float *out; long long *cyc, h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    probe<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
    probe<MODE><<<148, warps * 32>>>(out, cyc, iters, 1.0f);
    cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 8 * CH * per_iter * warps;       // warp-instructions per SM
    printf("%-44s warps/SM %2d: %.3f warp-instr / clk / SMSP\n", name, warps, instr / (double)h / 4.0);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {4, 8, 16}) {
        run<0>(w, "FFMA x = x*m + m (2 distinct regs)", 1);
        run<1>(w, "FFMA x = x*a_i + b_i (3 distinct regs)", 1);
        run<2>(w, "FFMA x = x*imm + b_i", 1);
        run<5>(w, "FMNMX", 1);
        run<6>(w, "PRMT", 1);
        run<3>(w, "FFMA(3 regs) + FMNMX interleaved", 2);
        run<4>(w, "FFMA(3 regs) + PRMT interleaved", 2);
    }
    return 0;
}
