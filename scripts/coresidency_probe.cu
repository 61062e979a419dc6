// co-residency probe: does a small kernel B share SMs with a big persistent kernel A?
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
__device__ unsigned int started, stop_flag;
__device__ unsigned int sm_hits[256];
__device__ __forceinline__ unsigned long long gt() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
template <int TMEM>
__global__ void __maxnreg__(96) kernA(unsigned long long ns, int use_bar) {
    extern __shared__ unsigned char sm[];
    __shared__ unsigned int tm;
    if (TMEM) {
        if (threadIdx.x < 32) {
            unsigned a = (unsigned)__cvta_generic_to_shared(&tm);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(a));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) atomicAdd(&started, 1u);
    sm[threadIdx.x] = 1;
    if (use_bar) { if (threadIdx.x < 128) asm volatile("bar.sync 1, 128;" ::: "memory"); else if (threadIdx.x < 256) asm volatile("bar.sync 2, 128;" ::: "memory"); else if (threadIdx.x < 384) asm volatile("bar.sync 3, 128;" ::: "memory"); }
    unsigned long long t0 = gt();
    while (gt() - t0 < ns) __nanosleep(200);
    __syncthreads();
    if (TMEM && threadIdx.x < 32) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm));
    if (blockIdx.x == 0 && threadIdx.x == 0) stop_flag = 1;
}
template <int SMEM>
__global__ void __launch_bounds__(256, 8) kernB() {
    __shared__ int s[SMEM / 4 > 0 ? SMEM / 4 : 1];
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    s[threadIdx.x % (SMEM / 4 > 0 ? SMEM / 4 : 1)] = smid;
    __syncthreads();
    if (threadIdx.x == 0 && *(volatile unsigned int *)&stop_flag == 0) atomicAdd(&sm_hits[smid], 1u);
    unsigned long long t0 = gt();
    while (gt() - t0 < 5000) __nanosleep(100);
}
int main(int argc, char **argv) {
    int smemA = argc > 1 ? atoi(argv[1]) : 199344, tmem = argc > 2 ? atoi(argv[2]) : 0, carveB = argc > 3 ? atoi(argv[3]) : 1;
    int bar = argc > 4 ? atoi(argv[4]) : 0, prio = argc > 5 ? atoi(argv[5]) : 1, gridA = argc > 6 ? atoi(argv[6]) : 147; int carveA = argc > 7 ? atoi(argv[7]) : 0;
    auto A = tmem ? kernA<1> : kernA<0>;
    cudaFuncSetAttribute(A, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 64);
    if (carveA) cudaFuncSetAttribute(A, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (carveB) cudaFuncSetAttribute(kernB<2256>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaStream_t sa, sb; int lo, hi; cudaDeviceGetStreamPriorityRange(&lo, &hi);
    cudaStreamCreateWithPriority(&sa, cudaStreamNonBlocking, prio ? hi : lo); cudaStreamCreateWithPriority(&sb, cudaStreamNonBlocking, lo);
    unsigned int z = 0; unsigned int zz[256] = {};
    cudaMemcpyToSymbol(started, &z, 4); cudaMemcpyToSymbol(stop_flag, &z, 4); cudaMemcpyToSymbol(sm_hits, zz, sizeof(zz));
    void *p_started; cudaGetSymbolAddress(&p_started, started);
    A<<<gridA, 512, smemA, sa>>>(3000000ull, bar);
    cuStreamWaitValue32((CUstream)sb, (CUdeviceptr)(uintptr_t)p_started, gridA, CU_STREAM_WAIT_VALUE_GEQ);
    kernB<2256><<<20000, 256, 0, sb>>>();
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(zz, sm_hits, sizeof(zz));
    int sms = 0; unsigned tot = 0, mx = 0; for (int i = 0; i < 256; ++i) { if (zz[i]) ++sms; tot += zz[i]; if (zz[i] > mx) mx = zz[i]; }
    printf("carveA=%d smemA=%d tmem=%d carveB=%d bar=%d prio=%d gridA=%d -> %s: B blocks while A ran: %u on %d SMs (max %u per SM)\n", carveA, smemA, tmem, carveB, bar, prio, gridA, cudaGetErrorString(e), tot, sms, mx);
    return 0;
}
