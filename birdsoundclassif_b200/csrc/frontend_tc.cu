// Tensor-core front-end for sm_100a: the sliding DFT of frontend.cu with its two contractions
// moved onto tcgen05.mma (accumulators in TMEM).
//
//   anchors   R_a[k]  = sum_n x[a*hop - N/2 + n] w^(kn)            one frame in 64, K = N/2 folded pairs
//   slides    D_t[k]  = sum_m (x[s_t+N+m] - x[s_t+m]) w^(km)       every frame, K = hop/2 folded pairs
//   R_{t+1} = w^(-hop k)(R_t + D_t),   X_t[k] = R_t[k]/2 - (R_t[k-1] + R_t[k+1])/4  (Hann),   dB
//
// GEMM orientation: M = 128 frequency bins (TMEM lanes), N = frames (TMEM columns), K = folded sample
// pairs.  A = twiddles (cos / sin), resident in shared memory for the CTA's whole life; B = folded
// samples built by the CTA from int16 PCM.  After tcgen05.ld each thread owns ONE bin and a run of
// consecutive frames in registers, so the per-frame recurrence is a register chain with no shuffles.
//
// Precision: PCM sums are 18-bit integers, exact as fp16 hi + lo (scaled by 1/8); twiddles are
// hi + 2^-11 lo'.  Three fp16 products (hi*hi, lo*hi, hi'*lo' with hi' = 2^-11 hi kept normal) reproduce
// the fp32 product to ~2^-24; accumulation is fp32 in TMEM.  The anchor contraction (K = 662) is split
// over 8 TMEM accumulators summed in the epilogue to keep fp32 accumulation chains short.
//
// Shared-memory operands use the canonical K-major, no-swizzle UMMA layout: 8x8 core matrices of
// 128 contiguous bytes, K-adjacent core matrices contiguous (LBO = 128 B), 8-row groups SBO apart.
#include <cuda_fp16.h>
#include <vector>
#include <cmath>

#include "frontend_tc.cuh"

namespace nbm {

constexpr int TC_THREADS = 256;
constexpr int BINS_PER_RANGE = 126;     // rows 1..126 of each 128-row range are emitted, 0 and 127 are Hann halos
constexpr int STAGE_LD = 33;            // float2 per bin row of the R stage (32 frames + pad)
constexpr int PADF = 8;                 // front padding (floats) of the sample buffer
constexpr int NA = 32;                  // anchors per anchor task (MMA N)
constexpr int KS = 4;                   // k-steps (of 16 pairs) per anchor stage
constexpr int NP = 8;                   // partial accumulators of the anchor contraction

struct TcParams {
    int N, hop, low_idx, n_bins, n_ranges;
    int npH, KP, nk;            // hop/2 pairs, padded K of the slide GEMM, k-steps
    int npN, n_stages;          // N/2 pairs, anchor stages of KS k-steps
    int off;                    // sample-buffer offset making the 8-pair vectors 16 B aligned
    int buf_len;                // floats in the sample buffer
    float min_level_sq;
    const __half *a_slide;      // [n_ranges][4][128 x KP]  (cos_h, cos_l, sin_h, sin_l) UMMA layout
    const __half *a_anchor;     // [n_ranges][n_stages][4][128 x 64]
    const float2 *cf, *gf, *gb, *rot;   // [n_ranges*128] per-bin constants
};

struct TcPlan {
    TcParams p;
    void *d_blob = nullptr;
    size_t smem_slide = 0, smem_anchor = 0;
    int grid_slide = 0;
};

// ------------------------------------------------------------------------- PTX helpers ---------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
// bounded wait: a lost arrival traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const long long t0 = clock64();
    while (!mbar_try_wait(bar, parity)) {
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// K-major, SWIZZLE_NONE shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
    d |= (uint64_t)1 << 46;          // descriptor version 1 (Blackwell)
    return d;
}
// kind::f16 instruction descriptor: fp16 x fp16 -> fp32, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc(int M, int Ncols) {
    return (1u << 4) | ((uint32_t)(Ncols >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
                 "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
                   "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
                   "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                 : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// log2 of a normal positive float: one MUFU.LG2 (the C intrinsic adds a denormal fix-up we never need)
__device__ __forceinline__ float fast_log2(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// byte offset of element (row, k) in a K-major no-swizzle operand with K extent kext
__host__ __device__ inline size_t umma_off(int row, int k, int kext) {
    return (size_t)(row >> 3) * (kext >> 3) * 128 + (size_t)(k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}

__device__ __forceinline__ float load_scaled(const void *pcm, int dtype, int channels, long long idx) {
    // sample in "int16 units / 8": exact for PCM16, so that 18-bit pair sums split exactly into fp16 hi + lo
    float s = 0.f;
    if (dtype == NBM_PCM_INT16) {
        const short *p = reinterpret_cast<const short *>(pcm) + idx * channels;
        for (int c = 0; c < channels; ++c) s += (float)__ldg(p + c);
        s *= 0.125f;
    } else {
        const float *p = reinterpret_cast<const float *>(pcm) + idx * channels;
        for (int c = 0; c < channels; ++c) s += __ldg(p + c);
        s *= 4096.0f;
    }
    return channels == 1 ? s : s / (float)channels;
}

// 8 values -> fp16 hi, lo (= v - hi) and hi' (= hi * 2^-11), each packed as one 16-byte vector
__device__ __forceinline__ void split8(const float (&v)[8], uint4 &h, uint4 &l, uint4 &hs) {
    uint32_t hh[4], ll[4], ss[4];
    const __half2 scale = __float2half2_rn(0.00048828125f);     // 2^-11
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const __half2 hp = __floats2half2_rn(v[2 * i], v[2 * i + 1]);
        const float2 hf = __half22float2(hp);
        const __half2 lp = __floats2half2_rn(v[2 * i] - hf.x, v[2 * i + 1] - hf.y);
        const __half2 sp = __hmul2(hp, scale);
        hh[i] = *reinterpret_cast<const uint32_t *>(&hp);
        ll[i] = *reinterpret_cast<const uint32_t *>(&lp);
        ss[i] = *reinterpret_cast<const uint32_t *>(&sp);
    }
    h = make_uint4(hh[0], hh[1], hh[2], hh[3]);
    l = make_uint4(ll[0], ll[1], ll[2], ll[3]);
    hs = make_uint4(ss[0], ss[1], ss[2], ss[3]);
}

__device__ __forceinline__ int find_seg(const SegDesc *segs, int n_segs, int tile) {
    int lo = 0, hi = n_segs - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (segs[mid].group0 <= tile) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// ------------------------------------------------------------------------- anchor kernel -------
// One task = NA consecutive anchor frames (frames 64*i) of one segment x one 128-bin range.
__global__ void __launch_bounds__(TC_THREADS, 1)
anchor_tc_kernel(TcParams P, const SegDesc *__restrict__ segs, int n_segs, const int *__restrict__ task_seg,
                 const int *__restrict__ task_first, const void *__restrict__ pcm, int dtype, int channels,
                 float2 *__restrict__ anchors) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int range = blockIdx.y;
    const int seg_idx = task_seg[blockIdx.x];
    const int first = task_first[blockIdx.x];          // first anchor (segment-local index) of this task
    const SegDesc sd = segs[seg_idx];
    const int n_anch_seg = (sd.n_frames + GF - 1) / GF + 1;
    const long long anchor_base = (long long)sd.group0 + seg_idx;      // global index of the segment's anchor 0

    unsigned char *sA = smem_raw;                                   // KS x 4 x [128 x 16]  = 64 KB, stage-contiguous
    unsigned char *sB = smem_raw + (size_t)KS * 4 * 128 * 16 * 2;   // 6 x [NA x 64]
    constexpr int A_STAGE_BYTES = KS * 4 * 128 * 16 * 2;
    constexpr int B_MAT_BYTES = NA * KS * 16 * 2;

    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc(128, NA);
    const int N = P.N, half = N / 2;

    for (int st = 0; st < P.n_stages; ++st) {
        // ---- A stage: straight 16-byte copies of the pre-laid-out twiddle block -------------------
        const uint4 *ga = reinterpret_cast<const uint4 *>(
            reinterpret_cast<const unsigned char *>(P.a_anchor) + ((size_t)range * P.n_stages + st) * A_STAGE_BYTES);
        uint4 *da = reinterpret_cast<uint4 *>(sA);
        for (int i = tid; i < A_STAGE_BYTES / 16; i += TC_THREADS) da[i] = __ldg(ga + i);
        // ---- B stage: thread = (anchor a, group of 8 pairs) --------------------------------------
        {
            const int a = tid % NA, jg = tid / NA;                // jg in [0, 8): pairs 8*jg .. 8*jg+7 of this stage
            const int ai = first + a;
            float fp[8], fm[8];
            const long long f0 = (long long)ai * GF * P.hop - half;   // segment-relative start of the anchor frame
            const bool live = ai < n_anch_seg;
#pragma unroll
            for (int jj = 0; jj < 8; ++jj) {
                const int j = st * (KS * 16) + jg * 8 + jj;
                float hi = 0.f, lo = 0.f;
                if (live && j < P.npN) {
                    const long long shi = f0 + half + j, slo = f0 + half - 1 - j;
                    if (shi >= 0 && shi < sd.n_samples) hi = load_scaled(pcm, dtype, channels, sd.pcm_start + shi);
                    if (slo >= 0 && slo < sd.n_samples) lo = load_scaled(pcm, dtype, channels, sd.pcm_start + slo);
                }
                fp[jj] = hi + lo;
                fm[jj] = hi - lo;
            }
            uint4 h, l, hs;
            const size_t o = umma_off(a, jg * 8, KS * 16);
            split8(fp, h, l, hs);
            *reinterpret_cast<uint4 *>(sB + 0 * B_MAT_BYTES + o) = h;
            *reinterpret_cast<uint4 *>(sB + 1 * B_MAT_BYTES + o) = l;
            *reinterpret_cast<uint4 *>(sB + 2 * B_MAT_BYTES + o) = hs;
            split8(fm, h, l, hs);
            *reinterpret_cast<uint4 *>(sB + 3 * B_MAT_BYTES + o) = h;
            *reinterpret_cast<uint4 *>(sB + 4 * B_MAT_BYTES + o) = l;
            *reinterpret_cast<uint4 *>(sB + 5 * B_MAT_BYTES + o) = hs;
        }
        fence_async_smem();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const int part = st % NP;
            const uint32_t d_cos = tmem_base + part * (2 * NA), d_sin = d_cos + NA;
            const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
            constexpr uint32_t A_MAT = 128 * KS * 16 * 2;          // one of the 4 A matrices of the stage
            constexpr uint32_t SBO = (KS * 16 / 8) * 128;
#pragma unroll
            for (int kk = 0; kk < KS; ++kk) {
                const uint32_t acc = (st >= NP || kk > 0) ? 1u : 0u;
                const uint32_t ko = kk * 256;
                const uint64_t ach = make_desc(a0 + 0 * A_MAT + ko, 128, SBO), acl = make_desc(a0 + 1 * A_MAT + ko, 128, SBO);
                const uint64_t ash = make_desc(a0 + 2 * A_MAT + ko, 128, SBO), asl = make_desc(a0 + 3 * A_MAT + ko, 128, SBO);
                const uint64_t bph = make_desc(b0 + 0 * B_MAT_BYTES + ko, 128, SBO), bpl = make_desc(b0 + 1 * B_MAT_BYTES + ko, 128, SBO);
                const uint64_t bps = make_desc(b0 + 2 * B_MAT_BYTES + ko, 128, SBO), bmh = make_desc(b0 + 3 * B_MAT_BYTES + ko, 128, SBO);
                const uint64_t bml = make_desc(b0 + 4 * B_MAT_BYTES + ko, 128, SBO), bms = make_desc(b0 + 5 * B_MAT_BYTES + ko, 128, SBO);
                umma_f16(d_cos, ach, bph, idesc, acc);
                umma_f16(d_cos, ach, bpl, idesc, 1u);
                umma_f16(d_cos, acl, bps, idesc, 1u);
                umma_f16(d_sin, ash, bmh, idesc, acc);
                umma_f16(d_sin, ash, bml, idesc, 1u);
                umma_f16(d_sin, asl, bms, idesc, 1u);
            }
            umma_commit(&bar);
        }
        mbar_wait(&bar, st & 1);           // MMAs of this stage have consumed the shared-memory operands
        tc_fence_after();
    }

    // ---- epilogue: thread = bin row (4 warps), sum the partials, rotate, store -------------------
    if (warp < 4) {
        const int row = warp * 32 + lane;
        const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
        float ac[NA], as[NA], v[32];
#pragma unroll
        for (int i = 0; i < NA; ++i) { ac[i] = 0.f; as[i] = 0.f; }
        const int used = min(NP, P.n_stages);
        for (int part = 0; part < used; ++part) {
            tmem_ld32(tl + part * (2 * NA), v);
#pragma unroll
            for (int i = 0; i < NA; ++i) ac[i] += v[i];
            tmem_ld32(tl + part * (2 * NA) + NA, v);
#pragma unroll
            for (int i = 0; i < NA; ++i) as[i] += v[i];
        }
        const float2 rot = P.rot[range * 128 + row];          // e^{-i theta (N-1)/2} * 2^-12 = (c, -s) form below
#pragma unroll
        for (int a = 0; a < NA; ++a) {
            const int ai = first + a;
            if (ai < n_anch_seg) {
                // R = (c0 - i s0)(A - iB)
                const float rr = rot.x * ac[a] - rot.y * as[a];
                const float ri = -(rot.x * as[a] + rot.y * ac[a]);
                anchors[(anchor_base + ai) * (P.n_ranges * 128) + range * 128 + row] = make_float2(rr, ri);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------- slide kernel --------
struct TileCtx {
    SegDesc sd;
    int t0;                 // first frame of the tile inside its segment
    long long anchor_row;   // float2 index of this thread's anchor for its direction
    long long s0;           // segment-relative sample index of buf[0]
    int fast;               // 16-byte vector loads possible for this tile
    unsigned full;          // bit e: prefetched vector e lies fully inside the segment
};

constexpr int PF = 6;       // prefetched 16-byte PCM vectors per thread (covers buf_len <= 8 * PF * TC_THREADS)

// 8 PCM16 samples (one 16-byte vector) -> 8 floats in "int16 / 8" units, exact:
// bits(2^20 + u/8) = 0x49800000 | u for u = x + 32768, so subtracting 2^20 + 4096 leaves x / 8.
__device__ __forceinline__ void cvt8_pcm16(const int4 raw, float4 &lo, float4 &hi) {
    const uint32_t w[4] = {(uint32_t)raw.x ^ 0x80008000u, (uint32_t)raw.y ^ 0x80008000u,
                           (uint32_t)raw.z ^ 0x80008000u, (uint32_t)raw.w ^ 0x80008000u};
    float f[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        f[2 * i] = __uint_as_float(0x49800000u | (w[i] & 0xffffu)) - 1052672.0f;
        f[2 * i + 1] = __uint_as_float(0x49800000u | (w[i] >> 16)) - 1052672.0f;
    }
    lo = make_float4(f[0], f[1], f[2], f[3]);
    hi = make_float4(f[4], f[5], f[6], f[7]);
}

// Persistent CTA: one 128-bin range, a strided sequence of 64-frame tiles.  The MMAs of tile i run while
// the CTA does the epilogue of tile i-1 (two TMEM accumulator buffers, one shared-memory B buffer).
__global__ void __launch_bounds__(TC_THREADS, 1)
slide_tc_kernel(TcParams P, const SegDesc *__restrict__ segs, int n_segs, int total_tiles,
                const void *__restrict__ pcm, int dtype, int channels, int vec_ok,
                const float2 *__restrict__ anchors, float *__restrict__ spec, float2 *__restrict__ tile_mm) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t bar[2];
    __shared__ uint32_t tmem_base_s;
    __shared__ unsigned int s_mm[2];            // this (tile, range)'s min / max, order-preserving encoding
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int range = blockIdx.x % P.n_ranges;
    const int q0 = blockIdx.x / P.n_ranges, qstride = gridDim.x / P.n_ranges;

    const int KP = P.KP, nk = P.nk, hop = P.hop, N = P.N;
    const size_t a_mat = (size_t)128 * KP * 2, b_mat = (size_t)GF * KP * 2;
    unsigned char *sA = smem_raw;                               // 4 x [128 x KP]
    unsigned char *sB = sA + 4 * a_mat;                         // 6 x [64 x KP]
    float *buf = reinterpret_cast<float *>(sB + 6 * b_mat);     // samples (scaled), PADF + off front padding
    float2 *stage = reinterpret_cast<float2 *>(buf + P.buf_len);    // [128][STAGE_LD]

    // ---- one-time: twiddles resident, B padding zeroed, TMEM, barriers --------------------------
    {
        const uint4 *ga = reinterpret_cast<const uint4 *>(reinterpret_cast<const unsigned char *>(P.a_slide) + (size_t)range * 4 * a_mat);
        uint4 *da = reinterpret_cast<uint4 *>(sA);
        for (int i = tid; i < (int)(4 * a_mat / 16); i += TC_THREADS) da[i] = __ldg(ga + i);
        uint4 *db = reinterpret_cast<uint4 *>(sB);
        for (int i = tid; i < (int)(6 * b_mat / 16); i += TC_THREADS) db[i] = make_uint4(0, 0, 0, 0);
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, 256);
    if (tid == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        s_mm[0] = 0xffffffffu; s_mm[1] = 0u;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t idesc = make_idesc(128, GF);
    const uint32_t SBO = (uint32_t)(KP / 8) * 128;

    // epilogue-1 role: bin row, direction
    const int quarter = warp & 3, dir = warp >> 2;
    const int row = quarter * 32 + lane;
    const float2 cf = P.cf[range * 128 + row];
    const float2 gg = dir == 0 ? P.gf[range * 128 + row] : P.gb[range * 128 + row];
    const uint32_t lane_sel = (uint32_t)(quarter * 32) << 16;
    const int njg = (P.npH + 7) / 8;
    const int half_hop = hop / 2;
    const int n_iters = q0 < total_tiles ? (total_tiles - q0 + qstride - 1) / qstride : 0;

    int seg_idx = 0;
    TileCtx cur{}, prev{}, nxt{};
    int4 pre[PF];
    const short *p16 = reinterpret_cast<const short *>(pcm);
    const int nv = P.buf_len / 8;
    const bool can_prefetch = vec_ok && nv <= PF * TC_THREADS;
    // tile context + issue of the 16-byte PCM loads (they land while the previous tile's epilogue runs)
    auto open_tile = [&](int it_) {
        const int tile = q0 + it_ * qstride;
        while (seg_idx + 1 < n_segs && segs[seg_idx + 1].group0 <= tile) ++seg_idx;
        nxt.sd = segs[seg_idx];
        const int lt = tile - nxt.sd.group0;
        nxt.t0 = lt * GF;
        nxt.anchor_row = ((long long)nxt.sd.group0 + seg_idx + lt + dir) * (P.n_ranges * 128) + range * 128 + row;
        nxt.s0 = (long long)nxt.t0 * hop - N / 2 - (PADF + P.off);
        const long long g0 = nxt.sd.pcm_start + nxt.s0;
        nxt.fast = can_prefetch && (g0 & 7) == 0;
        nxt.full = 0u;
        if (nxt.fast) {
#pragma unroll
            for (int e = 0; e < PF; ++e) {
                const int v = tid + e * TC_THREADS;
                const long long sv = nxt.s0 + 8 * v;
                if (v < nv && sv >= 0 && sv + 8 <= nxt.sd.n_samples) {
                    pre[e] = __ldg(reinterpret_cast<const int4 *>(p16 + g0 + 8 * v));
                    nxt.full |= 1u << e;
                }
            }
        }
    };
    float vmin = INFINITY, vmax = -INFINITY;

    if (n_iters > 0) open_tile(0);
    for (int it = 0; it <= n_iters; ++it) {
        if (it < n_iters) {
            cur = nxt;
            // the MMAs of tile it-1 (issued one iteration ago) must have finished reading sB
            if (it >= 1) mbar_wait(&bar[(it - 1) & 1], ((it - 1) >> 1) & 1);

            // ---- samples of the tile into shared memory (zero outside the segment) -------------
            if (cur.fast) {
                const long long g0 = cur.sd.pcm_start + cur.s0;
#pragma unroll
                for (int e = 0; e < PF; ++e) {
                    const int v = tid + e * TC_THREADS;
                    if (v < nv) {
                        float4 lo, hi;
                        if (cur.full & (1u << e)) {
                            cvt8_pcm16(pre[e], lo, hi);
                        } else {
                            const long long sv = cur.s0 + 8 * v;
                            float f[8];
#pragma unroll
                            for (int el = 0; el < 8; ++el)
                                f[el] = (sv + el >= 0 && sv + el < cur.sd.n_samples) ? (float)__ldg(p16 + g0 + 8 * v + el) * 0.125f : 0.f;
                            lo = make_float4(f[0], f[1], f[2], f[3]);
                            hi = make_float4(f[4], f[5], f[6], f[7]);
                        }
                        reinterpret_cast<float4 *>(buf)[2 * v] = lo;
                        reinterpret_cast<float4 *>(buf)[2 * v + 1] = hi;
                    }
                }
            } else {
                for (int i = tid; i < P.buf_len; i += TC_THREADS) {
                    const long long sx = cur.s0 + i;
                    buf[i] = (sx >= 0 && sx < cur.sd.n_samples) ? load_scaled(pcm, dtype, channels, cur.sd.pcm_start + sx) : 0.f;
                }
            }
            __syncthreads();
            // ---- B operand: folded differences, fp16 hi / lo / hi*2^-11 --------------------------
            for (int u = tid; u < GF * njg; u += TC_THREADS) {
                const int n = u % GF, jg = u / GF;
                const int base = PADF + P.off + n * hop;
                const float4 *nh = reinterpret_cast<const float4 *>(buf + base + N + half_hop + 8 * jg);
                const float4 *oh = reinterpret_cast<const float4 *>(buf + base + half_hop + 8 * jg);
                const float4 *nl = reinterpret_cast<const float4 *>(buf + base + N + half_hop - 8 - 8 * jg);
                const float4 *ol = reinterpret_cast<const float4 *>(buf + base + half_hop - 8 - 8 * jg);
                const float4 a0 = nh[0], a1 = nh[1], b0 = oh[0], b1 = oh[1];
                const float4 c0 = nl[0], c1 = nl[1], e0 = ol[0], e1 = ol[1];
                const float dh[8] = {a0.x - b0.x, a0.y - b0.y, a0.z - b0.z, a0.w - b0.w,
                                     a1.x - b1.x, a1.y - b1.y, a1.z - b1.z, a1.w - b1.w};
                // the lo sample of pair jj sits at position 7 - jj of the ascending vector
                const float dl[8] = {c1.w - e1.w, c1.z - e1.z, c1.y - e1.y, c1.x - e1.x,
                                     c0.w - e0.w, c0.z - e0.z, c0.y - e0.y, c0.x - e0.x};
                float ep[8], em[8];
#pragma unroll
                for (int jj = 0; jj < 8; ++jj) {
                    const bool ok = 8 * jg + jj < P.npH;
                    ep[jj] = ok ? dh[jj] + dl[jj] : 0.f;
                    em[jj] = ok ? dh[jj] - dl[jj] : 0.f;
                }
                uint4 h, l, hs;
                const size_t o = umma_off(n, jg * 8, KP);
                split8(ep, h, l, hs);
                *reinterpret_cast<uint4 *>(sB + 0 * b_mat + o) = h;
                *reinterpret_cast<uint4 *>(sB + 1 * b_mat + o) = l;
                *reinterpret_cast<uint4 *>(sB + 2 * b_mat + o) = hs;
                split8(em, h, l, hs);
                *reinterpret_cast<uint4 *>(sB + 3 * b_mat + o) = h;
                *reinterpret_cast<uint4 *>(sB + 4 * b_mat + o) = l;
                *reinterpret_cast<uint4 *>(sB + 5 * b_mat + o) = hs;
            }
            fence_async_smem();
            tc_fence_before();          // the epilogue's tcgen05.ld of two tiles ago precede this barrier
            __syncthreads();
            // ---- MMAs into TMEM buffer it&1: cos cols [0,64), sin cols [64,128) -------------------
            if (tid == 0) {
                tc_fence_after();
                const uint32_t d0 = tmem_base + (uint32_t)(it & 1) * 128;
                const uint32_t a0 = smem_u32(sA), b0 = smem_u32(sB);
                for (int kk = 0; kk < nk; ++kk) {
                    const uint32_t ko = kk * 256, acc = kk > 0 ? 1u : 0u;
                    const uint64_t ach = make_desc(a0 + 0 * (uint32_t)a_mat + ko, 128, SBO), acl = make_desc(a0 + 1 * (uint32_t)a_mat + ko, 128, SBO);
                    const uint64_t ash = make_desc(a0 + 2 * (uint32_t)a_mat + ko, 128, SBO), asl = make_desc(a0 + 3 * (uint32_t)a_mat + ko, 128, SBO);
                    const uint64_t bph = make_desc(b0 + 0 * (uint32_t)b_mat + ko, 128, SBO), bpl = make_desc(b0 + 1 * (uint32_t)b_mat + ko, 128, SBO);
                    const uint64_t bps = make_desc(b0 + 2 * (uint32_t)b_mat + ko, 128, SBO), bmh = make_desc(b0 + 3 * (uint32_t)b_mat + ko, 128, SBO);
                    const uint64_t bml = make_desc(b0 + 4 * (uint32_t)b_mat + ko, 128, SBO), bms = make_desc(b0 + 5 * (uint32_t)b_mat + ko, 128, SBO);
                    umma_f16(d0, ach, bph, idesc, acc);
                    umma_f16(d0, ach, bpl, idesc, 1u);
                    umma_f16(d0, acl, bps, idesc, 1u);
                    umma_f16(d0 + GF, ash, bmh, idesc, acc);
                    umma_f16(d0 + GF, ash, bml, idesc, 1u);
                    umma_f16(d0 + GF, asl, bms, idesc, 1u);
                }
                umma_commit(&bar[it & 1]);
            }
            if (it + 1 < n_iters) open_tile(it + 1);
        }
        if (it >= 1) {
            // ================= epilogue of tile it-1 (its MMAs were issued one iteration ago) =======
            const int pb = (it - 1) & 1;
            const SegDesc &sd = prev.sd;
            const int t0 = prev.t0;
            const float2 anc = __ldg(anchors + prev.anchor_row);
            mbar_wait(&bar[pb], ((it - 1) >> 1) & 1);
            tc_fence_after();
            const uint32_t tl = tmem_base + (uint32_t)pb * 128 + lane_sel;
            float Rr = anc.x, Ri = anc.y;
            float *spec_seg = spec + sd.spec_off;
            for (int hp = 0; hp < 2; ++hp) {
                float gc[16], gs[16];
                if (dir == 0) {
                    // frames 16hp .. 16hp+15 ; G columns: hp=0 -> 0..14 (frame 0 is the anchor), hp=1 -> 15..30
                    const int c0 = hp == 0 ? 0 : 15;
                    tmem_ld16(tl + c0, gc);
                    tmem_ld16(tl + GF + c0, gs);
                    if (hp == 0) stage[row * STAGE_LD + 0] = make_float2(Rr, Ri);
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        if (hp == 0 && i == 15) break;
                        const float nr = cf.x * Rr - cf.y * Ri + gg.x * gc[i] + gg.y * gs[i];
                        const float ni = cf.x * Ri + cf.y * Rr - gg.x * gs[i] + gg.y * gc[i];
                        Rr = nr; Ri = ni;
                        stage[row * STAGE_LD + (hp == 0 ? i + 1 : i)] = make_float2(Rr, Ri);
                    }
                } else {
                    // frames 63-16hp down to 48-16hp ; G columns 63-16hp .. 48-16hp
                    const int c0 = 48 - 16 * hp;
                    tmem_ld16(tl + c0, gc);
                    tmem_ld16(tl + GF + c0, gs);
#pragma unroll
                    for (int i = 15; i >= 0; --i) {
                        const float nr = cf.x * Rr + cf.y * Ri - gg.x * gc[i] + gg.y * gs[i];
                        const float ni = cf.x * Ri - cf.y * Rr + gg.x * gs[i] + gg.y * gc[i];
                        Rr = nr; Ri = ni;
                        stage[row * STAGE_LD + 16 + i] = make_float2(Rr, Ri);
                    }
                }
                __syncthreads();
                // ---- Hann + dB + store: lane = frame column, warp = 16 bin rows --------------------
                {
                    const int fl = lane < 16 ? 16 * hp + lane : 48 - 16 * hp + (lane - 16);     // frame inside the tile
                    const int r_lo = 1 + 16 * warp;
                    // rows r < ob_lim are inside the band (only the last range is cut short)
                    const int r_hi = min(min(r_lo + 16, 127), P.n_bins - (range * BINS_PER_RANGE - 1));
                    if (t0 + fl < sd.n_frames && r_lo < r_hi) {
                        float2 pv = stage[(r_lo - 1) * STAGE_LD + lane], cu = stage[r_lo * STAGE_LD + lane];
                        float *out = spec_seg + (long long)(range * BINS_PER_RANGE + r_lo - 1) * sd.row_stride + t0 + fl;
#pragma unroll 4
                        for (int r = r_lo; r < r_hi; ++r) {
                            const float2 nx = stage[(r + 1) * STAGE_LD + lane];
                            // 4 X = 2 R[k] - (R[k-1] + R[k+1]);  10 log10(|X|^2) = 10 log10(|4X|^2) - 10 log10(16)
                            const float xr = fmaf(2.0f, cu.x, -(pv.x + nx.x));
                            const float xi = fmaf(2.0f, cu.y, -(pv.y + nx.y));
                            const float pw = fmaxf(fmaf(xr, xr, xi * xi), 16.0f * P.min_level_sq);
                            const float db = fmaf(fast_log2(pw), 3.0102999566398120f, -12.041199826559248f);
                            *out = db;
                            vmin = fminf(vmin, db);
                            vmax = fmaxf(vmax, db);
                            out += sd.row_stride;
                            pv = cu; cu = nx;
                        }
                    }
                }
                __syncthreads();
            }
            // ---- this (tile, range)'s min / max for the whole-file reduction (refine_minmax_kernel) ----
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
                vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            }
            if (lane == 0 && vmin <= vmax) {
                atomicMin(&s_mm[0], float_to_ordered(vmin));
                atomicMax(&s_mm[1], float_to_ordered(vmax));
            }
            vmin = INFINITY; vmax = -INFINITY;
            __syncthreads();
            if (tid == 0) {
                const long long tile = q0 + (long long)(it - 1) * qstride;
                tile_mm[tile * P.n_ranges + range] = make_float2(ordered_to_float(s_mm[0]), ordered_to_float(s_mm[1]));
                s_mm[0] = 0xffffffffu; s_mm[1] = 0u;        // next use is behind at least one more barrier
            }
        }
        prev = cur;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 256);
}

}  // namespace nbm

// ----------------------------------------------------------------------------- host side -------
using namespace nbm;

namespace {
inline void put_split(std::vector<__half> &dst, size_t base_elems_h, size_t base_elems_l, size_t off_bytes, double w) {
    const __half h = __float2half_rn((float)w);
    const double rem = (w - (double)__half2float(h)) * 2048.0;
    dst[base_elems_h + off_bytes / 2] = h;
    dst[base_elems_l + off_bytes / 2] = __float2half_rn((float)rem);
}
}  // namespace

int nbm::tc_plan_create(const nbm_frontend_params &p, TcPlan **out) {
    *out = nullptr;
    const int N = p.n_fft, hop = p.hop;
    const int npH = hop / 2, KP = ((npH + 15) / 16) * 16;
    const bool ok = (N % 4 == 0) && (hop % 4 == 0) && KP <= 80 && hop >= 8 && p.n_bins <= 3 * BINS_PER_RANGE &&
                    p.low_idx >= 1 && N >= 2 * hop;
    if (!ok) return NBM_ERR_UNSUPPORTED;
    auto *pl = new TcPlan();
    TcParams &k = pl->p;
    k.N = N; k.hop = hop; k.low_idx = p.low_idx; k.n_bins = p.n_bins;
    k.n_ranges = (p.n_bins + BINS_PER_RANGE - 1) / BINS_PER_RANGE;
    k.npH = npH; k.KP = KP; k.nk = KP / 16;
    k.npN = N / 2; k.n_stages = (k.npN + KS * 16 - 1) / (KS * 16);
    k.off = (4 - (npH % 4)) % 4;
    k.buf_len = ((PADF + k.off + GF * hop + N + 16 + 7) / 8) * 8;   // +16: masked tail pairs read past the last block
    k.min_level_sq = (float)(p.min_level * p.min_level);
    const int R = k.n_ranges, N2 = 2 * N;

    const size_t slide_elems = (size_t)R * 4 * 128 * KP;
    const size_t anchor_elems = (size_t)R * k.n_stages * 4 * 128 * (KS * 16);
    std::vector<__half> a_slide(slide_elems, __float2half_rn(0.f)), a_anchor(anchor_elems, __float2half_rn(0.f));
    std::vector<float2> cf(R * 128), gf(R * 128), gb(R * 128), rot(R * 128);
    auto ang = [&](long long q) { return M_PI * (double)(q % N2) / (double)N; };
    const double s12 = 1.0 / 4096.0;
    for (int r = 0; r < R; ++r)
        for (int row = 0; row < 128; ++row) {
            const long long kbin = p.low_idx - 1 + (long long)r * BINS_PER_RANGE + row;
            // slide twiddles: pair j <-> u = j + 1/2
            const size_t mat = (size_t)128 * KP;
            for (int j = 0; j < npH; ++j) {
                const double a = ang(kbin * (2 * j + 1));
                const size_t o = umma_off(row, j, KP);
                put_split(a_slide, ((size_t)r * 4 + 0) * mat, ((size_t)r * 4 + 1) * mat, o, cos(a));
                put_split(a_slide, ((size_t)r * 4 + 2) * mat, ((size_t)r * 4 + 3) * mat, o, sin(a));
            }
            const size_t amat = (size_t)128 * (KS * 16);
            for (int j = 0; j < k.npN; ++j) {
                const double a = ang(kbin * (2 * j + 1));
                const int st = j / (KS * 16), jj = j % (KS * 16);
                const size_t base = ((size_t)r * k.n_stages + st) * 4 * amat;
                const size_t o = umma_off(row, jj, KS * 16);
                put_split(a_anchor, base + 0 * amat, base + 1 * amat, o, cos(a));
                put_split(a_anchor, base + 2 * amat, base + 3 * amat, o, sin(a));
            }
            const int i = r * 128 + row;
            double a = ang(kbin * 2 * hop);
            cf[i] = make_float2((float)cos(a), (float)sin(a));
            a = ang(kbin * (hop + 1));
            gf[i] = make_float2((float)(cos(a) * s12), (float)(sin(a) * s12));
            a = ang(kbin * (hop - 1));
            gb[i] = make_float2((float)(cos(a) * s12), (float)(sin(a) * s12));
            a = ang(kbin * (N - 1));
            rot[i] = make_float2((float)(cos(a) * s12), (float)(sin(a) * s12));
        }
    const size_t b_slide = align_up(slide_elems * 2, 256), b_anchor = align_up(anchor_elems * 2, 256);
    const size_t b_c = align_up((size_t)R * 128 * sizeof(float2), 256);
    const size_t total = b_slide + b_anchor + 4 * b_c;
    cudaError_t e = cudaMalloc(&pl->d_blob, total);
    if (e != cudaSuccess) { delete pl; return cuda_fail(e, "cudaMalloc(tc tables)"); }
    unsigned char *d = reinterpret_cast<unsigned char *>(pl->d_blob);
    e = cudaMemcpy(d, a_slide.data(), slide_elems * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + b_slide, a_anchor.data(), anchor_elems * 2, cudaMemcpyHostToDevice);
    const float2 *src[4] = {cf.data(), gf.data(), gb.data(), rot.data()};
    for (int i = 0; i < 4 && e == cudaSuccess; ++i)
        e = cudaMemcpy(d + b_slide + b_anchor + i * b_c, src[i], (size_t)R * 128 * sizeof(float2), cudaMemcpyHostToDevice);
    k.a_slide = reinterpret_cast<const __half *>(d);
    k.a_anchor = reinterpret_cast<const __half *>(d + b_slide);
    k.cf = reinterpret_cast<const float2 *>(d + b_slide + b_anchor);
    k.gf = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + b_c);
    k.gb = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + 2 * b_c);
    k.rot = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + 3 * b_c);
    pl->smem_slide = (size_t)4 * 128 * KP * 2 + (size_t)6 * GF * KP * 2 + (size_t)k.buf_len * 4 + (size_t)128 * STAGE_LD * 8;
    pl->smem_anchor = (size_t)KS * 4 * 128 * 16 * 2 + (size_t)6 * NA * KS * 16 * 2;
    int dev = 0, sms = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    pl->grid_slide = std::max(1, sms / R) * R;
    cudaFuncAttributes fa_s{}, fa_a{};
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa_s, slide_tc_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa_a, anchor_tc_kernel);
    if (e == cudaSuccess && ((size_t)max_smem < pl->smem_slide + fa_s.sharedSizeBytes ||
                             (size_t)max_smem < pl->smem_anchor + fa_a.sharedSizeBytes)) {
        tc_plan_destroy(pl);
        return NBM_ERR_UNSUPPORTED;
    }
    // per-function attribute (not per plan): allow the device maximum minus the kernel's static shared memory
    if (e == cudaSuccess) e = cudaFuncSetAttribute(slide_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   max_smem - (int)fa_s.sharedSizeBytes);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(anchor_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                   max_smem - (int)fa_a.sharedSizeBytes);
    if (e != cudaSuccess) { tc_plan_destroy(pl); return cuda_fail(e, "tc_plan_create"); }
    *out = pl;
    return NBM_OK;
}

void nbm::tc_plan_destroy(TcPlan *pl) {
    if (!pl) return;
    if (pl->d_blob) cudaFree(pl->d_blob);
    delete pl;
}

size_t nbm::tc_anchor_bytes(const TcPlan *pl, long long n_anchors) {
    return align_up((size_t)n_anchors * pl->p.n_ranges * 128 * sizeof(float2), 256);
}

int nbm::tc_anchor_group() { return NA; }
int nbm::tc_n_ranges(const TcPlan *pl) { return pl->p.n_ranges; }
int nbm::tc_bins_per_range() { return BINS_PER_RANGE; }

int nbm::tc_launch(const TcPlan *pl, const SegDesc *d_segs, int n_segs, int total_tiles, const int *d_task_seg,
                   const int *d_task_first, int n_tasks, const void *d_pcm, int dtype, int channels, float *d_spec,
                   float2 *d_tile_mm, void *d_anchors, cudaStream_t stream) {
    const TcParams &k = pl->p;
    float2 *anchors = reinterpret_cast<float2 *>(d_anchors);
    dim3 ga((unsigned)n_tasks, (unsigned)k.n_ranges);
    anchor_tc_kernel<<<ga, TC_THREADS, pl->smem_anchor, stream>>>(k, d_segs, n_segs, d_task_seg, d_task_first, d_pcm,
                                                                  dtype, channels, anchors);
    const int grid = std::min(pl->grid_slide, std::max(1, total_tiles) * k.n_ranges);
    // 16-byte vector loads of PCM16 need a mono int16 stream on a 16-byte aligned base
    const int vec_ok = (dtype == NBM_PCM_INT16 && channels == 1 && (reinterpret_cast<uintptr_t>(d_pcm) & 15) == 0) ? 1 : 0;
    slide_tc_kernel<<<(grid / k.n_ranges) * k.n_ranges, TC_THREADS, pl->smem_slide, stream>>>(
        k, d_segs, n_segs, total_tiles, d_pcm, dtype, channels, vec_ok, anchors, d_spec, d_tile_mm);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}
