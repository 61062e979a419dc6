// Tensor-core front-end for sm_100a: the sliding DFT of frontend.cu with its two contractions
// moved onto tcgen05.mma (accumulators in TMEM).  Mono PCM16 input.
//
//   anchors   R_a[k]  = sum_n x[a*hop - N/2 + n] w^(kn)            one frame in 64, K = N/2 folded pairs
//   slides    D_t[k]  = sum_m (x[s_t+N+m] - x[s_t+m]) w^(km)       every frame, K = hop/2 folded pairs
//   R_{t+1} = w^(-hop k)(R_t + D_t),   X_t[k] = R_t[k]/2 - (R_t[k-1] + R_t[k+1])/4  (Hann),   dB
//
// GEMM orientation: M = 128 frequency bins (TMEM lanes), N = frames (TMEM columns), K = folded sample
// pairs.  A = twiddles (cos / sin): resident in TMEM for the slide kernel, streamed through shared
// memory by cp.async.bulk for the anchor kernel.  B = folded samples built from int16 PCM.  After
// tcgen05.ld each thread owns ONE bin and a run of consecutive frames in registers, so the per-frame
// recurrence is a register chain with no shuffles.
//
// Precision: everything fed to the tensor core is exact.  Twiddles are 2^11 w = H + L (two fp16 at the
// same scale); samples are split into BYTE PLANES (a folded sum or difference of int16 samples is formed
// separately on the high and low bytes in packed half arithmetic, value / 256 = hi + lo), so the four
// products (H + L)(hi + lo) reproduce the fp32 product; accumulation is fp32 in TMEM, which truncates:
// the integer-valued H x hi products get their own accumulator (anchors) or are added last (slides).
//
// Shared-memory operands use the canonical K-major, no-swizzle UMMA layout: 8x8 core matrices of
// 128 contiguous bytes, K-adjacent core matrices contiguous (LBO = 128 B), 8-row groups SBO apart.
#include <cuda_fp16.h>
#include <vector>
#include <cmath>

#include "frontend_tc.cuh"

namespace nbm {

constexpr int BINS_PER_RANGE = 126;     // rows 1..126 of each 128-row range are emitted, 0 and 127 are Hann halos
constexpr int PADF = 8;                 // front padding (samples) of the sample buffer

struct TcParams {
    int N, hop, low_idx, n_bins, n_ranges;
    int npH, KP, nk;            // hop/2 pairs, padded K of the slide GEMM, k-steps
    int npN, n_stages;          // N/2 pairs, anchor k-blocks of AKB pairs
    int off;                    // sample-buffer offset making the 8-pair vectors 16 B aligned
    int buf_len;                // samples in the per-chain sample buffer
    float min_level_sq;
    const __half *a_slide;      // [n_ranges][128 rows][cos_h, cos_l, sin_h, sin_l][KP], row-contiguous (copied into TMEM lanes)
    const __half *a_anchor;     // [n_ranges][n_stages][cos_h, cos_l, sin_h, sin_l][128 x AKB] UMMA layout
    const float2 *cf, *gf, *gb, *rot, *cfix;   // [n_ranges*128] per-bin constants (cfix: see the recurrence)
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
    // the hardware parks the thread up to the time hint (ns) before reporting "not yet"
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
    return ok != 0;
}
// the same for waits that are expected to be long (a whole pipeline stage): sleep between polls so the
// polling warp does not take issue slots from the working ones
template <int NS>
__device__ __forceinline__ void mbar_wait_long(uint64_t *bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(NS);
        if (++spins > (1 << 22)) __trap();
    }
}
// bounded wait: a lost arrival traps (launch error) instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    int spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1 << 24)) __trap();
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




// ------------------------------------------------------------------------- shared helpers -----
// Re-alignment of 16-bit samples read with aligned 16-byte loads: W holds NW consecutive 32-bit words (two samples
// each) starting at an aligned address; the wanted run starts `sh` samples (0..7) later.  Returns words [0, NO).
template <int NW, int NO>
__device__ __forceinline__ void realign_words(const uint32_t (&W)[NW], int sh, uint32_t (&out)[NO]) {
    static_assert(NW >= NO + 4, "need four spare words");
    uint32_t t[NO + 2], u[NO + 1];
    const bool q2 = (sh & 4) != 0, q1 = (sh & 2) != 0;
    const int r = (sh & 1) * 16;
#pragma unroll
    for (int i = 0; i < NO + 2; ++i) t[i] = q2 ? W[i + 2] : W[i];
#pragma unroll
    for (int i = 0; i < NO + 1; ++i) u[i] = q1 ? t[i + 1] : t[i];
#pragma unroll
    for (int i = 0; i < NO; ++i) out[i] = __funnelshift_r(u[i], u[i + 1], r);
}

// one lane of the (converged) warp; the compiler treats the guarded region as single-threaded, so warp-uniform
// operands of tcgen05 instructions need no per-value election loop
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred P;\n\telect.sync _|P, 0xffffffff;\n\tselp.u32 %0, 1, 0, P;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ __half2 u32_as_h2(uint32_t u) { return *reinterpret_cast<__half2 *>(&u); }
__device__ __forceinline__ uint32_t h2_as_u32(__half2 h) { return *reinterpret_cast<uint32_t *>(&h); }
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(taddr) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint4 &lo, const uint4 &hi) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr), "r"(lo.x), "r"(lo.y), "r"(lo.z), "r"(lo.w), "r"(hi.x), "r"(hi.y), "r"(hi.z), "r"(hi.w) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// D[tmem] (+)= A[tmem] * B[smem]
__device__ __forceinline__ void umma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "r"(acc) : "memory");
}
// the same with the accumulate flag known at compile time (no predicate register to set up per instruction)
template <bool ACC>
__device__ __forceinline__ void umma_f16_ts_c(uint32_t tmem_d, uint32_t tmem_a, uint64_t bdesc, uint32_t idesc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(tmem_d), "r"(tmem_a), "l"(bdesc), "r"(idesc), "n"(ACC ? 1 : 0) : "memory");
}


// ------------------------------------------------------------------------- anchor kernel -------
// Anchors: the rectangular-window DFT of every 64th frame, R_a[k] = sum over the N/2 folded sample pairs.
// One CTA = AN consecutive anchors of one segment x one 128-bin range: a GEMM with M = 128 bins, N = AN anchors,
// K = N/2 pairs, streamed in k-blocks of AKB pairs through a two-stage shared-memory ring:
//   builder warps (8)  PCM16 straight from global memory (prefetched one k-block ahead) -> byte-plane fp16 B
//                      operand (same exact hi/lo split as the slide kernel); one elected thread also starts the
//                      cp.async.bulk of the k-block's pre-laid-out twiddle block (A operand, 32 KB).  (Cluster
//                      pairs sharing one multicast copy of A were tried: the L2 stream is not the limit and the
//                      pair's cluster barriers cost more than the halved traffic saved.)
//   MMA warp (1)       8 x tcgen05.mma per k-step (4 hi/lo products x cos/sin), accumulators in TMEM.  The hi x hi
//                      products are integers (twiddle x 2^11 times a byte-plane sum) and get an accumulator of
//                      their own, where fp32 addition is exact up to 2^24; the three small products (2^-8 .. 2^-19
//                      of the first) go to a second one, so they are not absorbed by the large partial sums
// Epilogue: thread = bin, add the two accumulators, rotate to the frame-start phase reference, store float2.
constexpr int AN = 128;                 // anchors per task (MMA N)
constexpr int AKB = 32;                 // pairs per k-block (2 k-steps)
constexpr int A_BUILD_WARPS = 4 * (AKB / 8);     // thread = (anchor, 8 pairs): AN / 32 warps per 8-pair slice of the k-block
constexpr int A_THREADS = 32 * (A_BUILD_WARPS + 1);
constexpr int A_MAT_BYTES = 128 * AKB * 2;          // one [128 x AKB] fp16 matrix (A: bins, B: anchors)
constexpr int A_STAGE_BYTES = 8 * A_MAT_BYTES;      // 4 twiddle + 4 sample matrices
constexpr int A_STAGES = 3;                         // the twiddle block of k-block kb+2 is in flight while kb is multiplied

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(A_THREADS, 1)
anchor_tc_kernel(TcParams P, const SegDesc *__restrict__ segs, int seg_lo, int seg_hi, long long anchor_begin,
                 long long anchor_end, const short *__restrict__ pcm, float2 *__restrict__ anchors) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint64_t full_a[A_STAGES], full_b[A_STAGES], empty[A_STAGES], done;
    __shared__ uint32_t tmem_base_s;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int range = blockIdx.y;
    // Anchors are numbered globally: segment s owns numbers group0[s] + s .. group0[s] + s + tiles[s] (frames 64 i and
    // the one just past its end), so a CTA's AN anchors may span several segments and every CTA but the last is full.
    const long long ag0 = anchor_begin + (long long)blockIdx.x * AN;
    const int n_kb = P.n_stages;

    if (warp == 0) tmem_alloc(&tmem_base_s, 512);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < A_STAGES; ++i) {
            mbar_init(&full_a[i], 1); mbar_init(&full_b[i], A_BUILD_WARPS); mbar_init(&empty[i], 1);
        }
        mbar_init(&done, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;

    if (warp == A_BUILD_WARPS) {
        // ================================ MMA issuer ===================================================
        const uint32_t idesc = make_idesc(128, AN);
        constexpr uint32_t SBO = (AKB / 8) * 128;
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % A_STAGES, ph = (kb / A_STAGES) & 1;
            mbar_wait(&full_a[s], ph);
            mbar_wait(&full_b[s], ph);
            tc_fence_after();
            if (elect_one()) {
                const uint32_t a0 = smem_u32(smem_raw + (size_t)s * A_STAGE_BYTES), b0 = a0 + 4 * A_MAT_BYTES;
                const uint32_t d_cos = tmem_base, d_sin = d_cos + AN;            // hi x hi (integer-valued) sums
                const uint32_t e_cos = tmem_base + 2 * AN, e_sin = e_cos + AN;   // the small products
#pragma unroll
                for (int kk = 0; kk < AKB / 16; ++kk) {
                    const uint32_t acc = (kb > 0 || kk > 0) ? 1u : 0u, ko = kk * 256;
                    const uint64_t ach = make_desc(a0 + 0 * A_MAT_BYTES + ko, 128, SBO), acl = make_desc(a0 + 1 * A_MAT_BYTES + ko, 128, SBO);
                    const uint64_t ash = make_desc(a0 + 2 * A_MAT_BYTES + ko, 128, SBO), asl = make_desc(a0 + 3 * A_MAT_BYTES + ko, 128, SBO);
                    const uint64_t bph = make_desc(b0 + 0 * A_MAT_BYTES + ko, 128, SBO), bpl = make_desc(b0 + 1 * A_MAT_BYTES + ko, 128, SBO);
                    const uint64_t bmh = make_desc(b0 + 2 * A_MAT_BYTES + ko, 128, SBO), bml = make_desc(b0 + 3 * A_MAT_BYTES + ko, 128, SBO);
                    umma_f16(d_cos, ach, bph, idesc, acc);
                    umma_f16(d_sin, ash, bmh, idesc, acc);
                    umma_f16(e_cos, ach, bpl, idesc, acc);
                    umma_f16(e_sin, ash, bml, idesc, acc);
                    umma_f16(e_cos, acl, bph, idesc, 1u);
                    umma_f16(e_sin, asl, bmh, idesc, 1u);
                    umma_f16(e_cos, acl, bpl, idesc, 1u);
                    umma_f16(e_sin, asl, bml, idesc, 1u);
                }
                umma_commit(&empty[s]);
                if (kb == n_kb - 1) umma_commit(&done);
            }
            __syncwarp();
        }
    } else {
        // ================================ builders: thread = (anchor a, 8 pairs of the k-block) ============
        // (16 warps: with 8, each building 16 pairs, a k-block took twice as long to build as to multiply)
        const int a = tid & (AN - 1), part = tid >> 7;              // part in 0 .. AKB/8 - 1
        const long long ag = ag0 + a;
        const bool live = ag < anchor_end;
        int lo_s = seg_lo, hi_s = seg_hi - 1;                       // this anchor's segment: last s with group0[s] + s <= ag
        while (lo_s < hi_s) {
            const int mid = (lo_s + hi_s + 1) >> 1;
            if ((long long)segs[mid].group0 + mid <= ag) lo_s = mid; else hi_s = mid - 1;
        }
        const SegDesc sd = segs[lo_s];
        const int ai = (int)(ag - ((long long)sd.group0 + lo_s));  // frame 64 ai of the segment
        const long long c = (long long)ai * GF * P.hop;            // segment-relative index of the frame centre
        // c and the pair offsets are multiples of 8 samples, so every 8-sample run of this thread starts `sh`
        // samples after a 16-byte boundary (sh = 0 when the file starts on one)
        const bool vec_ok = ((reinterpret_cast<uintptr_t>(pcm) & 15) == 0) && (GF * P.hop) % 8 == 0;
        const int sh = (int)(sd.pcm_start & 7);
        uint32_t hi_w[4], lo_w[4];          // 8 samples above / below the centre for this thread's pairs, offset binary
        int4 raw_h[2], raw_l[2];            // their raw 16-byte loads, in flight while the previous k-block is built
        bool fast_h = false, fast_l = false;
        // issue the loads of samples s0 .. s0+7 (segment-relative); nothing here waits for them
        auto issue8 = [&](long long s0, int4 (&raw)[2], bool &fastflag, uint32_t (&w)[4]) {
            const long long a0 = s0 - sh;                           // start of the aligned window (segment-relative)
            fastflag = live && vec_ok && a0 >= 0 && a0 + (sh ? 16 : 8) <= sd.n_samples;
            if (fastflag) {
                const int4 *p4 = reinterpret_cast<const int4 *>(pcm + sd.pcm_start + a0);
                raw[0] = __ldg(p4);
                if (sh) raw[1] = __ldg(p4 + 1);
            } else if (!live) {
#pragma unroll
                for (int i = 0; i < 4; ++i) w[i] = 0x80008000u;
            } else {                                                // first / last anchors of a segment: centre padding
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    uint32_t u[2];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const long long sx = s0 + 2 * i + h;
                        u[h] = (sx >= 0 && sx < sd.n_samples) ? ((uint32_t)(uint16_t)__ldg(pcm + sd.pcm_start + sx) ^ 0x8000u) : 0x8000u;
                    }
                    w[i] = u[0] | (u[1] << 16);
                }
            }
        };
        // first use of the loaded vectors: re-align (file not on a 16-byte boundary) and flip to offset binary
        auto finish8 = [&](const int4 (&raw)[2], bool fastflag, uint32_t (&w)[4]) {
            if (!fastflag) return;
            if (sh == 0) {
                w[0] = (uint32_t)raw[0].x; w[1] = (uint32_t)raw[0].y; w[2] = (uint32_t)raw[0].z; w[3] = (uint32_t)raw[0].w;
            } else {
                const uint32_t W[8] = {(uint32_t)raw[0].x, (uint32_t)raw[0].y, (uint32_t)raw[0].z, (uint32_t)raw[0].w,
                                       (uint32_t)raw[1].x, (uint32_t)raw[1].y, (uint32_t)raw[1].z, (uint32_t)raw[1].w};
                realign_words<8, 4>(W, sh, w);
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) w[i] ^= 0x80008000u;
        };
        auto prefetch = [&](int kb) {
            const int j0 = kb * AKB + 8 * part;                     // first pair of this thread in the k-block
            issue8(c + j0, raw_h, fast_h, hi_w);                    // pair j <-> sample c + j
            issue8(c - j0 - 8, raw_l, fast_l, lo_w);                //            and sample c - 1 - j
        };
        if (live) {
            // the anchor's whole window (N samples around c) into L2 now: the register loads below are issued one k-block
            // (~0.6 us) ahead, less than an HBM round trip; this thread takes every (AKB/8)-th 128-byte line
            const long long w0 = max(c - P.N / 2, 0ll), w1 = min(c + P.N / 2, sd.n_samples);
            const char *pb = reinterpret_cast<const char *>(pcm + sd.pcm_start + w0);
            for (long long o = (long long)part * 128; o < (w1 - w0) * 2; o += (AKB / 8) * 128)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pb + o));
        }
        prefetch(0);
        for (int kb = 0; kb < n_kb; ++kb) {
            const int s = kb % A_STAGES, ph = (kb / A_STAGES) & 1;
            mbar_wait(&empty[s], ph ^ 1);                           // the MMAs of k-block kb-A_STAGES have consumed this stage
            unsigned char *sA = smem_raw + (size_t)s * A_STAGE_BYTES, *sBm = sA + 4 * A_MAT_BYTES;
            if (tid == 0) {
                mbar_expect_tx(&full_a[s], 4 * A_MAT_BYTES);
                bulk_g2s(sA, reinterpret_cast<const unsigned char *>(P.a_anchor) + ((size_t)range * n_kb + kb) * 4 * A_MAT_BYTES,
                         4 * A_MAT_BYTES, &full_a[s]);
            }
            const int j0 = kb * AKB + 8 * part;
            finish8(raw_h, fast_h, hi_w);
            finish8(raw_l, fast_l, lo_w);
            {
                uint32_t ph4[4], pl4[4], mh4[4], ml4[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    // pairs j0 + 2q, +1: above the centre word hi_w[q]; below: mirrored, word 3 - q reversed
                    const uint32_t wa = hi_w[q], wc = lo_w[3 - q];
                    const __half2 ah = u32_as_h2(__byte_perm(wa, 0x64646464u, 0x4341)), al = u32_as_h2(__byte_perm(wa, 0x44444444u, 0x4240));
                    const __half2 ch = u32_as_h2(__byte_perm(wc, 0x64646464u, 0x4143)), cl = u32_as_h2(__byte_perm(wc, 0x44444444u, 0x4042));
                    // sums: (1024 + ua) + (1024 + uc) - 2304 = xa_h + xc_h (signed high bytes);  (4 + la) + (4 + lc) - 8 = la + lc
                    uint32_t v0 = h2_as_u32(__hadd2(ah, __hsub2(ch, __float2half2_rn(2304.0f))));
                    uint32_t v1 = h2_as_u32(__hadd2(al, __hsub2(cl, __float2half2_rn(8.0f))));
                    uint32_t v2 = h2_as_u32(__hsub2(ah, ch)), v3 = h2_as_u32(__hsub2(al, cl));
                    const int j = j0 + 2 * q;
                    const uint32_t keep = (j < P.npN ? 0x0000ffffu : 0u) | (j + 1 < P.npN ? 0xffff0000u : 0u);
                    ph4[q] = v0 & keep; pl4[q] = v1 & keep; mh4[q] = v2 & keep; ml4[q] = v3 & keep;
                }
                const size_t o = umma_off(a, 8 * part, AKB);
                *reinterpret_cast<uint4 *>(sBm + 0 * A_MAT_BYTES + o) = make_uint4(ph4[0], ph4[1], ph4[2], ph4[3]);
                *reinterpret_cast<uint4 *>(sBm + 1 * A_MAT_BYTES + o) = make_uint4(pl4[0], pl4[1], pl4[2], pl4[3]);
                *reinterpret_cast<uint4 *>(sBm + 2 * A_MAT_BYTES + o) = make_uint4(mh4[0], mh4[1], mh4[2], mh4[3]);
                *reinterpret_cast<uint4 *>(sBm + 3 * A_MAT_BYTES + o) = make_uint4(ml4[0], ml4[1], ml4[2], ml4[3]);
            }
            if (kb + 1 < n_kb) prefetch(kb + 1);
            fence_async_smem();
            __syncwarp();
            if (lane == 0) mbar_arrive(&full_b[s]);
        }
        // ---- epilogue: thread = bin row (warps 0..3), add the two accumulators, rotate, store ---------------
        if (warp < 4) {
            mbar_wait(&done, 0);
            tc_fence_after();
            const int row = warp * 32 + lane;
            const uint32_t tl = tmem_base + ((uint32_t)(warp * 32) << 16);
            const float2 rot = P.rot[range * 128 + row];          // e^{-i theta (N-1)/2} x scale
            for (int c0 = 0; c0 < AN; c0 += 16) {          // 16 columns at a time: 17 warps are allocated registers as 20
                float ac[16], as[16], v[16], u[16];
                tmem_ld16_nowait(tl + c0, ac);
                tmem_ld16_nowait(tl + AN + c0, as);
                tmem_ld16_nowait(tl + 2 * AN + c0, v);
                tmem_ld16_nowait(tl + 3 * AN + c0, u);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 16; ++i) { ac[i] += v[i]; as[i] += u[i]; }
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const long long an = ag0 + c0 + i;
                    if (an < anchor_end) {
                        // R = (c0 - i s0)(A - iB)
                        const float rr = rot.x * ac[i] - rot.y * as[i];
                        const float ri = -(rot.x * as[i] + rot.y * ac[i]);
                        anchors[an * (P.n_ranges * 128) + range * 128 + row] = make_float2(rr, ri);
                    }
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ------------------------------------------------------------------------- slide kernel --------
// Persistent CTA = one 128-bin range; work unit = a 32-frame CHAIN: even chains slide forward from the
// anchor at frame 64k (frames 64k .. 64k+31), odd chains slide backward from the anchor at frame 64(k+1)
// (frames 64k+63 .. 64k+32).  The CTA is G identical, independent warp GROUPS (4 warps = 128 threads = the
// 128 TMEM lanes); group g takes every G-th chain of the CTA and runs all stages on it:
//
//   fill      PCM16 -> offset-binary uint16 in shared memory (16-byte vectors prefetched one chain ahead)
//   build     folded differences of the chain's 32 hop blocks -> fp16 hi/lo B operand (UMMA K-major layout)
//   MMA       one elected thread: 30 x tcgen05.mma (M=128 bins, N=32 frames, K=16 pairs); A = the twiddles,
//             resident in TMEM (an SS-mode MMA at N=32 would re-read 4 KB of A per instruction from shared
//             memory, the binding resource of this kernel); B from shared memory; D = the group's accumulator
//   recur     thread = bin: tcgen05.ld the chain's D columns, 32-step recurrence in registers, R parked in
//             the group's shared-memory stage (float2 [128 bins][32 frames])
//   emit      thread = frame: Hann over neighbouring bins, |.|^2, dB, coalesced 128-byte row stores, min/max
//
// The MMA of chain i+1 is issued before the emit stage of chain i, so it runs under it; groups are out of
// phase with each other, which is what overlaps the latency-bound stages (named barriers are group-local,
// there is no CTA-wide barrier after start-up).
constexpr int CF = 32;                      // frames per chain
constexpr int ST_LD = CF + 2;               // float2 per bin row of the R stage: 272 B pitch, conflict-free STS.128
constexpr int ROWS_PER_EWARP = 32;          // emit rows per warp
constexpr int TM_ACC_PER_GROUP = 2 * CF;    // (cos, sin) accumulators of a group
constexpr int NK_T = 5;                     // k-steps whose twiddles fit in TMEM next to the accumulators (any further one is
                                            // read from shared memory, at 4 KB per MMA)
constexpr int TM_COLS = 512;
constexpr int FIX_EVERY = 4;                // recurrence steps between two corrections of the twiddle's rounding error

// group-local barrier of 128 threads.  The id must be a literal: with a register id the compiler reserves all 16
// hardware barriers for the CTA, and a kernel that needs a barrier of its own can then no longer share the SM.
__device__ __forceinline__ void named_bar_sync(int id, int /*nthreads = 128*/) {
    if (id == 1) asm volatile("bar.sync 1, 128;" ::: "memory");
    else if (id == 2) asm volatile("bar.sync 2, 128;" ::: "memory");
    else asm volatile("bar.sync 3, 128;" ::: "memory");
}

// Chain -> segment bookkeeping kept in registers; the descriptor is re-read only when the segment changes.
struct ChainWalk {
    const SegDesc *segs;
    int n_segs, seg_idx, next_chain0;
    SegDesc sd;
    __device__ void init(const SegDesc *s, int n, int start) {
        segs = s; n_segs = n; seg_idx = start; sd = s[start];
        next_chain0 = start + 1 < n ? 2 * s[start + 1].group0 : 0x7fffffff;
    }
    // chains are visited in increasing order
    __device__ __forceinline__ void seek(int chain) {
        while (chain >= next_chain0) {
            ++seg_idx;
            sd = segs[seg_idx];
            next_chain0 = seg_idx + 1 < n_segs ? 2 * segs[seg_idx + 1].group0 : 0x7fffffff;
        }
    }
};

#ifdef NBM_WS_TIMING
// diagnostic build only (scripts/ws_timing.py): cycles per pipeline phase, summed over the warps of CTA 0
__device__ unsigned long long ws_dbg[32];
#define WS_T0() long long t_mark = clock64(); unsigned long long t_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}
#define WS_MARK(i) do { const long long t_now = clock64(); t_acc[i] += (unsigned long long)(t_now - t_mark); t_mark = t_now; } while (0)
#define WS_FLUSH(base) do { if (blockIdx.x == 0 && lane == 0) for (int i_ = 0; i_ < 8; ++i_) atomicAdd(&ws_dbg[(base) + i_], t_acc[i_]); } while (0)
#else
#define WS_T0()
#define WS_MARK(i)
#define WS_FLUSH(base)
#endif

// Warp roles of the CTA (16 warps; registers are allocated in units of 4 warps, so 16 x 128 registers is the
// shape that fills the register file):
//   warps 0 .. 4G-1   G worker groups of 4 warps: recur -> build(next) -> emit
//   warp  4G          MMA issuer for every group (one elected thread)
//   warps 4G+1 .. +G  one fill warp per group: PCM16 of the group's next chain -> shared memory
constexpr int WS_G = 3;
#ifndef NBM_WORKER_WAIT_NS
#define NBM_WORKER_WAIT_NS 32
#endif
#ifndef NBM_MMA_WAIT_NS
#define NBM_MMA_WAIT_NS 256
#endif
constexpr int WS_WORKER_WAIT_NS = NBM_WORKER_WAIT_NS, WS_MMA_WAIT_NS = NBM_MMA_WAIT_NS;   // sleep between polls of an mbarrier
constexpr int WS_THREADS = 32 * (4 * WS_G + 1 + WS_G);
constexpr int PV = 11;                      // 16-byte PCM vectors per fill lane per round (two rounds per chain)

constexpr int PS = 4;                       // shifted-path vectors per fill lane per round (register budget)

// Pixels that a float32 transform cannot deliver within tolerance are FLAGGED here and recomputed in float64 by
// refine_groups_kernel (frontend.cu).  Measured (scripts/fe_outliers.py): the absolute error of a pixel is float32
// rounding, ~2^-24 rms and < ~2e-6 worst, of the LARGEST rectangular-window magnitude |R_t[k]| its bin carried along the
// chain -- the recurrence is an integrator, so what a loud call leaves behind when it sweeps through a bin stays in that
// bin until the next anchor.  In dB this is invisible unless the pixel itself lies ~60 dB below that magnitude (a deep
// null, or a quiet frame right after a loud one).  The recurrence therefore tracks max |R| per bin (one FMNMX3 per step);
// each emit warp takes the largest over its 32 rows and their halo up to the thread's quarter of the chain as its
// reference, and a thread whose minimum falls below the resulting level puts its block of pixels, with that level, on the
// `cand` list (list_entry) and leaves it out of the warp's min/max partial.  When the list is full the block keeps
// its float32 values (and stays in the partial).
struct FlagArgs {
    float rel_db;                           // how far below the reference magnitude a pixel is flagged (negative)
    uint4 *cand;                            // list_entry(pack_group(...), level)
    unsigned int *cand_count;
    unsigned int cand_cap;
};

__device__ __forceinline__ void
slide_ws_body(const TcParams &P, const SegDesc *__restrict__ segs, int n_segs, int seg_begin, int chain_begin, int total_chains,
              const short *__restrict__ pcm, const float2 *__restrict__ anchors,
              float *__restrict__ spec, float2 *__restrict__ chain_mm, const FlagArgs FA) {
    constexpr int G = WS_G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ uint64_t acc_full[G], b_ready[G], s_full[G], s_free[G];
    __shared__ uint32_t tmem_base_s;
    __shared__ float rmax_s[G][4][4];       // [warp = 32 bin rows][quarter of the chain]: largest |R| carried up to there, recur -> emit
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool is_worker = warp < 4 * G, is_mma_warp = warp == 4 * G;
    const int g = is_worker ? warp >> 2 : (is_mma_warp ? 0 : warp - 4 * G - 1);    // group served
    const int wq = warp & 3, gt = tid & 127;                       // worker: warp in group (= TMEM lane quarter), thread in group
    const int range = blockIdx.x % P.n_ranges;
    const int q0 = chain_begin + blockIdx.x / P.n_ranges, qstride = gridDim.x / P.n_ranges;   // chains [chain_begin, total_chains)

    const int KP = P.KP, nk = P.nk, hop = P.hop, N = P.N;
    const size_t b_mat = (size_t)CF * KP * 2;
    const size_t grp_bytes = 4 * b_mat + (size_t)128 * ST_LD * 8 + (size_t)P.buf_len * 2;
    unsigned char *gbase = smem_raw + (size_t)g * grp_bytes;
    unsigned char *sB = gbase;                                           // 4 x [32 x KP] fp16
    float2 *stage = reinterpret_cast<float2 *>(gbase + 4 * b_mat);       // [128][ST_LD]
    uint16_t *buf16 = reinterpret_cast<uint16_t *>(stage + 128 * ST_LD); // samples of one chain, offset binary
    unsigned char *sAt = smem_raw + (size_t)G * grp_bytes;               // twiddles of the k-steps beyond NK_T: 4 x [128 x 16]

    if (is_worker) {   // B zeroed once: the K padding columns are never written again
        uint4 *db = reinterpret_cast<uint4 *>(sB);
        for (int i = gt; i < (int)(4 * b_mat / 16); i += 128) db[i] = make_uint4(0, 0, 0, 0);
    }
    if (warp == 0) tmem_alloc(&tmem_base_s, TM_COLS);
    if (tid == 0) {
#pragma unroll
        for (int i = 0; i < G; ++i) {
            mbar_init(&acc_full[i], 1); mbar_init(&b_ready[i], 4);
            mbar_init(&s_full[i], 1); mbar_init(&s_free[i], 4);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fence_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_s;
    const uint32_t tmem_a = tmem_base + G * TM_ACC_PER_GROUP;      // 4 matrices x NK_T x 8 columns
    // ---- twiddles: thread = bin row.  k-steps < NK_T into the row's TMEM lane (16 fp16 = 8 columns per store),
    //      the rest into shared memory in UMMA K-major layout ---------------------------------------------
    if (warp < 4) {
        const uint4 *grow = reinterpret_cast<const uint4 *>(
            reinterpret_cast<const unsigned char *>(P.a_slide) + ((size_t)range * 128 + gt) * 4 * KP * 2);
        const uint32_t tl = tmem_a + ((uint32_t)(wq * 32) << 16);
        for (int m = 0; m < 4; ++m)
            for (int kk = 0; kk < nk; ++kk) {
                const uint4 lo = __ldg(grow + (m * KP + kk * 16) / 8), hi = __ldg(grow + (m * KP + kk * 16) / 8 + 1);
                if (kk < NK_T) {
                    tmem_st8(tl + (uint32_t)(m * (NK_T * 8) + kk * 8), lo, hi);
                } else {
                    unsigned char *d = sAt + (size_t)m * 128 * 16 * 2 * (nk - NK_T);
                    *reinterpret_cast<uint4 *>(d + umma_off(gt, (kk - NK_T) * 16, 16 * (nk - NK_T))) = lo;
                    *reinterpret_cast<uint4 *>(d + umma_off(gt, (kk - NK_T) * 16 + 8, 16 * (nk - NK_T))) = hi;
                }
            }
        tmem_st_wait();
        fence_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // chains of group g: q0 + (j * G + g) * qstride
    const int first = q0 + g * qstride, cstride = G * qstride;
    const int n_iters = first < total_chains ? (total_chains - first + cstride - 1) / cstride : 0;
    const uint32_t idesc = make_idesc(128, CF);
    const uint32_t SBO = (uint32_t)(KP / 8) * 128;
    if (is_mma_warp) {
        // ================================ MMA issuer (all groups, round robin) =========================
        // b_ready[g] (4 warp arrivals) says: B operand of group g's next chain is in shared memory AND the group
        // has pulled its previous accumulator into registers.
        WS_T0();
        for (int it = 0; it < n_iters; ++it) {          // n_iters of group 0 is the largest
#pragma unroll 1
            for (int gg = 0; gg < G; ++gg) {
                if (q0 + (it * G + gg) * qstride >= total_chains) break;
                mbar_wait_long<WS_MMA_WAIT_NS>(&b_ready[gg], it & 1);
                tc_fence_after();
                WS_MARK(0);
                if (elect_one()) {
                    const uint32_t b0 = smem_u32(smem_raw + (size_t)gg * grp_bytes);
                    const uint64_t dph = make_desc(b0 + 0 * (uint32_t)b_mat, 128, SBO), dpl = make_desc(b0 + 1 * (uint32_t)b_mat, 128, SBO);
                    const uint64_t dmh = make_desc(b0 + 2 * (uint32_t)b_mat, 128, SBO), dml = make_desc(b0 + 3 * (uint32_t)b_mat, 128, SBO);
                    // fp32 accumulation in TMEM truncates, so order matters: all the small products (2^-8 .. 2^-19 of
                    // the hi x hi ones) are summed first, while the accumulator is still small and they keep their low
                    // bits; the integer-valued hi x hi products are added last (5 roundings instead of 40).
                    const uint32_t d_c = tmem_base + (uint32_t)gg * TM_ACC_PER_GROUP, d_s = d_c + CF;
                    // Fully unrolled (nk <= NK_T is a condition of tc_plan_create, so every twiddle block is in TMEM): rolled,
                    // each MMA cost ~45 cycles of dependent uniform-datapath address arithmetic and branches, and this one
                    // warp, which serves all three groups, was busy 81 % of the time -- the groups queued for it.
                    constexpr uint32_t mt = NK_T * 8;                   // cos_h, cos_l, sin_h, sin_l at ka + {0,1,2,3} mt
#pragma unroll
                    for (int kk = 0; kk < NK_T; ++kk) {
                        if (kk < nk) {
                            const uint64_t ko = (uint64_t)(kk * 16);    // 256 B per k-step in the 16-byte address field
                            const uint32_t ka = tmem_a + kk * 8;
                            if (kk == 0) {
                                umma_f16_ts_c<false>(d_c, ka + 1 * mt, dpl + ko, idesc);
                                umma_f16_ts_c<false>(d_s, ka + 3 * mt, dml + ko, idesc);
                            } else {
                                umma_f16_ts_c<true>(d_c, ka + 1 * mt, dpl + ko, idesc);
                                umma_f16_ts_c<true>(d_s, ka + 3 * mt, dml + ko, idesc);
                            }
                            umma_f16_ts_c<true>(d_c, ka + 1 * mt, dph + ko, idesc);
                            umma_f16_ts_c<true>(d_s, ka + 3 * mt, dmh + ko, idesc);
                            umma_f16_ts_c<true>(d_c, ka + 0 * mt, dpl + ko, idesc);
                            umma_f16_ts_c<true>(d_s, ka + 2 * mt, dml + ko, idesc);
                        }
                    }
#pragma unroll
                    for (int kk = 0; kk < NK_T; ++kk) {
                        if (kk < nk) {
                            const uint64_t ko = (uint64_t)(kk * 16);
                            const uint32_t ka = tmem_a + kk * 8;
                            umma_f16_ts_c<true>(d_c, ka + 0 * mt, dph + ko, idesc);
                            umma_f16_ts_c<true>(d_s, ka + 2 * mt, dmh + ko, idesc);
                        }
                    }
                    umma_commit(&acc_full[gg]);
                }
                __syncwarp();
                WS_MARK(1);
            }
        }
        WS_FLUSH(8);
    } else if (!is_worker) {
        // ================================ fill warp of group g ============================================
        // Chain j's samples: 16-byte global loads issued first (they fly while the group still reads chain j-1's
        // samples), then, once the group has released the buffer, flipped to offset binary and stored.
        WS_T0();
        ChainWalk cw;
        cw.init(segs, n_segs, seg_begin);
        const int nv = P.buf_len / 8;
        // pull the PCM of a chain this far ahead into L2, so the register loads below see L2, not HBM, latency
        auto l2_prefetch = [&](int chain) {
            ChainWalk w2 = cw;
            w2.seek(chain);
            const int t0 = (chain - 2 * w2.sd.group0) * CF;
            long long a = (long long)t0 * hop - N / 2 - (PADF + P.off), b = a + P.buf_len;
            a = a < 0 ? 0 : a;
            b = b > w2.sd.n_samples ? w2.sd.n_samples : b;
            const long long ga = (w2.sd.pcm_start + a + 7) & ~7ll, gb = (w2.sd.pcm_start + b) & ~7ll;      // 16-byte granules
            if (gb > ga && elect_one())
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(pcm + ga), "r"((uint32_t)((gb - ga) * 2)) : "memory");
        };
        constexpr int L2_AHEAD = 3;
        for (int j = 1; j < L2_AHEAD && j < n_iters; ++j) l2_prefetch(first + j * cstride);
        for (int it = 0; it < n_iters; ++it) {
            const int chain = first + it * cstride;
            if (it + L2_AHEAD < n_iters) l2_prefetch(chain + L2_AHEAD * cstride);
            cw.seek(chain);
            const int t0 = (chain - 2 * cw.sd.group0) * CF;
            const long long s0 = (long long)t0 * hop - N / 2 - (PADF + P.off);
            const long long pcm0 = cw.sd.pcm_start, ns = cw.sd.n_samples;
            const long long g0 = pcm0 + s0;
            const bool fast = (reinterpret_cast<uintptr_t>(pcm) & 15) == 0;
            const int sh = (int)(g0 & 7);           // the chain starts sh samples after a 16-byte boundary of the batch buffer
            // vector v covers samples s0 + 8v .. +7 and is assembled from the aligned 16-byte loads at s0 - sh + 8v
            // (and the next one when sh != 0): fast iff those lie inside the segment, v_lo <= v < v_hi
            const long long a0 = s0 - sh;
            const int v_lo = a0 >= 0 ? 0 : (int)((-a0 + 7) >> 3);
            const long long room = ns - a0 - (sh ? 8 : 0);
            const int v_hi = room <= 0 ? 0 : (int)(room >> 3 < nv ? room >> 3 : nv);
            auto slow_vector = [&](int v) {         // straddles or lies outside the segment: centre padding (first / last chains)
                const long long sv = s0 + 8 * v;
                uint32_t u[8];
#pragma unroll
                for (int el = 0; el < 8; ++el)
                    u[el] = (sv + el >= 0 && sv + el < ns) ? ((uint32_t)(uint16_t)__ldg(pcm + pcm0 + sv + el) ^ 0x8000u) : 0x8000u;
                reinterpret_cast<uint4 *>(buf16)[v] =
                    make_uint4(u[0] | (u[1] << 16), u[2] | (u[3] << 16), u[4] | (u[5] << 16), u[6] | (u[7] << 16));
            };
            if (fast && sh == 0) {
                for (int vb = 0; vb < nv; vb += 32 * PV) {
                    int4 pre[PV];
                    const int4 *src = reinterpret_cast<const int4 *>(pcm + g0) + vb + lane;
#pragma unroll
                    for (int e = 0; e < PV; ++e) {
                        const int v = vb + lane + 32 * e;
                        if (v >= v_lo && v < v_hi) pre[e] = __ldg(src + 32 * e);
                    }
                    WS_MARK(0);
                    if (vb == 0 && it > 0) mbar_wait_long<2048>(&s_free[g], (it - 1) & 1);   // every warp is done with chain it-1's samples
                    WS_MARK(1);
#pragma unroll
                    for (int e = 0; e < PV; ++e) {
                        const int v = vb + lane + 32 * e;
                        if (v >= v_lo && v < v_hi)
                            reinterpret_cast<uint4 *>(buf16)[v] =
                                make_uint4((uint32_t)pre[e].x ^ 0x80008000u, (uint32_t)pre[e].y ^ 0x80008000u,
                                           (uint32_t)pre[e].z ^ 0x80008000u, (uint32_t)pre[e].w ^ 0x80008000u);
                    }
                    if (v_lo > 0 || v_hi < nv) {    // first / last chains of a segment
#pragma unroll 1
                        for (int v = vb + lane; v < min(nv, vb + 32 * PV); v += 32)
                            if (!(v >= v_lo && v < v_hi)) slow_vector(v);
                    }
                }
            } else if (fast) {
                // the file does not start on a 16-byte boundary of the batch buffer: aligned loads, re-aligned in registers
                for (int vb = 0; vb < nv; vb += 32 * PS) {
                    int4 pa[PS], pb[PS];
                    const int4 *src = reinterpret_cast<const int4 *>(pcm + g0 - sh) + vb + lane;
#pragma unroll
                    for (int e = 0; e < PS; ++e) {
                        const int v = vb + lane + 32 * e;
                        if (v >= v_lo && v < v_hi) { pa[e] = __ldg(src + 32 * e); pb[e] = __ldg(src + 32 * e + 1); }
                    }
                    if (vb == 0 && it > 0) mbar_wait_long<2048>(&s_free[g], (it - 1) & 1);
#pragma unroll
                    for (int e = 0; e < PS; ++e) {
                        const int v = vb + lane + 32 * e;
                        if (v >= v_lo && v < v_hi) {
                            const uint32_t W[8] = {(uint32_t)pa[e].x, (uint32_t)pa[e].y, (uint32_t)pa[e].z, (uint32_t)pa[e].w,
                                                   (uint32_t)pb[e].x, (uint32_t)pb[e].y, (uint32_t)pb[e].z, (uint32_t)pb[e].w};
                            uint32_t o[4];
                            realign_words<8, 4>(W, sh, o);
                            reinterpret_cast<uint4 *>(buf16)[v] =
                                make_uint4(o[0] ^ 0x80008000u, o[1] ^ 0x80008000u, o[2] ^ 0x80008000u, o[3] ^ 0x80008000u);
                        }
                    }
                    if (v_lo > 0 || v_hi < nv) {
#pragma unroll 1
                        for (int v = vb + lane; v < min(nv, vb + 32 * PS); v += 32)
                            if (!(v >= v_lo && v < v_hi)) slow_vector(v);
                    }
                }
            } else {                                // the batch buffer itself is not 16-byte aligned
                if (it > 0) mbar_wait_long<2048>(&s_free[g], (it - 1) & 1);
                for (int i = lane; i < P.buf_len; i += 32) {
                    const long long sx = s0 + i;
                    buf16[i] = (sx >= 0 && sx < ns) ? (uint16_t)((uint16_t)__ldg(pcm + pcm0 + sx) ^ 0x8000u) : (uint16_t)0x8000u;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&s_full[g]);
            WS_MARK(2);
        }
        WS_FLUSH(16);
    } else {
    // ================================ worker group g ======================================================
    ChainWalk cw;
    cw.init(segs, n_segs, seg_begin);
    const int bar_id = 1 + g;
    const int half_hop = hop / 2;
    const int njg = (P.npH + 7) / 8;
    const int n = gt & 31;                                         // build: frame column of this thread
    const int base = PADF + P.off + n * hop;
    const uint32_t acc_col = tmem_base + (uint32_t)g * TM_ACC_PER_GROUP;
    const uint32_t lane_sel = (uint32_t)(wq * 32) << 16;
    const float2 cf = P.cf[range * 128 + gt], cfx = P.cfix[range * 128 + gt];
    const float2 gF = P.gf[range * 128 + gt], gB = P.gb[range * 128 + gt];
    const int r_lo = 1 + ROWS_PER_EWARP * wq;
    // rows inside the band (only the last range is cut short)
    const int r_hi = min(min(r_lo + ROWS_PER_EWARP, 127), P.n_bins - (range * BINS_PER_RANGE - 1));
    const float floor_pw = 16.0f * P.min_level_sq;

    // Byte-plane split.  With u = x + 32768 = 256 uh + ul (offset binary, the offsets cancel in differences),
    //   ep = (a - b) + (c - d),  em = (a - b) - (c - d)     (a, b: new/old sample above the block centre; c, d: below)
    // are formed separately on the high and the low bytes in packed fp16 arithmetic: a byte b becomes the half
    // 0x6400 | b = 1024 + b (high plane) or 0x4400 | b = 4 + b/256 (low plane, pre-scaled), so every difference and
    // sum is an exactly representable small number.  B = (ep_hi, ep_lo/256, em_hi, em_lo/256), value ep/256 = hi + lo.
    auto planes = [&](uint32_t wa, uint32_t wb, uint32_t wc, uint32_t we, uint32_t &ph, uint32_t &pl, uint32_t &mh, uint32_t &ml) {
        // wa, wb: samples (2q, 2q+1) above the centre (new, old), bytes [l0 h0 l1 h1]
        const __half2 ah = u32_as_h2(__byte_perm(wa, 0x64646464u, 0x4341)), al = u32_as_h2(__byte_perm(wa, 0x44444444u, 0x4240));
        const __half2 bh = u32_as_h2(__byte_perm(wb, 0x64646464u, 0x4341)), bl = u32_as_h2(__byte_perm(wb, 0x44444444u, 0x4240));
        // wc, we: the mirrored samples below the centre, stored ascending: half 0 <- sample 1 of the word, half 1 <- sample 0
        const __half2 ch = u32_as_h2(__byte_perm(wc, 0x64646464u, 0x4143)), cl = u32_as_h2(__byte_perm(wc, 0x44444444u, 0x4042));
        const __half2 eh = u32_as_h2(__byte_perm(we, 0x64646464u, 0x4143)), el = u32_as_h2(__byte_perm(we, 0x44444444u, 0x4042));
        const __half2 dhh = __hsub2(ah, bh), dhl = __hsub2(al, bl), dlh = __hsub2(ch, eh), dll = __hsub2(cl, el);
        ph = h2_as_u32(__hadd2(dhh, dlh)); pl = h2_as_u32(__hadd2(dhl, dll));
        mh = h2_as_u32(__hsub2(dhh, dlh)); ml = h2_as_u32(__hsub2(dhl, dll));
    };
    const bool new_al8 = (N & 3) == 0;      // N % 4 == 2 (e.g. n_fft 4410): the NEW samples sit 4-byte, not 8-byte, aligned
    auto build_unit = [&](int jg) {         // thread = (frame n, pairs 8 jg .. 8 jg + 7), all live
        // 8 consecutive samples = two 8-byte loads (frames are 8-byte, not 16-byte, aligned)
        const uint2 *nh = reinterpret_cast<const uint2 *>(buf16 + base + N + half_hop + 8 * jg);
        const uint2 *oh = reinterpret_cast<const uint2 *>(buf16 + base + half_hop + 8 * jg);
        const uint2 *nl = reinterpret_cast<const uint2 *>(buf16 + base + N + half_hop - 8 - 8 * jg);
        const uint2 *ol = reinterpret_cast<const uint2 *>(buf16 + base + half_hop - 8 - 8 * jg);
        uint2 a0, a1, c0, c1;
        if (new_al8) {
            a0 = nh[0]; a1 = nh[1]; c0 = nl[0]; c1 = nl[1];
        } else {
            const uint32_t *nh4 = reinterpret_cast<const uint32_t *>(nh), *nl4 = reinterpret_cast<const uint32_t *>(nl);
            a0 = make_uint2(nh4[0], nh4[1]); a1 = make_uint2(nh4[2], nh4[3]);
            c0 = make_uint2(nl4[0], nl4[1]); c1 = make_uint2(nl4[2], nl4[3]);
        }
        const uint2 b0 = oh[0], b1 = oh[1], e0 = ol[0], e1 = ol[1];
        const uint32_t wa[4] = {a0.x, a0.y, a1.x, a1.y}, wb[4] = {b0.x, b0.y, b1.x, b1.y};
        // the lo sample of pair jj sits at position 7 - jj of the ascending vector: words in reverse
        const uint32_t wc[4] = {c1.y, c1.x, c0.y, c0.x}, we[4] = {e1.y, e1.x, e0.y, e0.x};
        uint32_t ph[4], pl[4], mh[4], ml[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) planes(wa[q], wb[q], wc[q], we[q], ph[q], pl[q], mh[q], ml[q]);
        const size_t o = umma_off(n, jg * 8, KP);
        *reinterpret_cast<uint4 *>(sB + 0 * b_mat + o) = make_uint4(ph[0], ph[1], ph[2], ph[3]);
        *reinterpret_cast<uint4 *>(sB + 1 * b_mat + o) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
        *reinterpret_cast<uint4 *>(sB + 2 * b_mat + o) = make_uint4(mh[0], mh[1], mh[2], mh[3]);
        *reinterpret_cast<uint4 *>(sB + 3 * b_mat + o) = make_uint4(ml[0], ml[1], ml[2], ml[3]);
    };
    auto build_tail = [&](int jg) {         // the last, partial group of pairs (npH % 8 in {2, 4, 6}), word by word
        const uint32_t *nh = reinterpret_cast<const uint32_t *>(buf16 + base + N + half_hop + 8 * jg);
        const uint32_t *oh = reinterpret_cast<const uint32_t *>(buf16 + base + half_hop + 8 * jg);
        const uint32_t *nl = reinterpret_cast<const uint32_t *>(buf16 + base + N + half_hop - 8 - 8 * jg);
        const uint32_t *ol = reinterpret_cast<const uint32_t *>(buf16 + base + half_hop - 8 - 8 * jg);
        const size_t o = umma_off(n, jg * 8, KP);
        for (int q = 0; 8 * jg + 2 * q < P.npH; ++q) {
            uint32_t ph, pl, mh, ml;
            planes(nh[q], oh[q], nl[3 - q], ol[3 - q], ph, pl, mh, ml);
            *reinterpret_cast<uint32_t *>(sB + 0 * b_mat + o + 4 * q) = ph;
            *reinterpret_cast<uint32_t *>(sB + 1 * b_mat + o + 4 * q) = pl;
            *reinterpret_cast<uint32_t *>(sB + 2 * b_mat + o + 4 * q) = mh;
            *reinterpret_cast<uint32_t *>(sB + 3 * b_mat + o + 4 * q) = ml;
        }
    };
    // B operand of the group's chain `it`; hands the operand to the MMA warp and the samples back to the fill warp
    auto build = [&](int it) {
        mbar_wait_long<WS_WORKER_WAIT_NS>(&s_full[g], it & 1);
        const int n_full = P.npH / 8;                               // units with all 8 pairs live
        for (int jg = wq; jg < n_full; jg += 4) build_unit(jg);
        if (n_full < njg && wq == (it & 3)) build_tail(n_full);
        fence_async_smem();
        __syncwarp();
        if (lane == 0) { mbar_arrive(&b_ready[g]); mbar_arrive(&s_free[g]); }
    };
    auto anchor_of = [&](int chain) {       // the chain's anchor for this thread's bin (look-ahead walk)
        ChainWalk w2 = cw;
        w2.seek(chain);
        const int lc = chain - 2 * w2.sd.group0;
        return __ldg(anchors + ((long long)w2.sd.group0 + w2.seg_idx + ((lc + 1) >> 1)) * (P.n_ranges * 128) + range * 128 + gt);
    };

    float2 anc_next = make_float2(0.f, 0.f);
    if (n_iters > 0) {
        anc_next = anchor_of(first);
        build(0);
    }
    WS_T0();
    for (int it = 0; it < n_iters; ++it) {
        const int chain = first + it * cstride;
        cw.seek(chain);
        const int lc = chain - 2 * cw.sd.group0;
        const int t0 = lc * CF;
        const bool fwd = (lc & 1) == 0;
        const float2 anc = anc_next;
        if (it + 1 < n_iters) anc_next = anchor_of(chain + cstride);
        WS_MARK(0);
        // ---- recur: accumulators -> registers -> 32-step recurrence -> stage, 16 columns at a time ----------
        {
            mbar_wait_long<WS_WORKER_WAIT_NS>(&acc_full[g], it & 1);
            WS_MARK(1);
            tc_fence_after();
            float4 *st4 = reinterpret_cast<float4 *>(stage + (size_t)gt * ST_LD);
            float Rr = anc.x, Ri = anc.y;
            float pr = Rr, pi = Ri;
            float mR = fmaxf(fabsf(Rr), fabsf(Ri));                          // largest |R| (inf-norm) this bin carries along the chain
#pragma unroll 1
            for (int hh = 0; hh < 2; ++hh) {          // not unrolled: the straight-line recurrence is instruction-fetch bound
                const int c0 = fwd ? 16 * hh : 16 - 16 * hh;             // forward chains walk the columns up, backward ones down
                float gc[16], gs[16];
                tmem_ld16_nowait(acc_col + lane_sel + c0, gc);
                tmem_ld16_nowait(acc_col + lane_sel + CF + c0, gs);
                tmem_ld_wait();
                if (fwd) {
                    // frames t0 .. t0+31 ; column c+1 from column c with D_c
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const int c = c0 + i;
                        if (c == CF - 1) break;
                        const float gr = gF.x * gc[i] + gF.y * gs[i];
                        const float gi = gF.y * gc[i] - gF.x * gs[i];
                        float nr = fmaf(cf.x, Rr, fmaf(-cf.y, Ri, gr));
                        float ni = fmaf(cf.x, Ri, fmaf(cf.y, Rr, gi));
#ifndef NBM_NO_CFIX
                        if ((i & (FIX_EVERY - 1)) == FIX_EVERY - 1)
#else
                        if (false)
#endif
                        {          // R *= rho^FIX_EVERY = 1 + cfx
                            const float tr = fmaf(cfx.x, nr, fmaf(-cfx.y, ni, nr));
                            ni = fmaf(cfx.x, ni, fmaf(cfx.y, nr, ni));
                            nr = tr;
                        }
                        Rr = nr; Ri = ni;
#ifndef NBM_NO_FLAG
                        mR = fmaxf(mR, fmaxf(fabsf(nr), fabsf(ni)));
#endif
                        if (c & 1) { pr = Rr; pi = Ri; }                       // column c+1 even: first of a pair
                        else st4[c >> 1] = make_float4(pr, pi, Rr, Ri);       // columns (c, c+1)
                        if (i == 7) {                                          // 8 (24) steps from the anchor
                            const unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(mR));   // |R| >= 0: uint order
                            if (lane == 0) rmax_s[g][wq][2 * hh] = __uint_as_float(m);
                        }
                    }
                } else {
                    // frames t0+31 down to t0 ; column c from column c+1 with D_c ; column 32 is the anchor
#pragma unroll
                    for (int i = 15; i >= 0; --i) {
                        const int c = c0 + i;
                        const float gr = gB.y * gs[i] - gB.x * gc[i];
                        const float gi = gB.x * gs[i] + gB.y * gc[i];
                        float nr = fmaf(cf.x, Rr, fmaf(cf.y, Ri, gr));
                        float ni = fmaf(cf.x, Ri, fmaf(-cf.y, Rr, gi));
#ifndef NBM_NO_CFIX
                        if ((i & (FIX_EVERY - 1)) == 0)
#else
                        if (false)
#endif
                        {                      // R *= conj(rho)^FIX_EVERY
                            const float tr = fmaf(cfx.x, nr, fmaf(cfx.y, ni, nr));
                            ni = fmaf(cfx.x, ni, fmaf(-cfx.y, nr, ni));
                            nr = tr;
                        }
                        Rr = nr; Ri = ni;
#ifndef NBM_NO_FLAG
                        mR = fmaxf(mR, fmaxf(fabsf(nr), fabsf(ni)));
#endif
                        if (c & 1) { pr = Rr; pi = Ri; }                       // odd column: second of a pair
                        else st4[c >> 1] = make_float4(Rr, Ri, pr, pi);       // columns (c, c+1)
                        if (i == 8) {
                            const unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(mR));
                            if (lane == 0) rmax_s[g][wq][2 * hh] = __uint_as_float(m);
                        }
                    }
                }
                {                                                              // 16 (32) steps from the anchor
                    const unsigned int m = __reduce_max_sync(0xffffffffu, __float_as_uint(mR));
                    if (lane == 0) rmax_s[g][wq][2 * hh + 1] = __uint_as_float(m);
                }
            }
            tc_fence_before();
        }
        WS_MARK(2);
        named_bar_sync(bar_id, 128);        // stage complete; accumulator and B operand free (MMA(it) done, D loaded)
        WS_MARK(3);
        // ---- next chain's B operand: its MMAs run under this chain's emit stage ------------------------
        if (it + 1 < n_iters) build(it + 1);
        WS_MARK(4);
        // ---- emit: Hann + dB + store.  Warp = 32 bin rows; a thread owns TWO adjacent frames (one LDS.128 /
        //      STG.64 per row) and the half-warps walk the two 16-row halves, so each row store is one 128-byte line
        {
            float vmin = INFINITY, vmax = -INFINITY;
            const int hw = lane >> 4, fc = 2 * (lane & 15);             // half-warp, first of the two frame columns
            const int ra = r_lo + 16 * hw, rb = min(ra + 16, r_hi);
            const float4 *st = reinterpret_cast<const float4 *>(stage + fc);
            constexpr int LD4 = ST_LD / 2;                               // row pitch in float4
            const int nfr = cw.sd.n_frames - (t0 + fc);                 // frames of this thread inside the segment
            if (nfr > 0 && ra < rb) {
                float4 pv = st[(ra - 1) * LD4], cu = st[ra * LD4];
                const int stride = cw.sd.row_stride;
                char *out = reinterpret_cast<char *>(spec + cw.sd.spec_off + (long long)(range * BINS_PER_RANGE + ra - 1) * stride + t0 + fc);
                const long long stride_b = (long long)stride * 4;
                // Flag level of this thread's two frames: 20 log10(|R|max) + rel_db, -inf for a silent chain.  |R|max is the
                // largest magnitude the warp's rows (r_lo .. r_lo + 31 and their halo: recur warps wq and wq + 1) carried
                // from the anchor up to the end of the frames' quarter of the chain -- what comes later on the way cannot
                // have left an error in them.  Column c is c steps from the anchor in a forward chain, 32 - c in a backward one.
                const int qi = fwd ? fc >> 3 : (31 - fc) >> 3;
                const float th = fmaf(fast_log2(fmaxf(rmax_s[g][wq][qi], rmax_s[g][min(wq + 1, 3)][qi])), 6.0205999132796239f, FA.rel_db);
                const bool pair_ok = (cw.sd.spec_off & 1) == 0;          // a later STFT chunk of a long file may start on an odd column
                // 4 X = 2 R[k] - (R[k-1] + R[k+1]);  10 log10(|X|^2) = 10 log10(|4X|^2) - 10 log10(16)
                auto row_db = [&](const float4 &nx, float &db0, float &db1) {
                    const float xr0 = fmaf(2.0f, cu.x, -(pv.x + nx.x)), xi0 = fmaf(2.0f, cu.y, -(pv.y + nx.y));
                    const float xr1 = fmaf(2.0f, cu.z, -(pv.z + nx.z)), xi1 = fmaf(2.0f, cu.w, -(pv.w + nx.w));
                    const float pw0 = fmaxf(fmaf(xr0, xr0, xi0 * xi0), floor_pw);
                    const float pw1 = fmaxf(fmaf(xr1, xr1, xi1 * xi1), floor_pw);
                    db0 = fmaf(fast_log2(pw0), 3.0102999566398120f, -12.041199826559248f);
                    db1 = fmaf(fast_log2(pw1), 3.0102999566398120f, -12.041199826559248f);
                };
                if (nfr > 1 && pair_ok) {
                    // The common case on its own loop: with the width / alignment tests inside the loop the compiler
                    // merged the 8-byte store with the scalar fallback into predicated 4-byte stores (three STG and four
                    // predicated min/max per row instead of one STG.64 and two FMNMX3).
                    // unroll 4 measured best: 2 lacks ILP, 8 (or a fixed-trip fully unrolled loop) loses more to instruction fetch
#pragma unroll 4
                    for (int r = ra; r < rb; ++r) {
                        const float4 nx = st[(r + 1) * LD4];
                        float db0, db1;
                        row_db(nx, db0, db1);
                        *reinterpret_cast<float2 *>(out) = make_float2(db0, db1);
                        vmin = fminf(vmin, fminf(db0, db1));
                        vmax = fmaxf(vmax, fmaxf(db0, db1));
                        out += stride_b;
                        pv = cu; cu = nx;
                    }
                } else {
                    // last (odd) frame column of a segment, or a later STFT chunk of a long file that starts on an odd column
#pragma unroll 1
                    for (int r = ra; r < rb; ++r) {
                        const float4 nx = st[(r + 1) * LD4];
                        float db0, db1;
                        row_db(nx, db0, db1);
                        reinterpret_cast<float *>(out)[0] = db0;
                        vmin = fminf(vmin, db0);
                        vmax = fmaxf(vmax, db0);
                        if (nfr > 1) {
                            reinterpret_cast<float *>(out)[1] = db1;
                            vmin = fminf(vmin, db1);
                            vmax = fmaxf(vmax, db1);
                        }
                        out += stride_b;
                        pv = cu; cu = nx;
                    }
                }
#ifdef NBM_NO_FLAG
                if (false) {
#else
                if (vmin < th) {
#endif
                    // rare: some pixel of this thread lies below the chain's flag level.  The thread's block of pixels
                    // (rb - ra rows x min(nfr, 2) frames) goes on the list as ONE entry -- refine_groups_kernel finds the
                    // flagged pixels in it, recomputes them in float64 and folds the block's exact minimum into the file's --
                    // and the block stays out of this warp's min/max partial.  No loop here: the other three warps of the
                    // group wait for this one at the barrier below.
                    const unsigned int at = atomicAdd(FA.cand_count, 1u);
                    if (at < FA.cand_cap) {
                        FA.cand[at] = list_entry(pack_group(cw.seg_idx, range * BINS_PER_RANGE + ra - 1, rb - ra, nfr > 1 ? 2 : 1, t0 + fc), th);
                        vmin = INFINITY;
                    }
                }
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
                vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
            }
            if (lane == 0) chain_mm[((size_t)chain * P.n_ranges + range) * 4 + wq] = make_float2(vmin, vmax);
        }
        WS_MARK(5);
        named_bar_sync(bar_id, 128);        // stage free for the next recurrence
        WS_MARK(6);
    }
    WS_FLUSH(0);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, TM_COLS);
}

__global__ void __launch_bounds__(WS_THREADS, 1)
slide_ws_kernel(TcParams P, const SegDesc *__restrict__ segs, int n_segs, int seg_begin, int chain_begin, int total_chains,
                const short *__restrict__ pcm, const float2 *__restrict__ anchors,
                float *__restrict__ spec, float2 *__restrict__ chain_mm, FlagArgs FA) {
    slide_ws_body(P, segs, n_segs, seg_begin, chain_begin, total_chains, pcm, anchors, spec, chain_mm, FA);
}

}  // namespace nbm

// ----------------------------------------------------------------------------- host side -------
using namespace nbm;

namespace {
// twiddle tables: 2^11 w = H + L, both at the same scale, so (H, L) x (hi, lo) needs no rescaled operand copy
inline void put_split_scaled(std::vector<__half> &dst, size_t base_elems_h, size_t base_elems_l, size_t off_bytes, double w) {
    const double ws = w * 2048.0;
    const __half h = __float2half_rn((float)ws);
    dst[base_elems_h + off_bytes / 2] = h;
    dst[base_elems_l + off_bytes / 2] = __float2half_rn((float)(ws - (double)__half2float(h)));
}
}  // namespace

int nbm::tc_plan_create(const nbm_frontend_params &p, TcPlan **out) {
    *out = nullptr;
    const int N = p.n_fft, hop = p.hop;
    const int npH = hop / 2, KP = ((npH + 15) / 16) * 16;
    const bool ok = (N % 2 == 0) && (hop % 4 == 0) && KP / 16 <= NK_T && hop >= 8 && p.n_bins <= 3 * BINS_PER_RANGE &&
                    p.low_idx >= 1 && N >= 2 * hop;
    if (!ok) return NBM_ERR_UNSUPPORTED;
    auto *pl = new TcPlan();
    TcParams &k = pl->p;
    k.N = N; k.hop = hop; k.low_idx = p.low_idx; k.n_bins = p.n_bins;
    k.n_ranges = (p.n_bins + BINS_PER_RANGE - 1) / BINS_PER_RANGE;
    k.npH = npH; k.KP = KP; k.nk = KP / 16;
    k.npN = N / 2; k.n_stages = (k.npN + AKB - 1) / AKB;
    k.off = (4 - (npH % 4)) % 4;
    k.buf_len = ((PADF + k.off + CF * hop + N + 16 + 7) / 8) * 8;   // +16: masked tail pairs read past the last block
    k.min_level_sq = (float)(p.min_level * p.min_level);
    const int R = k.n_ranges, N2 = 2 * N;

    const size_t slide_elems = (size_t)R * 4 * 128 * KP;
    const size_t anchor_elems = (size_t)R * k.n_stages * 4 * 128 * AKB;
    std::vector<__half> a_slide(slide_elems, __float2half_rn(0.f)), a_anchor(anchor_elems, __float2half_rn(0.f));
    std::vector<float2> cf(R * 128), gf(R * 128), gb(R * 128), rot(R * 128), cfix(R * 128);
    auto ang = [&](long long q) { return M_PI * (double)(q % N2) / (double)N; };
    const double s18 = 1.0 / 262144.0;        // byte planes (value / 256) x twiddles x 2^11, samples / 32768
    for (int r = 0; r < R; ++r)
        for (int row = 0; row < 128; ++row) {
            const long long kbin = p.low_idx - 1 + (long long)r * BINS_PER_RANGE + row;
            // slide twiddles: pair j <-> u = j + 1/2 ; [range][row][cos_h, cos_l, sin_h, sin_l][KP], row-contiguous
            // because each thread copies its own row into its TMEM lane
            const size_t rbase = ((size_t)r * 128 + row) * 4 * KP;
            for (int j = 0; j < npH; ++j) {
                const double a = ang(kbin * (2 * j + 1));
                put_split_scaled(a_slide, rbase + 0 * KP, rbase + 1 * KP, (size_t)j * 2, cos(a));
                put_split_scaled(a_slide, rbase + 2 * KP, rbase + 3 * KP, (size_t)j * 2, sin(a));
            }
            const size_t amat = (size_t)128 * AKB;
            for (int j = 0; j < k.npN; ++j) {
                const double a = ang(kbin * (2 * j + 1));
                const int st = j / AKB, jj = j % AKB;
                const size_t base = ((size_t)r * k.n_stages + st) * 4 * amat;
                const size_t o = umma_off(row, jj, AKB);
                put_split_scaled(a_anchor, base + 0 * amat, base + 1 * amat, o, cos(a));
                put_split_scaled(a_anchor, base + 2 * amat, base + 3 * amat, o, sin(a));
            }
            const int i = r * 128 + row;
            double a = ang(kbin * 2 * hop);
            cf[i] = make_float2((float)cos(a), (float)sin(a));
            {
                // the float32 twiddle is off by rho = cf_true / cf_f32 in EVERY step of the recurrence, a systematic error
                // that grows linearly along the chain; every FIX_EVERY steps the recurrence multiplies by rho^FIX_EVERY
                const double cr = (double)cf[i].x, ci = (double)cf[i].y, d2 = cr * cr + ci * ci;
                double pr = (cos(a) * cr + sin(a) * ci) / d2, pi = (sin(a) * cr - cos(a) * ci) / d2;      // rho
                double qr = 1.0, qi = 0.0;
                for (int e = 0; e < FIX_EVERY; ++e) { const double t = qr * pr - qi * pi; qi = qr * pi + qi * pr; qr = t; }
                cfix[i] = make_float2((float)(qr - 1.0), (float)qi);
            }
            a = ang(kbin * (hop + 1));
            gf[i] = make_float2((float)(cos(a) * s18), (float)(sin(a) * s18));
            a = ang(kbin * (hop - 1));
            gb[i] = make_float2((float)(cos(a) * s18), (float)(sin(a) * s18));
            a = ang(kbin * (N - 1));
            rot[i] = make_float2((float)(cos(a) * s18), (float)(sin(a) * s18));
        }
    const size_t b_slide = align_up(slide_elems * 2, 256), b_anchor = align_up(anchor_elems * 2, 256);
    const size_t b_c = align_up((size_t)R * 128 * sizeof(float2), 256);
    const size_t total = b_slide + b_anchor + 5 * b_c;
    cudaError_t e = cudaMalloc(&pl->d_blob, total);
    if (e != cudaSuccess) { delete pl; return cuda_fail(e, "cudaMalloc(tc tables)"); }
    unsigned char *d = reinterpret_cast<unsigned char *>(pl->d_blob);
    e = cudaMemcpy(d, a_slide.data(), slide_elems * 2, cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaMemcpy(d + b_slide, a_anchor.data(), anchor_elems * 2, cudaMemcpyHostToDevice);
    const float2 *src[5] = {cf.data(), gf.data(), gb.data(), rot.data(), cfix.data()};
    for (int i = 0; i < 5 && e == cudaSuccess; ++i)
        e = cudaMemcpy(d + b_slide + b_anchor + i * b_c, src[i], (size_t)R * 128 * sizeof(float2), cudaMemcpyHostToDevice);
    k.a_slide = reinterpret_cast<const __half *>(d);
    k.a_anchor = reinterpret_cast<const __half *>(d + b_slide);
    k.cf = reinterpret_cast<const float2 *>(d + b_slide + b_anchor);
    k.gf = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + b_c);
    k.gb = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + 2 * b_c);
    k.rot = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + 3 * b_c);
    k.cfix = reinterpret_cast<const float2 *>(d + b_slide + b_anchor + 4 * b_c);
    pl->smem_slide = WS_G * ((size_t)4 * CF * KP * 2 + (size_t)128 * ST_LD * 8 + (size_t)k.buf_len * 2) +
                     (size_t)4 * 128 * 16 * 2 * std::max(0, k.nk - NK_T);
    pl->smem_anchor = (size_t)A_STAGES * A_STAGE_BYTES + 128;
    int dev = 0, sms = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    pl->grid_slide = std::max(1, sms / R) * R;
    cudaFuncAttributes fa_s{}, fa_a{};
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa_s, slide_ws_kernel);
    if (e == cudaSuccess) e = cudaFuncGetAttributes(&fa_a, anchor_tc_kernel);
    if (e == cudaSuccess && ((size_t)max_smem < pl->smem_slide + fa_s.sharedSizeBytes ||
                             (size_t)max_smem < pl->smem_anchor + fa_a.sharedSizeBytes)) {
        tc_plan_destroy(pl);
        return NBM_ERR_UNSUPPORTED;
    }
    // per-function attribute (not per plan): allow the device maximum minus the kernel's static shared memory
    if (e == cudaSuccess) e = cudaFuncSetAttribute(slide_ws_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
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

#ifdef NBM_WS_TIMING
extern "C" int nbm_debug_ws_timing(unsigned long long *out, int reset) {
    unsigned long long z[32] = {0};
    if (out && cudaMemcpyFromSymbol(out, ws_dbg, sizeof(z)) != cudaSuccess) return -1;
    if (reset && cudaMemcpyToSymbol(ws_dbg, z, sizeof(z)) != cudaSuccess) return -1;
    return 0;
}
#endif

int nbm::tc_n_ranges(const TcPlan *pl) { return pl->p.n_ranges; }
int nbm::tc_bins_per_range() { return BINS_PER_RANGE; }
int nbm::tc_chain_frames() { return CF; }
int nbm::tc_slots_per_range() { return 4; }
int nbm::tc_bins_per_slot() { return ROWS_PER_EWARP; }

int nbm::tc_launch_anchors(const TcPlan *pl, const SegDesc *d_segs, int seg_lo, int seg_hi, long long anchor_begin,
                           long long anchor_end, const void *d_pcm, void *d_anchors, cudaStream_t stream) {
    const TcParams &k = pl->p;
    if (anchor_end <= anchor_begin) return NBM_OK;
    dim3 ga((unsigned)((anchor_end - anchor_begin + AN - 1) / AN), (unsigned)k.n_ranges);
    anchor_tc_kernel<<<ga, A_THREADS, pl->smem_anchor, stream>>>(k, d_segs, seg_lo, seg_hi, anchor_begin, anchor_end,
                                                                 reinterpret_cast<const short *>(d_pcm),
                                                                 reinterpret_cast<float2 *>(d_anchors));
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}

int nbm::tc_launch_slides(const TcPlan *pl, const SegDesc *d_segs, int n_segs, int seg_begin, int group_begin, int group_end,
                          const void *d_pcm, float *d_spec, float2 *d_tile_mm, const void *d_anchors,
                          float rel_db, uint4 *d_cand, unsigned int *d_cand_count, unsigned int cand_cap, cudaStream_t stream) {
    const TcParams &k = pl->p;
    const int chain_begin = 2 * group_begin, chain_end = 2 * group_end;
    const int grid = std::min(pl->grid_slide, std::max(1, chain_end - chain_begin) * k.n_ranges);
    FlagArgs fa{rel_db, d_cand, d_cand_count, cand_cap};
    slide_ws_kernel<<<(grid / k.n_ranges) * k.n_ranges, WS_THREADS, pl->smem_slide, stream>>>(
        k, d_segs, n_segs, seg_begin, chain_begin, chain_end, reinterpret_cast<const short *>(d_pcm),
        reinterpret_cast<const float2 *>(d_anchors), d_spec, d_tile_mm, fa);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}
