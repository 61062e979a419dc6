// Detector post-processing kernels: anchor/RoI delta decode, clamp, min-size and score filters,
// and the reference's in-order greedy NMS as a bitmask kernel pair.
//
// Replaces nbm_model/nets/util/nets_utils.py:35-59,169-245, ProposalLayer.forward
// (nbm_model/nets/layers.py:226-303), the FastRCNN inference tail (layers.py:688-778) and
// merge_images (nbm_model/run_detection.py:163-249).
//
// Bit-exactness: boxes are integer-valued float32; every IoU operation is a single IEEE
// round-to-nearest op in the reference's order (no FMA contraction), the comparison is
// `iou >= (float)thresh`, NaN never suppresses, and the greedy order is the INPUT order.
#include <cub/cub.cuh>
#include <vector>
#include <climits>

#include "common.cuh"

namespace nbm {

// ------------------------------------------------------------------------------ decode ------
__device__ __forceinline__ float4 decode_one(float4 d, float4 a) {
    // bbox_reg_to_coord, nets_utils.py:171-186; eager PyTorch = one rounding per op
    const float wa = __fadd_rn(__fsub_rn(a.z, a.x), 1.0f);
    const float ha = __fadd_rn(__fsub_rn(a.w, a.y), 1.0f);
    const float xa = __fadd_rn(a.x, __fmul_rn(0.5f, wa));
    const float ya = __fadd_rn(a.y, __fmul_rn(0.5f, ha));
    const float x = __fadd_rn(__fmul_rn(d.x, wa), xa);
    const float y = __fadd_rn(__fmul_rn(d.y, ha), ya);
    const float w = __fmul_rn(expf(d.z), wa);
    const float h = __fmul_rn(expf(d.w), ha);
    const float hw = __fmul_rn(0.5f, w), hh = __fmul_rn(0.5f, h);
    return make_float4(rintf(__fsub_rn(x, hw)), rintf(__fsub_rn(y, hh)),
                       rintf(__fadd_rn(x, hw)), rintf(__fadd_rn(y, hh)));     // half-to-even == torch.round
}

__device__ __forceinline__ float4 clamp_box(float4 b, float clip_w, float clip_h) {
    if (clip_w > 0.f) { b.x = fminf(fmaxf(b.x, 0.f), clip_w - 1.f); b.z = fminf(fmaxf(b.z, 0.f), clip_w - 1.f); }
    if (clip_h > 0.f) { b.y = fminf(fmaxf(b.y, 0.f), clip_h - 1.f); b.w = fminf(fmaxf(b.w, 0.f), clip_h - 1.f); }
    return b;
}

__device__ __forceinline__ bool big_enough(float4 b, float min_size) {
    return (__fadd_rn(__fsub_rn(b.z, b.x), 1.0f) >= min_size) && (__fadd_rn(__fsub_rn(b.w, b.y), 1.0f) >= min_size);
}

__global__ void decode_kernel(const float4 *__restrict__ deltas, const float4 *__restrict__ anchors, int B, int N,
                              int per_image, float clip_w, float clip_h, float min_size,
                              float4 *__restrict__ boxes, uint8_t *__restrict__ valid) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)B * N) return;
    const float4 a = per_image ? anchors[i] : anchors[i % N];
    const float4 b = clamp_box(decode_one(deltas[i], a), clip_w, clip_h);
    boxes[i] = b;
    if (valid) valid[i] = big_enough(b, min_size) ? 1 : 0;
}

// --------------------------------------------------------------------------------- NMS -------
__device__ __forceinline__ bool iou_ge(float4 a, float4 b, float area_a, float area_b, float thresh) {
    // batch_self_overlap, nets_utils.py:193-205
    const float xi = fmaxf(__fadd_rn(__fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)), 1.0f), 0.0f);
    const float yi = fmaxf(__fadd_rn(__fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)), 1.0f), 0.0f);
    const float inter = __fmul_rn(xi, yi);
    const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
    return __fdiv_rn(inter, uni) >= thresh;       // NaN -> false
}

__device__ __forceinline__ float box_area(float4 b) {
    return __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.0f), __fadd_rn(__fsub_rn(b.w, b.y), 1.0f));
}

// mask[b][i][w] bit j' set  <=>  j = 64 w + j' > i, j < n, IoU(i, j) >= thresh.   Upper triangle only.
// Block = one 64 x 64 tile of the matrix, 256 threads: four adjacent lanes share a row and take 16 columns each (the
// serial loop of IEEE divisions per thread is what this kernel's time consists of), their 16-bit pieces are joined by
// two shuffles.
constexpr int NMS_MASK_THREADS = 256;
__global__ void __launch_bounds__(NMS_MASK_THREADS)
nms_mask_kernel(const float4 *__restrict__ boxes, const int *__restrict__ n_valid, const int *__restrict__ n_cap,
                int N, float thresh, unsigned long long *__restrict__ mask, int words) {
    const int b = blockIdx.z, rb = blockIdx.y, cb = blockIdx.x;
    if (cb < rb) return;
    int n = n_valid ? n_valid[b] : N;
    if (n_cap) n = min(n, *n_cap);
    n = min(n, N);
    if (rb * 64 >= n || cb * 64 >= n) return;
    __shared__ float4 cbox[64];
    __shared__ float carea[64];
    const float4 *bx = boxes + (long long)b * N;
    const int t = threadIdx.x;
    const int j0 = cb * 64;
    if (t < 64 && j0 + t < n) { cbox[t] = bx[j0 + t]; carea[t] = box_area(cbox[t]); }
    __syncthreads();
    const int r = t >> 2, q = t & 3;
    const int i = rb * 64 + r;
    unsigned long long bits = 0;
    if (i < n) {
        const float4 me = bx[i];
        const float ma = box_area(me);
        const int lim = min(64, n - j0);
        const int lo = max(16 * q, cb == rb ? r + 1 : 0), hi = min(16 * q + 16, lim);
        for (int jj = lo; jj < hi; ++jj)
            if (iou_ge(cbox[jj], me, carea[jj], ma, thresh)) bits |= 1ull << jj;
    }
    bits |= __shfl_xor_sync(0xffffffffu, bits, 1);
    bits |= __shfl_xor_sync(0xffffffffu, bits, 2);
    if (q == 0 && i < n) mask[((long long)b * N + i) * words + cb] = bits;
}

// Sequential greedy resolution, one block per image, 64 boxes per step.
__global__ void __launch_bounds__(256)
nms_scan_kernel(const unsigned long long *__restrict__ mask, const int *__restrict__ n_valid,
                const int *__restrict__ n_cap, int N, int words, int *__restrict__ keep_idx,
                int *__restrict__ keep_cnt) {
    extern __shared__ unsigned long long sm[];
    unsigned long long *removed = sm;            // [words]
    unsigned long long *kept = sm + words;       // [words]
    __shared__ unsigned long long diag[64];
    __shared__ unsigned long long keepbits;
    __shared__ int total;
    const int b = blockIdx.x, tid = threadIdx.x;
    int n = n_valid ? n_valid[b] : N;
    if (n_cap) n = min(n, *n_cap);
    n = min(n, N);
    const unsigned long long *m = mask + (long long)b * N * words;
    const int chunks = (n + 63) / 64;
    for (int w = tid; w < words; w += blockDim.x) { removed[w] = 0; kept[w] = 0; }
    __syncthreads();
    for (int c = 0; c < chunks; ++c) {
        if (tid < 64) {
            const int i = c * 64 + tid;
            diag[tid] = i < n ? m[(long long)i * words + c] : 0ull;
        }
        __syncthreads();
        if (tid == 0) {
            unsigned long long cur = removed[c], kb = 0;
            const int lim = min(64, n - c * 64);
            for (int t = 0; t < lim; ++t)
                if (!((cur >> t) & 1ull)) { kb |= 1ull << t; cur |= diag[t]; }
            keepbits = kb;
            kept[c] = kb;
        }
        __syncthreads();
        const unsigned long long kb = keepbits;
        for (int w = c + 1 + tid; w < chunks; w += blockDim.x) {
            unsigned long long acc = 0, bits = kb;
            while (bits) {
                const int t = __ffsll((long long)bits) - 1;
                bits &= bits - 1;
                acc |= m[(long long)(c * 64 + t) * words + w];
            }
            removed[w] |= acc;
        }
        __syncthreads();
    }
    // compaction: ascending kept indices
    if (tid == 0) {
        int off = 0;
        for (int c = 0; c < chunks; ++c) { const int pc = __popcll(kept[c]); removed[c] = (unsigned long long)off; off += pc; }
        total = off;
        keep_cnt[b] = off;
    }
    __syncthreads();
    for (int c = tid; c < chunks; c += blockDim.x) {
        unsigned long long bits = kept[c];
        int off = (int)removed[c];
        while (bits) {
            const int t = __ffsll((long long)bits) - 1;
            bits &= bits - 1;
            keep_idx[(long long)b * N + off++] = c * 64 + t;
        }
    }
    for (int i = total + tid; i < N; i += blockDim.x) keep_idx[(long long)b * N + i] = -1;
}

static inline int nms_words(int N) { return (N + 63) / 64; }

// The same resolution for N <= 2048 (32 words) when the image's bit matrix fits in shared memory -- every call of
// the detector (500 proposals, <= 50 final boxes, a file's merge list): the matrix is staged once, then ONE WARP walks
// the chunks with no block barrier.  Lane w owns word w of the removed-mask; the 64-box diagonal is resolved by
// following the alive bits (find-first-set, broadcast read of that row's diagonal word), so its cost is the number of
// boxes KEPT in the chunk, not 64; the kept rows are then OR-ed into the later words lane-parallel from shared memory.
constexpr int NMS_SMALL_THREADS = 128;
__global__ void __launch_bounds__(NMS_SMALL_THREADS)
nms_scan_small_kernel(const unsigned long long *__restrict__ mask, const int *__restrict__ n_valid,
                      const int *__restrict__ n_cap, int N, int words, int *__restrict__ keep_idx,
                      int *__restrict__ keep_cnt) {
    extern __shared__ unsigned long long sm[];           // [n][words], upper triangle
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    int n = n_valid ? n_valid[b] : N;
    if (n_cap) n = min(n, *n_cap);
    n = min(n, N);
    const unsigned long long *m = mask + (long long)b * N * words;
    const int chunks = (n + 63) / 64;
    {
        // stage rows [0, n): words left of the diagonal were never written by nms_mask_kernel and are never read here, so
        // the copy need not tell them apart -- plain 16-byte vectors, four in flight per thread
        const int n16 = n * words / 2;                   // words is even or the matrix is copied word by word below
        if ((words & 1) == 0) {
            const uint4 *src = reinterpret_cast<const uint4 *>(m);
            uint4 *dst = reinterpret_cast<uint4 *>(sm);
            int i = tid;
            for (; i + 3 * NMS_SMALL_THREADS < n16; i += 4 * NMS_SMALL_THREADS) {
                const uint4 v0 = src[i], v1 = src[i + NMS_SMALL_THREADS], v2 = src[i + 2 * NMS_SMALL_THREADS], v3 = src[i + 3 * NMS_SMALL_THREADS];
                dst[i] = v0; dst[i + NMS_SMALL_THREADS] = v1; dst[i + 2 * NMS_SMALL_THREADS] = v2; dst[i + 3 * NMS_SMALL_THREADS] = v3;
            }
            for (; i < n16; i += NMS_SMALL_THREADS) dst[i] = src[i];
        } else {
            for (int i = tid; i < n * words; i += NMS_SMALL_THREADS) sm[i] = m[i];
        }
    }
    __syncthreads();
    if (tid >= 32) return;
    unsigned long long removed = 0, kept = 0;            // word `lane`
    for (int c = 0; c < chunks; ++c) {
        unsigned long long cur = __shfl_sync(0xffffffffu, removed, c);
        const int lim = min(64, n - c * 64);
        if (lim < 64) cur |= ~0ull << lim;               // boxes past the end count as removed
        const unsigned long long *rows = sm + (size_t)c * 64 * words;
        // The 64-box diagonal, in input order: box t is kept iff no kept earlier box of the chunk (nor an earlier chunk)
        // removed it.  Row t only has bits above t, so bit t of `cur` is final once step t - 1 is done and the kept set is
        // simply ~cur at the end.  The diagonal words are fetched ahead of the dependent chain (broadcast reads, 16 at a
        // time); a step of the chain is a bit test and a predicated OR.
        // (rows past `lim` hold stale shared memory, but their own bit in `cur` is set, so they are never OR-ed in)
        const unsigned long long *dg = rows + c;
#pragma unroll
        for (int t0 = 0; t0 < 64; t0 += 16) {            // fully unrolled: bit positions are immediates
            unsigned long long d[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) d[u] = dg[(size_t)(t0 + u) * words];
#pragma unroll
            for (int u = 0; u < 16; ++u)
                if (!((cur >> (t0 + u)) & 1ull)) cur |= d[u];
        }
        const unsigned long long kb = ~cur;
        if (lane == c) kept = kb;
        // OR the kept rows into the later words.  Lane l contributes rows l and l + 32 of the chunk; the OR over the lanes is
        // a warp reduction (redux.sync on the two halves), so a word costs a dozen instructions, not a walk over 64 rows.
        const bool k0 = (kb >> lane) & 1ull, k1 = (kb >> (lane + 32)) & 1ull;
        for (int w = c + 1; w < chunks; ++w) {
            unsigned long long v = (k0 ? rows[(size_t)lane * words + w] : 0ull) | (k1 ? rows[(size_t)(lane + 32) * words + w] : 0ull);
            const unsigned int lo = __reduce_or_sync(0xffffffffu, (unsigned int)v);
            const unsigned int hi = __reduce_or_sync(0xffffffffu, (unsigned int)(v >> 32));
            if (lane == w) removed |= ((unsigned long long)hi << 32) | lo;
        }
    }
    // compaction: ascending kept indices; exclusive prefix of the per-word counts over the lanes
    const int pc = lane < chunks ? __popcll(kept) : 0;
    int off = pc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, off, o);
        if (lane >= o) off += v;
    }
    const int total = __shfl_sync(0xffffffffu, off, 31);
    off -= pc;
    while (kept) {
        const int t = __ffsll((long long)kept) - 1;
        kept &= kept - 1;
        keep_idx[(long long)b * N + off++] = lane * 64 + t;
    }
    if (lane == 0) keep_cnt[b] = total;
    for (int i = total + lane; i < N; i += 32) keep_idx[(long long)b * N + i] = -1;
}
constexpr size_t NMS_SMALL_SMEM_MAX = 200 * 1024;

static int launch_nms(const float *d_boxes, const int *d_n, const int *d_cap, int B, int N, float thresh,
                      int *d_keep_idx, int *d_keep_cnt, void *ws, size_t ws_bytes, cudaStream_t s) {
    const int words = nms_words(N);
    const size_t need = (size_t)B * N * words * sizeof(unsigned long long);
    if (ws_bytes < need) { set_error("nms workspace too small: %zu < %zu", ws_bytes, need); return NBM_ERR_WORKSPACE; }
    dim3 grid(words, words, B);
    nms_mask_kernel<<<grid, NMS_MASK_THREADS, 0, s>>>(reinterpret_cast<const float4 *>(d_boxes), d_n, d_cap, N, thresh,
                                        reinterpret_cast<unsigned long long *>(ws), words);
    const size_t smem_small = (size_t)words * 64 * words * sizeof(unsigned long long);   // whole 64-row chunks: the resolve reads full chunks
    if (words <= 32 && smem_small <= NMS_SMALL_SMEM_MAX) {
        static bool attr_set = false;                   // per function, not per call
        if (!attr_set) {
            NBM_CUDA(cudaFuncSetAttribute(nms_scan_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          (int)NMS_SMALL_SMEM_MAX));
            attr_set = true;
        }
        nms_scan_small_kernel<<<B, NMS_SMALL_THREADS, smem_small, s>>>(reinterpret_cast<const unsigned long long *>(ws), d_n,
                                                                      d_cap, N, words, d_keep_idx, d_keep_cnt);
        NBM_CUDA(cudaGetLastError());
        return NBM_OK;
    }
    const size_t smem = (size_t)2 * words * sizeof(unsigned long long);
    NBM_REQUIRE(smem <= 200 * 1024, "N too large for the NMS scan kernel");
    if (smem > 48 * 1024)
        NBM_CUDA(cudaFuncSetAttribute(nms_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    nms_scan_kernel<<<B, 256, smem, s>>>(reinterpret_cast<const unsigned long long *>(ws), d_n, d_cap, N, words,
                                         d_keep_idx, d_keep_cnt);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}

// --------------------------------------------------------------------------- proposals -------
// scores/deltas from the RPN's [B, C, H, W] layout (layers.py:264-267), decode + clamp + min-size.
// Block = RPN_CELLS consecutive cells of one image.  The RPN tensors are [B, C, H, W]: the 5 A planes a box needs (4 A
// deltas, A foreground scores) are read as rows of RPN_CELLS floats (coalesced) into shared memory; then the boxes are
// decoded and written in the reference's order n = cell * A + a (consecutive threads, consecutive n).
constexpr int RPN_CELLS = 64, RPN_THREADS = 256;
__global__ void __launch_bounds__(RPN_THREADS)
rpn_decode_kernel(const float *__restrict__ cls, const float *__restrict__ reg,
                  const float4 *__restrict__ anchors, int B, int A, int H, int W, float clip_w,
                  float clip_h, float min_size, float4 *__restrict__ boxes, float *__restrict__ keys,
                  int *__restrict__ idx, int *__restrict__ n_valid) {
    extern __shared__ float planes[];                    // [5 A][RPN_CELLS]: reg planes 0 .. 4A-1, then the fg score planes
    const int hw = H * W, N = A * hw;
    const int b = blockIdx.y, c0 = blockIdx.x * RPN_CELLS, nc = min(RPN_CELLS, hw - c0);
    const float *rb = reg + (long long)b * 4 * A * hw + c0;
    const float *cb = cls + (long long)b * 2 * A * hw + c0;
    for (int i = threadIdx.x; i < 5 * A * RPN_CELLS; i += RPN_THREADS) {
        const int pl = i / RPN_CELLS, c = i - pl * RPN_CELLS;
        if (c < nc) planes[i] = pl < 4 * A ? rb[(long long)pl * hw + c] : cb[(long long)(2 * (pl - 4 * A) + 1) * hw + c];
    }
    __syncthreads();
    int ok_cnt = 0;
    for (int l = threadIdx.x; l < nc * A; l += RPN_THREADS) {
        const int c = l / A, a = l - c * A;
        const int n = c0 * A + l;
        const float4 d = make_float4(planes[(4 * a + 0) * RPN_CELLS + c], planes[(4 * a + 1) * RPN_CELLS + c],
                                     planes[(4 * a + 2) * RPN_CELLS + c], planes[(4 * a + 3) * RPN_CELLS + c]);
        const float score = planes[(4 * A + a) * RPN_CELLS + c];
        const float4 bx = clamp_box(decode_one(d, anchors[n]), clip_w, clip_h);
        const long long g = (long long)b * N + n;
        boxes[g] = bx;
        const bool ok = big_enough(bx, min_size);
        keys[g] = ok ? score : -INFINITY;        // invalid boxes sort last (scores are probabilities)
        if (idx) idx[g] = n;
        ok_cnt += ok ? 1 : 0;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ok_cnt += __shfl_xor_sync(0xffffffffu, ok_cnt, o);
    if ((threadIdx.x & 31) == 0 && ok_cnt) atomicAdd(n_valid + b, ok_cnt);
}

// Top-K of one image's scores in stable descending order (ties: lower index first) -- what the reference gets from
// argsort(descending) + filter + [:pre_nms_topN] (layers.py:292-297) -- without sorting all A*H*W candidates: one block
// per image keeps the keys in shared memory, finds the K-th largest by an 8-bit radix select (4 histogram passes), collects
// the keys above it plus the lowest-index ties, and bitonic-sorts those K <= TOPK_MAX.
constexpr int TOPK_THREADS = 1024, TOPK_MAX = 1024;
__global__ void __launch_bounds__(TOPK_THREADS)
proposal_topk_kernel(const float4 *__restrict__ boxes, const float *__restrict__ keys, const int *__restrict__ pre, int N, int cap,
                     float4 *__restrict__ out_boxes, float *__restrict__ out_scores) {
    extern __shared__ unsigned int skey[];               // [N] order-preserving keys
    __shared__ unsigned long long cand[TOPK_MAX];        // (~key) << 32 | index: ascending = score descending, index ascending
    __shared__ unsigned int hist[256];
    __shared__ unsigned int s_prefix, s_want;
    __shared__ unsigned int wsum_gt[32], wsum_eq[32];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int K = min(min(*pre, cap), TOPK_MAX);
    if (K <= 0) return;
    const float *kb = keys + (long long)b * N;
    for (int i = tid; i < N; i += TOPK_THREADS) skey[i] = float_to_ordered(kb[i]);
    if (tid == 0) { s_prefix = 0; s_want = (unsigned int)K; }
    __syncthreads();
    // ---- radix select: after the pass for `shift`, s_prefix holds the top bits of the K-th largest key --------------
    for (int shift = 24; shift >= 0; shift -= 8) {
        for (int i = tid; i < 256; i += TOPK_THREADS) hist[i] = 0;
        __syncthreads();
        const unsigned int prefix = s_prefix, himask = shift == 24 ? 0u : (0xffffffffu << (shift + 8));
        for (int i = tid; i < N; i += TOPK_THREADS) {
            const unsigned int k = skey[i];
            if ((k & himask) == prefix) atomicAdd(&hist[(k >> shift) & 255u], 1u);
        }
        __syncthreads();
        if (warp == 0) {
            // lane l owns bins 255 - 8 l .. 248 - 8 l (descending); the K-th largest lies in the first bin where the running
            // count from the top reaches `want`
            unsigned int h[8], sum = 0;
#pragma unroll
            for (int j = 0; j < 8; ++j) { h[j] = hist[255 - 8 * lane - j]; sum += h[j]; }
            unsigned int incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
            const unsigned int want = s_want;
            unsigned int above = incl - sum;             // keys in higher bins
            const bool here = above < want && incl >= want;
            if (here) {
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    if (above + h[j] >= want) {
                        s_prefix = prefix | ((unsigned int)(255 - 8 * lane - j) << shift);
                        s_want = want - above;           // rank inside that bin
                        break;
                    }
                    above += h[j];
                }
            }
        }
        __syncthreads();
    }
    const unsigned int T = s_prefix, need_eq = s_want;   // K-th largest key; that many of the keys == T belong to the top K
    // ---- collect: thread t scans a contiguous index range, so ranks inside the tie class follow the index order ----------
    const int chunk = (N + TOPK_THREADS - 1) / TOPK_THREADS;
    const int i0 = min(tid * chunk, N), i1 = min(i0 + chunk, N);
    unsigned int n_gt = 0, n_eq = 0;
    for (int i = i0; i < i1; ++i) { const unsigned int k = skey[i]; n_gt += k > T; n_eq += k == T; }
    unsigned int inc_gt = n_gt, inc_eq = n_eq;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int a = __shfl_up_sync(0xffffffffu, inc_gt, o), e = __shfl_up_sync(0xffffffffu, inc_eq, o);
        if (lane >= o) { inc_gt += a; inc_eq += e; }
    }
    if (lane == 31) { wsum_gt[warp] = inc_gt; wsum_eq[warp] = inc_eq; }
    __syncthreads();
    if (warp == 0) {
        unsigned int a = wsum_gt[lane], e = wsum_eq[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned int va = __shfl_up_sync(0xffffffffu, a, o), ve = __shfl_up_sync(0xffffffffu, e, o);
            if (lane >= o) { a += va; e += ve; }
        }
        wsum_gt[lane] = a; wsum_eq[lane] = e;            // inclusive over warps
    }
    __syncthreads();
    const unsigned int total_gt = wsum_gt[31];           // == K - need_eq
    unsigned int at_gt = inc_gt - n_gt + (warp ? wsum_gt[warp - 1] : 0u);
    unsigned int at_eq = inc_eq - n_eq + (warp ? wsum_eq[warp - 1] : 0u);
    for (int i = i0; i < i1; ++i) {
        const unsigned int k = skey[i];
        if (k > T) cand[at_gt++] = ((unsigned long long)(~k) << 32) | (unsigned int)i;
        else if (k == T) { if (at_eq < need_eq) cand[total_gt + at_eq] = ((unsigned long long)(~k) << 32) | (unsigned int)i; ++at_eq; }
    }
    for (int i = K + tid; i < TOPK_MAX; i += TOPK_THREADS) cand[i] = ~0ull;
    __syncthreads();
    // ---- bitonic sort of the TOPK_MAX slots, ascending --------------------------------------------------------------------
    for (int k2 = 2; k2 <= TOPK_MAX; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            const int ixj = tid ^ j;
            if (ixj > tid) {
                const unsigned long long x = cand[tid], y = cand[ixj];
                const bool up = (tid & k2) == 0;
                if ((x > y) == up) { cand[tid] = y; cand[ixj] = x; }
            }
            __syncthreads();
        }
    if (tid < K) {
        const int i = (int)(unsigned int)cand[tid];
        out_boxes[(long long)b * cap + tid] = boxes[(long long)b * N + i];
        out_scores[(long long)b * cap + tid] = kb[i];
    }
}

__global__ void fill_offsets_kernel(int *seg, int n, int stride) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) seg[i] = i * stride;
}

__global__ void proposal_pre_kernel(const int *n_valid, int B, int pre_topN, int rcnn_bs, int *pre, int *status) {
    int m = INT_MAX;
    for (int b = 0; b < B; ++b) m = min(m, n_valid[b]);
    m = min(m, pre_topN);
    *pre = m;
    *status = (m < rcnn_bs) ? -1 : 0;
}

__global__ void gather_sorted_kernel(const float4 *__restrict__ boxes, const float *__restrict__ keys_sorted,
                                     const int *__restrict__ idx_sorted, const int *__restrict__ pre, int N, int cap,
                                     float4 *__restrict__ out_boxes, float *__restrict__ out_scores) {
    const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= min(*pre, cap)) return;
    out_boxes[(long long)b * cap + i] = boxes[(long long)b * N + idx_sorted[(long long)b * N + i]];
    out_scores[(long long)b * cap + i] = keys_sorted[(long long)b * N + i];
}

__global__ void proposal_out_kernel(const float4 *__restrict__ boxes, const float *__restrict__ scores,
                                    const int *__restrict__ keep_idx, const int *__restrict__ keep_cnt, int B, int cap,
                                    int post_topN, const int *__restrict__ status, float4 *__restrict__ rois,
                                    float *__restrict__ out_scores, int *__restrict__ M_out) {
    int m = INT_MAX;
    for (int b = 0; b < B; ++b) m = min(m, keep_cnt[b]);
    m = min(m, post_topN);                         // nets_utils.py:236
    if (*status < 0) m = -1;
    if (blockIdx.x == 0 && threadIdx.x == 0) *M_out = m;
    const int b = blockIdx.x;
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int k = keep_idx[(long long)b * cap + i];
        rois[(long long)b * post_topN + i] = boxes[(long long)b * cap + k];
        out_scores[(long long)b * post_topN + i] = scores[(long long)b * cap + k];
    }
}

// ------------------------------------------------------------------------- final tail --------
// One block per image, thread per RoI (R <= 256).  layers.py:688-766.
constexpr int TAIL_MAX_R = 256;
__global__ void __launch_bounds__(TAIL_MAX_R)
final_detections_kernel(const float *__restrict__ bbox_reg, const float *__restrict__ probs,
                        const float4 *__restrict__ rois, int R, int C, float img_w, float img_h, float nms_thresh,
                        float min_score, float4 *__restrict__ det_boxes, float *__restrict__ det_scores,
                        int *__restrict__ det_class, int *__restrict__ det_count) {
    __shared__ float s_score[TAIL_MAX_R];
    __shared__ int s_cls[TAIL_MAX_R];
    __shared__ float4 s_box[TAIL_MAX_R];
    __shared__ float4 o_box[TAIL_MAX_R];      // score-descending, class 0 removed
    __shared__ float o_score[TAIL_MAX_R];
    __shared__ int o_cls[TAIL_MAX_R];
    __shared__ unsigned char o_dead[TAIL_MAX_R];
    __shared__ int s_pos[TAIL_MAX_R];
    __shared__ int n_fg;
    const int b = blockIdx.x, r = threadIdx.x;
    const int C1 = C + 1;
    if (r < R) {
        const float *p = probs + ((long long)b * R + r) * C1;
        float best = p[0];
        int arg = 0;
        for (int c = 1; c < C1; ++c) { const float v = p[c]; if (v > best) { best = v; arg = c; } }   // first max
        const float *d = bbox_reg + ((long long)b * R + r) * 4 * C1 + 4 * arg;
        const float4 box = clamp_box(decode_one(make_float4(d[0], d[1], d[2], d[3]), rois[(long long)b * R + r]),
                                     img_w, img_h);
        s_score[r] = best; s_cls[r] = arg; s_box[r] = box;
    }
    __syncthreads();
    if (r < R) {
        // rank in the stable descending order (ties: lower index first)
        // The order must be TOTAL or two rows share a rank and a slot of s_pos stays unwritten: NaN scores (a diverged
        // head) sort first, as torch.argsort(descending=True) places them, ties among them by index.
        const float me = s_score[r];
        const bool me_nan = me != me;
        int rank = 0;
        for (int j = 0; j < R; ++j) {
            const float v = s_score[j];
            const bool v_nan = v != v;
            const bool before = (v_nan || me_nan) ? (v_nan && !me_nan) : (v > me);
            const bool same = (v_nan && me_nan) || v == me;
            rank += before || (same && j < r);
        }
        s_pos[rank] = r;
    }
    __syncthreads();
    if (r == 0) {
        int n = 0;
        for (int i = 0; i < R; ++i) {
            const int src = s_pos[i];
            if (s_cls[src] > 0) { o_box[n] = s_box[src]; o_score[n] = s_score[src]; o_cls[n] = s_cls[src]; o_dead[n] = 0; ++n; }
        }
        n_fg = n;
    }
    __syncthreads();
    const int n = n_fg;
    // greedy NMS over n <= R boxes: step i, all threads j > i test against box i if it is alive
    for (int i = 0; i < n; ++i) {
        if (!o_dead[i] && r > i && r < n && !o_dead[r]) {
            if (iou_ge(o_box[r], o_box[i], box_area(o_box[r]), box_area(o_box[i]), nms_thresh)) o_dead[r] = 1;
        }
        __syncthreads();
    }
    if (r == 0) {
        int m = 0;
        for (int i = 0; i < n; ++i) {
            if (!o_dead[i] && o_score[i] > min_score) {
                det_boxes[(long long)b * R + m] = o_box[i];
                det_scores[(long long)b * R + m] = o_score[i];
                det_class[(long long)b * R + m] = o_cls[i];
                ++m;
            }
        }
        det_count[b] = m;
    }
}

// ------------------------------------------------------------------------------ merge ---------
// run_detection.py:180-230: border filter, x offset, end-of-file filter; key = class (dropped -> INT_MAX)
__global__ void merge_filter_kernel(const float4 *__restrict__ boxes, const int *__restrict__ cls,
                                    const int *__restrict__ tile, int n, int n_tiles, float w_pix, float hop_spectro,
                                    float spec_len, float min_border, float4 *__restrict__ shifted,
                                    int *__restrict__ keys, int *__restrict__ idx) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 b = boxes[i];
    const int t = tile[i];
    const float width = __fsub_rn(b.z, b.x);
    const bool right = b.z >= w_pix - 5.f, left = b.x <= 4.f;
    const bool edge = (t == 0) ? right : ((t == n_tiles - 1) ? left : (left || right));
    bool drop = edge && (width < min_border);
    const float off = __fmul_rn(hop_spectro, (float)t);
    b.x = __fadd_rn(b.x, off);
    b.z = __fadd_rn(b.z, off);
    drop = drop || (b.z >= spec_len);
    shifted[i] = b;
    keys[i] = drop ? INT_MAX : cls[i];
    idx[i] = i;
}

__global__ void merge_gather_kernel(const float4 *__restrict__ shifted, const int *__restrict__ keys_sorted,
                                    const int *__restrict__ idx_sorted, int n, float4 *__restrict__ cand,
                                    int *__restrict__ n_cand) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const bool ok = keys_sorted[i] != INT_MAX;
    if (ok) cand[i] = shifted[idx_sorted[i]];
    if (ok && (i == n - 1 || keys_sorted[i + 1] == INT_MAX)) *n_cand = i + 1;
    if (i == 0 && !ok) *n_cand = 0;
}

__global__ void merge_out_kernel(const float4 *__restrict__ cand, const float *__restrict__ scores,
                                 const int *__restrict__ keys_sorted, const int *__restrict__ idx_sorted,
                                 const int *__restrict__ keep_idx, const int *__restrict__ keep_cnt,
                                 float4 *__restrict__ out_boxes, float *__restrict__ out_scores,
                                 int *__restrict__ out_class, int *__restrict__ out_count) {
    const int m = *keep_cnt;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) *out_count = m;
    if (i >= m) return;
    const int k = keep_idx[i];
    out_boxes[i] = cand[k];
    out_scores[i] = scores[idx_sorted[k]];
    out_class[i] = keys_sorted[k];
}

}  // namespace nbm

using namespace nbm;

// --------------------------------------------------------------------------- C entry points --
extern "C" int nbm_make_anchors(int32_t base_size, const double *ratios, int32_t n_ratios, const int64_t *scales,
                                int32_t n_scales, int32_t width, int32_t height, int32_t stride, float *out) {
    NBM_REQUIRE(ratios && scales && out && n_ratios > 0 && n_scales > 0 && width > 0 && height > 0, "bad argument");
    const int A = n_ratios * n_scales;
    std::vector<long long> base((size_t)A * 4);
    const double side = sqrt((double)base_size * (double)base_size);
    const long long centre = (long long)(base_size / 2);        // int(base_size / 2)
    for (int s = 0; s < n_scales; ++s)
        for (int r = 0; r < n_ratios; ++r) {
            const double w = sqrt(ratios[r]) * side * (double)scales[s];
            const double h = (1.0 / sqrt(ratios[r])) * side * (double)scales[s];
            long long *a = &base[(size_t)(s * n_ratios + r) * 4];
            // numpy .astype(int) truncates toward zero
            a[0] = (long long)(-w / 2 + (double)centre);
            a[1] = (long long)(-h / 2 + (double)centre);
            a[2] = (long long)(w / 2 + (double)centre);
            a[3] = (long long)(h / 2 + (double)centre);
        }
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x)
            for (int a = 0; a < A; ++a) {
                float *o = out + (((size_t)y * width + x) * A + a) * 4;
                o[0] = (float)(base[a * 4 + 0] + (long long)x * stride);
                o[1] = (float)(base[a * 4 + 1] + (long long)y * stride);
                o[2] = (float)(base[a * 4 + 2] + (long long)x * stride);
                o[3] = (float)(base[a * 4 + 3] + (long long)y * stride);
            }
    return NBM_OK;
}

extern "C" int nbm_decode_boxes(const float *d_deltas, const float *d_anchors, int32_t B, int32_t N,
                                int32_t anchors_per_image, float clip_w, float clip_h, float min_size,
                                float *d_boxes, uint8_t *d_valid, void *stream) {
    NBM_REQUIRE(d_deltas && d_anchors && d_boxes && B >= 0 && N >= 0, "bad argument");
    const long long total = (long long)B * N;
    if (total == 0) return NBM_OK;
    decode_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4 *>(d_deltas), reinterpret_cast<const float4 *>(d_anchors), B, N,
        anchors_per_image, clip_w, clip_h, min_size, reinterpret_cast<float4 *>(d_boxes), d_valid);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}

extern "C" size_t nbm_nms_workspace_bytes(int32_t B, int32_t N) {
    return (size_t)std::max(B, 0) * std::max(N, 0) * nms_words(std::max(N, 1)) * sizeof(unsigned long long);
}

extern "C" int nbm_nms_greedy(const float *d_boxes, const int32_t *d_n, int32_t B, int32_t N, float thresh,
                              int32_t *d_keep_idx, int32_t *d_keep_cnt, void *ws, size_t ws_bytes, void *stream) {
    NBM_REQUIRE(d_keep_cnt && B >= 0 && N >= 0, "bad argument");
    if (B == 0) return NBM_OK;
    if (N == 0) { NBM_CUDA(cudaMemsetAsync(d_keep_cnt, 0, sizeof(int) * B, (cudaStream_t)stream)); return NBM_OK; }
    NBM_REQUIRE(d_boxes && d_keep_idx && ws, "null argument");
    return launch_nms(d_boxes, d_n, nullptr, B, N, thresh, d_keep_idx, d_keep_cnt, ws, ws_bytes, (cudaStream_t)stream);
}

namespace {
struct ProposalWs {
    size_t boxes, keys, keys_sorted, idx, idx_sorted, seg, n_valid, pre, status, M, top_boxes, top_scores,
        keep_idx, keep_cnt, nms, cub, total, cub_bytes;
};
ProposalWs proposal_ws(const nbm_proposal_params &p, int B) {
    ProposalWs w{};
    const size_t N = (size_t)p.A * p.H * p.W, cap = (size_t)p.pre_nms_topN;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    w.boxes = take(B * N * 16); w.keys = take(B * N * 4); w.keys_sorted = take(B * N * 4);
    w.idx = take(B * N * 4); w.idx_sorted = take(B * N * 4); w.seg = take((B + 1) * 4);
    w.n_valid = take(B * 4); w.pre = take(4); w.status = take(4); w.M = take(4);
    w.top_boxes = take(B * cap * 16); w.top_scores = take(B * cap * 4);
    w.keep_idx = take(B * cap * 4); w.keep_cnt = take(B * 4);
    w.nms = take(nbm_nms_workspace_bytes(B, (int)cap));
    size_t cub_bytes = 0;
    cub::DeviceSegmentedRadixSort::SortPairsDescending(nullptr, cub_bytes, (const float *)nullptr, (float *)nullptr,
                                                       (const int *)nullptr, (int *)nullptr, (int)(B * N), B,
                                                       (const int *)nullptr, (const int *)nullptr);
    w.cub_bytes = cub_bytes;
    w.cub = take(cub_bytes);
    w.total = o;
    return w;
}
}  // namespace

extern "C" size_t nbm_proposals_workspace_bytes(const nbm_proposal_params *p, int32_t B) {
    if (!p || B <= 0) return 0;
    return proposal_ws(*p, B).total;
}

namespace {
int proposals_impl(const nbm_proposal_params *p, const float *d_cls, const float *d_reg, const float *d_anchors, int32_t B,
                   float *d_rois, float *d_scores, int32_t *h_M, int32_t *d_M, void *ws_, size_t ws_bytes, void *stream_);
}

extern "C" int nbm_proposals(const nbm_proposal_params *p, const float *d_cls, const float *d_reg,
                             const float *d_anchors, int32_t B, float *d_rois, float *d_scores, int32_t *h_M,
                             void *ws_, size_t ws_bytes, void *stream_) {
    NBM_REQUIRE(h_M, "bad argument");
    return proposals_impl(p, d_cls, d_reg, d_anchors, B, d_rois, d_scores, h_M, nullptr, ws_, ws_bytes, stream_);
}

extern "C" int nbm_proposals_async(const nbm_proposal_params *p, const float *d_cls, const float *d_reg,
                                   const float *d_anchors, int32_t B, float *d_rois, float *d_scores, int32_t *d_M,
                                   void *ws_, size_t ws_bytes, void *stream_) {
    NBM_REQUIRE(d_M, "bad argument");
    return proposals_impl(p, d_cls, d_reg, d_anchors, B, d_rois, d_scores, nullptr, d_M, ws_, ws_bytes, stream_);
}

namespace {
int proposals_impl(const nbm_proposal_params *p, const float *d_cls, const float *d_reg, const float *d_anchors, int32_t B,
                   float *d_rois, float *d_scores, int32_t *h_M, int32_t *d_M, void *ws_, size_t ws_bytes, void *stream_) {
    NBM_REQUIRE(p && d_cls && d_reg && d_anchors && d_rois && d_scores && ws_ && B >= 1, "bad argument");
    NBM_REQUIRE(p->pre_nms_topN >= 1 && p->post_nms_topN >= 1, "topN must be positive");
    cudaStream_t s = (cudaStream_t)stream_;
    const ProposalWs w = proposal_ws(*p, B);
    if (ws_bytes < w.total) { set_error("proposal workspace too small: %zu < %zu", ws_bytes, w.total); return NBM_ERR_WORKSPACE; }
    char *ws = reinterpret_cast<char *>(ws_);
    const int N = p->A * p->H * p->W, cap = p->pre_nms_topN;
    auto *boxes = reinterpret_cast<float4 *>(ws + w.boxes);
    auto *keys = reinterpret_cast<float *>(ws + w.keys);
    auto *keys_sorted = reinterpret_cast<float *>(ws + w.keys_sorted);
    auto *idx = reinterpret_cast<int *>(ws + w.idx);
    auto *idx_sorted = reinterpret_cast<int *>(ws + w.idx_sorted);
    auto *seg = reinterpret_cast<int *>(ws + w.seg);
    auto *n_valid = reinterpret_cast<int *>(ws + w.n_valid);
    auto *pre = reinterpret_cast<int *>(ws + w.pre);
    auto *status = reinterpret_cast<int *>(ws + w.status);
    auto *M = d_M ? d_M : reinterpret_cast<int *>(ws + w.M);
    auto *top_boxes = reinterpret_cast<float4 *>(ws + w.top_boxes);
    auto *top_scores = reinterpret_cast<float *>(ws + w.top_scores);
    auto *keep_idx = reinterpret_cast<int *>(ws + w.keep_idx);
    auto *keep_cnt = reinterpret_cast<int *>(ws + w.keep_cnt);

    NBM_CUDA(cudaMemsetAsync(n_valid, 0, sizeof(int) * B, s));
    const long long total = (long long)B * N;
    const size_t topk_smem = (size_t)N * sizeof(unsigned int);
    const bool use_topk = cap <= TOPK_MAX && topk_smem <= 160 * 1024;      // else: full segmented sort (cub)
    {
        const size_t smem = (size_t)5 * p->A * RPN_CELLS * sizeof(float);
        NBM_REQUIRE(smem <= 48 * 1024, "too many anchors per cell for rpn_decode_kernel");
        dim3 gd((p->H * p->W + RPN_CELLS - 1) / RPN_CELLS, B);
        rpn_decode_kernel<<<gd, RPN_THREADS, smem, s>>>(
            d_cls, d_reg, reinterpret_cast<const float4 *>(d_anchors), B, p->A, p->H, p->W, p->img_width, p->img_height,
            p->min_size, boxes, keys, use_topk ? nullptr : idx, n_valid);
    }
    proposal_pre_kernel<<<1, 1, 0, s>>>(n_valid, B, p->pre_nms_topN, p->rcnn_batch_size, pre, status);
    if (use_topk) {
        static bool attr_set = false;
        if (!attr_set) {
            NBM_CUDA(cudaFuncSetAttribute(proposal_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
            attr_set = true;
        }
        proposal_topk_kernel<<<B, TOPK_THREADS, topk_smem, s>>>(boxes, keys, pre, N, cap, top_boxes, top_scores);
    } else {
        fill_offsets_kernel<<<(B + 1 + 127) / 128, 128, 0, s>>>(seg, B + 1, N);
        size_t cub_bytes = w.cub_bytes;
        NBM_CUDA(cub::DeviceSegmentedRadixSort::SortPairsDescending(ws + w.cub, cub_bytes, keys, keys_sorted, idx,
                                                                    idx_sorted, (int)total, B, seg, seg + 1, 0, 32, s));
        dim3 gg((cap + 127) / 128, B);
        gather_sorted_kernel<<<gg, 128, 0, s>>>(boxes, keys_sorted, idx_sorted, pre, N, cap, top_boxes, top_scores);
    }
    int rc = launch_nms(reinterpret_cast<const float *>(top_boxes), nullptr, pre, B, cap, p->nms_thresh, keep_idx,
                        keep_cnt, ws + w.nms, nbm_nms_workspace_bytes(B, cap), s);
    if (rc != NBM_OK) return rc;
    proposal_out_kernel<<<B, 128, 0, s>>>(top_boxes, top_scores, keep_idx, keep_cnt, B, cap, p->post_nms_topN, status,
                                          reinterpret_cast<float4 *>(d_rois), d_scores, M);
    if (h_M) {
        NBM_CUDA(cudaMemcpyAsync(h_M, M, sizeof(int), cudaMemcpyDeviceToHost, s));
        NBM_CUDA(cudaStreamSynchronize(s));
    }
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}
}  // namespace

extern "C" int nbm_final_detections(const float *d_bbox_reg, const float *d_probs, const float *d_rois, int32_t B,
                                    int32_t R, int32_t num_classes, float img_width, float img_height,
                                    float nms_thresh, float min_score, float *d_det_boxes, float *d_det_scores,
                                    int32_t *d_det_class, int32_t *d_det_count, void *stream) {
    NBM_REQUIRE(d_bbox_reg && d_probs && d_rois && d_det_boxes && d_det_scores && d_det_class && d_det_count,
                "null argument");
    NBM_REQUIRE(B >= 1 && R >= 1 && num_classes >= 1, "bad sizes");
    if (R > TAIL_MAX_R) { set_error("R=%d exceeds the fused tail limit %d", R, TAIL_MAX_R); return NBM_ERR_UNSUPPORTED; }
    final_detections_kernel<<<B, TAIL_MAX_R, 0, (cudaStream_t)stream>>>(
        d_bbox_reg, d_probs, reinterpret_cast<const float4 *>(d_rois), R, num_classes, img_width, img_height,
        nms_thresh, min_score, reinterpret_cast<float4 *>(d_det_boxes), d_det_scores, d_det_class, d_det_count);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}

namespace {
struct MergeWs { size_t shifted, keys, keys_sorted, idx, idx_sorted, cand, n_cand, keep_idx, keep_cnt, nms, cub, total, cub_bytes; };
MergeWs merge_ws(int n) {
    MergeWs w{};
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t at = o; o = align_up(o + bytes, 256); return at; };
    const size_t N = (size_t)std::max(n, 1);
    w.shifted = take(N * 16); w.keys = take(N * 4); w.keys_sorted = take(N * 4); w.idx = take(N * 4);
    w.idx_sorted = take(N * 4); w.cand = take(N * 16); w.n_cand = take(4); w.keep_idx = take(N * 4);
    w.keep_cnt = take(4); w.nms = take(nbm_nms_workspace_bytes(1, (int)N));
    size_t cb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, cb, (const int *)nullptr, (int *)nullptr, (const int *)nullptr,
                                    (int *)nullptr, (int)N);
    w.cub_bytes = cb;
    w.cub = take(cb);
    w.total = o;
    return w;
}
}  // namespace

extern "C" size_t nbm_merge_workspace_bytes(int32_t n) { return merge_ws(n).total; }

extern "C" int nbm_merge_detections(const float *d_boxes, const float *d_scores, const int32_t *d_class,
                                    const int32_t *d_tile, int32_t n, int32_t n_tiles, int32_t w_pix,
                                    int32_t hop_spectro, int64_t spectrogram_length, float nms_thresh,
                                    float *d_out_boxes, float *d_out_scores, int32_t *d_out_class,
                                    int32_t *d_out_count, void *ws_, size_t ws_bytes, void *stream_) {
    NBM_REQUIRE(d_out_count && n >= 0 && n_tiles >= 1, "bad argument");
    cudaStream_t s = (cudaStream_t)stream_;
    if (n == 0) { NBM_CUDA(cudaMemsetAsync(d_out_count, 0, sizeof(int), s)); return NBM_OK; }
    NBM_REQUIRE(d_boxes && d_scores && d_class && d_tile && d_out_boxes && d_out_scores && d_out_class && ws_,
                "null argument");
    const MergeWs w = merge_ws(n);
    if (ws_bytes < w.total) { set_error("merge workspace too small: %zu < %zu", ws_bytes, w.total); return NBM_ERR_WORKSPACE; }
    char *ws = reinterpret_cast<char *>(ws_);
    auto *shifted = reinterpret_cast<float4 *>(ws + w.shifted);
    auto *keys = reinterpret_cast<int *>(ws + w.keys);
    auto *keys_sorted = reinterpret_cast<int *>(ws + w.keys_sorted);
    auto *idx = reinterpret_cast<int *>(ws + w.idx);
    auto *idx_sorted = reinterpret_cast<int *>(ws + w.idx_sorted);
    auto *cand = reinterpret_cast<float4 *>(ws + w.cand);
    auto *n_cand = reinterpret_cast<int *>(ws + w.n_cand);
    auto *keep_idx = reinterpret_cast<int *>(ws + w.keep_idx);
    auto *keep_cnt = reinterpret_cast<int *>(ws + w.keep_cnt);
    const float min_border = (float)(0.9 * (double)(w_pix - hop_spectro));      // run_detection.py:165
    const int blocks = (n + 255) / 256;
    merge_filter_kernel<<<blocks, 256, 0, s>>>(reinterpret_cast<const float4 *>(d_boxes), d_class, d_tile, n, n_tiles,
                                               (float)w_pix, (float)hop_spectro, (float)spectrogram_length, min_border,
                                               shifted, keys, idx);
    size_t cb = w.cub_bytes;
    NBM_CUDA(cub::DeviceRadixSort::SortPairs(ws + w.cub, cb, keys, keys_sorted, idx, idx_sorted, n, 0, 32, s));
    merge_gather_kernel<<<blocks, 256, 0, s>>>(shifted, keys_sorted, idx_sorted, n, cand, n_cand);
    int rc = launch_nms(reinterpret_cast<const float *>(cand), n_cand, nullptr, 1, n, nms_thresh, keep_idx, keep_cnt,
                        ws + w.nms, nbm_nms_workspace_bytes(1, n), s);
    if (rc != NBM_OK) return rc;
    merge_out_kernel<<<blocks, 256, 0, s>>>(cand, d_scores, keys_sorted, idx_sorted, keep_idx, keep_cnt,
                                            reinterpret_cast<float4 *>(d_out_boxes), d_out_scores, d_out_class,
                                            d_out_count);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}
