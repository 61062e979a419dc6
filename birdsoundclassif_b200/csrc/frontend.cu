// Front-end kernels: PCM -> dB band spectrogram (+ per-file min/max) -> normalised detector tiles.
//
// Replaces File_Processor.spectrogram / split_power_spec of the reference
// (nbm_model/nbm_datasets/prepare_dataset.py:228-294) and librosa.stft as called at :237.
//
// Algorithm (exact DFT bins of an n_fft that need not be a power of two; 1324 = 4*331 here):
//   R_t[k]  = sum_n x[t*hop - n_fft/2 + n] w^(kn),  w = exp(-2 pi i / n_fft)   (rectangular window)
//   X_t[k]  = R_t[k]/2 - (R_t[k-1] + R_t[k+1])/4                               (periodic Hann, exact)
//   R_{t+1} = w^(-hop k) (R_t + D_t),  D_t[k] = sum_{m<hop} (x[s_t+n_fft+m] - x[s_t+m]) w^(km)
// Frames overlap by 1 - hop/n_fft (90 %), so instead of one n_fft-point transform per frame each
// group of 64 frames computes ONE direct anchor DFT (middle frame) and slides it 32 frames forward
// and 32 backward with hop-point difference DFTs.  Both the anchor and the hop-point sums are
// folded about their centre (cos part uses x[c+u]+x[c-u], sin part x[c+u]-x[c-u]) which halves the
// multiply-adds.  Thread = frequency bin, so the per-frame recurrence is register-resident; the
// Hann combination uses warp shuffles (each warp computes 32 bins, emits the inner 30).
#include <algorithm>
#include <mutex>
#include <vector>
#include <cmath>

#include <cstdlib>
#include <type_traits>

#include <cuda.h>

#include "common.cuh"
#include "frontend_tc.cuh"

namespace nbm {

constexpr int PASS = 32;    // frames slid per direction
constexpr int BINS_PER_WARP = 30;
constexpr int STAGE_LD = PASS + 1;
constexpr int TILE_ROWS = 15;   // spectrogram rows per block in the tiling kernel

struct FileDesc {
    long long spec_off;    // float offset of the file's S[0][0]
    long long tile0;       // index of the file's first tile in d_tiles
    int row_stride;
    int n_tiles;
    int total_frames;
    int last_width;        // valid columns of the last tile (rest is reflect padding)
    int group0, n_groups;  // the file's 64-frame groups (contiguous)
    int seg0, n_segs;      // the file's STFT chunks (contiguous)
};

struct KParams {
    int N, hop, low_idx, n_bins, w_pix, hop_spectro;
    int npN, npH;          // folded pair counts, ceil(N/2), ceil(hop/2)
    int buf_len;           // N + (GF-1)*hop
    float min_level_sq;
    const float2 *tw;      // [2N] (cos, sin)(pi q / N)
};

__device__ __forceinline__ float load_sample(const void *pcm, int dtype, int channels, long long idx) {
    float s = 0.f;
    if (dtype == NBM_PCM_INT16) {
        const short *p = reinterpret_cast<const short *>(pcm) + idx * channels;
        for (int c = 0; c < channels; ++c) s += (float)__ldg(p + c) * (1.0f / 32768.0f);
    } else {
        const float *p = reinterpret_cast<const float *>(pcm) + idx * channels;
        for (int c = 0; c < channels; ++c) s += __ldg(p + c);
    }
    return channels == 1 ? s : s / (float)channels;
}

// pair j of a length-L fold: 2u = 2j + (L even), hi = (L-1+2u)/2, lo = (L-1-2u)/2
__device__ __forceinline__ void fold_idx(int L, int j, int &hi, int &lo) {
    int u2 = 2 * j + ((L & 1) ? 0 : 1);
    hi = (L - 1 + u2) >> 1;
    lo = (L - 1 - u2) >> 1;
}

__global__ void __launch_bounds__(512, 1)
stft_db_kernel(KParams P, const SegDesc *__restrict__ segs, int n_segs, const void *__restrict__ pcm,
               int dtype, int channels, float *__restrict__ spec, float2 *__restrict__ tile_mm, int group_begin,
               float rel_db, uint4 *__restrict__ cand,
               unsigned int *__restrict__ cand_count, unsigned int cand_cap) {
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nthr = blockDim.x;

    // ---- which segment / group -------------------------------------------------------------
    const int group = (int)blockIdx.x + group_begin;
    int lo_s = 0, hi_s = n_segs - 1;
    while (lo_s < hi_s) {
        int mid = (lo_s + hi_s + 1) >> 1;
        if (segs[mid].group0 <= group) lo_s = mid; else hi_s = mid - 1;
    }
    const SegDesc sd = segs[lo_s];
    const int t0 = (group - sd.group0) * GF;
    const int nf = min(GF, sd.n_frames - t0);
    const int a = nf >> 1;                       // anchor frame (local)
    const int N = P.N, hop = P.hop, N2 = 2 * P.N;

    const int buf_pad = (P.buf_len + 3) & ~3;
    float *buf = smem;
    float2 *F = reinterpret_cast<float2 *>(buf + buf_pad);
    float2 *E = F + ((P.npN + 1) & ~1);
    float *stage = reinterpret_cast<float *>(E + P.npH * PASS);
    __shared__ float bin_rmax[16 * 32];     // largest |R| each thread's bin carried in the pass (thread = bin), see refine_groups_kernel

    // ---- samples of the group (zero outside the segment: centre padding, pad_mode='constant') --
    const long long s0 = (long long)t0 * hop - N / 2;
    const int need = N + (nf - 1) * hop;
    for (int i = tid; i < need; i += nthr) {
        long long s = s0 + i;
        buf[i] = (s >= 0 && s < sd.n_samples) ? load_sample(pcm, dtype, channels, sd.pcm_start + s) : 0.f;
    }
    __syncthreads();

    // ---- anchor: folded pairs of frame a -------------------------------------------------------
    const float *xa = buf + a * hop;
    for (int j = tid; j < P.npN; j += nthr) {
        int hi, lo;
        fold_idx(N, j, hi, lo);
        F[j] = (hi == lo) ? make_float2(xa[hi], 0.f) : make_float2(xa[hi] + xa[lo], xa[hi] - xa[lo]);
    }
    __syncthreads();

    const int k = P.low_idx - 1 + warp * BINS_PER_WARP + lane;      // this thread's DFT bin
    const int kk = k % N2;
    const int step = (2 * kk) % N2;
    float Rr, Ri;
    {
        int q = (N & 1) ? 0 : kk;
        double dr = 0.0, di = 0.0;
        for (int j0 = 0; j0 < P.npN; j0 += 32) {
            float ar = 0.f, ai = 0.f;
            const int j1 = min(j0 + 32, P.npN);
            for (int j = j0; j < j1; ++j) {
                const float2 f = F[j];
                const float2 w = __ldg(P.tw + q);
                ar = fmaf(f.x, w.x, ar);
                ai = fmaf(f.y, w.y, ai);
                q += step;
                if (q >= N2) q -= N2;
            }
            dr += (double)ar;
            di += (double)ai;
        }
        // R_a = exp(-i theta (N-1)/2) (A - iB)
        const float2 c0 = __ldg(P.tw + (int)(((long long)kk * (N - 1)) % N2));
        Rr = (float)((double)c0.x * dr - (double)c0.y * di);
        Ri = (float)(-((double)c0.x * di + (double)c0.y * dr));
    }
    const float2 cf = __ldg(P.tw + (int)(((long long)kk * 2 * hop) % N2));     // e^{+i theta hop}
    const float2 gf = __ldg(P.tw + (int)(((long long)kk * (hop + 1)) % N2));   // e^{+i theta (hop+1)/2}
    const float2 gb = __ldg(P.tw + (int)(((long long)kk * (hop - 1)) % N2));   // e^{-i theta (hop-1)/2} = (c,-s)
    const float Ra_r = Rr, Ra_i = Ri;
    // The float32 twiddle cf is off by rho = cf_true / cf in every step of the recurrence (a systematic error that grows
    // linearly along the pass); every 4 steps the recurrence multiplies by rho^4 = 1 + fx (see frontend_tc.cu).
    float2 fx;
    {
        double sn, cs;
        sincospi((double)(((long long)kk * 2 * hop) % N2) / (double)N, &sn, &cs);
        const double cr = (double)cf.x, ci = (double)cf.y, d2 = cr * cr + ci * ci;
        const double pr = (cs * cr + sn * ci) / d2, pi = (sn * cr - cs * ci) / d2;
        double qr = pr * pr - pi * pi, qi = 2.0 * pr * pi;               // rho^2
        const double q4r = qr * qr - qi * qi, q4i = 2.0 * qr * qi;       // rho^4
        fx = make_float2((float)(q4r - 1.0), (float)q4i);
    }

    const int out_bin = warp * BINS_PER_WARP + lane - 1;
    const bool emit = lane >= 1 && lane <= BINS_PER_WARP && out_bin < P.n_bins;
    float vmin = INFINITY, vmax = -INFINITY;
    float *spec_seg = spec + sd.spec_off;
    const int n_warps = nthr >> 5;

    auto emit_frame = [&](float r, float im, int col) {
        const float lr = __shfl_up_sync(0xffffffffu, r, 1), li = __shfl_up_sync(0xffffffffu, im, 1);
        const float rr = __shfl_down_sync(0xffffffffu, r, 1), ri = __shfl_down_sync(0xffffffffu, im, 1);
        const float xr = 0.5f * r - 0.25f * (lr + rr);
        const float xi = 0.5f * im - 0.25f * (li + ri);
        const float p = fmaxf(fmaf(xr, xr, xi * xi), P.min_level_sq);
        const float db = 3.0102999566398120f * __log2f(p);     // 10 log10(p)
        if (emit) stage[out_bin * STAGE_LD + col] = db;
    };

    for (int pass = 0; pass < 2; ++pass) {
        const bool fwd = pass == 0;
        const int cnt = fwd ? (nf - 1 - a) : a;          // hop-DFTs needed in this direction
        // ---- folded hop-point differences, E[j][i] for the pass's frames ---------------------
        for (int idx = tid; idx < P.npH * PASS; idx += nthr) {
            const int j = idx / PASS, i = idx - j * PASS;
            float2 e = make_float2(0.f, 0.f);
            if (i < cnt) {
                const int t = fwd ? (a + i) : (a - 1 - i);
                const float *x = buf + t * hop;
                int hi, lo;
                fold_idx(hop, j, hi, lo);
                const float dh = x[N + hi] - x[hi];
                if (hi == lo) e = make_float2(dh, 0.f);
                else { const float dl = x[N + lo] - x[lo]; e = make_float2(dh + dl, dh - dl); }
            }
            E[idx] = e;
        }
        __syncthreads();
        // ---- G_t[k] = sum_j E+ cos - i sum_j E- sin, 32 frames in registers -------------------
        float gr[PASS], gi[PASS];
#pragma unroll
        for (int i = 0; i < PASS; ++i) { gr[i] = 0.f; gi[i] = 0.f; }
        if (cnt > 0) {
            int q = (hop & 1) ? 0 : kk;
            const float4 *E4 = reinterpret_cast<const float4 *>(E);
            for (int j = 0; j < P.npH; ++j) {
                const float2 w = __ldg(P.tw + q);
#pragma unroll
                for (int i2 = 0; i2 < PASS / 2; ++i2) {
                    const float4 e = E4[j * (PASS / 2) + i2];
                    gr[2 * i2] = fmaf(e.x, w.x, gr[2 * i2]);
                    gi[2 * i2] = fmaf(e.y, w.y, gi[2 * i2]);
                    gr[2 * i2 + 1] = fmaf(e.z, w.x, gr[2 * i2 + 1]);
                    gi[2 * i2 + 1] = fmaf(e.w, w.y, gi[2 * i2 + 1]);
                }
                q += step;
                if (q >= N2) q -= N2;
            }
        }
        // ---- slide, Hann (shuffles), dB, stage -------------------------------------------------
        Rr = Ra_r; Ri = Ra_i;
        float mR = fmaxf(fabsf(Rr), fabsf(Ri));
        if (fwd) {
            emit_frame(Rr, Ri, 0);
#pragma unroll
            for (int i = 0; i < PASS; ++i) {
                if (i < cnt) {       // block-uniform
                    const float Gr = gr[i], Gi = -gi[i];
                    float nr = cf.x * Rr - cf.y * Ri + gf.x * Gr - gf.y * Gi;
                    float ni = cf.x * Ri + cf.y * Rr + gf.x * Gi + gf.y * Gr;
                    if ((i & 3) == 3) {
                        const float tr = fmaf(fx.x, nr, fmaf(-fx.y, ni, nr));
                        ni = fmaf(fx.x, ni, fmaf(fx.y, nr, ni));
                        nr = tr;
                    }
                    Rr = nr; Ri = ni;
                    mR = fmaxf(mR, fmaxf(fabsf(nr), fabsf(ni)));
                    emit_frame(Rr, Ri, i + 1);
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < PASS; ++i) {
                if (i < cnt) {
                    const float Gr = gr[i], Gi = -gi[i];
                    // R_t = conj(cf) R_{t+1} - (gb.x - i gb.y) G_t
                    float nr = cf.x * Rr + cf.y * Ri - (gb.x * Gr + gb.y * Gi);
                    float ni = cf.x * Ri - cf.y * Rr - (gb.x * Gi - gb.y * Gr);
                    if ((i & 3) == 3) {
                        const float tr = fmaf(fx.x, nr, fmaf(fx.y, ni, nr));
                        ni = fmaf(fx.x, ni, fmaf(-fx.y, nr, ni));
                        nr = tr;
                    }
                    Rr = nr; Ri = ni;
                    mR = fmaxf(mR, fmaxf(fabsf(nr), fabsf(ni)));
                    emit_frame(Rr, Ri, a - 1 - i);
                }
            }
        }
        bin_rmax[tid] = mR;
        __syncthreads();
        // ---- coalesced rows out -----------------------------------------------------------------
        const int ncols = fwd ? (cnt + 1) : cnt;
        const int col0 = t0 + (fwd ? a : 0);
        // min / max of the unflagged pixels; pixels below their frame's flag level go to the float64 pass
        // (refine_groups_kernel), exactly as in the tensor-core kernel
        for (int h = 0; h < 2; ++h) {
            const int c = lane + 32 * h;
            if (c >= ncols) continue;
            for (int b = warp; b < P.n_bins; b += n_warps) {
                // bin b is thread (b / 30) * 32 + b % 30 + 1; its Hann neighbours are the threads next to it
                const int tb = (b / BINS_PER_WARP) * 32 + b % BINS_PER_WARP + 1;
                const float m3 = fmaxf(bin_rmax[tb], fmaxf(bin_rmax[tb - 1], bin_rmax[tb + 1]));
                const float th = m3 > 0.f ? fmaf(__log2f(m3), 6.0205999132796239f, rel_db) : -INFINITY;
                const float v = stage[b * STAGE_LD + c];
                spec_seg[(long long)b * sd.row_stride + col0 + c] = v;
                vmax = fmaxf(vmax, v);
                bool listed = false;
                if (v < th) {
                    const unsigned int at = atomicAdd(cand_count, 1u);
                    if (at < cand_cap) {
                        cand[at] = list_entry(pack_group(lo_s, b, 1, 1, col0 + c), INFINITY);   // tested against its own bin's level already
                        listed = true;
                    }
                }
                if (!listed) vmin = fminf(vmin, v);
            }
        }
        __syncthreads();
    }

    // ---- per-file min / max ---------------------------------------------------------------------
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    float *red = stage;
    if (lane == 0) { red[warp] = vmin; red[32 + warp] = vmax; }
    __syncthreads();
    if (warp == 0) {
        vmin = lane < n_warps ? red[lane] : INFINITY;
        vmax = lane < n_warps ? red[32 + lane] : -INFINITY;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
            vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
        }
        if (lane == 0) tile_mm[group] = make_float2(vmin, vmax);          // one slot per group on this path
    }
}

// ---------------------------------------------------------------------------------------------
// Float64 refinement of the pixels float32 cannot deliver within tolerance, and the whole-file min / max
// (prepare_dataset.py:248-250).
//
// The transform's absolute error in a pixel is float32 rounding (~2^-24 rms, < ~2e-6 worst) of the largest rectangular-
// window magnitude |R_t[k]| its bin carried along the sliding recurrence since the anchor (measured,
// scripts/fe_outliers.py) -- inherent to any float32 sliding DFT.  In dB that is invisible except where the pixel lies
// ~60 dB below that magnitude: deep spectral nulls, and quiet frames a few hops after a loud call swept through the bin.
// The deepest null of a file is s_min, which offsets every normalised pixel.  So:
//   transform kernels        track max |R| per bin along the chain, list the pixels (tensor-core kernel: the thread's block of
//                            pixels) that fall `rel_db` below it and keep them out of the min/max partials;
//   refine_groups_kernel     recomputes each listed pixel the way the reference does -- float64 windowed DFT ->
//                            complex64 -> float32 |.| -> float64 log10 (prepare_dataset.py:237-240) -- patches the dB
//                            band and folds the exact value into the file's minimum (ordered-integer atomicMin);
//   minmax_kernel            file min = min(partials of the unlisted pixels, exact minimum of the listed blocks), with the
//                            unlisted pixels within 0.25 dB of it recomputed too, so that s_min is the reference's value.
struct RefineParams {
    int N, hop, low_idx, n_bins;
    int mm_frames;          // frames per min/max partial group (a divisor of GF)
    int mm_per_group;       // partials per mm_frames group (ranges x slots)
    int slots_per_range, bins_per_range, bins_per_slot;
    double min_level;
    float margin_db;        // unlisted pixels this close to the file minimum are recomputed as well
    int n_ranges;           // tensor-core path: 128-row ranges (flag levels are per (chain, range, emit warp)), else 0
    const double2 *tw64;    // [N] (cos, sin)(2 pi q / N)
    const double *hann64;   // [N] periodic Hann window / 32768
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// One warp per listed block (pack_group): lane = pixel (row = lane / 2, frame = lane % 2).  Pixels below the flag level
// are recomputed one after the other by the whole warp.  Sample i = 32 j + l of the window belongs to lane l, and
// e^{-i theta (32 j + l)} = e^{-i theta 32 j} e^{-i theta l}: inside the loop the twiddle is the same for every lane (one
// broadcast load from the float64 table per step), the lane's own factor is applied once at the end.  Per sample and lane
// that leaves four float64 operations: the int16 sample becomes a double by the 2^52 trick (no conversion unit), times the
// window (exactly the reference's float32 x float64 product, rounded once), and two fused multiply-adds.
struct GroupKey {
    int seg, bin0, rows, nfr, frame0;
    __device__ explicit GroupKey(unsigned long long key)
        : seg((int)(key >> 42)), bin0((int)((key >> 32) & 1023)), rows((int)((key >> 28) & 15) + 1),
          nfr((int)((key >> 27) & 1) + 1), frame0((int)(key & GROUP_MAX_FRAME)) {}
};

__global__ void __launch_bounds__(256, 4)
refine_groups_kernel(RefineParams R, const SegDesc *__restrict__ segs, const uint4 *__restrict__ cand,
                     const unsigned int *__restrict__ cand_count, unsigned int cand_cap,
                     float *__restrict__ spec, const void *__restrict__ pcm, int dtype, int channels,
                     unsigned int *__restrict__ file_min, unsigned int *__restrict__ n_recomputed) {
    const int lane = threadIdx.x & 31;
    const unsigned int n = min(*cand_count, cand_cap);
    const unsigned int warps = gridDim.x * (blockDim.x >> 5);
    const bool fast = dtype == NBM_PCM_INT16 && channels == 1;
    constexpr double MAGIC = 4503601774854144.0;            // 2^52 + 2^31
    const float floor_mag = (float)R.min_level;
    unsigned int c = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    uint4 ent = c < n ? cand[c] : make_uint4(0, 0, 0, 0);
    // The kernel is latency-bound (a block's pixels and its 2.9 KB of samples are scattered reads that left L2 long ago):
    // while a block is being worked on, the next one's samples and pixel rows are pulled into L2.
    auto prefetch = [&](const uint4 &e2) {
        const GroupKey g(entry_key(e2));
        const SegDesc sd = segs[g.seg];
        const long long s_lo = (long long)g.frame0 * R.hop - R.N / 2;
        const long long a = max(s_lo, 0ll), b = min(s_lo + R.N + R.hop, sd.n_samples);
        const size_t esz = (dtype == NBM_PCM_INT16 ? 2 : 4) * (size_t)channels;
        const char *p0 = reinterpret_cast<const char *>(pcm) + (size_t)(sd.pcm_start + a) * esz;
        const long long bytes = (b - a) * (long long)esz;
        for (long long o = (long long)lane * 128; o < bytes; o += 32 * 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(p0 + o));
        if (lane < g.rows)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(spec + sd.spec_off + (long long)(g.bin0 + lane) * sd.row_stride + g.frame0));
    };
    if (c < n) prefetch(ent);
    while (c < n) {
        const unsigned int cn = c + warps;
        const uint4 ent_next = cn < n ? cand[cn] : make_uint4(0, 0, 0, 0);
        if (cn < n) prefetch(ent_next);
        const GroupKey g(entry_key(ent));
        const int bin0 = g.bin0, frame0 = g.frame0;
        const SegDesc sd = segs[g.seg];
        const float th = __uint_as_float(ent.z);         // the level the transform kernel tested the block against
        const int row = lane >> 1, e = lane & 1;
        const bool mine = row < g.rows && e < g.nfr;
        float *px = spec + sd.spec_off + (long long)(bin0 + row) * sd.row_stride + frame0 + e;
        float v = mine ? *px : INFINITY;
        unsigned int todo = __ballot_sync(0xffffffffu, mine && v < th);
        const short *p16 = reinterpret_cast<const short *>(pcm) + sd.pcm_start;
        if (lane == 0 && todo) atomicAdd(n_recomputed, (unsigned int)__popc(todo));
        while (todo) {
            const int l = __ffs((int)todo) - 1;
            todo &= todo - 1;
            const long long k = R.low_idx + bin0 + (l >> 1);
            const long long s0 = (long long)(frame0 + (l & 1)) * R.hop - R.N / 2;
            const int qstep = (int)((k * 32) % R.N);
            int q = 0;                                       // (32 j k) mod N
            double ar = 0.0, ai = 0.0;
            constexpr int UB = 11;                          // samples in flight per lane: four rounds for n_fft = 1324
            if (fast && s0 >= 0 && s0 + R.N <= sd.n_samples) {
                // window inside the segment (all but its first and last five frames): no per-sample bounds tests, pointers
                // that advance by a constant, loads of a whole round issued before the first use
                const short *ps = p16 + s0 + lane;
                const double *ph = R.hann64 + lane;
                const double2 *tw = R.tw64;
                for (int i = lane; i < R.N; i += 32 * UB, ps += 32 * UB, ph += 32 * UB) {
                    int xi[UB];
                    double h[UB];
#pragma unroll
                    for (int u = 0; u < UB; ++u) {          // the last round is partial: a zero weight drops the sample
                        const bool in = i + 32 * u < R.N;
                        xi[u] = in ? (int)__ldg(ps + 32 * u) : 0;
                        h[u] = in ? ph[32 * u] : 0.0;
                    }
#pragma unroll
                    for (int u = 0; u < UB; ++u) {
                        const double2 w = tw[q];            // same address in every lane
                        const double x = (__hiloint2double(0x43300000, xi[u] ^ 0x80000000) - MAGIC) * h[u];
                        ar = fma(x, w.x, ar);
                        ai = fma(-x, w.y, ai);
                        q += qstep;
                        q -= q >= R.N ? R.N : 0;
                    }
                }
            } else {
                for (int i = lane; i < R.N; i += 32) {
                    const long long sx = s0 + i;
                    if (sx >= 0 && sx < sd.n_samples) {     // centre padding, pad_mode='constant'
                        // the reference multiplies the float32 sample (int16 / 32768, exact) by the float64 window: one rounding
                        const double xs = fast ? __hiloint2double(0x43300000, (int)__ldg(p16 + sx) ^ 0x80000000) - MAGIC
                                               : 32768.0 * (double)load_sample(pcm, dtype, channels, sd.pcm_start + sx);
                        const double x = xs * R.hann64[i];  // hann64 = window / 32768 (a power of two: still one rounding)
                        const double2 w = R.tw64[q];
                        ar = fma(x, w.x, ar);
                        ai = fma(-x, w.y, ai);
                    }
                    q += qstep;
                    q -= q >= R.N ? R.N : 0;
                }
            }
            {
                const double2 w = R.tw64[(int)((k * lane) % R.N)];                  // the lane's own factor e^{-i theta l}
                const double tr = ar * w.x + ai * w.y;
                ai = ai * w.x - ar * w.y;
                ar = tr;
            }
            ar = warp_sum(ar);
            ai = warp_sum(ai);
            const float re = (float)ar, im = (float)ai;                             // stored complex64
            // np.abs -> float32 (re^2 + im^2 is exact in double), then 20 log10(max(min_level, .)): the dB value only has to
            // be good to float32 (it is normalised and handed to the model as float32), so log2f, not the float64 routine
            const float mag = (float)sqrt((double)re * (double)re + (double)im * (double)im);
            const float db = 6.0205999132796239f * log2f(fmaxf(floor_mag, mag));
            if (lane == l) { v = db; *px = db; }
        }
        // the block's exact minimum (it was left out of the min/max partial)
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
        if (lane == 0) atomicMin(file_min + sd.file, float_to_ordered(v));
        c = cn;
        ent = ent_next;
    }
}

__device__ __forceinline__ double block_sum_256(double v, double *red) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();                                 // red[] free again
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w];
    return t;
}

// Whole-file min / max, one block per file.  s_min offsets EVERY normalised pixel, so it has to be the reference's own
// value, not a float32 approximation of it: the minimum is min(exact minimum of the listed blocks, partials of the rest),
// and the few unlisted pixels within `margin_db` of it (their float32 values are good to ~0.01 dB, the deep nulls having
// been listed) are recomputed like the listed ones, patched, and the minimum is taken over the exact values.
constexpr int MINMAX_CAP = 64;
__global__ void __launch_bounds__(256)
minmax_kernel(RefineParams R, const SegDesc *__restrict__ segs, const FileDesc *__restrict__ files,
              const float2 *__restrict__ tile_mm, const unsigned int *__restrict__ file_min, float *__restrict__ spec,
              const void *__restrict__ pcm, int dtype, int channels, float *__restrict__ out, int file_begin) {
    __shared__ float s_red[16];
    __shared__ double d_red[16];
    __shared__ int c_seg[MINMAX_CAP], c_bin[MINMAX_CAP], c_frame[MINMAX_CAP];
    __shared__ int n_cand, n_hot, hot[MINMAX_CAP];
    const int file = blockIdx.x + file_begin;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const FileDesc fd = files[file];
    const int per_gf = GF / R.mm_frames;
    const int n_ent = fd.n_groups * per_gf * R.mm_per_group;
    const float2 *mm = tile_mm + (size_t)fd.group0 * per_gf * R.mm_per_group;
    float vmin = INFINITY, vmax = -INFINITY;
    for (int i = tid; i < n_ent; i += 256) {
        const float2 v = mm[i];
        vmin = fminf(vmin, v.x);
        vmax = fmaxf(vmax, v.y);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        vmin = fminf(vmin, __shfl_xor_sync(0xffffffffu, vmin, o));
        vmax = fmaxf(vmax, __shfl_xor_sync(0xffffffffu, vmax, o));
    }
    if (lane == 0) { s_red[warp] = vmin; s_red[8 + warp] = vmax; }
    if (tid == 0) { n_cand = 0; n_hot = 0; }
    __syncthreads();
    vmin = s_red[0]; vmax = s_red[8];
#pragma unroll
    for (int w = 1; w < 8; ++w) { vmin = fminf(vmin, s_red[w]); vmax = fmaxf(vmax, s_red[8 + w]); }
    const unsigned int fm = file_min[file];
    const float listed_min = fm != 0xffffffffu ? ordered_to_float(fm) : INFINITY;     // exact
    vmax = fmaxf(vmax, listed_min == INFINITY ? -INFINITY : listed_min);             // a file whose every pixel was listed
    const float thr = fminf(vmin, listed_min) + R.margin_db;

    // partial groups whose minimum is within the margin (a handful), then their pixels
    for (int i = tid; i < n_ent; i += 256)
        if (mm[i].x <= thr) {
            const int at = atomicAdd(&n_hot, 1);
            if (at < MINMAX_CAP) hot[at] = i;
        }
    __syncthreads();
    const int nh = min(n_hot, MINMAX_CAP);
    const int n_slots = R.mm_per_group, slots_per_range = R.slots_per_range;
    for (int h = 0; h < nh; ++h) {
        const int i = hot[h];
        const int mg = fd.group0 * per_gf + i / n_slots, slot = i % n_slots;     // partial group (mm_frames frames)
        int si = fd.seg0 + fd.n_segs - 1;
        while (si > fd.seg0 && segs[si].group0 * per_gf > mg) --si;
        const SegDesc sd = segs[si];
        const int t0 = (mg - sd.group0 * per_gf) * R.mm_frames, nf = min(R.mm_frames, sd.n_frames - t0);
        const int rng = slot / slots_per_range, sl = slot % slots_per_range;
        const int b0 = rng * R.bins_per_range + sl * R.bins_per_slot;
        const int nb = min(min(R.bins_per_slot, R.bins_per_range - sl * R.bins_per_slot), R.n_bins - b0);
        for (int p = tid; p < nb * nf; p += 256) {
            const int b = b0 + p / nf, t = t0 + p % nf;
            if (spec[sd.spec_off + (long long)b * sd.row_stride + t] <= thr) {
                const int at = atomicAdd(&n_cand, 1);
                if (at < MINMAX_CAP) { c_seg[at] = si; c_bin[at] = b; c_frame[at] = t; }
            }
        }
    }
    __syncthreads();
    const int nc = min(n_cand, MINMAX_CAP);
    double best = INFINITY;
    for (int c = 0; c < nc; ++c) {
        const SegDesc sd = segs[c_seg[c]];
        const long long k = R.low_idx + c_bin[c];
        const long long s0 = (long long)c_frame[c] * R.hop - R.N / 2;
        double ar = 0.0, ai = 0.0;
        for (int n = tid; n < R.N; n += 256) {
            const long long s = s0 + n;
            if (s < 0 || s >= sd.n_samples) continue;           // centre padding, pad_mode='constant'
            const double x = (double)load_sample(pcm, dtype, channels, sd.pcm_start + s);
            const double xw = x * (0.5 - 0.5 * R.tw64[n].x);                        // periodic Hann
            const double2 w = R.tw64[(int)((k * n) % R.N)];
            ar = fma(xw, w.x, ar);
            ai = fma(-xw, w.y, ai);
        }
        ar = block_sum_256(ar, d_red);
        ai = block_sum_256(ai, d_red);
        const float re = (float)ar, im = (float)ai;                                 // stored complex64
        const float mag = (float)hypot((double)re, (double)im);                    // np.abs -> float32
        const double db = 20.0 * log10(fmax(R.min_level, (double)mag));
        best = fmin(best, db);
        if (tid == 0) spec[sd.spec_off + (long long)c_bin[c] * sd.row_stride + c_frame[c]] = (float)db;
    }
    if (tid == 0) {
        // every unlisted pixel within the margin was recomputed unless the list overflowed (e.g. digital silence)
        float smin = fminf(vmin, listed_min);
        if (nc > 0 && n_cand <= MINMAX_CAP && n_hot <= MINMAX_CAP) smin = fminf((float)best, listed_min);
        else if (nc > 0) smin = fminf(smin, (float)best);
        out[2 * file] = smin;
        out[2 * file + 1] = vmax;
    }
}

// Descriptor upload without the copy engine: a DMA copy on the caller's stream would queue behind whatever large
// H2D transfer (the next chunk of PCM) is in flight on the same engine and stall the kernels that wait for it.
// Pinned host memory is device-addressable under UVA, so a few warps simply read it.
__global__ void upload_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

// Normalise by the file's min/max and cut detector windows (prepare_dataset.py:248-250, 255-294);
// columns past the file's end mirror numpy's iterated 'reflect' pad of the partial window.
// (x - s_min) / (s_max - s_min) as reciprocal + one FMA Newton step: correctly rounded for these operands
// (in particular exactly 0 and 1 at the extremes) at a third of the cost of the IEEE divide sequence.
__device__ __forceinline__ float norm_div(float num, float den, float inv) {
    const float q = num * inv;
    const float r = fmaf(-q, den, num);
    // the reference's image lies in [0, 1]; only unrefined floor pixels of digital silence can leave it by an ulp
    return fminf(fmaxf(fmaf(r, inv, q), 0.0f), 1.0f);
}

__device__ __forceinline__ int reflect_src(int c, int width, int period) {
    if (c < width) return c;
    if (width == 1) return 0;
    const int r = c % period;
    return r < width ? r : period - r;
}

template <bool VEC4, bool L2_ONLY>
__device__ __forceinline__ void
tile_block(const KParams &P, const FileDesc &fd, long long tile, int row_block, float smin, float smax,
           const float *__restrict__ spec, float *__restrict__ tiles) {
    const int kt = (int)(tile - fd.tile0);
    const int start = kt * P.hop_spectro;
    const int width = (kt == fd.n_tiles - 1) ? fd.last_width : P.w_pix;
    const float range = smax - smin;
    const float inv = 1.0f / range;
    const int r0 = row_block * TILE_ROWS;
    const int r1 = min(r0 + TILE_ROWS, P.n_bins);
    const int period = 2 * (width - 1);
    const float *sbase = spec + fd.spec_off + start;
    float *tbase = tiles + (tile * P.n_bins) * P.w_pix;
    if (!(range > 0.0f)) {
        // A recording whose band is constant (digital silence: s_max == s_min): the reference's 0 / 0 is NaN in every pixel
        // (prepare_dataset.py:250), where norm_div's clamp would give 0.  Block-uniform, so the pixel loops below stay as
        // they are (one extra instruction per pixel there cost the pass 2.5 %).
        const float nanv = __int_as_float(0x7fc00000);
        for (int r = r0; r < r1; ++r)
            for (int c = threadIdx.x; c < P.w_pix; c += blockDim.x) tbase[(long long)r * P.w_pix + c] = nanv;
        return;
    }
    // L2_ONLY: the band was written by other SMs during this very launch; read it from L2, never through L1
    auto ld = [](const float *p) { return L2_ONLY ? __ldcg(p) : *p; };
    if (VEC4 && width == P.w_pix) {
        // Full tile: four scalar loads per thread would each touch a quarter of every 32-byte sector (4x the L1 look-ups,
        // and 4x the L2 -> SM traffic where L1 is bypassed).  Rows are 128-byte aligned (row_stride % 32 == 0), so
        // the tile's misalignment is start & 3 for every row: two aligned 16-byte loads, re-aligned in registers.
        const int sh = start & 3;
        auto rows = [&](auto SH) {
            constexpr int S = decltype(SH)::value;
            for (int c = 4 * threadIdx.x; c < P.w_pix; c += 4 * blockDim.x) {
#pragma unroll 5
                for (int r = r0; r < r1; ++r) {
                    const float4 *q = reinterpret_cast<const float4 *>(sbase + (long long)r * fd.row_stride + c - S);
                    const float4 a = L2_ONLY ? __ldcg(q) : __ldg(q);
                    float4 b = a;
                    if (S != 0) b = L2_ONLY ? __ldcg(q + 1) : __ldg(q + 1);
                    float4 v;
                    if (S == 0) v = a;
                    else if (S == 1) v = make_float4(a.y, a.z, a.w, b.x);
                    else if (S == 2) v = make_float4(a.z, a.w, b.x, b.y);
                    else v = make_float4(a.w, b.x, b.y, b.z);
                    float4 o;
                    o.x = norm_div(v.x - smin, range, inv);
                    o.y = norm_div(v.y - smin, range, inv);
                    o.z = norm_div(v.z - smin, range, inv);
                    o.w = norm_div(v.w - smin, range, inv);
                    __stcs(reinterpret_cast<float4 *>(tbase + (long long)r * P.w_pix + c), o);
                }
            }
        };
        if (sh == 0) rows(std::integral_constant<int, 0>{});
        else if (sh == 1) rows(std::integral_constant<int, 1>{});
        else if (sh == 2) rows(std::integral_constant<int, 2>{});
        else rows(std::integral_constant<int, 3>{});
    } else if (VEC4) {
        for (int c = 4 * threadIdx.x; c < P.w_pix; c += 4 * blockDim.x) {
            int src[4];
            if (c + 3 < width) { src[0] = c; src[1] = c + 1; src[2] = c + 2; src[3] = c + 3; }
            else {
#pragma unroll
                for (int e = 0; e < 4; ++e) src[e] = reflect_src(c + e, width, period);
            }
#pragma unroll 5
            for (int r = r0; r < r1; ++r) {
                const float *sp = sbase + (long long)r * fd.row_stride;
                float4 o;
                o.x = norm_div(ld(sp + src[0]) - smin, range, inv);
                o.y = norm_div(ld(sp + src[1]) - smin, range, inv);
                o.z = norm_div(ld(sp + src[2]) - smin, range, inv);
                o.w = norm_div(ld(sp + src[3]) - smin, range, inv);
                __stcs(reinterpret_cast<float4 *>(tbase + (long long)r * P.w_pix + c), o);
            }
        }
    } else {
        for (int c = threadIdx.x; c < P.w_pix; c += blockDim.x) {
            const int src = reflect_src(c, width, period);
            for (int r = r0; r < r1; ++r)
                tbase[(long long)r * P.w_pix + c] = norm_div(ld(sbase + (long long)r * fd.row_stride + src) - smin, range, inv);
        }
    }
}

__device__ __forceinline__ int file_of_tile(const FileDesc *__restrict__ files, int n_files, long long tile) {
    int lo_f = 0, hi_f = n_files - 1;
    while (lo_f < hi_f) {
        int mid = (lo_f + hi_f + 1) >> 1;
        if (files[mid].tile0 <= tile) lo_f = mid; else hi_f = mid - 1;
    }
    return lo_f;
}

template <bool VEC4>
__global__ void __launch_bounds__(256, 8)
tile_kernel(KParams P, const FileDesc *__restrict__ files, int n_files, const float *__restrict__ spec,
            const float *__restrict__ minmax, float *__restrict__ tiles, long long tile_begin) {
    const long long tile = blockIdx.x + tile_begin;
    const int f = file_of_tile(files, n_files, tile);
    tile_block<VEC4, false>(P, files[f], tile, blockIdx.y, minmax[2 * f], minmax[2 * f + 1], spec, tiles);
}

// Dataset images (prepare_dataset.py:85: np.round(img * 255).astype(np.uint8)): round-half-even of the normalised
// tile value times 255.  Pure streaming pass, 4 B read + 1 B written per pixel.
__global__ void __launch_bounds__(256)
tiles_to_u8_kernel(const float4 *__restrict__ src, uchar4 *__restrict__ dst, long long n4,
                   const float *__restrict__ tail_src, unsigned char *__restrict__ tail_dst, int n_tail) {
    auto q = [](float v) { return (unsigned char)__float2int_rn(fminf(fmaxf(v, 0.0f), 1.0f) * 255.0f); };
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
        const float4 v = __ldcs(src + i);
        dst[i] = make_uchar4(q(v.x), q(v.y), q(v.z), q(v.w));
    }
    if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) tail_dst[threadIdx.x] = q(tail_src[threadIdx.x]);
}

}  // namespace nbm

// ------------------------------------------------------------------------------ host side ------
using namespace nbm;

struct nbm_frontend_plan {
    nbm_frontend_params prm;
    KParams kp;
    float2 *d_tw = nullptr;
    int device = 0;
    int n_threads = 0;
    size_t smem_bytes = 0;
    TcPlan *tc = nullptr;           // tensor-core path (nullptr -> CUDA-core stft_db_kernel)
    // pinned staging for descriptor uploads
    std::mutex mu;
    void *h_stage = nullptr;
    size_t h_stage_bytes = 0;
    cudaEvent_t staged = nullptr;
    // optional per-kernel timing (nbm_frontend_set_profiling)
    bool profiling = false;
    cudaEvent_t ev[5] = {};                    // anchors | slides (or the CUDA-core STFT) | refinement + min/max | tiles

    double acc_ms[4] = {0.0, 0.0, 0.0, 0.0};
    long long acc_runs = 0;
    bool ev_pending = false;
    // the transform runs on a high-priority side stream
    cudaStream_t s_hi = nullptr;
    cudaEvent_t ev_in = nullptr, ev_done = nullptr;
    double2 *d_tw64 = nullptr;  // [n_fft] float64 twiddles of the refinement pass
    double *d_hann64 = nullptr; // [n_fft] float64 periodic Hann window / 32768
    float flag_rel_db = -62.0f; // a pixel this far below the largest |R| its bin carried is recomputed in float64 (NBM_REFINE_REL_DB)
    double cand_frac = 0.004;   // candidate list capacity as a fraction of the batch's pixels
    const unsigned int *last_count = nullptr;   // device counter of the last run's list (nbm_frontend_last_listed)
    unsigned int last_cap = 0;
    cudaStream_t last_stream = nullptr;
};

namespace {

// Everything the host derives from the per-file sample counts: STFT chunks (prepare_dataset.py:236),
// detector windows (:266) with the last window's valid width (:268-278 incl. the seam quirk), the
// 64-frame groups, and the workspace carve-up
//   [segs | files | per-file exact minima + list counters | min/max partials | refinement list | anchors | dB bands].
struct BatchLayout {
    std::vector<SegDesc> segs;
    std::vector<FileDesc> files;
    std::vector<int64_t> n_frames;
    size_t spec_floats = 0;
    long long tiles = 0, n_anchors = 0;
    int groups = 0;
    size_t o_segs = 0, o_files = 0, o_cnt = 0, o_mm = 0, o_cand = 0, o_anchors = 0, o_spec = 0, total = 0;
    unsigned int cand_cap = 0;
    size_t upload_bytes = 0;       // segs and files are uploaded from the host
};

int build_layout(const nbm_frontend_plan *pl, const int64_t *sizes, const int64_t *offsets, int n_files, BatchLayout &B) {
    const nbm_frontend_params &p = pl->prm;
    B.files.resize(n_files);
    B.n_frames.resize(n_files);
    for (int f = 0; f < n_files; ++f) {
        const int64_t n = sizes ? sizes[f] : offsets[f + 1] - offsets[f];
        NBM_REQUIRE(n >= 0, "negative sample count for file %d", f);
        FileDesc &fd = B.files[f];
        fd.spec_off = (long long)B.spec_floats;
        fd.tile0 = B.tiles;
        fd.group0 = B.groups;
        fd.seg0 = (int)B.segs.size();
        long long T = 0, col = 0, s = offsets ? offsets[f] : 0;
        const int64_t n_chunks = n / p.stft_chunk + 1;              // range(int(len/max_l)+1)
        const size_t seg_first = B.segs.size();
        for (int64_t c = 0; c < n_chunks; ++c) {
            const int64_t len = std::max<int64_t>(0, std::min<int64_t>(n, (c + 1) * p.stft_chunk) - c * p.stft_chunk);
            SegDesc sd;
            sd.pcm_start = s;
            sd.n_samples = len;
            sd.spec_off = 0;                                         // patched below (needs row_stride)
            sd.n_frames = (int)(1 + len / p.hop);
            NBM_REQUIRE(sd.n_frames <= GROUP_MAX_FRAME, "STFT chunk of file %d has too many frames", f);
            sd.row_stride = 0;
            sd.file = f;
            sd.group0 = B.groups;
            const int tiles_seg = (sd.n_frames + GF - 1) / GF;
            B.groups += tiles_seg;
            B.n_anchors += tiles_seg + 1;
            B.segs.push_back(sd);
            T += sd.n_frames;
            s += len;
        }
        NBM_REQUIRE(T < (1ll << 31) - 64, "file %d too long", f);
        const long long stride = (long long)align_up((size_t)T, 32);
        for (size_t i = seg_first; i < B.segs.size(); ++i) {
            B.segs[i].spec_off = fd.spec_off + col;
            B.segs[i].row_stride = (int)stride;
            col += B.segs[i].n_frames;
        }
        long long nt = 1;                                            // max(1, int(1 + ceil((T - w_pix) / hop_spectro)))
        if (T > p.w_pix) nt = 1 + (T - p.w_pix + p.hop_spectro - 1) / p.hop_spectro;
        const long long start = (nt - 1) * (long long)p.hop_spectro;
        int last_width = p.w_pix;
        if (start + p.w_pix > T) {
            // a window that starts in chunk c and runs past the end of the FILE keeps chunk c's columns only
            long long edge = 0;
            size_t i = seg_first;
            for (; i < B.segs.size(); ++i) {
                if (start < edge + B.segs[i].n_frames) break;
                edge += B.segs[i].n_frames;
            }
            last_width = (int)(edge + B.segs[i].n_frames - start);
        }
        fd.n_groups = B.groups - fd.group0;
        fd.n_segs = (int)B.segs.size() - fd.seg0;
        fd.row_stride = (int)stride;
        fd.n_tiles = (int)nt;
        fd.total_frames = (int)T;
        fd.last_width = last_width;
        B.n_frames[f] = T;
        B.spec_floats += (size_t)stride * p.n_bins;
        B.tiles += nt;
    }
    size_t o = 0;
    auto take = [&](size_t bytes) { const size_t at = o; o = align_up(o + bytes, 256); return at; };
    B.o_segs = take(B.segs.size() * sizeof(SegDesc));
    B.o_files = take(B.files.size() * sizeof(FileDesc));
    B.upload_bytes = o;
    B.o_cnt = take(((size_t)n_files + 2) * sizeof(unsigned int));    // per-file exact minimum (ordered uint) | blocks listed | pixels recomputed
    B.o_mm = take((size_t)B.groups * (pl->tc ? (GF / tc_chain_frames()) * tc_n_ranges(pl->tc) * tc_slots_per_range() : 1) *
                  sizeof(float2));                                  // min/max partials
    B.cand_cap = (unsigned int)std::min<double>(64.0 * 1024 * 1024, std::max<double>(65536.0, pl->cand_frac * (double)B.spec_floats));
    B.o_cand = take((size_t)B.cand_cap * sizeof(uint4));
    B.o_anchors = take(pl->tc ? tc_anchor_bytes(pl->tc, B.n_anchors) : 0);
    B.o_spec = take(B.spec_floats * sizeof(float) + 16);
    B.total = o;
    return NBM_OK;
}

}  // namespace

extern "C" int nbm_frontend_plan_create(const nbm_frontend_params *p, nbm_frontend_plan **out) {
    NBM_REQUIRE(p && out, "null argument");
    NBM_REQUIRE(p->n_fft >= 8 && p->hop >= 1 && p->hop <= p->n_fft, "need 1 <= hop <= n_fft, n_fft >= 8");
    NBM_REQUIRE(p->low_idx >= 1 && p->n_bins >= 1 && p->low_idx + p->n_bins <= p->n_fft / 2 + 1,
                "band [low_idx, low_idx+n_bins) must lie in [1, n_fft/2]");
    NBM_REQUIRE(p->w_pix >= 1 && p->hop_spectro >= 1 && p->stft_chunk >= p->n_fft, "bad tiling parameters");
    if (p->pad_mode != 0) {
        set_error("pad_mode %d not supported (only 0 = 'constant', the librosa>=0.10 default)", p->pad_mode);
        return NBM_ERR_UNSUPPORTED;
    }
    const int n_warps = (p->n_bins + BINS_PER_WARP - 1) / BINS_PER_WARP;
    NBM_REQUIRE(n_warps <= 16, "n_bins too large for one block (max 480)");
    auto *pl = new nbm_frontend_plan();
    pl->prm = *p;
    KParams &k = pl->kp;
    k.N = p->n_fft; k.hop = p->hop; k.low_idx = p->low_idx; k.n_bins = p->n_bins;
    k.w_pix = p->w_pix; k.hop_spectro = p->hop_spectro;
    k.npN = (p->n_fft + 1) / 2; k.npH = (p->hop + 1) / 2;
    k.buf_len = p->n_fft + (GF - 1) * p->hop;
    k.min_level_sq = (float)(p->min_level * p->min_level);
    pl->n_threads = 32 * n_warps;
    pl->smem_bytes = (size_t)((k.buf_len + 3) & ~3) * 4 + (size_t)((k.npN + 1) & ~1) * 8 +
                     (size_t)k.npH * PASS * 8 + (size_t)std::max(p->n_bins * STAGE_LD, 64) * 4;
    cudaError_t e = cudaGetDevice(&pl->device);
    if (e != cudaSuccess) { delete pl; return cuda_fail(e, "cudaGetDevice"); }
    int max_smem = 0;
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, pl->device);
    {
        cudaFuncAttributes fa{};
        if (cudaFuncGetAttributes(&fa, stft_db_kernel) == cudaSuccess) max_smem -= (int)fa.sharedSizeBytes;   // static part
    }
    if ((size_t)max_smem < pl->smem_bytes) {
        set_error("front-end needs %zu B of shared memory per block, device offers %d", pl->smem_bytes, max_smem);
        delete pl;
        return NBM_ERR_UNSUPPORTED;
    }
    const int N2 = 2 * p->n_fft;
    std::vector<float2> tw(N2);
    for (int q = 0; q < N2; ++q) {
        const double ang = M_PI * (double)q / (double)p->n_fft;
        tw[q] = make_float2((float)cos(ang), (float)sin(ang));
    }
    if ((e = cudaMalloc(&pl->d_tw, N2 * sizeof(float2))) != cudaSuccess ||
        (e = cudaMemcpy(pl->d_tw, tw.data(), N2 * sizeof(float2), cudaMemcpyHostToDevice)) != cudaSuccess ||
        (e = cudaEventCreateWithFlags(&pl->staged, cudaEventDisableTiming)) != cudaSuccess ||
        // the attribute is per function, not per plan: always allow the device maximum
        (e = cudaFuncSetAttribute(stft_db_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  max_smem)) != cudaSuccess) {
        int rc = cuda_fail(e, "plan_create");
        nbm_frontend_plan_destroy(pl);
        return rc;
    }
    k.tw = pl->d_tw;
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);             // hi = numerically lowest = greatest priority
        e = cudaStreamCreateWithPriority(&pl->s_hi, cudaStreamNonBlocking, hi);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_in, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&pl->ev_done, cudaEventDisableTiming);
        if (e != cudaSuccess) { int rc = cuda_fail(e, "plan_create(streams)"); nbm_frontend_plan_destroy(pl); return rc; }
        const char *rd = getenv("NBM_REFINE_REL_DB");          // diagnostic: how far below the frame's level a pixel is flagged
        if (rd) pl->flag_rel_db = (float)atof(rd);
        const char *cf = getenv("NBM_REFINE_CAND_FRAC");
        if (cf) pl->cand_frac = std::min(0.5, std::max(1e-6, atof(cf)));
    }
    {
        // float64 twiddles of the refinement pass
        std::vector<double2> t64(p->n_fft);
        for (int q = 0; q < p->n_fft; ++q) {
            const double ang = 2.0 * M_PI * (double)q / (double)p->n_fft;
            t64[q] = make_double2(cos(ang), sin(ang));
        }
        // scipy.signal.get_window('hann', N, fftbins=True) = 0.5 - 0.5 cos(2 pi n / N), here pre-divided by 32768
        std::vector<double> h64(p->n_fft);
        for (int q = 0; q < p->n_fft; ++q) h64[q] = (0.5 - 0.5 * cos(2.0 * M_PI * (double)q / (double)p->n_fft)) / 32768.0;
        if ((e = cudaMalloc(&pl->d_tw64, t64.size() * sizeof(double2))) != cudaSuccess ||
            (e = cudaMemcpy(pl->d_tw64, t64.data(), t64.size() * sizeof(double2), cudaMemcpyHostToDevice)) != cudaSuccess ||
            (e = cudaMalloc(&pl->d_hann64, h64.size() * sizeof(double))) != cudaSuccess ||
            (e = cudaMemcpy(pl->d_hann64, h64.data(), h64.size() * sizeof(double), cudaMemcpyHostToDevice)) != cudaSuccess) {
            int rc = cuda_fail(e, "plan_create(tw64)");
            nbm_frontend_plan_destroy(pl);
            return rc;
        }
    }
    // tensor-core path unless the parameters do not fit it or NBM_FRONTEND_IMPL=cuda-core asks for the other one
    const char *impl = getenv("NBM_FRONTEND_IMPL");
    if (!(impl && strcmp(impl, "cuda-core") == 0)) {
        int rc = tc_plan_create(*p, &pl->tc);
        if (rc != NBM_OK && rc != NBM_ERR_UNSUPPORTED) { nbm_frontend_plan_destroy(pl); return rc; }
    }
    *out = pl;
    return NBM_OK;
}

extern "C" int nbm_frontend_plan_destroy(nbm_frontend_plan *pl) {
    if (!pl) return NBM_OK;
    if (pl->d_tw) cudaFree(pl->d_tw);
    if (pl->h_stage) cudaFreeHost(pl->h_stage);
    if (pl->staged) cudaEventDestroy(pl->staged);
    for (auto &e : pl->ev) if (e) cudaEventDestroy(e);
    if (pl->ev_in) cudaEventDestroy(pl->ev_in);
    if (pl->ev_done) cudaEventDestroy(pl->ev_done);
    if (pl->d_tw64) cudaFree(pl->d_tw64);
    if (pl->d_hann64) cudaFree(pl->d_hann64);
    if (pl->s_hi) cudaStreamDestroy(pl->s_hi);
    tc_plan_destroy(pl->tc);
    delete pl;
    return NBM_OK;
}

extern "C" int nbm_frontend_impl(const nbm_frontend_plan *pl) { return pl && pl->tc ? 1 : 0; }

extern "C" int nbm_frontend_query_batch(const nbm_frontend_plan *pl, const int64_t *n_samples, int32_t n_files,
                                        int64_t *n_frames, int64_t *tile_offsets, size_t *workspace_bytes) {
    NBM_REQUIRE(pl && n_samples && n_files >= 1, "bad argument");
    BatchLayout B;
    int rc = build_layout(pl, n_samples, nullptr, n_files, B);
    if (rc != NBM_OK) return rc;
    for (int f = 0; f < n_files; ++f) {
        if (n_frames) n_frames[f] = B.n_frames[f];
        if (tile_offsets) tile_offsets[f] = B.files[f].tile0;
    }
    if (tile_offsets) tile_offsets[n_files] = B.tiles;
    if (workspace_bytes) *workspace_bytes = B.total;
    return NBM_OK;
}

extern "C" int nbm_frontend_query(const nbm_frontend_plan *pl, int64_t n_samples, int64_t *n_frames,
                                  int64_t *n_tiles, size_t *workspace_bytes) {
    int64_t off[2] = {0, 0};
    int rc = nbm_frontend_query_batch(pl, &n_samples, 1, n_frames, off, workspace_bytes);
    if (rc == NBM_OK && n_tiles) *n_tiles = off[1];
    return rc;
}

extern "C" int nbm_frontend_spectrogram_view(const nbm_frontend_plan *pl, const int64_t *n_samples, int32_t n_files,
                                             int32_t file_index, size_t *offset_bytes, int64_t *row_stride) {
    NBM_REQUIRE(pl && n_samples && file_index >= 0 && file_index < n_files, "bad argument");
    BatchLayout B;
    int rc = build_layout(pl, n_samples, nullptr, n_files, B);
    if (rc != NBM_OK) return rc;
    if (offset_bytes) *offset_bytes = B.o_spec + (size_t)B.files[file_index].spec_off * sizeof(float);
    if (row_stride) *row_stride = B.files[file_index].row_stride;
    return NBM_OK;
}

// fold the previous profiled run's event pairs into the accumulators (waits for that run)
static int collect_profile(nbm_frontend_plan *pl) {
    if (!pl->ev_pending) return NBM_OK;
    NBM_CUDA(cudaEventSynchronize(pl->ev[4]));
    for (int i = 0; i < 4; ++i) {
        float ms = 0.f;
        NBM_CUDA(cudaEventElapsedTime(&ms, pl->ev[i], pl->ev[i + 1]));
        pl->acc_ms[i] += ms;
    }
    pl->acc_runs += 1;
    pl->ev_pending = false;
    return NBM_OK;
}

extern "C" int nbm_frontend_set_profiling(nbm_frontend_plan *pl, int32_t enable) {
    NBM_REQUIRE(pl, "null plan");
    std::lock_guard<std::mutex> lock(pl->mu);
    if (enable && !pl->ev[0])
        for (auto &e : pl->ev) NBM_CUDA(cudaEventCreate(&e));
    pl->profiling = enable != 0;
    for (auto &a : pl->acc_ms) a = 0.0;
    pl->acc_runs = 0; pl->ev_pending = false;
    return NBM_OK;
}

extern "C" int nbm_frontend_get_profile(nbm_frontend_plan *pl, double *stft_ms, double *tile_ms, int64_t *runs) {
    NBM_REQUIRE(pl, "null plan");
    std::lock_guard<std::mutex> lock(pl->mu);
    int rc = collect_profile(pl);
    if (rc != NBM_OK) return rc;
    if (stft_ms) *stft_ms = pl->acc_ms[0] + pl->acc_ms[1] + pl->acc_ms[2];
    if (tile_ms) *tile_ms = pl->acc_ms[3];
    if (runs) *runs = pl->acc_runs;
    return NBM_OK;
}

extern "C" int nbm_frontend_get_profile_kernels(nbm_frontend_plan *pl, double *ms4, int64_t *runs) {
    NBM_REQUIRE(pl && ms4, "null argument");
    std::lock_guard<std::mutex> lock(pl->mu);
    int rc = collect_profile(pl);
    if (rc != NBM_OK) return rc;
    for (int i = 0; i < 4; ++i) ms4[i] = pl->acc_ms[i];
    if (runs) *runs = pl->acc_runs;
    return NBM_OK;
}

extern "C" int nbm_frontend_run_batch(const nbm_frontend_plan *cpl, const void *d_pcm, int32_t pcm_dtype,
                                      int32_t channels, const int64_t *sample_offsets, int32_t n_files,
                                      float *d_tiles, float *d_minmax, void *d_workspace, size_t workspace_bytes,
                                      void *stream_) {
    auto *pl = const_cast<nbm_frontend_plan *>(cpl);
    NBM_REQUIRE(pl && d_pcm && sample_offsets && d_tiles && d_minmax && d_workspace, "null argument");
    NBM_REQUIRE(n_files >= 1 && channels >= 1, "need n_files >= 1, channels >= 1");
    NBM_REQUIRE(pcm_dtype == NBM_PCM_INT16 || pcm_dtype == NBM_PCM_FLOAT32, "unknown pcm_dtype %d", pcm_dtype);
    cudaStream_t stream = (cudaStream_t)stream_;
    const nbm_frontend_params &p = pl->prm;

    BatchLayout B;
    int rc = build_layout(pl, nullptr, sample_offsets, n_files, B);
    if (rc != NBM_OK) return rc;
    if (workspace_bytes < B.total) {
        set_error("workspace too small: %zu < %zu", workspace_bytes, B.total);
        return NBM_ERR_WORKSPACE;
    }
    char *ws = reinterpret_cast<char *>(d_workspace);
    SegDesc *d_segs = reinterpret_cast<SegDesc *>(ws + B.o_segs);
    FileDesc *d_files = reinterpret_cast<FileDesc *>(ws + B.o_files);
    float2 *d_tile_mm = reinterpret_cast<float2 *>(ws + B.o_mm);
    float *d_spec = reinterpret_cast<float *>(ws + B.o_spec);

    std::lock_guard<std::mutex> lock(pl->mu);
    {
        const size_t up = B.upload_bytes;
        if (pl->h_stage_bytes < up) {
            if (pl->h_stage) { NBM_CUDA(cudaEventSynchronize(pl->staged)); cudaFreeHost(pl->h_stage); pl->h_stage = nullptr; }
            NBM_CUDA(cudaMallocHost(&pl->h_stage, up + 16));
            pl->h_stage_bytes = up;
        } else {
            NBM_CUDA(cudaEventSynchronize(pl->staged));     // previous upload has left the staging buffer
        }
        char *h = reinterpret_cast<char *>(pl->h_stage);
        memcpy(h + B.o_segs, B.segs.data(), B.segs.size() * sizeof(SegDesc));
        memcpy(h + B.o_files, B.files.data(), B.files.size() * sizeof(FileDesc));
        const size_t n16 = (up + 15) / 16;
        upload_kernel<<<(unsigned)std::min<size_t>(64, (n16 + 255) / 256), 256, 0, stream>>>(
            reinterpret_cast<uint4 *>(ws), reinterpret_cast<const uint4 *>(pl->h_stage), n16);
        NBM_CUDA(cudaEventRecord(pl->staged, stream));
    }
    const bool prof = pl->profiling;
    if (prof) {
        rc = collect_profile(pl);
        if (rc != NBM_OK) return rc;
    }
    // the tensor-core kernels stage mono PCM16; float or multi-channel input takes the CUDA-core kernel
    const bool use_tc = pl->tc && pcm_dtype == NBM_PCM_INT16 && channels == 1;
    RefineParams rp;
    rp.N = p.n_fft; rp.hop = p.hop; rp.low_idx = p.low_idx; rp.n_bins = p.n_bins;
    rp.mm_frames = use_tc ? tc_chain_frames() : GF;
    rp.mm_per_group = use_tc ? tc_n_ranges(pl->tc) * tc_slots_per_range() : 1;
    rp.slots_per_range = use_tc ? tc_slots_per_range() : 1;
    rp.bins_per_range = use_tc ? tc_bins_per_range() : p.n_bins;
    rp.bins_per_slot = use_tc ? tc_bins_per_slot() : p.n_bins;
    rp.margin_db = 0.25f;
    rp.min_level = p.min_level;
    rp.n_ranges = use_tc ? tc_n_ranges(pl->tc) : 0;
    rp.tw64 = pl->d_tw64;
    rp.hann64 = pl->d_hann64;
    auto *d_cand = reinterpret_cast<uint4 *>(ws + B.o_cand);
    unsigned int *d_file_min = reinterpret_cast<unsigned int *>(ws + B.o_cnt), *d_cand_count = d_file_min + n_files;

    // caller's stream: descriptors ............................................................. | tiles
    // side stream:                 | anchors -> slides -> float64 refinement -> min/max |
    NBM_CUDA(cudaEventRecord(pl->ev_in, stream));              // inputs and descriptors are ready on the caller's stream
    cudaStream_t sc = pl->s_hi;
    NBM_CUDA(cudaStreamWaitEvent(sc, pl->ev_in, 0));
    NBM_CUDA(cudaMemsetAsync(d_file_min, 0xff, (size_t)n_files * sizeof(unsigned int), sc));
    NBM_CUDA(cudaMemsetAsync(d_cand_count, 0, 2 * sizeof(unsigned int), sc));
    if (prof) NBM_CUDA(cudaEventRecord(pl->ev[0], sc));
    if (use_tc) {
        // globally numbered anchors: segment s owns group0[s] + s .. group0[s] + s + tiles[s]
        rc = tc_launch_anchors(pl->tc, d_segs, 0, (int)B.segs.size(), 0, B.n_anchors, d_pcm, ws + B.o_anchors, sc);
        if (rc != NBM_OK) return rc;
        if (prof) NBM_CUDA(cudaEventRecord(pl->ev[1], sc));
        rc = tc_launch_slides(pl->tc, d_segs, (int)B.segs.size(), 0, 0, B.groups, d_pcm, d_spec, d_tile_mm,
                              ws + B.o_anchors, pl->flag_rel_db, d_cand, d_cand_count, B.cand_cap, sc);
        if (rc != NBM_OK) return rc;
    } else {
        if (prof) NBM_CUDA(cudaEventRecord(pl->ev[1], sc));
        stft_db_kernel<<<B.groups, pl->n_threads, pl->smem_bytes, sc>>>(pl->kp, d_segs, (int)B.segs.size(), d_pcm, pcm_dtype,
                                                                        channels, d_spec, d_tile_mm, 0, pl->flag_rel_db, d_cand,
                                                                        d_cand_count, B.cand_cap);
    }
    if (prof) NBM_CUDA(cudaEventRecord(pl->ev[2], sc));
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, pl->device);
        refine_groups_kernel<<<sms * 4, 256, 0, sc>>>(rp, d_segs, d_cand, d_cand_count, B.cand_cap, d_spec, d_pcm,
                                                     pcm_dtype, channels, d_file_min, d_cand_count + 1);
        minmax_kernel<<<n_files, 256, 0, sc>>>(rp, d_segs, d_files, d_tile_mm, d_file_min, d_spec, d_pcm, pcm_dtype, channels,
                                               d_minmax, 0);
    }
    pl->last_count = d_cand_count; pl->last_cap = B.cand_cap; pl->last_stream = stream;
    NBM_CUDA(cudaEventRecord(pl->ev_done, sc));
    NBM_CUDA(cudaStreamWaitEvent(stream, pl->ev_done, 0));
    if (prof) NBM_CUDA(cudaEventRecord(pl->ev[3], stream));
    dim3 grid((unsigned)B.tiles, (unsigned)((p.n_bins + TILE_ROWS - 1) / TILE_ROWS));
    if (p.w_pix % 4 == 0) tile_kernel<true><<<grid, 256, 0, stream>>>(pl->kp, d_files, n_files, d_spec, d_minmax, d_tiles, 0);
    else tile_kernel<false><<<grid, 256, 0, stream>>>(pl->kp, d_files, n_files, d_spec, d_minmax, d_tiles, 0);
    if (prof) { NBM_CUDA(cudaEventRecord(pl->ev[4], stream)); pl->ev_pending = true; }
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}

extern "C" int nbm_frontend_last_listed(nbm_frontend_plan *pl, int64_t *listed, int64_t *capacity, int64_t *recomputed) {
    NBM_REQUIRE(pl && listed && capacity, "null argument");
    std::lock_guard<std::mutex> lock(pl->mu);
    NBM_REQUIRE(pl->last_count, "no run yet");
    unsigned int n[2] = {0, 0};
    NBM_CUDA(cudaStreamSynchronize(pl->last_stream));
    NBM_CUDA(cudaMemcpy(n, pl->last_count, sizeof(n), cudaMemcpyDeviceToHost));
    *listed = n[0];
    *capacity = pl->last_cap;
    if (recomputed) *recomputed = n[1];
    return NBM_OK;
}

extern "C" int nbm_frontend_run(const nbm_frontend_plan *pl, const void *d_pcm, int32_t pcm_dtype, int32_t channels,
                                int64_t n_samples, float *d_tiles, float *d_minmax, void *d_workspace,
                                size_t workspace_bytes, void *stream) {
    int64_t off[2] = {0, n_samples};
    return nbm_frontend_run_batch(pl, d_pcm, pcm_dtype, channels, off, 1, d_tiles, d_minmax, d_workspace,
                                  workspace_bytes, stream);
}

extern "C" int nbm_tiles_to_u8(const float *d_tiles, int64_t n_values, uint8_t *d_out, void *stream_) {
    NBM_REQUIRE(n_values >= 0, "bad argument");
    if (n_values == 0) return NBM_OK;
    NBM_REQUIRE(d_tiles && d_out, "bad argument");
    NBM_REQUIRE((reinterpret_cast<uintptr_t>(d_tiles) & 15) == 0 && (reinterpret_cast<uintptr_t>(d_out) & 3) == 0,
                "tile buffers must be 16-byte (float) / 4-byte (uint8) aligned");
    cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
    const long long n4 = n_values / 4;
    const int n_tail = (int)(n_values - 4 * n4);
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long want = (n4 + 255) / 256;
    const int grid = (int)std::max<long long>(1, std::min<long long>(want, (long long)sms * 8));
    nbm::tiles_to_u8_kernel<<<grid, 256, 0, stream>>>(reinterpret_cast<const float4 *>(d_tiles), reinterpret_cast<uchar4 *>(d_out), n4,
                                                      d_tiles + 4 * n4, d_out + 4 * n4, n_tail);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}
