// Tensor-core (tcgen05 / TMEM) front-end path, see frontend_tc.cu.
#pragma once
#include "common.cuh"

namespace nbm {

constexpr int GF = 64;      // frames per group / tile (one anchor per tile boundary)

struct SegDesc {
    long long pcm_start;   // per-channel sample index of the segment's first sample
    long long n_samples;   // samples in this STFT chunk (prepare_dataset.py:236-237)
    long long spec_off;    // float offset of S[0][first column of this segment]
    int n_frames;
    int row_stride;
    int file;
    int group0;            // index of the segment's first 64-frame group (tile)
};

// One entry of the refinement list: a block of `rows` bins x `nfr` frames of segment `seg`, first pixel (bin0, frame).
// seg 22 bits | bin0 10 | rows-1 4 | nfr-1 1 | frame 27
__host__ __device__ inline unsigned long long pack_group(int seg, int bin0, int rows, int nfr, int frame) {
    return ((unsigned long long)seg << 42) | ((unsigned long long)bin0 << 32) | ((unsigned long long)(rows - 1) << 28) |
           ((unsigned long long)(nfr - 1) << 27) | (unsigned long long)(unsigned int)frame;
}
constexpr int GROUP_MAX_FRAME = (1 << 27) - 1;
// list entry = the packed block + the flag level it was tested against (the refinement pass tests the block's pixels
// against the same level)
__host__ __device__ inline uint4 list_entry(unsigned long long key, float level) {
#ifdef __CUDA_ARCH__
    return make_uint4((unsigned int)key, (unsigned int)(key >> 32), __float_as_uint(level), 0u);
#else
    unsigned int b;
    memcpy(&b, &level, 4);
    return make_uint4((unsigned int)key, (unsigned int)(key >> 32), b, 0u);
#endif
}
__host__ __device__ inline unsigned long long entry_key(const uint4 &e) { return ((unsigned long long)e.y << 32) | e.x; }

struct TcPlan;

// NBM_OK, or NBM_ERR_UNSUPPORTED when (n_fft, hop, n_bins) do not fit the tensor-core formulation
int tc_plan_create(const nbm_frontend_params &p, TcPlan **out);
void tc_plan_destroy(TcPlan *pl);
// bytes of anchor scratch for `n_anchors` anchor frames (sum over segments of tiles + 1)
size_t tc_anchor_bytes(const TcPlan *pl, long long n_anchors);
int tc_n_ranges(const TcPlan *pl);      // 128-row bin ranges (min/max slots per group)
int tc_bins_per_range();               // output bins per range
// min/max partials: one float2 per (32-frame chain, range, slot); slot s of a range covers tc_bins_per_slot() bins
int tc_chain_frames();
int tc_slots_per_range();
int tc_bins_per_slot();
// anchors: tcgen05 GEMM over the N/2 folded pairs of every 64th frame (mono PCM16 only), for the globally numbered
// anchors [anchor_begin, anchor_end) of segments [seg_lo, seg_hi) (segment s owns group0[s] + s .. + tiles[s])
int tc_launch_anchors(const TcPlan *pl, const SegDesc *d_segs, int seg_lo, int seg_hi, long long anchor_begin,
                      long long anchor_end, const void *d_pcm, void *d_anchors, cudaStream_t stream);
// slides: tcgen05 GEMM over hop/2 pairs + recurrence + Hann + dB for every frame, per-(chain, range, slot) min/max
// for the 64-frame groups [group_begin, group_end), which start in segment seg_begin.  Pixels below their frame's flag
// level (rel_db below the largest |R| the rows carried along the chain so far) are appended to d_cand as blocks together with
// that level (list_entry; at most cand_cap, *d_cand_count counts every attempt) and left out of the min/max partials;
// see refine_groups_kernel.
int tc_launch_slides(const TcPlan *pl, const SegDesc *d_segs, int n_segs, int seg_begin, int group_begin, int group_end,
                     const void *d_pcm, float *d_spec, float2 *d_tile_mm, const void *d_anchors,
                     float rel_db, uint4 *d_cand, unsigned int *d_cand_count, unsigned int cand_cap, cudaStream_t stream);

}  // namespace nbm
