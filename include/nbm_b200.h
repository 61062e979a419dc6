/* nbm_b200.h -- C ABI of libnbm_b200.so: the B200 (sm_100a) implementation of the NBM
 * audio hot path (waveform -> detector tiles; box decode / threshold / greedy NMS).
 *
 * The reference (LouisBearing/BirdSoundClassif) has no FFI layer: its seams are Python
 * symbols.  Each entry point below names the reference symbol (file:line under
 * /root/reference) whose arithmetic it replaces; INTEGRATION.md shows the ctypes stub a
 * maintainer adds on the reference side.
 *
 * Conventions
 *   - every call returns 0 on success or a negative nbm_status; nbm_last_error() gives a
 *     thread-local message.  No C++ exception crosses this boundary.
 *   - every `d_*` pointer is DEVICE memory owned by the caller (e.g. a torch tensor's
 *     data_ptr); the library allocates nothing per call except inside plan objects.
 *   - `stream` is a cudaStream_t passed as void*; all work is asynchronous on it and no
 *     call synchronises the device unless its comment says so.
 *   - one process per GPU; plans belong to the device current at creation time.
 */
#ifndef NBM_B200_H
#define NBM_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NBM_B200_VERSION 100   /* 0.1.0 */

typedef enum nbm_status {
    NBM_OK = 0,
    NBM_ERR_INVALID = -1,      /* bad argument */
    NBM_ERR_CUDA = -2,         /* CUDA runtime error (message has the cudaError string) */
    NBM_ERR_UNSUPPORTED = -3,  /* valid but not implemented (e.g. pad_mode=reflect) */
    NBM_ERR_WORKSPACE = -4     /* workspace too small */
} nbm_status;

int nbm_version(void);
const char *nbm_last_error(void);

/* ------------------------------------------------------------------ front-end ---------
 * Replaces File_Processor.process_file() = load -> spectrogram -> split_power_spec
 * (nbm_model/nbm_datasets/prepare_dataset.py:108-157, 160-165, 228-294) and the batch
 * assembly of run_detection (nbm_model/run_detection.py:49-55): int16/fp32 PCM in HBM ->
 * float32 tiles [n_tiles, 1, n_bins, w_pix] in HBM, normalised by the whole-file min/max.
 */
typedef struct nbm_frontend_params {
    int32_t sample_rate;   /* File_Processor.FREQ = 44100              prepare_dataset.py:98  */
    int32_t n_fft;         /* WIN_LENGTH  = int(FREQ / freq_accuracy)  prepare_dataset.py:125 */
    int32_t hop;           /* HOP_LENGTH  = int(FREQ * dt)             prepare_dataset.py:126 */
    int32_t low_idx;       /* LOW_IDX                                  prepare_dataset.py:134 */
    int32_t n_bins;        /* H_PIX = 375                              prepare_dataset.py:96  */
    int32_t w_pix;         /* W_PIX                                    prepare_dataset.py:114 */
    int32_t hop_spectro;   /* HOP_SPECTRO                              prepare_dataset.py:115 */
    int32_t pad_mode;      /* 0 = 'constant' (librosa >= 0.10 default); others unsupported    */
    int64_t stft_chunk;    /* max_l = int(5e7)                         prepare_dataset.py:234 */
    double  min_level;     /* exp(-100/20*ln 10)                       prepare_dataset.py:229 */
} nbm_frontend_params;

typedef struct nbm_frontend_plan nbm_frontend_plan;

/* Builds the device twiddle tables for (n_fft, hop).  Synchronises once. */
int nbm_frontend_plan_create(const nbm_frontend_params *params, nbm_frontend_plan **out_plan);
int nbm_frontend_plan_destroy(nbm_frontend_plan *plan);

/* Which device implementation the plan uses: 1 = tensor cores (tcgen05/TMEM sliding DFT), 0 = CUDA cores
 * (parameters outside the tensor-core formulation, or NBM_FRONTEND_IMPL=cuda-core in the environment). */
int nbm_frontend_impl(const nbm_frontend_plan *plan);

/* Pure host arithmetic (prepare_dataset.py:236,266): STFT columns (summed over the
 * <= stft_chunk-sample STFT chunks), detector windows, and the scratch bytes
 * nbm_frontend_run needs for a file of n_samples (per-channel samples). */
int nbm_frontend_query(const nbm_frontend_plan *plan, int64_t n_samples,
                       int64_t *n_frames, int64_t *n_tiles, size_t *workspace_bytes);

/* Same for a batch of files; any out pointer may be NULL.  tile_offsets has n_files+1 entries. */
int nbm_frontend_query_batch(const nbm_frontend_plan *plan, const int64_t *n_samples, int32_t n_files,
                             int64_t *n_frames, int64_t *tile_offsets, size_t *workspace_bytes);

#define NBM_PCM_INT16 0    /* interleaved little-endian PCM16, scaled by 1/32768 on the device */
#define NBM_PCM_FLOAT32 1  /* interleaved float32 already in [-1, 1) */

/* One file.  d_pcm: [n_samples * channels]; d_tiles: [n_tiles, 1, n_bins, w_pix] float32;
 * d_minmax: float32[2] = (s_min, s_max) in dB of the cropped band (prepare_dataset.py:248-249).
 * Channels are averaged to mono (librosa to_mono) before the STFT. */
int nbm_frontend_run(const nbm_frontend_plan *plan, const void *d_pcm, int32_t pcm_dtype,
                     int32_t channels, int64_t n_samples, float *d_tiles, float *d_minmax,
                     void *d_workspace, size_t workspace_bytes, void *stream);

/* Many files in one launch sequence (files stay independent: per-file min/max, per-file
 * tiling).  sample_offsets: host int64[n_files+1], per-channel sample index of each file's
 * first sample inside d_pcm (file f has sample_offsets[f+1]-sample_offsets[f] samples);
 * tiles of file f start at tile index tile_offsets[f] (nbm_frontend_query_batch);
 * d_minmax: float32[n_files, 2]. */
int nbm_frontend_run_batch(const nbm_frontend_plan *plan, const void *d_pcm, int32_t pcm_dtype,
                           int32_t channels, const int64_t *sample_offsets, int32_t n_files,
                           float *d_tiles, float *d_minmax, void *d_workspace,
                           size_t workspace_bytes, void *stream);

/* The un-normalised dB band [n_bins, n_frames] (row stride = *row_stride floats) that the last
 * run left in the workspace for file `file_index`; for tests and diagnostics. */
int nbm_frontend_spectrogram_view(const nbm_frontend_plan *plan, const int64_t *n_samples,
                                  int32_t n_files, int32_t file_index, size_t *offset_bytes,
                                  int64_t *row_stride);

/* Per-kernel device timing for benchmarks: when enabled, every run records CUDA events on the
 * caller's stream around the STFT/dB kernel and the tiling kernel; get_profile waits for the
 * last profiled run and returns the accumulated milliseconds and the number of runs since
 * profiling was (re-)enabled. */
int nbm_frontend_set_profiling(nbm_frontend_plan *plan, int32_t enable);
int nbm_frontend_get_profile(nbm_frontend_plan *plan, double *stft_ms, double *tile_ms, int64_t *runs);
/* The same, per kernel: ms4 = {anchor GEMM, slide/STFT kernel, whole-file min/max, tiling}. */
int nbm_frontend_get_profile_kernels(nbm_frontend_plan *plan, double *ms4, int64_t *runs);

/* Diagnostics of the float64 refinement pass of the LAST run (waits for it; valid while that run's workspace is alive):
 * *listed = pixel blocks the transform put on the refinement list (counted even when the list was full),
 * *capacity = entries the list could hold (listed > capacity means some pixels kept their float32 values),
 * *recomputed (optional) = pixels of those blocks that were below the flag level and recomputed in float64. */
int nbm_frontend_last_listed(nbm_frontend_plan *plan, int64_t *listed, int64_t *capacity, int64_t *recomputed);

/* Dataset images from detector tiles: out = uint8(round_half_even(tile * 255)), the quantisation
 * prepare_dataset() applies before writing a PNG (prepare_dataset.py:85).  d_tiles 16-byte aligned,
 * n_values = number of pixels (any count), asynchronous on `stream`. */
int nbm_tiles_to_u8(const float *d_tiles, int64_t n_values, uint8_t *d_out, void *stream);

/* ------------------------------------------------------------- post-processing --------
 * Anchor table: generate_anchors_frcnn + get_anchor_shifts_frcnn combined as in
 * ProposalLayer.forward (nets_utils.py:35-59, layers.py:252-258).  Host arithmetic;
 * out: float32[height*width*n_ratios*n_scales, 4], index (y*width + x)*A + a. */
int nbm_make_anchors(int32_t base_size, const double *ratios, int32_t n_ratios,
                     const int64_t *scales, int32_t n_scales, int32_t width, int32_t height,
                     int32_t stride, float *h_out);

/* bbox_reg_to_coord (nets_utils.py:169-186) fused with the clamp and min-size test of
 * ProposalLayer.forward (layers.py:279-285).  d_deltas [B, N, 4]; d_anchors [N, 4] when
 * anchors_per_image == 0 else [B, N, 4]; clip_w/clip_h <= 0 disables clamping;
 * d_valid (may be NULL) [B, N] = both sides >= min_size. */
int nbm_decode_boxes(const float *d_deltas, const float *d_anchors, int32_t B, int32_t N,
                     int32_t anchors_per_image, float clip_w, float clip_h, float min_size,
                     float *d_boxes, uint8_t *d_valid, void *stream);

/* Greedy in-order NMS = batch_self_overlap + the loop of nms (nets_utils.py:189-232):
 * box i (ascending, i < d_n[b]) is kept unless an earlier KEPT box has IoU >= thresh with it
 * (float32 IoU, +1 pixel convention).  Does NOT sort.  d_boxes [B, N, 4]; d_n int32[B] or NULL
 * (all N valid); d_keep_idx int32[B, N] (first d_keep_cnt[b] entries valid, ascending);
 * workspace: nbm_nms_workspace_bytes(B, N).  The caller applies the batch-min truncation
 * (nets_utils.py:236-238). */
size_t nbm_nms_workspace_bytes(int32_t B, int32_t N);
int nbm_nms_greedy(const float *d_boxes, const int32_t *d_n, int32_t B, int32_t N, float thresh,
                   int32_t *d_keep_idx, int32_t *d_keep_cnt, void *d_workspace,
                   size_t workspace_bytes, void *stream);

/* ProposalLayer.forward, eval branch (layers.py:226-303), fused: scores/deltas in the RPN's
 * native layouts d_cls [B, 2A, H, W] (softmaxed pairs, fg = odd channel), d_reg [B, 4A, H, W];
 * decode, clamp, min-size filter, stable descending sort, top pre_nms_topN (batch-min
 * coupled), NMS(nms_thresh), batch-min truncation to <= post_nms_topN.
 * Outputs: d_rois [B, post_nms_topN, 4], d_scores [B, post_nms_topN] (first *M rows valid),
 * h_M (host int32, written after an internal stream synchronise): M >= 0, or -1 for the
 * reference's "RPN failed" branch (fewer than rcnn_batch_size candidates, layers.py:288-290). */
typedef struct nbm_proposal_params {
    int32_t A, H, W;               /* anchors per cell, feature map size (15, 24, 64) */
    float   img_width, img_height; /* clamp to [0, img-1]                             */
    float   min_size;              /* config.min_threshold                            */
    float   nms_thresh;            /* config.nms_thresh                               */
    int32_t pre_nms_topN;          /* config.pre_nms_topN_eval                        */
    int32_t post_nms_topN;         /* config.post_nms_topN_eval                       */
    int32_t rcnn_batch_size;       /* config.rcnn_batch_size                          */
} nbm_proposal_params;
size_t nbm_proposals_workspace_bytes(const nbm_proposal_params *p, int32_t B);
int nbm_proposals(const nbm_proposal_params *p, const float *d_cls, const float *d_reg,
                  const float *d_anchors, int32_t B, float *d_rois, float *d_scores,
                  int32_t *h_M, void *d_workspace, size_t workspace_bytes, void *stream);
/* The same without the host read: M goes to d_M (device int32) and the call neither copies to the
 * host nor synchronises, so it can be recorded into a CUDA graph together with the network that
 * produces d_cls / d_reg (the caller reads d_M when it needs the RoI count). */
int nbm_proposals_async(const nbm_proposal_params *p, const float *d_cls, const float *d_reg,
                        const float *d_anchors, int32_t B, float *d_rois, float *d_scores,
                        int32_t *d_M, void *d_workspace, size_t workspace_bytes, void *stream);

/* FastRCNN.forward inference branch (layers.py:688-778), fused per image: argmax class,
 * class-specific delta gather, decode against rois, clamp, stable descending sort, drop
 * class 0, NMS(nms_thresh), score > min_score (strict).  d_bbox_reg [B*R, 4*(C+1)],
 * d_probs [B*R, C+1], d_rois [B, R, 4].  Output records, in surviving (score-descending)
 * order per image: d_det_boxes [B, R, 4], d_det_scores [B, R], d_det_class int32[B, R],
 * d_det_count int32[B].  Python regroups by class (stable), which reproduces the
 * reference's per-class dicts. */
int nbm_final_detections(const float *d_bbox_reg, const float *d_probs, const float *d_rois,
                         int32_t B, int32_t R, int32_t num_classes, float img_width,
                         float img_height, float nms_thresh, float min_score,
                         float *d_det_boxes, float *d_det_scores, int32_t *d_det_class,
                         int32_t *d_det_count, void *stream);

/* merge_images (run_detection.py:163-249) on flat records: n candidate boxes of one file in
 * tile-major order with their class and tile index.  Border filter, x offset by
 * hop_spectro*tile, end-of-file filter, reorder class-major (stable), one class-agnostic
 * greedy NMS(nms_thresh).  Outputs compacted survivors in NMS order: d_out_boxes [n,4],
 * d_out_scores [n], d_out_class int32[n], d_out_count int32[1].
 * workspace: nbm_merge_workspace_bytes(n). */
size_t nbm_merge_workspace_bytes(int32_t n);
int nbm_merge_detections(const float *d_boxes, const float *d_scores, const int32_t *d_class,
                         const int32_t *d_tile, int32_t n, int32_t n_tiles, int32_t w_pix,
                         int32_t hop_spectro, int64_t spectrogram_length, float nms_thresh,
                         float *d_out_boxes, float *d_out_scores, int32_t *d_out_class,
                         int32_t *d_out_count, void *d_workspace, size_t workspace_bytes,
                         void *stream);

/* ------------------------------------------------------------- second-stage pooling ----
 * ROIPooling.forward (layers.py:399-497): pyramid-level assignment, adaptive average pooling of the
 * level's feature-map crop, and the pooled RoI positional encoding, for all B*R RoIs in one launch
 * (the reference loops in Python with five .item() syncs per RoI).  d_feat: HOST array of n_layers
 * device pointers to [B, C, heights[l], widths[l]] float32 maps.  d_pe_freq [img_h, C/2] and d_pe_time
 * [img_w, C/2]: one_dimension_positional_encoding tables (position_encoding.py:10-15).  Outputs:
 * d_pool_out, d_pe_out [B, R, C, pool_h, pool_w]; d_level_out int32 [B, R].  Averages follow ATen's
 * summation order, so the result is bit-identical to the reference on the same inputs. */
int nbm_roi_pool(const float *d_rois, int32_t B, int32_t R, const float *const *d_feat, const int32_t *heights,
                 const int32_t *widths, int32_t n_layers, int32_t C, int32_t pool_h, int32_t pool_w,
                 int32_t img_h, int32_t img_w, const float *d_pe_freq, const float *d_pe_time,
                 float *d_pool_out, float *d_pe_out, int32_t *d_level_out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* NBM_B200_H */
