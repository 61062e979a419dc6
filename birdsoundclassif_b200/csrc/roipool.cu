// RoI pooling + RoI positional encoding of the detector's second stage
// (nbm_model/nets/layers.py:399-497, ROIPooling.forward), one launch instead of a Python loop over
// batch x RoIs with five .item() host syncs per RoI.
//
// Per RoI (x1, y1, x2, y2) in image pixels:
//   side  = sqrt((x2 - x1)(y2 - y1));  level = clamp(int(log(0.1 side) / log 2), 0, n_layers - 1)   :408-416
//   stride s = 2^(level + 1);  feature-map box = round(coord / s) (half to even)                     :418-427
//   y2 = min(y2, H - 1); grow the box by one cell per side until it spans pool_h x pool_w cells     :455-464
//   pooled = AdaptiveAvgPool2d(feature[level][b, :, y1:y2+1, x1:x2+1])  (python slicing clamps)      :480
//   pe     = AdaptiveAvgPool2d(cat(freq_pe[s y1 : s y2] broadcast over time,
//                                  time_pe[: s (x2 - x1)] broadcast over frequency))                 :483-489
// The averages replicate ATen's order (window summed row by row in float, then "/ kh / kw"), so results
// are bit-identical to the reference on the same feature maps.  That order makes every output one serial chain of
// float additions (up to ~7 500 for a positional-encoding window), so the kernel is latency-bound and wants many
// resident warps: blockIdx.y splits a RoI's C x pool_h x pool_w outputs over several blocks.
#include <algorithm>
#include "common.cuh"

namespace nbm {

constexpr int ROI_MAX_LEVELS = 8;

struct RoiPoolParams {
    int B, R, C, n_layers, pool_h, pool_w, img_h, img_w;
    int H[ROI_MAX_LEVELS], W[ROI_MAX_LEVELS];
    const float *feat[ROI_MAX_LEVELS];      // [B, C, H_l, W_l]
};

// ATen's adaptive pooling windows: [floor(a c / b), ceil((a + 1) c / b))
__device__ __forceinline__ int start_index(int a, int b, int c) { return (a * c) / b; }
__device__ __forceinline__ int end_index(int a, int b, int c) { return ((a + 1) * c + b - 1) / b; }

__global__ void __launch_bounds__(256)
roi_pool_kernel(RoiPoolParams P, const float4 *__restrict__ rois, const float *__restrict__ pe_freq,
                const float *__restrict__ pe_time, float *__restrict__ pool_out, float *__restrict__ pe_out,
                int *__restrict__ lvl_out) {
    __shared__ int g[8];        // level, x1, y1, x2, y2, stride
    const int roi = blockIdx.x, b = roi / P.R;
    if (threadIdx.x == 0) {
        const float4 r = rois[roi];
        const float side = sqrtf(__fmul_rn(__fsub_rn(r.z, r.x), __fsub_rn(r.w, r.y)));
        // torch: (log(side * 0.1) / np.log(2)).int(): float log, divided by float(ln 2), truncated
        const float lv = __fdiv_rn(logf(__fmul_rn(side, 0.1f)), 0.6931471805599453f);
        int level = (int)lv;                                   // NaN / -inf (degenerate RoIs) clamp to 0 below
        if (!(lv >= 0.f)) level = 0;
        level = min(max(level, 0), P.n_layers - 1);
        const int s = 2 << level;
        const float fs = (float)s;
        int x1 = (int)rintf(__fdiv_rn(r.x, fs)), y1 = (int)rintf(__fdiv_rn(r.y, fs));
        int x2 = (int)rintf(__fdiv_rn(r.z, fs)), y2 = (int)rintf(__fdiv_rn(r.w, fs));
        const int H = P.H[level], W = P.W[level];
        y2 = min(y2, H - 1);
        while (y2 - y1 + 1 < P.pool_h) { y1 = max(0, y1 - 1); y2 = min(H - 1, y2 + 1); if (y1 == 0 && y2 == H - 1) break; }
        while (x2 - x1 + 1 < P.pool_w) { x1 = max(0, x1 - 1); x2 = min(W - 1, x2 + 1); if (x1 == 0 && x2 == W - 1) break; }
        g[0] = level; g[1] = x1; g[2] = y1; g[3] = x2; g[4] = y2; g[5] = s;
        lvl_out[roi] = level;
    }
    __syncthreads();
    const int level = g[0], x1 = g[1], y1 = g[2], x2 = g[3], y2 = g[4], s = g[5];
    const int H = P.H[level], W = P.W[level];
    const int n_out = P.C * P.pool_h * P.pool_w, per_c = P.pool_h * P.pool_w;
    // ---- feature crop (python slice semantics: the end is clamped to the map) -----------------------
    const int ch = min(y2 + 1, H) - y1, cw = min(x2 + 1, W) - x1;
    const float *fm = P.feat[level] + ((size_t)b * P.C) * H * W;
    const int idx0 = blockIdx.y * blockDim.x + threadIdx.x, idx_step = gridDim.y * blockDim.x;
    for (int idx = idx0; idx < n_out; idx += idx_step) {
        const int c = idx / per_c, oh = (idx % per_c) / P.pool_w, ow = idx % P.pool_w;
        float v = 0.f;
        if (ch > 0 && cw > 0) {
            const int ih0 = start_index(oh, P.pool_h, ch), ih1 = end_index(oh, P.pool_h, ch);
            const int iw0 = start_index(ow, P.pool_w, cw), iw1 = end_index(ow, P.pool_w, cw);
            const float *p = fm + (size_t)c * H * W + (size_t)y1 * W + x1;
            float sum = 0.f;
            for (int ih = ih0; ih < ih1; ++ih)
                for (int iw = iw0; iw < iw1; ++iw) sum = __fadd_rn(sum, __ldg(p + (size_t)ih * W + iw));
            v = __fdiv_rn(__fdiv_rn(sum, (float)(ih1 - ih0)), (float)(iw1 - iw0));
        }
        pool_out[(size_t)roi * n_out + idx] = v;
    }
    // ---- positional encoding: [C/2 frequency channels | C/2 time channels] over an (Hf, Wt) pixel box ---
    const int C2 = P.C / 2;
    const int Hf = max(0, min(s * y2, P.img_h) - min(s * y1, P.img_h)), Wt = max(0, min(s * (x2 - x1), P.img_w));
    for (int idx = idx0; idx < n_out; idx += idx_step) {
        const int c = idx / per_c, oh = (idx % per_c) / P.pool_w, ow = idx % P.pool_w;
        float v = 0.f;
        if (Hf > 0 && Wt > 0) {
            const int ih0 = start_index(oh, P.pool_h, Hf), ih1 = end_index(oh, P.pool_h, Hf);
            const int iw0 = start_index(ow, P.pool_w, Wt), iw1 = end_index(ow, P.pool_w, Wt);
            float sum = 0.f;
            if (c < C2) {
                for (int ih = ih0; ih < ih1; ++ih) {
                    const float f = __ldg(pe_freq + (size_t)(s * y1 + ih) * C2 + c);
                    for (int iw = iw0; iw < iw1; ++iw) sum = __fadd_rn(sum, f);
                }
            } else {
                for (int ih = ih0; ih < ih1; ++ih)
                    for (int iw = iw0; iw < iw1; ++iw) sum = __fadd_rn(sum, __ldg(pe_time + (size_t)iw * C2 + (c - C2)));
            }
            v = __fdiv_rn(__fdiv_rn(sum, (float)(ih1 - ih0)), (float)(iw1 - iw0));
        }
        pe_out[(size_t)roi * n_out + idx] = v;
    }
}

}  // namespace nbm

using namespace nbm;

extern "C" int nbm_roi_pool(const float *d_rois, int32_t B, int32_t R, const float *const *d_feat, const int32_t *heights,
                            const int32_t *widths, int32_t n_layers, int32_t C, int32_t pool_h, int32_t pool_w,
                            int32_t img_h, int32_t img_w, const float *d_pe_freq, const float *d_pe_time,
                            float *d_pool_out, float *d_pe_out, int32_t *d_level_out, void *stream) {
    NBM_REQUIRE(d_rois && d_feat && heights && widths && d_pe_freq && d_pe_time && d_pool_out && d_pe_out && d_level_out,
                "null argument");
    NBM_REQUIRE(B >= 1 && R >= 1 && C >= 2 && C % 2 == 0 && pool_h >= 1 && pool_w >= 1, "bad shape");
    NBM_REQUIRE(n_layers >= 1 && n_layers <= ROI_MAX_LEVELS, "n_layers must be in [1, %d]", ROI_MAX_LEVELS);
    RoiPoolParams P;
    P.B = B; P.R = R; P.C = C; P.n_layers = n_layers; P.pool_h = pool_h; P.pool_w = pool_w; P.img_h = img_h; P.img_w = img_w;
    for (int l = 0; l < ROI_MAX_LEVELS; ++l) {
        P.H[l] = l < n_layers ? heights[l] : 0; P.W[l] = l < n_layers ? widths[l] : 0; P.feat[l] = l < n_layers ? d_feat[l] : nullptr;
        if (l < n_layers) NBM_REQUIRE(P.H[l] >= 1 && P.W[l] >= 1 && P.feat[l], "bad feature map %d", l);
    }
    // enough blocks for up to 8 per SM (the grid of B x R = 200 RoIs alone leaves most SMs with one block of serial chains)
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int n_out = C * pool_h * pool_w;
    const int chunks = std::max(1, std::min((8 * sms) / (B * R), (n_out + 255) / 256));      // one wave of 256-thread blocks
    roi_pool_kernel<<<dim3((unsigned)(B * R), (unsigned)chunks), 256, 0, (cudaStream_t)stream>>>(P, reinterpret_cast<const float4 *>(d_rois), d_pe_freq, d_pe_time,
                                                             d_pool_out, d_pe_out, d_level_out);
    NBM_CUDA(cudaGetLastError());
    return NBM_OK;
}
