/*
 * htd_b200 - C ABI of the B200-native (sm_100a) HTD RoI-head hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 *   - returns HTD_OK (0) or an HTD_ERR_* code; htd_last_error() gives the message
 *     (thread-local, valid until the next failing call on that thread);
 *   - takes DEVICE pointers owned by the caller (outputs and workspaces included - nothing is
 *     allocated, freed or synchronised inside), is enqueued on `stream` (a cudaStream_t) and is
 *     re-entrant per stream;
 *   - uses channels-last feature maps: level l is [B, H_l, W_l, C]; RoI features are
 *     [K, P, P, C] (= a torch [K, C, P, P] tensor in channels_last memory format).
 *
 * What each function replaces in the reference (paths relative to /root/reference; the native
 * RoIAlign lives in the un-vendored dependency mmcv-full 1.2.1, README.md:11):
 *   htd_level_assign     SingleRoIExtractor.map_roi_levels
 *                        mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:32-51
 *                        (duplicate: bbox_heads/htd_bbox_head.py:129-135)
 *   htd_roi_align_fwd    mmcv ext_module.roi_align_forward as reached from
 *                        roi_extractors/base_roi_extractor.py:49-55, called at
 *                        single_level_roi_extractor.py:81-98 (per-level nonzero/gather/scatter
 *                        loop collapsed into one launch, roi_level != NULL) and
 *                        adaptative_roi_extractor.py:71-74,87 (all levels for every RoI,
 *                        roi_level == NULL); the only in-tree record of the native signature is
 *                        build/lib/mmdet/ops/roi_align/roi_align.py:28-30 (forward_v2).
 *                        `bias` folds HTDRoIHead._fuse_global (htd_roi_head.py:133-141).
 *   htd_roi_align_bwd    mmcv ext_module.roi_align_backward (atomicAdd scatter) -
 *                        build/lib/mmdet/ops/roi_align/roi_align.py:67-71 (backward_v2); here an
 *                        atomic-free, deterministic pixel-tile gather.
 *   htd_layout_convert   the NCHW tensors the reference hands over (two_stage.py:80-87)
 *   htd_ba_*             AdptRoIExtractor.forward, adaptative_roi_extractor.py:76-91
 *   htd_iou_graph_build, htd_pgraph_*   HTDBBoxHead.forward graph loop, htd_bbox_head.py:198-219
 *                        with bbox_overlaps (core/bbox/iou_calculators/iou2d_calculator.py:129-150)
 */
#ifndef HTD_B200_H_
#define HTD_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTD_ABI_VERSION 1

#define HTD_OK 0
#define HTD_ERR_INVALID_ARGUMENT 1
#define HTD_ERR_CUDA 2
#define HTD_ERR_UNSUPPORTED 3

#define HTD_F32 0
#define HTD_BF16 1

#define HTD_MAX_LEVELS 8
#define HTD_MAX_POOLED 8

typedef void* htd_stream_t; /* cudaStream_t */

typedef struct HtdLevel {
    void* data;          /* device pointer, [B, H, W, C] channels-last */
    int32_t H;
    int32_t W;
    float spatial_scale; /* 1 / stride */
    int32_t reserved;
} HtdLevel;

int htd_abi_version(void);
const char* htd_last_error(void);

/* levels[k] = clamp(floor(log2(sqrt(w*h)/finest_scale + 1e-6)), 0, num_levels-1) in fp32
 * semantics; -1 for NaN scale.  rois: [K,5] = (batch, x1, y1, x2, y2) fp32. */
int htd_level_assign(const float* rois, int K, int num_levels, float finest_scale,
                     int32_t* levels, htd_stream_t stream);

/* Pixel footprint of every (level, roi): boxes[l*K + k] = (row0, row1, col0, col1) inclusive,
 * row1 < row0 when the RoI does not touch level l (not assigned to it, outside, degenerate).
 * roi_level == NULL: every RoI on every level.  pixel_count (nullable, [L] uint64, caller
 * zeroes) accumulates sum of fh*fw per level - the algorithmic-traffic unit of SURVEY 8(d). */
int htd_roi_footprints(const HtdLevel* levels, int L, int B, const float* rois, int K,
                       const int32_t* roi_level, int pooled, int sampling_ratio, int32_t* boxes,
                       unsigned long long* pixel_count, htd_stream_t stream);

/* RoIAlign forward (aligned=True, avg).  roi_level != NULL: out[k] sampled from level
 * roi_level[k] (zeros when -1), out is [K, P, P, C].  roi_level == NULL: out is
 * [L, K, P, P, C], every RoI on every level.  bias (nullable): [B, C] fp32 added to every bin
 * of RoI k with batch index b (SFA fuse). */
int htd_roi_align_fwd(const HtdLevel* levels, int L, int B, int C, int in_dtype,
                      const float* rois, int K, const int32_t* roi_level, int pooled,
                      int sampling_ratio, const float* bias, void* out, int out_dtype,
                      htd_stream_t stream);

/* RoIAlign backward.  grad_levels[l].data receives dX_l [B,H,W,C] (fully written, zeros where
 * no RoI lands).  boxes from htd_roi_footprints with the same roi_level.  dy: [K,P,P,C], or
 * [L,K,P,P,C] when dy_per_level != 0.  Effective gradient of RoI k on level l, bin (ph,pw):
 *     (scale[l*K+k] (1 if NULL) + ring(ph,pw)) * dy + addvec[(l*K+k)*C + c] (0 if NULL)
 * ring_edge < 0: ring == 0; ring_edge = e >= 0: on level 0 only, ring == 1 outside the interior
 * [e, P-e) x [e, P-e) (the BA border term, adaptative_roi_extractor.py:87-88). */
int htd_roi_align_bwd(const HtdLevel* grad_levels, int L, int B, int C, int dx_dtype,
                      const float* rois, int K, const int32_t* boxes, int pooled,
                      int sampling_ratio, const void* dy, int dy_dtype, int dy_per_level,
                      const float* scale, int ring_edge, const float* addvec,
                      htd_stream_t stream);

/* Layout / dtype conversion: src [N, R, S] -> dst [N, S, R] (NCHW->NHWC with R=C, S=H*W and
 * back with R=H*W, S=C).  dtypes HTD_F32 / HTD_BF16 independently for src and dst. */
int htd_layout_convert(const void* src, int src_dtype, void* dst, int dst_dtype, long long N,
                       int R, int S, htd_stream_t stream);

/* BA: mean over the P*P bins.  x: [N, PP, C] -> mean [N, C] fp32. */
int htd_ba_bin_mean(const void* x, int x_dtype, long long N, int PP, int C, float* mean,
                    htd_stream_t stream);

/* BA fuse forward: R [L,K,PP,C], logits a [L,K] fp32 ->
 *   w = softmax_l(a) [L,K] (written),  out[k] = sum_l w[l,k] R[l,k] + ring * R[0,k]
 * plus optional residual terms fused for HTDBBoxHead (htd_bbox_head.py:161-184):
 *   + add[k] ([K,PP,C], nullable) + bias[batch(k)] ([B,C] fp32, nullable, needs rois). */
int htd_ba_fuse_fwd(const void* R, int r_dtype, const float* logits, int L, int K, int P, int C,
                    int ring_edge, const void* add, int add_dtype, const float* bias,
                    const float* rois, int B, float* w, void* out, int out_dtype,
                    htd_stream_t stream);

/* BA fuse backward (attention part): dw[l,k] = <dout[k], R[l,k]> over PP*C, then
 * da[l,k] = w[l,k] * (dw[l,k] - sum_m w[m,k] dw[m,k]). */
int htd_ba_fuse_bwd(const void* R, int r_dtype, const void* dout, int dout_dtype, const float* w,
                    int L, int K, int PP, int C, float* da, htd_stream_t stream);

/* Segmented sum of a [K,PP,C] gradient over bins and RoIs of the same image:
 * dbias[b,c] = sum_{k: batch(k)=b} sum_bin g[k,bin,c]  (backward of the fused SFA bias). */
int htd_bias_grad(const void* g, int g_dtype, const float* rois, int K, int PP, int C, int B,
                  float* partial /* workspace [ceil(K/32), B, C] fp32 */, float* dbias,
                  htd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HTD_B200_H_ */
