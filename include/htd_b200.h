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
 *   htd_roi_plan         the coordinate arithmetic of mmcv's roi_align kernels (sample grid,
 *                        bilinear taps), hoisted out of the per-output-element loops
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

/* Sampling plan shared by forward and backward, built once per extractor call:
 *   boxes   [L*K][4]  footprint boxes as in htd_roi_footprints;
 *   offsets [L*K+1]   first table row of entry e = l*K + k (its fh + fw rows follow); entries sit at a
 *                     fixed stride inside the htd_roi_plan_rows_bound capacity (one launch, no scan) -
 *                     or densely packed (exclusive scan) when a pixel count is requested;
 *   ranges  [L*K][4*HTD_MAX_POOLED]  per output bin p: first/last feature row (y lo[8], y hi[8]) and
 *                     column (x lo[8], x hi[8]) it samples (hi < lo: empty bin);
 *   weights [rows_cap][HTD_MAX_POOLED] fp32 separable axis weights: row offsets[e] + (r - row0)
 *                     holds Wy[p][r] for the P bins, row offsets[e] + fh + (c - col0) holds Wx[p][c]
 *                     (aligned=True, avg pooling; sampling_ratio 0 = adaptive ceil(roi/P) grid);
 *                     with pooled < HTD_MAX_POOLED the last entry of every row holds the sum of
 *                     the row's P weights (read by the bf16 backward's add-vector term), the
 *                     entries P .. HTD_MAX_POOLED-2 are 0.
 * rows_cap >= htd_roi_plan_rows_bound(levels, L, K, roi_level != NULL). */
long long htd_roi_plan_rows_bound(const HtdLevel* levels, int L, int K, int single_level);
int htd_roi_plan(const HtdLevel* levels, int L, int B, const float* rois, int K,
                 const int32_t* roi_level, int pooled, int sampling_ratio, int32_t* boxes,
                 int32_t* offsets, int32_t* ranges, float* weights, long long rows_cap,
                 unsigned long long* pixel_count, htd_stream_t stream);

/* RoIAlign forward (aligned=True, avg) from a plan.  roi_level != NULL: out[k] sampled from level
 * roi_level[k] (zeros when -1), out is [K, P, P, C].  roi_level == NULL: out is
 * [L, K, P, P, C], every RoI on every level.  bias (nullable): [B, C] fp32 added to every bin
 * of RoI k with batch index b (SFA fuse). */
int htd_roi_align_fwd(const HtdLevel* levels, int L, int B, int C, int in_dtype,
                      const float* rois, int K, const int32_t* roi_level, int pooled,
                      const int32_t* boxes, const int32_t* offsets, const int32_t* ranges,
                      const float* weights, const float* bias, void* out, int out_dtype,
                      htd_stream_t stream);

/* RoIAlign backward from the same plan.  grad_levels[l].data receives dX_l [B,H,W,C] (fully
 * written, zeros where no RoI lands).  dy: [K,P,P,C], or [L,K,P,P,C] when dy_per_level != 0.
 * Effective gradient of RoI k on level l, bin (ph,pw):
 *     (scale[l*K+k] (1 if NULL) + ring(ph,pw)) * dy + addvec[(l*K+k)*C + c] (0 if NULL)
 * ring_edge < 0: ring == 0; ring_edge = e >= 0: on level 0 only, ring == 1 outside the interior
 * [e, P-e) x [e, P-e) (the BA border term, adaptative_roi_extractor.py:87-88). */
int htd_roi_align_bwd(const HtdLevel* grad_levels, int L, int B, int C, int dx_dtype,
                      const float* rois, int K, const int32_t* boxes, const int32_t* offsets,
                      const int32_t* ranges, const float* weights, int pooled, const void* dy,
                      int dy_dtype, int dy_per_level, const float* scale, int ring_edge,
                      const float* addvec, htd_stream_t stream);

/* Several extractor calls whose gradients land in the same pyramid (an HTD step has three: the
 * stage-0 and stage-1 single-level extractions and the BA extraction), gathered in ONE pass over
 * the tiles of dX.  Per source: the fields of htd_roi_align_bwd.  dx_nchw != 0 writes
 * grad_levels[l].data as [B,C,H,W] (the layout the reference's FPN expects) instead of
 * channels-last. */
#define HTD_MAX_BWD_SOURCES 4
typedef struct HtdBwdSource {
    const float* rois;
    const int32_t* boxes;
    const int32_t* offsets;
    const int32_t* ranges;
    const float* weights;
    const void* dy;
    const float* scale;
    const void* addvec;       /* fp32, or bf16 when addvec_dtype == HTD_BF16 */
    int32_t K;
    int32_t dy_per_level;
    int32_t ring_edge;
    int32_t addvec_dtype;     /* HTD_F32 (0) | HTD_BF16: bf16 needs dy bf16, pooled < 8, C % 64 == 0, C <= 256 */
} HtdBwdSource;
int htd_roi_align_bwd_multi(const HtdLevel* grad_levels, int L, int B, int C, int dx_dtype,
                            int dx_nchw, const HtdBwdSource* sources, int nsrc, int pooled,
                            int dy_dtype, htd_stream_t stream);

/* 1 when htd_roi_align_bwd(_multi) with these sizes runs the bf16 tensor-pipe gather (bf16 dy, C a
 * multiple of 64 in 64..256, not switched off by HTD_BWD_KERNEL=scalar / htd_debug_set_bwd_variant),
 * else 0 (the exact scalar gather).  A bf16 add vector (addvec_dtype == HTD_BF16) is accepted only
 * when this returns 1 and pooled < HTD_MAX_POOLED. */
int htd_roi_align_bwd_uses_tensor_pipe(int C, int pooled, int dy_dtype);

#ifdef HTD_DEBUG_HOOKS
/* Measurement hooks: exported only by the library built with -DHTD_DEBUG_HOOKS
 * (htd_b200/_lib/libhtd_b200_hooks.so, used by tests and tools that compare kernel variants).
 * The product library does not export them and ignores the HTD_*_KERNEL / HTD_DENSE_* environment
 * variables.  Not part of the reference's interface.
 *
 * htd_debug_set_bwd_trace (tools/trace_bwd.py): when `records` is non-null the bf16 backward gather
 * writes per CTA six uint64 {globaltimer at start, at end, hits, K-step blocks, SM id, level} at
 * records[6 * blockIdx]; pass NULL to switch it off. */
void htd_debug_set_bwd_trace(unsigned long long* records);

/* Kernel used for bf16 dy by htd_roi_align_bwd(_multi).  0 = the scalar FFMA gather (also the fp32
 * path), 1..5 = warp layouts / ring depths of the tensor-pipe gather (3 is the default), -1 = back
 * to the default (or the HTD_BWD_KERNEL environment variable). */
void htd_debug_set_bwd_variant(int variant);

/* Dense-kernel options: "dense_pair" (0 off, 1 conv fprop / dgrad in the CTA-pair form, 2 also the
 * GEMM kinds), "dense_debug" (experiment bits, csrc/dense_gemm.cu), "pair_stages" (ring depth). */
void htd_debug_set_option(const char* name, int value);
#endif

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

/* RoI maps [K, PP, C] (channels-last) -> the FC flatten order [K, C, PP] of the reference
 * (x.flatten(1) of an NCHW tensor, convfc_bbox_head.py:145), optionally adding the SFA vector of
 * the RoI's image on the way: dst[k, c, p] = src[k, p, c] + bias[image(k), c]
 * (HTDRoIHead._fuse_global, htd_roi_head.py:133-141, for stage 0: the extraction then does not
 * wait for the global-context head).  bias [B, C] fp32 or NULL; rois [K, 5] (image index first).
 * PP * C <= 12800 and a multiple of 8. */
int htd_roi_flatten(const void* src, int src_dtype, void* dst, int dst_dtype, int K, int PP, int C,
                    const float* bias, const float* rois, int B, htd_stream_t stream);

/* Attention MLP of the BA extractor on the bin means m [rows, C] (rows = levels * RoIs):
 * h = tanh(m W1^T + b1) [rows, H], logits = h W2^T + b2 [rows]
 * (AdptRoIExtractor's conv1 / tanh / conv2 on the globally pooled RoI maps,
 * adaptative_roi_extractor.py:60-74; the 1x1 convs on 1x1 maps are these two products).
 * Parameters are read in their own dtype (p_dtype: HTD_F32 / HTD_BF16, one for all four);
 * H must be 128 and C a multiple of 32 up to 256 (htd_ba_mlp_supported).
 * Backward: dm = inv_pp * (da W2 (1 - h^2)) W1 and the four parameter gradients (written in
 * p_dtype); workspace = htd_ba_mlp_workspace_floats(rows, C) floats.  Deterministic. */
int htd_ba_mlp_supported(int C, int H);
long long htd_ba_mlp_workspace_floats(long long rows, int C);
int htd_ba_mlp_fwd(const float* m, long long rows, int C, int H, const void* w1, const void* b1,
                   const void* w2, const void* b2, int p_dtype, float* h, float* logits,
                   htd_stream_t stream);
int htd_ba_mlp_bwd(const float* da, const float* h, const float* m, long long rows, int C, int H,
                   const void* w1, const void* w2, int p_dtype, float inv_pp, float* dm,
                   float* workspace, void* dw1, void* db1, void* dw2, void* db2,
                   htd_stream_t stream);

/* Segmented sum of a [K,PP,C] gradient over bins and RoIs of the same image:
 * dbias[b,c] = sum_{k: batch(k)=b} sum_bin g[k,bin,c]  (backward of the fused SFA bias). */
int htd_bias_grad(const void* g, int g_dtype, const float* rois, int K, int PP, int C, int B,
                  float* partial /* workspace max(K, ceil(K/4) * B) * C fp32 */, float* dbias,
                  htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * PGraph (HTDBBoxHead.forward graph loop, htd_bbox_head.py:195-219).
 *
 * RoIs are regrouped into "sorted space": a stable counting sort by key = level * B + image
 * (row order inside a group = ascending original RoI index, i.e. the reference's boolean-mask
 * order, htd_bbox_head.py:199-207).  Every LEVEL block starts on a multiple of `align` rows and
 * every group on a multiple of HTD_GROUP_ALIGN rows (TMA needs 16-byte aligned K offsets); the
 * gaps are pad rows (perm == -1) and are zero in every sorted-space buffer.  A group g with span
 * (off, n) owns sorted rows [off, off+n); its n x n matrices (adjacency, similarity, attention)
 * are stored as rows off..off+n-1, columns 0..n-1 of one [Npad, ld] buffer.
 * ------------------------------------------------------------------------------------------ */
#define HTD_MAX_GROUPS 64
#define HTD_GROUP_ALIGN 8
#define HTD_PLAN_ROW_CAPACITY(K, L, B, align) ((K) + (L) * (align) + (L) * (B) * HTD_GROUP_ALIGN)

/* table layout (int32): [2*g + 0/1] = (off, n) of group g = level*B + image, g < L*B;
 * [2*L*B + 2*l + 0/1] = (off, n) of level block l (n = rows from the block start to its last
 * group row, inner pad rows included); [2*L*B + 2*L] = Npad (multiple of align). */
#define HTD_PLAN_TABLE_INTS(L, B) (2 * (L) * (B) + 2 * (L) + 1)

/* levels: [K] int32 from htd_level_assign (-1 rows and rows with image id outside [0,B) join no
 * group).  Ncap >= HTD_PLAN_ROW_CAPACITY(K, L, B, align) rows of capacity.  Outputs (device): perm [Ncap] sorted row ->
 * original index (-1 pad), pos [K] original -> sorted row (-1 none), rowspan [Ncap][2] = (off, n)
 * of the row's group (n = 0 for pad rows), boxes [Ncap][4] = the row's (x1,y1,x2,y2), table. */
int htd_pgraph_plan(const float* rois, const int32_t* levels, int K, int B, int L, int align,
                    int Ncap, int32_t* perm, int32_t* pos, int32_t* rowspan, float* boxes,
                    int32_t* table, htd_stream_t stream);

/* Gather rows into sorted space with dtype conversion: v[p, c] = src[perm[p], c] (* [gate[perm[p],
 * c] > 0] when gate != NULL; 0 for pad rows).  dst (nullable): [Npad, ldd] row-major, columns
 * D..ldd-1 zero-filled.  dstT (nullable): [D, ldt], dstT[c, p] = v[p, c]. */
int htd_pgraph_pack(const void* src, int src_dtype, long long lds, const void* gate,
                    int gate_dtype, long long ldg, const int32_t* perm, int Npad, int D, void* dst,
                    long long ldd, void* dstT, long long ldt, int dst_dtype, htd_stream_t stream);

/* Local adjacency of every group (htd_bbox_head.py:207-210, bbox_overlaps =
 * core/bbox/iou_calculators/iou2d_calculator.py:129-150):
 *   M[p, j] = (IoU(box_p, box_{off+j}) > 0) or (off + j == p)          j < n
 *   bits[p, j/32] bit j%32 = M (ldb uint32 words per row, zero-padded), deg[p] = sum_j M[p, j]
 *   adj[p, j] = deg[p]^-1/2 * M[p, j] * deg[off+j]^-1/2  (0 for n <= j < ldn and for pad rows)
 * IoU uses the reference's fp32 expression order; the mask is bit-exact. */
int htd_iou_graph_build(const float* boxes, const int32_t* rowspan, int Npad, uint32_t* bits,
                        int ldb, int32_t* deg, void* adj, int adj_dtype, long long ldn,
                        htd_stream_t stream);

/* Global attention (htd_bbox_head.py:211,215): out[p, j] = softmax_j((1 - M[p, j]) * S[p, j]),
 * j < n - masked entries keep logit 0, NOT -inf; zeros for n <= j < ldo and pad rows. */
int htd_pgraph_masked_softmax(const float* S, long long lds, const uint32_t* bits, int ldb,
                              const int32_t* rowspan, int Npad, void* out, int out_dtype,
                              long long ldo, htd_stream_t stream);

/* Backward of the above: dS[p, j] = (1 - M[p, j]) * A[p, j] * (dA[p, j] - sum_j A[p, j] dA[p, j]),
 * fp32, zeros beyond n / pad rows. */
int htd_pgraph_softmax_bwd(const void* A, int a_dtype, long long lda, const float* dA,
                           long long ldda, const uint32_t* bits, int ldb, const int32_t* rowspan,
                           int Npad, float* dS, long long ldds, htd_stream_t stream);

/* Per-group transpose / symmetrise: out[p, j] = alpha * in[p, j] + beta * in[off + j, p - off],
 * j < n; zeros beyond n / pad rows. */
int htd_pgraph_group_transpose(const void* in, int in_dtype, long long ldi, const int32_t* rowspan,
                               int Npad, float alpha, float beta, void* out, int out_dtype,
                               long long ldo, htd_stream_t stream);

/* Column sums over row segments: out[s, c] = sum_{p in [seg[2s], seg[2s]+seg[2s+1])} x[p, c];
 * seg is a DEVICE pointer (e.g. the level part of the plan table). */
int htd_pgraph_segment_colsum(const void* x, int x_dtype, long long ldx, const int32_t* seg,
                              int num_seg, int D, float* out, htd_stream_t stream);

/* One problem of a grouped contraction D = A * B^T (both operands K-major / row-major [rows, K]). */
typedef struct HtdGemmGroup {
    int32_t M, N, K;
    int32_t a_row, a_k0; /* A block origin: rows a_row.., K columns a_k0..   */
    int32_t b_row, b_k0; /* B block origin: rows b_row.. (the N index), K columns b_k0.. */
    int32_t d_row, d_col; /* D[(d_row + m), d_col + n]  (d_row + m goes through d_rowmap if given) */
    int32_t dt_row, dt_col; /* DT[(dt_row + n), dt_col + m] */
    int32_t bias_off;    /* bias[bias_off + n] added before relu (ignored when bias == NULL) */
} HtdGemmGroup;

/* Grouped GEMM for the PGraph contractions (htd_bbox_head.py:210-216 and their backward,
 * SURVEY.md Appendix D).  ab_dtype HTD_BF16: tcgen05 tensor cores (TMA-fed, TMEM accumulators,
 * fp32 accumulate); HTD_F32: exact-fp32 FFMA tiles (the fp32 parity configuration).  For BF16
 * the K extent of at least ONE operand must be zero beyond the group's K up to the next multiple
 * of 64 and the other finite (sorted-space buffers guarantee it); leading dimensions and the K
 * offsets a_k0 / b_k0 must be multiples of 8 elements and base pointers 16-byte aligned.
 * Outputs (either may be NULL): D [.., ldd] (d_dtype) with optional row scatter map
 * (d_rowmap[d_row + m] < 0 drops the row), DT [.., ldt] (dt_dtype) = transposed copy.
 * v = acc (+ bias[bias_off + n]) (relu) is what both receive. */
int htd_pgraph_gemm(const void* A, long long a_rows, long long a_ld, const void* B,
                    long long b_rows, long long b_ld, int ab_dtype, const HtdGemmGroup* groups,
                    int G, void* D, int d_dtype, long long ldd, const int32_t* d_rowmap, void* DT,
                    int dt_dtype, long long ldt, const float* bias, int relu,
                    htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Dense bf16 contractions of the head on the tcgen05 tensor cores:
 * the FC stacks (convfc_bbox_head.py:141-148 `shared_fcs`, htd_bbox_head.py:114-121,191-192 `fcs`,
 * :227-228 fc_cls / fc_reg) and the 3x3 regression conv tower on the 7x7 RoI maps
 * (htd_bbox_head.py:75-113,186), forward + data gradient + weight gradient - the work the
 * reference hands to cuBLAS (torch.nn.Linear) and cuDNN (mmcv ConvModule -> nn.Conv2d).
 * All operands bf16, fp32 accumulation in TMEM; D is [M, N]:
 *   HTD_DENSE_NT  D = A[M,K] . B[N,K]^T            lda, ldb = row pitch of the [.,K] matrices
 *   HTD_DENSE_NN  D = A[M,K] . B[K,N]              ldb = row pitch of the [K,N] matrix
 *   HTD_DENSE_TN  D = A[K,M]^T . B[K,N]            lda, ldb = row pitches of the [K,.] matrices
 *   HTD_DENSE_CONV_FPROP  A = W [Cout,3,3,Cin] (a channels-last conv weight), B = X [P,7,7,Cin]
 *                         -> D = Y [P,7,7,Cout]  (stride 1, padding 1, no bias)
 *   HTD_DENSE_CONV_DGRAD  A = W, B = dY [P,7,7,Cout] -> D = dX [P,7,7,Cin]
 *   HTD_DENSE_CONV_WGRAD  A = dY [P,7,7,Cout], B = X [P,7,7,Cin] -> D = dW [Cout,3,3,Cin]
 * Row pitches must be multiples of 8 elements, base pointers 16-byte aligned; nothing has to be
 * padded (the TMA unit zero-fills every tile tail).  Epilogue (GEMM kinds; conv kinds take relu /
 * gate only, on the channels-last output):  v = acc + bias[n];  D2[m,n] = act(v +
 * row_bias[row_class[m], n]) (optional second output, bf16, pitch ldd - htd_bbox_head.py:161-164:
 * the FC of `x + global_feat` shares the product with the FC of `x`);  D[m,n] = act(v) * [gate[m,n]
 * > 0] (gate = the forward activation of the layer whose ReLU the gradient passes).
 * splits: 0 = choose (fill the SMs), n > 0 = n k-slices; partial sums go through `workspace`
 * (htd_dense_gemm_workspace_bytes) and are added in slice order: results are deterministic.
 * htd_gate_colsum: dz = dy * [y > 0] (y, dz optional) and out[n] = sum_m dz[m,n] (out_dtype) -
 * the bias gradient of an FC layer and its ReLU backward in one pass; partial: [ceil(rows/64), N]
 * fp32. */
#define HTD_DENSE_NT 0
#define HTD_DENSE_NN 1
#define HTD_DENSE_TN 2
#define HTD_DENSE_CONV_FPROP 3
#define HTD_DENSE_CONV_DGRAD 4
#define HTD_DENSE_CONV_WGRAD 5
typedef struct HtdDenseGemm {
    int32_t kind;
    int32_t M, N, K;              /* GEMM kinds */
    int32_t P, Cin, Cout, pooled; /* conv kinds: RoIs, channels, map size (7) */
    int32_t d_dtype, relu, splits, bias_dtype; /* bias_dtype: HTD_F32 (default) or HTD_BF16 */
    const void* A;
    const void* B;
    void* D;
    void* D2;
    const void* bias;
    const float* row_bias;
    const int32_t* row_class;
    const void* gate;
    long long lda, ldb, ldd, ldg, ld_row_bias;
} HtdDenseGemm;
long long htd_dense_gemm_workspace_bytes(const HtdDenseGemm* g);
int htd_dense_gemm(const HtdDenseGemm* g, void* workspace, long long workspace_bytes,
                   htd_stream_t stream);
int htd_gate_colsum(const void* dy, long long ld_dy, const void* y, long long ld_y, int rows, int N,
                    void* dz, long long ld_dz, float* partial, void* out, int out_dtype,
                    htd_stream_t stream);
/* Backward glue of the dual-output FC (D / D2 above): dH, H [2M, N] bf16 (rows M.. = the D2 part),
 * cls [M] < R <= 8.  dz [M, N] = dH_a * [H_a > 0] + dH_b * [H_b > 0]; out [1 + R, N] (out_dtype):
 * row 0 = column sums of dz (bias gradient), row 1 + r = column sums of dH_b * [H_b > 0] over the
 * rows of class r (gradient of row_bias).  partial: [ceil(M/64), 1 + R, N] fp32. */
int htd_dual_gate(const void* dH, const void* H, const int32_t* cls, int M, int N, int R, void* dz,
                  float* partial, void* out, int out_dtype, htd_stream_t stream);
/* out = a + alpha * b + g[image(roi)] on channels-last bf16 RoI maps [P, PP, C] (g [B, C] or NULL,
 * rois [P, 5] give the image index): the regression-branch input of HTDBBoxHead
 * (htd_bbox_head.py:163,184: x_reg + global_feat + alpha * enhanced_feat). */
int htd_add3(const void* a, const void* b, float alpha, const void* g, const float* rois, int P,
             int PP, int C, int B, void* out, htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Target / loss / decode glue, one kernel per job, no host sync.
 *
 * htd_bbox_targets: BBoxHead._get_target_single (bbox_head.py:85-118) for all sampled RoIs at
 *   once.  boxes [K,4]; gt_boxes [K,4] / gt_labels [K] int64 valid on rows with is_pos[k] != 0;
 *   labels[k] = gt label or num_classes (background); label_weights = pos_weight (1 if <= 0) or 1;
 *   bbox_targets = DeltaXYWHBBoxCoder.encode (delta_xywh_bbox_coder.py:98-120) on positives, 0 else;
 *   bbox_weights = 1 on positives.  is_pos[k] == 2 marks a pad row of htd_assign_sample:
 *   background label, label_weight 0, no box target.
 * htd_bbox_decode: delta2bbox (delta_xywh_bbox_coder.py:123-204), class-agnostic [K,4] deltas;
 *   rois rows of roi_stride 4 or 5 floats (5: batch index first, copied to out when out_stride 5);
 *   clip != 0 clamps to [0,max_w] x [0,max_h].
 * htd_rcnn_loss_fwd: BBoxHead.loss (bbox_head.py:141-186) for class-agnostic regression:
 *   out4 = (w_cls * sum_k lw_k CE_k / max(#{lw > 0}, 1),  top-1 accuracy in %,
 *           w_bbox * sum_{k positive} sum_j bw_kj smoothL1_beta(pred - target) / K,  1 / avg_factor);
 *   dcls [K,num_cls1] / dbbox [K,4] receive the unnormalised gradients, partial is a
 *   [ceil(K/8), 4] fp32 workspace.  htd_rcnn_loss_bwd writes them, scaled by the incoming
 *   gradients g_cls / g_bbox (device scalars, NULL = 0), to dcls_out / dbbox_out; the saved
 *   buffers are only read, so the backward may run more than once.
 *   pad_rows != 0: rows with label_weight == 0 are the pad rows of htd_assign_sample (static
 *   shapes), not samples - they are left out of the accuracy, and the accuracy / loss_bbox
 *   denominators are the number of real rows max(#{lw > 0}, 1) instead of K (the reference's
 *   `bbox_targets.size(0)` counts sampled RoIs only, bbox_head.py:176-183). */
int htd_bbox_targets(const float* boxes, const float* gt_boxes, const long long* gt_labels,
                     const unsigned char* is_pos, int K, int num_classes, float pos_weight,
                     const float* means4, const float* stds4, long long* labels,
                     float* label_weights, float* bbox_targets, float* bbox_weights,
                     htd_stream_t stream);
int htd_bbox_decode(const float* rois, int roi_stride, const void* deltas, int delta_dtype, int K,
                    const float* means4, const float* stds4, float wh_ratio_clip, int clip,
                    float max_h, float max_w, float* out, int out_stride, htd_stream_t stream);
int htd_rcnn_loss_fwd(const void* cls_score, int num_cls1, const void* bbox_pred, int dtype,
                      const long long* labels, const float* label_weights,
                      const float* bbox_targets, const float* bbox_weights, int K, int num_classes,
                      float beta, float w_cls, float w_bbox, int pad_rows, void* dcls, void* dbbox,
                      float* partial, float* out4, htd_stream_t stream);
int htd_rcnn_loss_bwd(const void* dcls, long long ncls, const void* dbbox, long long nbox, int dtype,
                      const float* g_cls, const float* g_bbox, const float* out4, float w_cls,
                      float w_bbox, int K, int pad_rows, void* dcls_out, void* dbbox_out,
                      htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Proposal -> gt assignment + random sampling of one RoI-head stage, all images in one launch,
 * static output shapes, no host sync.  Replaces, per image,
 * MaxIoUAssigner.assign (core/bbox/assigners/max_iou_assigner.py:84-212; float thresholds,
 * gt_max_assign_all=True, no ignore regions), RandomSampler.sample (core/bbox/samplers/
 * base_sampler.py:34-101, random_sampler.py:56-78) and SamplingResult (sampling_result.py), called
 * from HTDRoIHead.forward_train (htd_roi_head.py:254-264,300-310).
 *   props [B,N,4]; valid [B,N] (NULL = all) marks the proposals that take part; gt_boxes [B,G,4],
 *   gt_labels [B,G] padded to G slots, num_gt [B] device integers.
 *   Candidate index c in [0, G+N): gt slot c (a candidate iff add_gt_as_proposals and c < num_gt)
 *   or proposal c - G - the reference's cat([gt_bboxes, bboxes]) order.  keys [B, G+N]: uniform
 *   random numbers >= 0; a class with more members than wanted keeps those with the smallest
 *   (key, c) - the distribution of randperm(n)[:want] - and, like the reference after `.unique()`,
 *   lists them in ascending candidate order.
 * Outputs, `num` rows per image = positives (<= num_pos), negatives, then pad rows:
 *   rois [B*num,5] (image index first; pad rows are zero-area boxes at the origin),
 *   kind [B*num] 1 positive / 0 negative / 2 pad, row_gt_boxes [B*num,4] / row_gt_labels [B*num] /
 *   row_gt_index [B*num] the matched gt of positive rows (0 / 0 / -1 elsewhere), row_is_gt [B*num]
 *   1 when the row is an appended gt box, row_cand [B*num] index into the reference's
 *   cat([gt_bboxes[:num_gt], bboxes]) (-1 on pad rows), counts [B,4] = sampled positives, sampled
 *   negatives, positive candidates, negative candidates.  Optional gt_inds / max_overlaps [B,G+N]:
 *   the AssignResult after add_gt_ (-2 / 0 on slots that are no candidates). */
#define HTD_MAX_GT 1024
#define HTD_MAX_CANDIDATES 16384
int htd_assign_sample(const float* props, const unsigned char* valid, int B, int N,
                      const float* gt_boxes, const long long* gt_labels, const int32_t* num_gt,
                      int G, const float* keys, float pos_iou_thr, float neg_iou_thr,
                      float min_pos_iou, int match_low_quality, int add_gt_as_proposals, int num,
                      int num_pos, float neg_pos_ub, float* rois, unsigned char* kind,
                      float* row_gt_boxes, long long* row_gt_labels, unsigned char* row_is_gt,
                      int32_t* row_cand, int32_t* row_gt_index, int32_t* counts, int32_t* gt_inds,
                      float* max_overlaps, htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Multi-class NMS of the test branch: multiclass_nms
 * (core/post_processing/bbox_nms.py:7-71, called from BBoxHead.get_bboxes, bbox_heads/
 * bbox_head.py:188-225) on top of mmcv.ops.batched_nms / nms (mmcv-full 1.2.1, un-vendored).
 *   boxes [K,4] (box_classes == 1, class-agnostic regression) or [K,C,4] (box_classes == C);
 *   scores [K,C+1], last column = background (ignored).  Candidates are the (k, c) with
 *   score > score_thr; their boxes are shifted by c * (max candidate coordinate + 1) in fp32 as
 *   batched_nms does; greedy suppression in descending score order when IoU > iou_thr; output =
 *   the survivors in descending score order (ties: ascending k*C + c), the first max_num.
 *   det [max_num,5] (x1,y1,x2,y2,score) and labels [max_num] are written for rows < count[0];
 *   workspace: htd_multiclass_nms_workspace_bytes(K, C) bytes.  No host sync, three launches
 *   (five for C <= 8 and K > 512 - the RPN's level-as-class call - where the pair tests are a
 *   bit matrix computed by the whole GPU and resolved sequentially per class). */
#define HTD_NMS_MAX_ROIS 4096
#define HTD_NMS_MAX_CLASSES 1024
long long htd_multiclass_nms_workspace_bytes(int K, int C);
int htd_multiclass_nms(const float* boxes, int box_classes, const float* scores, int K, int C,
                       float score_thr, float iou_thr, int max_num, float* det, long long* labels,
                       int32_t* count, void* workspace, htd_stream_t stream);

/* Soft-NMS form of the same step: nms=dict(type='soft_nms', iou_thr=0.5, min_score=0.05) of
 * configs/htd/htd_resnet101_2x.py:298 (and the three other R-101 / X-101 configs), dispatched at
 * mmdet/core/post_processing/bbox_nms.py:61 to mmcv.ops.batched_nms -> mmcv.ops.soft_nms
 * (mmcv-full 1.2.1, un-vendored; algorithm restated in oracle/soft_nms_ref.c).  Same candidates
 * and coordinate shift as above; then mmcv's sequential loop reproduced exactly (first maximum
 * of the current order, swap to the front, linear decay `score *= 1 - iou` where iou >= iou_thr
 * (method 1; method 0: naive = score 0), removal below min_score by overwriting with the last
 * box), so detections, decayed scores, labels AND the order among equal scores are the
 * reference's.  det [max_num,5] holds the un-shifted boxes and the decayed scores; rows <
 * count[0] are written.  One launch (one CTA), no host sync; the loop ends after max_num picks. */
long long htd_multiclass_soft_nms_workspace_bytes(int K, int C);
int htd_multiclass_soft_nms(const float* boxes, int box_classes, const float* scores, int K, int C,
                            float score_thr, float iou_thr, float min_score, int method,
                            int max_num, float* det, long long* labels, int32_t* count,
                            void* workspace, htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * FPN glue on channels-last maps - the producer of the head's pyramid (SURVEY.md section 8 row f4;
 * mmdet/models/necks/fpn.py:165-216).  Maps are [B, H, W, C] in memory, fp32 or bf16 (dtype), C a
 * multiple of 4 (fp32) / 8 (bf16).  The 1x1 lateral convolutions are htd_dense_gemm (kind NT).
 *   htd_fpn_topdown_fwd: out = fine + interpolate(coarse, size=(Hf, Wf), mode='nearest')
 *                        (fpn.py:187-190; ATen's source index min(floor(dst * in/out), in - 1));
 *                        out may alias fine.
 *   htd_fpn_topdown_bwd: dcoarse = the nearest-neighbour backward of dout (every coarse pixel sums
 *                        its fine pixels in ascending order, fp32 accumulation: deterministic);
 *                        the gradient of `fine` is dout itself.
 *   htd_fpn_subsample:   backward == 0: out [B, (H-1)/2+1, (W-1)/2+1, C] = in[:, ::2, ::2, :]
 *                        (max_pool2d(x, 1, stride=2), fpn.py:201); backward != 0: `in` is the
 *                        gradient of that output, out [B, H, W, C] receives it at the even pixels
 *                        and zero elsewhere. */
int htd_fpn_topdown_fwd(const void* fine, const void* coarse, void* out, int dtype, int B, int Hf,
                        int Wf, int Hc, int Wc, int C, htd_stream_t stream);
int htd_fpn_topdown_bwd(const void* dout, void* dcoarse, int dtype, int B, int Hf, int Wf, int Hc,
                        int Wc, int C, htd_stream_t stream);
int htd_fpn_subsample(const void* in, void* out, int dtype, int B, int H, int W, int C, int backward,
                      htd_stream_t stream);

/* Ranking step of the RPN proposal path (mmdet/models/dense_heads/rpn_head.py:125-134:
 * `scores.sort(descending=True)` and the first nms_pre entries): for each of `rows` rows of n fp32
 * keys (row r at keys + r * row_stride) the k largest in descending order - out_keys [rows, k] - and
 * their positions - out_idx [rows, k]; equal keys in ascending position (a stable sort; the
 * reference's sort is unstable).  1 <= k <= min(n, HTD_TOPK_MAX).  One CTA per row: radix select of
 * the k-th largest key, one collecting pass, bitonic sort in shared memory. */
#define HTD_TOPK_MAX 4096
int htd_topk_sorted(const float* keys, long long row_stride, int rows, int n, int k, float* out_keys,
                    int32_t* out_idx, htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Fused GroupNorm + ReLU of the regression conv tower (mmcv ConvModule conv -> GN -> ReLU,
 * htd_bbox_head.py:75-113,186).  x, y, dy, dx: [N, HW, C] channels-last, C % G == 0 and
 * (C / G) % 8 == 0; gamma / beta / dgamma / dbeta in the tensor dtype; mean / rstd [N*G]: fp32 for HTD_BF16 tensors, fp64 for
 * HTD_F32 tensors (the fp32 configuration computes its statistics and residuals in fp64).
 *   y = relu((x - mean) * rstd * gamma + beta), mean / rstd per (n, group), biased variance + eps
 * Backward recomputes the ReLU mask from x; part is a [2, N, C] fp32 workspace; dgamma / dbeta
 * are fully written. */
int htd_gn_relu_fwd(const void* x, int dtype, int N, int HW, int C, int G, const void* gamma,
                    const void* beta, float eps, void* y, void* mean, void* rstd,
                    htd_stream_t stream);
int htd_gn_relu_bwd(const void* x, const void* dy, int dtype, const void* mean, const void* rstd,
                    const void* gamma, const void* beta, int N, int HW, int C, int G, void* dx,
                    float* part, void* dgamma, void* dbeta, htd_stream_t stream);

/* ReLU + global average pool of the last tower conv (ConvModule without norm, htd_bbox_head.py:
 * 109-113, then avg_pool :188-189), channels-last x [N, HW, C], C % 8 == 0:
 *   y[n,c] = mean_hw relu(x[n,hw,c]);   dx[n,hw,c] = x[n,hw,c] > 0 ? g[n,c] / HW : 0. */
int htd_relu_mean_fwd(const void* x, int dtype, int N, int HW, int C, void* y, htd_stream_t stream);
int htd_relu_mean_bwd(const void* x, const void* g, int dtype, int N, int HW, int C, void* dx,
                      htd_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Device-side scheduling of the PGraph contractions: no host read of the plan table, so the whole
 * head step can be captured in a CUDA graph.  htd_pgraph_schedule derives, from the plan table on
 * the DEVICE, the problem descriptors of the six contraction shapes of the forward/backward pass
 * (one HtdGemmGroup per (level, image) group or per level block) and their tile prefix sums for
 * tiles of bm x bn outputs.  htd_pgraph_gemm_scheduled launches `max_tiles` CTAs (a static upper
 * bound from htd_pgraph_max_tiles); CTAs beyond the real tile count exit at once.
 *   set HTD_SCHED_GROUP_ND  : per group  M=n  N=d  K=n   a_row=off b_k0=off d_row=off dt_col=off
 *   set HTD_SCHED_GROUP_NN_S: per group  M=n  N=n  K=ds  a_row=off b_row=off d_row=off
 *   set HTD_SCHED_GROUP_NN_D: per group  M=n  N=n  K=d   a_row=off b_row=off d_row=off
 *   set HTD_SCHED_GROUP_NS  : per group  M=n  N=ds K=n   a_row=off b_k0=off d_row=off
 *   set HTD_SCHED_LEVEL_ND  : per level  M=ln N=d  K=d   a_row=loff b_row=l*d d_row=loff
 *                                        dt_col=loff bias_off=l*d
 *   set HTD_SCHED_LEVEL_DD  : per level  M=d  N=d  K=ln  a_k0=loff b_k0=loff d_row=l*d
 * sched layout: groups [HTD_SCHED_SETS][HTD_MAX_GROUPS] then int32 tile_start
 * [HTD_SCHED_SETS][HTD_MAX_GROUPS + 1]; HTD_SCHED_BYTES bytes in total. */
#define HTD_SCHED_GROUP_ND 0
#define HTD_SCHED_GROUP_NN_S 1
#define HTD_SCHED_GROUP_NN_D 2
#define HTD_SCHED_GROUP_NS 3
#define HTD_SCHED_LEVEL_ND 4
#define HTD_SCHED_LEVEL_DD 5
#define HTD_SCHED_SETS 6
#define HTD_SCHED_BYTES                                                     \
    (HTD_SCHED_SETS * HTD_MAX_GROUPS * (int)sizeof(HtdGemmGroup) +         \
     HTD_SCHED_SETS * (HTD_MAX_GROUPS + 1) * (int)sizeof(int32_t))

int htd_pgraph_schedule(const int32_t* table, int B, int L, int d, int ds, int ab_dtype,
                        void* sched, htd_stream_t stream);

/* Static upper bound of the tile count of a set: K RoIs in total, groups of at most max_group
 * RoIs, Ncap sorted rows of capacity. */
long long htd_pgraph_max_tiles(int set, int K, int B, int L, int max_group, int Ncap, int d,
                               int ds, int ab_dtype);

int htd_pgraph_gemm_scheduled(const void* A, long long a_rows, long long a_ld, const void* B,
                              long long b_rows, long long b_ld, int ab_dtype, const void* sched,
                              int set, long long max_tiles, void* D, int d_dtype, long long ldd,
                              const int32_t* d_rowmap, void* DT, int dt_dtype, long long ldt,
                              const float* bias, int relu, htd_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* HTD_B200_H_ */
