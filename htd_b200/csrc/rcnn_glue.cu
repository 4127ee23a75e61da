// Box-target / loss / decode glue of the RoI head as fused kernels.
//
// Reference path: BBoxHead.get_targets / _get_target_single
// (bbox_head.py:85-139) with DeltaXYWHBBoxCoder.encode (delta_xywh_bbox_coder.py:78-120),
// BBoxHead.loss (bbox_head.py:141-186) = CrossEntropyLoss + accuracy + SmoothL1Loss with boolean-mask
// row selection and `.item()` / `.any()` host syncs, and regress_by_class / delta2bbox
// (bbox_head.py:306-335, delta_xywh_bbox_coder.py:123-204).  In PyTorch these are ~200 kernels of a
// few microseconds per training step; here: one kernel per job, no host sync, fp32 arithmetic.
#include "common.cuh"

namespace htd {

// ------------------------------------------------------------------------------------------
// targets: labels / label_weights / bbox_targets / bbox_weights of every sampled RoI
// ------------------------------------------------------------------------------------------
// boxes [K,4] sampled boxes; gt_boxes [K,4] and gt_labels [K] hold the matched gt of POSITIVE rows
// (is_pos[k] != 0), anything elsewhere.  bbox_head.py:85-139 + delta_xywh_bbox_coder.py:98-120.
__global__ void __launch_bounds__(256) bbox_targets_kernel(
    const float* __restrict__ boxes, const float* __restrict__ gt_boxes,
    const long long* __restrict__ gt_labels, const unsigned char* __restrict__ is_pos, int K,
    int num_classes, float pos_weight, float m0, float m1, float m2, float m3, float s0, float s1,
    float s2, float s3, long long* __restrict__ labels, float* __restrict__ label_weights,
    float* __restrict__ bbox_targets, float* __restrict__ bbox_weights) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const bool pos = is_pos[k] == 1, pad = is_pos[k] == 2;
    float t[4] = {0.f, 0.f, 0.f, 0.f};
    if (pos) {
        const float4 p = *reinterpret_cast<const float4*>(boxes + (size_t)k * 4);
        const float4 g = *reinterpret_cast<const float4*>(gt_boxes + (size_t)k * 4);
        const float px = (p.x + p.z) * 0.5f, py = (p.y + p.w) * 0.5f;
        const float pw = p.z - p.x, ph = p.w - p.y;
        const float gx = (g.x + g.z) * 0.5f, gy = (g.y + g.w) * 0.5f;
        const float gw = g.z - g.x, gh = g.w - g.y;
        t[0] = ((gx - px) / pw - m0) / s0;
        t[1] = ((gy - py) / ph - m1) / s1;
        t[2] = (logf(gw / pw) - m2) / s2;
        t[3] = (logf(gh / ph) - m3) / s3;
    }
    labels[k] = pos ? gt_labels[k] : (long long)num_classes;
    label_weights[k] = pos ? (pos_weight <= 0.f ? 1.f : pos_weight) : (pad ? 0.f : 1.f);
    const float w = pos ? 1.f : 0.f;
    *reinterpret_cast<float4*>(bbox_targets + (size_t)k * 4) = make_float4(t[0], t[1], t[2], t[3]);
    *reinterpret_cast<float4*>(bbox_weights + (size_t)k * 4) = make_float4(w, w, w, w);
}

// ------------------------------------------------------------------------------------------
// decode: new_rois[k] = (batch, clip(delta2bbox(rois[k,1:5], deltas[k])))
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) bbox_decode_kernel(
    const float* __restrict__ rois, int roi_stride, const T* __restrict__ deltas, int K, float m0,
    float m1, float m2, float m3, float s0, float s1, float s2, float s3, float max_ratio,
    int clip, float max_h, float max_w, float* __restrict__ out, int out_stride) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float* r = rois + (size_t)k * roi_stride + (roi_stride - 4);
    const float dx = ldv<T>(deltas + (size_t)k * 4 + 0) * s0 + m0;
    const float dy = ldv<T>(deltas + (size_t)k * 4 + 1) * s1 + m1;
    float dw = ldv<T>(deltas + (size_t)k * 4 + 2) * s2 + m2;
    float dh = ldv<T>(deltas + (size_t)k * 4 + 3) * s3 + m3;
    dw = fminf(fmaxf(dw, -max_ratio), max_ratio);
    dh = fminf(fmaxf(dh, -max_ratio), max_ratio);
    const float px = (r[0] + r[2]) * 0.5f, py = (r[1] + r[3]) * 0.5f;
    const float pw = r[2] - r[0], ph = r[3] - r[1];
    const float gw = pw * expf(dw), gh = ph * expf(dh);
    const float gx = px + pw * dx, gy = py + ph * dy;
    float x1 = gx - gw * 0.5f, y1 = gy - gh * 0.5f, x2 = gx + gw * 0.5f, y2 = gy + gh * 0.5f;
    if (clip) {
        x1 = fminf(fmaxf(x1, 0.f), max_w); x2 = fminf(fmaxf(x2, 0.f), max_w);
        y1 = fminf(fmaxf(y1, 0.f), max_h); y2 = fminf(fmaxf(y2, 0.f), max_h);
    }
    float* o = out + (size_t)k * out_stride;
    if (out_stride == 5) { o[0] = rois[(size_t)k * roi_stride]; ++o; }
    o[0] = x1; o[1] = y1; o[2] = x2; o[3] = y2;
}

// ------------------------------------------------------------------------------------------
// loss: softmax cross entropy (+accuracy) and class-agnostic smooth-L1, with their gradients
// ------------------------------------------------------------------------------------------
constexpr int kLossWarps = 8;

// one warp per RoI.  partial[blockIdx.x][4] = sums of (weighted CE, [label_weight > 0], top-1 hits,
// weighted smooth-L1) over the block's rows; dcls / dbbox are the UNNORMALISED gradients
//   dcls[k][c] = lw_k (softmax_c - [c == label_k]),  dbbox[k][j] = bw_kj pos_k dsmoothl1(pred - tgt).
template <typename T>
__global__ void __launch_bounds__(kLossWarps * 32) rcnn_loss_rows_kernel(
    const T* __restrict__ cls_score, int num_cls1, const T* __restrict__ bbox_pred,
    const long long* __restrict__ labels, const float* __restrict__ label_weights,
    const float* __restrict__ bbox_targets, const float* __restrict__ bbox_weights, int K,
    int num_classes, float beta, int pad_rows, T* __restrict__ dcls, T* __restrict__ dbbox,
    float* __restrict__ partial) {
    __shared__ float s_part[kLossWarps][4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int k = blockIdx.x * kLossWarps + warp;
    float ce = 0.f, cnt = 0.f, hit = 0.f, sl1 = 0.f;
    if (k < K) {
        const T* z = cls_score + (size_t)k * num_cls1;
        const long long lab = labels[k];
        const float lw = label_weights[k];
        float mx = -INFINITY;
        int amax = 0;
        for (int c = lane; c < num_cls1; c += 32) {
            const float v = ldv<T>(z + c);
            if (v > mx) { mx = v; amax = c; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {                    // max with lowest-index tie break
            const float om = __shfl_xor_sync(0xffffffffu, mx, o);
            const int oa = __shfl_xor_sync(0xffffffffu, amax, o);
            if (om > mx || (om == mx && oa < amax)) { mx = om; amax = oa; }
        }
        float se = 0.f;
        for (int c = lane; c < num_cls1; c += 32) se += expf(ldv<T>(z + c) - mx);
        se = warp_sum(se);
        const float lse = logf(se) + mx;
        const float inv = 1.f / se;
        for (int c = lane; c < num_cls1; c += 32) {
            const float p = expf(ldv<T>(z + c) - mx) * inv;
            stv<T>(dcls + (size_t)k * num_cls1 + c, lw * (p - (c == (int)lab ? 1.f : 0.f)));
        }
        if (lane == 0) {
            const float zl = (lab >= 0 && lab < num_cls1) ? ldv<T>(z + lab) : 0.f;
            ce = (lse - zl) * lw;
            cnt = lw > 0.f ? 1.f : 0.f;
            hit = (amax == (int)lab && !(pad_rows && !(lw > 0.f))) ? 1.f : 0.f;
        }
        const bool pos = lab >= 0 && lab < num_classes;
        if (lane < 4) {
            const float d = ldv<T>(bbox_pred + (size_t)k * 4 + lane) - bbox_targets[(size_t)k * 4 + lane];
            const float w = pos ? bbox_weights[(size_t)k * 4 + lane] : 0.f;
            const float ad = fabsf(d);
            const float l = ad < beta ? 0.5f * ad * ad / beta : ad - 0.5f * beta;
            const float g = ad < beta ? d / beta : (d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f));
            sl1 = l * w;
            stv<T>(dbbox + (size_t)k * 4 + lane, g * w);
        }
        sl1 += __shfl_xor_sync(0xffffffffu, sl1, 1);
        sl1 += __shfl_xor_sync(0xffffffffu, sl1, 2);
    }
    if (lane == 0) { s_part[warp][0] = ce; s_part[warp][1] = cnt; s_part[warp][2] = hit; s_part[warp][3] = sl1; }
    __syncthreads();
    if (threadIdx.x < 4) {
        float t = 0.f;
#pragma unroll
        for (int w = 0; w < kLossWarps; ++w) t += s_part[w][threadIdx.x];
        partial[(size_t)blockIdx.x * 4 + threadIdx.x] = t;
    }
}

// out[0] = loss_cls, out[1] = acc (%), out[2] = loss_bbox, out[3] = 1 / avg_factor
__global__ void __launch_bounds__(256) rcnn_loss_final_kernel(const float* __restrict__ partial,
                                                              int nblk, int K, float w_cls,
                                                              float w_bbox, int pad_rows,
                                                              float* __restrict__ out) {
    __shared__ float s_red[4][8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = threadIdx.x; i < nblk; i += 256)
#pragma unroll
        for (int j = 0; j < 4; ++j) a[j] += partial[(size_t)i * 4 + j];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        a[j] = warp_sum(a[j]);
        if (lane == 0) s_red[j][warp] = a[j];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t[j] = 0.f;
            for (int w = 0; w < 8; ++w) t[j] += s_red[j][w];
        }
        const float avg = fmaxf(t[1], 1.f);
        out[0] = w_cls * t[0] / avg;
        const float rows = pad_rows ? avg : (float)K;
        out[1] = K > 0 ? 100.f * t[2] / rows : 0.f;
        out[2] = K > 0 ? w_bbox * t[3] / rows : 0.f;
        out[3] = 1.f / avg;
    }
}

// dcls_out = dcls * g_cls * w_cls / avg_factor ; dbbox_out = dbbox * g_bbox * w_bbox / K  (g_* are
// device scalars).  The saved unnormalised gradients are only READ, so the node can run backward
// more than once (retain_graph) without scaling them twice.
template <typename T>
__global__ void __launch_bounds__(256) rcnn_loss_bwd_kernel(const T* __restrict__ dcls, long long ncls,
                                                            const T* __restrict__ dbbox, long long nbox,
                                                            T* __restrict__ dcls_out,
                                                            T* __restrict__ dbbox_out,
                                                            const float* __restrict__ g_cls,
                                                            const float* __restrict__ g_bbox,
                                                            const float* __restrict__ out, float w_cls,
                                                            float w_bbox, int K, int pad_rows) {
    const float sc = (g_cls ? *g_cls : 0.f) * w_cls * out[3];
    const float sb = (g_bbox ? *g_bbox : 0.f) * w_bbox *
                     (pad_rows ? out[3] : 1.f / (float)(K > 0 ? K : 1));
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < ncls + nbox;
         i += (long long)gridDim.x * blockDim.x) {
        if (i < ncls) stv<T>(dcls_out + i, ldv<T>(dcls + i) * sc);
        else stv<T>(dbbox_out + (i - ncls), ldv<T>(dbbox + (i - ncls)) * sb);
    }
}

}  // namespace htd

using namespace htd;

extern "C" {

int htd_bbox_targets(const float* boxes, const float* gt_boxes, const long long* gt_labels,
                     const unsigned char* is_pos, int K, int num_classes, float pos_weight,
                     const float* means4, const float* stds4, long long* labels,
                     float* label_weights, float* bbox_targets, float* bbox_weights,
                     htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && num_classes >= 1 && means4 && stds4, "htd_bbox_targets: bad arguments");
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(boxes && gt_boxes && gt_labels && is_pos && labels && label_weights &&
                      bbox_targets && bbox_weights, "htd_bbox_targets: null pointer");
    bbox_targets_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(
        boxes, gt_boxes, gt_labels, is_pos, K, num_classes, pos_weight, means4[0], means4[1],
        means4[2], means4[3], stds4[0], stds4[1], stds4[2], stds4[3], labels, label_weights,
        bbox_targets, bbox_weights);
    HTD_CHECK_LAUNCH("htd_bbox_targets");
    return HTD_OK;
}

int htd_bbox_decode(const float* rois, int roi_stride, const void* deltas, int delta_dtype, int K,
                    const float* means4, const float* stds4, float wh_ratio_clip, int clip,
                    float max_h, float max_w, float* out, int out_stride, htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && (roi_stride == 4 || roi_stride == 5) && (out_stride == 4 || out_stride == 5) &&
                      means4 && stds4 && wh_ratio_clip > 0.f && out_stride <= roi_stride,
                  "htd_bbox_decode: bad arguments");
    HTD_CHECK_ARG(delta_dtype == HTD_F32 || delta_dtype == HTD_BF16, "htd_bbox_decode: bad dtype");
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(rois && deltas && out, "htd_bbox_decode: null pointer");
    const float max_ratio = fabsf(logf(wh_ratio_clip));
    cudaStream_t st = (cudaStream_t)stream;
    if (delta_dtype == HTD_F32)
        bbox_decode_kernel<float><<<(K + 255) / 256, 256, 0, st>>>(
            rois, roi_stride, static_cast<const float*>(deltas), K, means4[0], means4[1], means4[2],
            means4[3], stds4[0], stds4[1], stds4[2], stds4[3], max_ratio, clip, max_h, max_w, out,
            out_stride);
    else
        bbox_decode_kernel<__nv_bfloat16><<<(K + 255) / 256, 256, 0, st>>>(
            rois, roi_stride, static_cast<const __nv_bfloat16*>(deltas), K, means4[0], means4[1],
            means4[2], means4[3], stds4[0], stds4[1], stds4[2], stds4[3], max_ratio, clip, max_h,
            max_w, out, out_stride);
    HTD_CHECK_LAUNCH("htd_bbox_decode");
    return HTD_OK;
}

int htd_rcnn_loss_fwd(const void* cls_score, int num_cls1, const void* bbox_pred, int dtype,
                      const long long* labels, const float* label_weights,
                      const float* bbox_targets, const float* bbox_weights, int K, int num_classes,
                      float beta, float w_cls, float w_bbox, int pad_rows, void* dcls, void* dbbox,
                      float* partial, float* out4, htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && num_cls1 >= 2 && num_classes >= 1 && beta > 0.f,
                  "htd_rcnn_loss_fwd: bad arguments");
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_rcnn_loss_fwd: bad dtype");
    HTD_CHECK_ARG(out4 && (K == 0 || (cls_score && bbox_pred && labels && label_weights &&
                                      bbox_targets && bbox_weights && dcls && dbbox && partial)),
                  "htd_rcnn_loss_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int nblk = (K + kLossWarps - 1) / kLossWarps;
    if (nblk > 0) {
        if (dtype == HTD_F32)
            rcnn_loss_rows_kernel<float><<<nblk, kLossWarps * 32, 0, st>>>(
                static_cast<const float*>(cls_score), num_cls1, static_cast<const float*>(bbox_pred),
                labels, label_weights, bbox_targets, bbox_weights, K, num_classes, beta, pad_rows,
                static_cast<float*>(dcls), static_cast<float*>(dbbox), partial);
        else
            rcnn_loss_rows_kernel<__nv_bfloat16><<<nblk, kLossWarps * 32, 0, st>>>(
                static_cast<const __nv_bfloat16*>(cls_score), num_cls1,
                static_cast<const __nv_bfloat16*>(bbox_pred), labels, label_weights, bbox_targets,
                bbox_weights, K, num_classes, beta, pad_rows, static_cast<__nv_bfloat16*>(dcls),
                static_cast<__nv_bfloat16*>(dbbox), partial);
        HTD_CHECK_LAUNCH("htd_rcnn_loss_fwd(rows)");
    }
    rcnn_loss_final_kernel<<<1, 256, 0, st>>>(partial, nblk, K, w_cls, w_bbox, pad_rows, out4);
    HTD_CHECK_LAUNCH("htd_rcnn_loss_fwd(final)");
    return HTD_OK;
}

int htd_rcnn_loss_bwd(const void* dcls, long long ncls, const void* dbbox, long long nbox, int dtype,
                      const float* g_cls, const float* g_bbox, const float* out4, float w_cls,
                      float w_bbox, int K, int pad_rows, void* dcls_out, void* dbbox_out,
                      htd_stream_t stream) {
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_rcnn_loss_bwd: bad dtype");
    HTD_CHECK_ARG(ncls >= 0 && nbox >= 0 && out4, "htd_rcnn_loss_bwd: bad arguments");
    if (ncls + nbox == 0) return HTD_OK;
    HTD_CHECK_ARG(dcls && dbbox && dcls_out && dbbox_out, "htd_rcnn_loss_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = ncls + nbox;
    const unsigned blocks = (unsigned)((n + 255) / 256 < 592 ? (n + 255) / 256 : 592);
    if (dtype == HTD_F32)
        rcnn_loss_bwd_kernel<float><<<blocks, 256, 0, st>>>(
            static_cast<const float*>(dcls), ncls, static_cast<const float*>(dbbox), nbox,
            static_cast<float*>(dcls_out), static_cast<float*>(dbbox_out), g_cls, g_bbox, out4, w_cls,
            w_bbox, K, pad_rows);
    else
        rcnn_loss_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(dcls), ncls, static_cast<const __nv_bfloat16*>(dbbox),
            nbox, static_cast<__nv_bfloat16*>(dcls_out), static_cast<__nv_bfloat16*>(dbbox_out), g_cls,
            g_bbox, out4, w_cls, w_bbox, K, pad_rows);
    HTD_CHECK_LAUNCH("htd_rcnn_loss_bwd");
    return HTD_OK;
}

}  // extern "C"
