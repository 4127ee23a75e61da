/*
 * ORACLE (test infrastructure, not product code).
 *
 * Plain-C restatement of the RoIAlign arithmetic the reference reaches through
 * `mmcv.ops.RoIAlign` (mmcv-full 1.2.1, pinned in /root/reference/README.md:11, NOT
 * vendored in the reference tree).  Call sites it stands in for:
 *   mmdet/models/roi_heads/roi_extractors/base_roi_extractor.py:49-55  (construction)
 *   mmdet/models/roi_heads/roi_extractors/single_level_roi_extractor.py:93
 *   mmdet/models/roi_heads/roi_extractors/adaptative_roi_extractor.py:72,87
 * Algorithm: the published detectron2 / mmcv / torchvision "ROIAlign, aligned=True,
 * sampling_ratio=0 (adaptive grid), avg pool" — SURVEY.md §8 row a4/a5.  Pinned in
 * tests/test_oracle_cpu.py against torchvision.ops.roi_align (CPU) forward and its
 * autograd backward, in fp32 and fp64.
 *
 * Layout: input [B,C,H,W] contiguous (NCHW), rois [K,5] = (batch, x1, y1, x2, y2),
 * output [K,C,PH,PW].  One OpenMP thread per RoI (forward) / per (batch, channel)
 * plane (backward, so no atomics are needed on the host either).
 */
#include <math.h>
#include <stddef.h>
#include <string.h>

#define DEFINE_ROI_ALIGN(T, SUFFIX, CEIL)                                                  \
                                                                                           \
/* bilinear tap set for one sample point; w[]=0 and idx=-1 when out of range */            \
static void taps_##SUFFIX(T y, T x, int H, int W, int pos[4], T w[4]) {                    \
    if (y < (T)-1.0 || y > (T)H || x < (T)-1.0 || x > (T)W) {                              \
        pos[0] = pos[1] = pos[2] = pos[3] = -1;                                            \
        w[0] = w[1] = w[2] = w[3] = 0;                                                     \
        return;                                                                            \
    }                                                                                      \
    if (y <= 0) y = 0;                                                                     \
    if (x <= 0) x = 0;                                                                     \
    int y_low = (int)y, x_low = (int)x, y_high, x_high;                                    \
    if (y_low >= H - 1) { y_high = y_low = H - 1; y = (T)y_low; } else y_high = y_low + 1; \
    if (x_low >= W - 1) { x_high = x_low = W - 1; x = (T)x_low; } else x_high = x_low + 1; \
    T ly = y - y_low, lx = x - x_low, hy = (T)1.0 - ly, hx = (T)1.0 - lx;                  \
    pos[0] = y_low * W + x_low;  w[0] = hy * hx;                                           \
    pos[1] = y_low * W + x_high; w[1] = hy * lx;                                           \
    pos[2] = y_high * W + x_low; w[2] = ly * hx;                                           \
    pos[3] = y_high * W + x_high; w[3] = ly * lx;                                          \
}                                                                                          \
                                                                                           \
void roi_align_fwd_##SUFFIX(const T *input, const T *rois, T *output, int B, int C,        \
                            int H, int W, int K, int PH, int PW, T spatial_scale,          \
                            int sampling_ratio, int aligned) {                             \
    (void)B;                                                                               \
    _Pragma("omp parallel for schedule(dynamic, 1)")                                       \
    for (int n = 0; n < K; ++n) {                                                          \
        const T *r = rois + (size_t)n * 5;                                                 \
        int b = (int)r[0];                                                                 \
        T off = aligned ? (T)0.5 : (T)0.0;                                                 \
        T sw = r[1] * spatial_scale - off, sh = r[2] * spatial_scale - off;                \
        T ew = r[3] * spatial_scale - off, eh = r[4] * spatial_scale - off;                \
        T rw = ew - sw, rh = eh - sh;                                                      \
        if (!aligned) { if (rw < (T)1.0) rw = (T)1.0; if (rh < (T)1.0) rh = (T)1.0; }      \
        T bh = rh / (T)PH, bw = rw / (T)PW;                                                \
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)CEIL(rh / (T)PH);              \
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)CEIL(rw / (T)PW);              \
        T count = (T)((gh * gw) > 1 ? (gh * gw) : 1);                                      \
        for (int c = 0; c < C; ++c) {                                                      \
            const T *plane = input + ((size_t)b * C + c) * H * W;                          \
            T *out = output + ((size_t)n * C + c) * PH * PW;                               \
            for (int ph = 0; ph < PH; ++ph)                                                \
                for (int pw = 0; pw < PW; ++pw) {                                          \
                    T acc = 0;                                                             \
                    for (int iy = 0; iy < gh; ++iy) {                                      \
                        T y = sh + ph * bh + ((T)iy + (T)0.5) * bh / (T)gh;                \
                        for (int ix = 0; ix < gw; ++ix) {                                  \
                            T x = sw + pw * bw + ((T)ix + (T)0.5) * bw / (T)gw;            \
                            int pos[4]; T w[4];                                            \
                            taps_##SUFFIX(y, x, H, W, pos, w);                             \
                            if (pos[0] < 0) continue;                                      \
                            acc += w[0] * plane[pos[0]] + w[1] * plane[pos[1]] +           \
                                   w[2] * plane[pos[2]] + w[3] * plane[pos[3]];            \
                        }                                                                  \
                    }                                                                      \
                    out[ph * PW + pw] = acc / count;                                       \
                }                                                                          \
        }                                                                                  \
    }                                                                                      \
}                                                                                          \
                                                                                           \
/* grad_input must be zero-filled by the caller (the reference does new_zeros). */         \
void roi_align_bwd_##SUFFIX(const T *grad_output, const T *rois, T *grad_input, int B,     \
                            int C, int H, int W, int K, int PH, int PW, T spatial_scale,   \
                            int sampling_ratio, int aligned) {                             \
    _Pragma("omp parallel for collapse(2) schedule(dynamic, 4)")                           \
    for (int b = 0; b < B; ++b)                                                            \
        for (int c = 0; c < C; ++c) {                                                      \
            T *plane = grad_input + ((size_t)b * C + c) * H * W;                           \
            for (int n = 0; n < K; ++n) {                                                  \
                const T *r = rois + (size_t)n * 5;                                         \
                if ((int)r[0] != b) continue;                                              \
                T off = aligned ? (T)0.5 : (T)0.0;                                         \
                T sw = r[1] * spatial_scale - off, sh = r[2] * spatial_scale - off;        \
                T ew = r[3] * spatial_scale - off, eh = r[4] * spatial_scale - off;        \
                T rw = ew - sw, rh = eh - sh;                                              \
                if (!aligned) {                                                            \
                    if (rw < (T)1.0) rw = (T)1.0;                                          \
                    if (rh < (T)1.0) rh = (T)1.0;                                          \
                }                                                                          \
                T bh = rh / (T)PH, bw = rw / (T)PW;                                        \
                int gh = sampling_ratio > 0 ? sampling_ratio : (int)CEIL(rh / (T)PH);      \
                int gw = sampling_ratio > 0 ? sampling_ratio : (int)CEIL(rw / (T)PW);      \
                T count = (T)((gh * gw) > 1 ? (gh * gw) : 1);                              \
                const T *go = grad_output + ((size_t)n * C + c) * PH * PW;                 \
                for (int ph = 0; ph < PH; ++ph)                                            \
                    for (int pw = 0; pw < PW; ++pw) {                                      \
                        T g = go[ph * PW + pw] / count;                                    \
                        for (int iy = 0; iy < gh; ++iy) {                                  \
                            T y = sh + ph * bh + ((T)iy + (T)0.5) * bh / (T)gh;            \
                            for (int ix = 0; ix < gw; ++ix) {                              \
                                T x = sw + pw * bw + ((T)ix + (T)0.5) * bw / (T)gw;        \
                                int pos[4]; T w[4];                                        \
                                taps_##SUFFIX(y, x, H, W, pos, w);                         \
                                if (pos[0] < 0) continue;                                  \
                                plane[pos[0]] += g * w[0];                                 \
                                plane[pos[1]] += g * w[1];                                 \
                                plane[pos[2]] += g * w[2];                                 \
                                plane[pos[3]] += g * w[3];                                 \
                            }                                                              \
                        }                                                                  \
                    }                                                                      \
            }                                                                              \
        }                                                                                  \
}

DEFINE_ROI_ALIGN(float, f32, ceilf)
DEFINE_ROI_ALIGN(double, f64, ceil)

/*
 * FPN level assignment, SingleRoIExtractor.map_roi_levels
 * (single_level_roi_extractor.py:32-51; duplicate in htd_bbox_head.py:129-135):
 *   lvl = clamp(floor(log2(sqrt((x2-x1)*(y2-y1)) / finest_scale + 1e-6)), 0, L-1)
 * evaluated in fp32 exactly as torch does (sqrtf, IEEE division, log2f, floorf).
 * NaN scale (negative area) follows torch: clamp keeps NaN, .long() of NaN is
 * implementation-defined; we return 0 like x86 cvttss2si -> INT_MIN -> clamp does not
 * apply after the cast in the reference.  Such boxes never reach the head (clipped
 * proposals have x2>=x1, y2>=y1) and are excluded from parity tests.
 */
void map_roi_levels_f32(const float *rois, long long *lvls, int K, int num_levels,
                        float finest_scale) {
    for (int n = 0; n < K; ++n) {
        const float *r = rois + (size_t)n * 5;
        float a = (r[3] - r[1]) * (r[4] - r[2]);
        float s = sqrtf(a);
        float t = floorf(log2f(s / finest_scale + 1e-6f));
        if (t < 0.f) t = 0.f;
        if (t > (float)(num_levels - 1)) t = (float)(num_levels - 1);
        lvls[n] = (t == t) ? (long long)t : 0;
    }
}
