// TEST INFRASTRUCTURE: host build of htd_b200/csrc/roi_axis.h (the exact header the CUDA
// kernels include) driving a straightforward loop nest with the same separable algorithm the
// kernels implement.  Lets the CPU suite check the axis-weight / bin-range / footprint / level
// math against the oracle without a GPU.  Not product code; never loaded by htd_b200.
#include <cstddef>
#include <vector>
#include "../../htd_b200/csrc/roi_axis.h"

using namespace htd;

extern "C" {

// feat: [B,H,W,C] fp32 NHWC; out: [K,P*P,C]
void emu_roi_align_fwd(const float* feat, const float* rois, float* out, int B, int C, int H,
                       int W, int K, int P, double scale, int sampling_ratio) {
#pragma omp parallel for schedule(dynamic, 1)
    for (int k = 0; k < K; ++k) {
        const float* r = rois + (size_t)k * 5;
        int b = (int)r[0];
        Axis ay = make_axis(r[2], r[4], scale, P, H, sampling_ratio, 1);
        Axis ax = make_axis(r[1], r[3], scale, P, W, sampling_ratio, 1);
        for (int ph = 0; ph < P; ++ph)
            for (int pw = 0; pw < P; ++pw) {
                float* o = out + ((size_t)k * P * P + ph * P + pw) * C;
                for (int c = 0; c < C; ++c) o[c] = 0.f;
                if (b < 0 || b >= B) continue;
                int r0, r1, c0, c1;
                bin_range(ay, ph, r0, r1);
                bin_range(ax, pw, c0, c1);
                for (int y = r0; y <= r1; ++y) {
                    float wy = axis_weight(ay, ph, y);
                    for (int x = c0; x <= c1; ++x) {
                        float w = wy * axis_weight(ax, pw, x);
                        const float* f = feat + (((size_t)b * H + y) * W + x) * C;
                        for (int c = 0; c < C; ++c) o[c] += w * f[c];
                    }
                }
            }
    }
}

// Checks that weights vanish outside bin_range / roi_range (what the kernels rely on):
// returns the number of (roi, bin, pixel) triples with non-zero weight outside the range.
int emu_check_ranges(const float* rois, int K, int P, int H, int W, double scale,
                     int sampling_ratio) {
    int bad = 0;
    for (int k = 0; k < K; ++k) {
        const float* r = rois + (size_t)k * 5;
        for (int axis = 0; axis < 2; ++axis) {
            int L = axis ? W : H;
            Axis a = axis ? make_axis(r[1], r[3], scale, P, W, sampling_ratio, 1)
                          : make_axis(r[2], r[4], scale, P, H, sampling_ratio, 1);
            int ulo, uhi;
            roi_range(a, P, ulo, uhi);
            for (int p = 0; p < P; ++p) {
                int lo, hi;
                bin_range(a, p, lo, hi);
                for (int j = 0; j < L; ++j) {
                    float w = axis_weight(a, p, j);
                    bool inside = (j >= lo && j <= hi);
                    if (!inside && w != 0.f) ++bad;
                    if (w != 0.f && !(j >= ulo && j <= uhi)) ++bad;
                    if (w < 0.f) ++bad;
                }
                if (hi >= lo && (axis_weight(a, p, lo) == 0.f && axis_weight(a, p, hi) == 0.f &&
                                 hi - lo > 1)) ++bad;
            }
        }
    }
    return bad;
}

// dY: [K,P*P,C]; dX: [B,H,W,C] (fully written)
void emu_roi_align_bwd(const float* dy, const float* rois, float* dx, int B, int C, int H, int W,
                       int K, int P, double scale, int sampling_ratio) {
    for (size_t i = 0; i < (size_t)B * H * W * C; ++i) dx[i] = 0.f;
    for (int k = 0; k < K; ++k) {
        const float* r = rois + (size_t)k * 5;
        int b = (int)r[0];
        if (b < 0 || b >= B) continue;
        Axis ay = make_axis(r[2], r[4], scale, P, H, sampling_ratio, 1);
        Axis ax = make_axis(r[1], r[3], scale, P, W, sampling_ratio, 1);
        int r0, r1, c0, c1;
        roi_range(ay, P, r0, r1);
        roi_range(ax, P, c0, c1);
        for (int y = r0; y <= r1; ++y)
            for (int x = c0; x <= c1; ++x) {
                float* d = dx + (((size_t)b * H + y) * W + x) * C;
                for (int ph = 0; ph < P; ++ph) {
                    float wy = axis_weight(ay, ph, y);
                    if (wy == 0.f) continue;
                    for (int pw = 0; pw < P; ++pw) {
                        float w = wy * axis_weight(ax, pw, x);
                        if (w == 0.f) continue;
                        const float* g = dy + ((size_t)k * P * P + ph * P + pw) * C;
                        for (int c = 0; c < C; ++c) d[c] += w * g[c];
                    }
                }
            }
    }
}

void emu_levels(const float* rois, int K, int num_levels, float finest, int* out) {
    for (int k = 0; k < K; ++k) {
        const float* r = rois + (size_t)k * 5;
        out[k] = roi_level(r[1], r[2], r[3], r[4], finest, num_levels);
    }
}
}
