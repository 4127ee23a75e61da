/* ORACLE (test infrastructure; never linked into the product).
 *
 * Soft-NMS of mmcv-full 1.2.1 (`mmcv.ops.soft_nms` -> ext `softnms`, CPU implementation
 * `softnms_cpu`; the dependency is pinned at README.md:11 of the reference and is NOT vendored
 * there).  Restated here from its published algorithm - the Bodla et al. reference loop that mmcv
 * took over from the original Cython code: select the current maximum, swap it to the front,
 * decay the scores of the remaining boxes by their overlap with it, drop a box whose score falls
 * below min_score by overwriting it with the last box.  Called by the reference through
 * mmcv.ops.batched_nms from mmdet/core/post_processing/bbox_nms.py:61 with
 * nms_cfg = dict(type='soft_nms', iou_thr=0.5, min_score=0.05) (configs/htd/htd_resnet101_2x.py:298),
 * i.e. method 'linear' (mmcv default), sigma 0.5, offset 0.
 *
 * All arithmetic in fp32 in the order written (compile with -ffp-contract=off).
 *   boxes  [n,4], scores [n]   inputs (not modified)
 *   dets   [n,5]  out: selected boxes + decayed scores, selection order
 *   inds   [n]    out: original indices of the selected boxes
 *   method 0 naive, 1 linear, 2 gaussian
 * returns the number of selected boxes. */
#include <math.h>
#include <stdlib.h>
#include <string.h>

long long oracle_soft_nms(const float* boxes, const float* scores, long long n, float iou_threshold,
                          float sigma, float min_score, int method, int offset, float* dets,
                          long long* inds) {
    float* x1 = malloc(sizeof(float) * (n > 0 ? n : 1));
    float* y1 = malloc(sizeof(float) * (n > 0 ? n : 1));
    float* x2 = malloc(sizeof(float) * (n > 0 ? n : 1));
    float* y2 = malloc(sizeof(float) * (n > 0 ? n : 1));
    float* sc = malloc(sizeof(float) * (n > 0 ? n : 1));
    float* areas = malloc(sizeof(float) * (n > 0 ? n : 1));
    const float off = (float)offset;
    for (long long i = 0; i < n; ++i) {
        x1[i] = boxes[4 * i + 0];
        y1[i] = boxes[4 * i + 1];
        x2[i] = boxes[4 * i + 2];
        y2[i] = boxes[4 * i + 3];
        sc[i] = scores[i];
        areas[i] = (x2[i] - x1[i] + off) * (y2[i] - y1[i] + off);
        inds[i] = i;
    }
    long long nboxes = n;
    for (long long i = 0; i < nboxes; ++i) {
        float max_score = sc[i];
        long long max_pos = i;
        long long pos = i + 1;
        while (pos < nboxes) {                 /* first maximum in the current order */
            if (max_score < sc[pos]) {
                max_score = sc[pos];
                max_pos = pos;
            }
            pos = pos + 1;
        }
        /* swap the maximum to position i, record it */
        const float ix1 = dets[i * 5 + 0] = x1[max_pos];
        const float iy1 = dets[i * 5 + 1] = y1[max_pos];
        const float ix2 = dets[i * 5 + 2] = x2[max_pos];
        const float iy2 = dets[i * 5 + 3] = y2[max_pos];
        const float iscore = dets[i * 5 + 4] = sc[max_pos];
        const float iarea = areas[max_pos];
        const long long iind = inds[max_pos];
        x1[max_pos] = x1[i];
        y1[max_pos] = y1[i];
        x2[max_pos] = x2[i];
        y2[max_pos] = y2[i];
        sc[max_pos] = sc[i];
        areas[max_pos] = areas[i];
        inds[max_pos] = inds[i];
        x1[i] = ix1;
        y1[i] = iy1;
        x2[i] = ix2;
        y2[i] = iy2;
        sc[i] = iscore;
        areas[i] = iarea;
        inds[i] = iind;

        pos = i + 1;
        while (pos < nboxes) {
            const float xx1 = fmaxf(ix1, x1[pos]);
            const float yy1 = fmaxf(iy1, y1[pos]);
            const float xx2 = fminf(ix2, x2[pos]);
            const float yy2 = fminf(iy2, y2[pos]);
            const float w = fmaxf(0.f, xx2 - xx1 + off);
            const float h = fmaxf(0.f, yy2 - yy1 + off);
            const float inter = w * h;
            const float ovr = inter / (iarea + areas[pos] - inter);
            float weight = 1.f;
            if (method == 0) {
                if (ovr >= iou_threshold) weight = 0.f;
            } else if (method == 1) {
                if (ovr >= iou_threshold) weight = 1.f - ovr;
            } else if (method == 2) {
                weight = expf(-(ovr * ovr) / sigma);
            }
            sc[pos] *= weight;
            if (sc[pos] < min_score) {         /* drop: overwrite with the last box, re-examine */
                x1[pos] = x1[nboxes - 1];
                y1[pos] = y1[nboxes - 1];
                x2[pos] = x2[nboxes - 1];
                y2[pos] = y2[nboxes - 1];
                sc[pos] = sc[nboxes - 1];
                areas[pos] = areas[nboxes - 1];
                inds[pos] = inds[nboxes - 1];
                nboxes = nboxes - 1;
                pos = pos - 1;
            }
            pos = pos + 1;
        }
    }
    free(x1); free(y1); free(x2); free(y2); free(sc); free(areas);
    return nboxes;
}
