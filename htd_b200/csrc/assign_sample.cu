// Proposal -> ground-truth assignment and random sampling between / before the two RoI-head stages,
// as ONE kernel per stage with static output shapes.
//
// Reference path, per image, in Python with host syncs (nonzero / randperm / unique):
//   MaxIoUAssigner.assign / assign_wrt_overlaps   core/bbox/assigners/max_iou_assigner.py:84-212
//   bbox_overlaps(gt, proposals)                  core/bbox/iou_calculators/iou2d_calculator.py:129-150
//   BaseSampler.sample                            core/bbox/samplers/base_sampler.py:34-101
//   RandomSampler._sample_pos/_sample_neg         core/bbox/samplers/random_sampler.py:56-78
//   SamplingResult                                core/bbox/samplers/sampling_result.py
// called from HTDRoIHead.forward_train (htd_roi_head.py:254-264, 300-310).
//
// Differences that keep the result identical and the shapes static:
//  * the random subset is chosen by caller-provided uniform keys: "the `want` candidates with the
//    smallest (key, index)" is the same distribution as randperm(n)[:want]; like the reference
//    (`.unique()` sorts), the chosen indices are emitted in ascending candidate order;
//  * every image yields exactly `num` rows: positives, then negatives, then zero-area pad rows
//    (kind 2) when the image has fewer than `num` candidates to give; counts are device integers.
#include "common.cuh"

namespace htd {

constexpr int kAsThreads = 1024;

struct AsParams {
    const float4* props;            // [B, N]
    const unsigned char* valid;     // [B, N] or null
    const float4* gt_boxes;         // [B, G]
    const long long* gt_labels;     // [B, G]
    const int* num_gt;              // [B]
    const float* keys;              // [B, G + N]
    int B, N, G;
    float pos_thr, neg_thr, min_pos_iou, neg_pos_ub;
    int match_low_quality, add_gt, num, num_pos;
    float* rois;                    // [B * num, 5]
    unsigned char* kind;            // [B * num]
    float4* row_gt_box;             // [B * num]
    long long* row_gt_label;        // [B * num]
    unsigned char* row_is_gt;       // [B * num]
    int* row_cand;                  // [B * num]
    int* row_gt_index;              // [B * num]
    int* counts;                    // [B, 4]
    int* gt_inds;                   // [B, G + N] or null
    float* max_ov;                  // [B, G + N] or null
};

// bbox_overlaps(mode='iou'), fp32, the reference's operation order; _rn intrinsics keep ptxas
// from contracting a*b-c into an fma, so thresholds compare the same bits as the oracle.
__device__ __forceinline__ float iou_rn(const float4 a, const float4 b) {
    const float area1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float ltx = fmaxf(a.x, b.x), lty = fmaxf(a.y, b.y);
    const float rbx = fminf(a.z, b.z), rby = fminf(a.w, b.w);
    const float w = fmaxf(__fsub_rn(rbx, ltx), 0.f), h = fmaxf(__fsub_rn(rby, lty), 0.f);
    const float overlap = __fmul_rn(w, h);
    const float uni = fmaxf(__fsub_rn(__fadd_rn(area1, area2), overlap), 1e-6f);
    return __fdiv_rn(overlap, uni);
}

// exclusive scan of one int per thread over the CTA, thread order; `total` = sum over the CTA
__device__ __forceinline__ int block_excl_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                                   // s_warp free for reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int t = s_warp[lane], ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += u;
        }
        s_warp[lane] = ti - t;
        if (lane == 31) s_warp[32] = ti;
    }
    __syncthreads();
    total = s_warp[32];
    return incl - v + s_warp[warp];
}

// One CTA per image.  Candidate c in [0, G + N): c < G is gt slot c (a candidate iff add_gt and
// c < num_gt), c >= G is proposal c - G.  Thread t owns the contiguous candidates
// [t * per, (t + 1) * per), so thread order == candidate order for the scans.
__global__ void __launch_bounds__(kAsThreads) assign_sample_kernel(const AsParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    __shared__ int s_warp[33];
    __shared__ int s_hist[256];
    __shared__ unsigned s_sel[2];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int G = p.G, N = p.N, n = G + N;
    float4* s_gt = reinterpret_cast<float4*>(smem);                       // [G]
    int* s_ind = reinterpret_cast<int*>(s_gt + G);                        // [n]  -2 not a candidate
    unsigned* s_key = reinterpret_cast<unsigned*>(s_ind + n);             // [n]
    unsigned* s_gtmax = s_key + n;                                        // [G]
    const int g = min(max(p.num_gt[b], 0), G);
    for (int i = tid; i < G; i += kAsThreads) {
        s_gt[i] = i < g ? p.gt_boxes[(size_t)b * G + i] : make_float4(0.f, 0.f, 0.f, 0.f);
        s_gtmax[i] = 0u;
    }
    __syncthreads();
    const float4* props = p.props + (size_t)b * N;
    const unsigned char* valid = p.valid ? p.valid + (size_t)b * N : nullptr;
    const int per = (n + kAsThreads - 1) / kAsThreads;
    const int c_lo = min(tid * per, n), c_hi = min(c_lo + per, n);

    if (p.match_low_quality && g > 0) {               // gt_max over the proposals (:185-186)
        for (int c = max(c_lo, G); c < c_hi; ++c) {
            if (valid && !valid[c - G]) continue;
            const float4 bx = props[c - G];
            for (int i = 0; i < g; ++i)
                atomicMax(&s_gtmax[i], __float_as_uint(iou_rn(s_gt[i], bx)));   // IoU >= 0
        }
        __syncthreads();
    }

    // ---- assignment (max_iou_assigner.py:159-212), then add_gt_ (assign_result.py, base_sampler.py:73-81)
    int my_pos = 0, my_neg = 0;
    for (int c = c_lo; c < c_hi; ++c) {
        int ind = -2;
        float mo = 0.f;
        if (c < G) {
            if (p.add_gt && c < g) { ind = c + 1; mo = 1.f; }
        } else if (!valid || valid[c - G]) {
            if (g == 0) {
                ind = 0;
            } else {
                const float4 bx = props[c - G];
                float best = -1.f;
                int arg = 0;
                for (int i = 0; i < g; ++i) {
                    const float v = iou_rn(s_gt[i], bx);
                    if (v > best) { best = v; arg = i; }        // first maximum
                }
                mo = best;
                ind = -1;
                if (best >= 0.f && best < p.neg_thr) ind = 0;
                if (best >= p.pos_thr) ind = arg + 1;
                if (p.match_low_quality)
                    for (int i = 0; i < g; ++i) {               // later gts override earlier ones
                        const float gm = __uint_as_float(s_gtmax[i]);
                        if (gm >= p.min_pos_iou && iou_rn(s_gt[i], bx) == gm) ind = i + 1;
                    }
            }
        }
        s_ind[c] = ind;
        const float key = p.keys[(size_t)b * n + c];
        s_key[c] = __float_as_uint(fmaxf(key, 0.f));
        my_pos += ind > 0;
        my_neg += ind == 0;
        if (p.gt_inds) p.gt_inds[(size_t)b * n + c] = ind;
        if (p.max_ov) p.max_ov[(size_t)b * n + c] = mo;
    }
    int npos_c, nneg_c;
    block_excl_scan(my_pos, s_warp, npos_c);
    block_excl_scan(my_neg, s_warp, nneg_c);

    const int want_pos = min(npos_c, p.num_pos);
    int num_neg = p.num - want_pos;                             // base_sampler.py:89-94
    if (p.neg_pos_ub >= 0.f) num_neg = min(num_neg, (int)(p.neg_pos_ub * (float)max(1, want_pos)));
    const int want_neg = min(nneg_c, max(num_neg, 0));

    // ---- select `want` members of a class by smallest (key, index); emit in index order
    auto select_and_emit = [&](const bool positives, const int have, const int want, const int base) {
        auto member = [&](int c) { return positives ? s_ind[c] > 0 : s_ind[c] == 0; };
        unsigned T = 0u;                                        // want == 0: nothing is below T
        int rem = 0;                                            // members with key == T to take
        const bool all = want >= have;
        if (!all && want > 0) {                                 // 4 x 8-bit radix select of the
            unsigned prefix = 0u;                               // want-th smallest key
            int remaining = want;
            for (int pass = 3; pass >= 0; --pass) {
                const int shift = pass * 8;
                const unsigned hi = pass == 3 ? 0u : (0xffffffffu << (shift + 8));
                for (int i = tid; i < 256; i += kAsThreads) s_hist[i] = 0;
                __syncthreads();
                for (int c = c_lo; c < c_hi; ++c)
                    if (member(c) && (s_key[c] & hi) == (prefix & hi))
                        atomicAdd(&s_hist[(s_key[c] >> shift) & 255u], 1);
                __syncthreads();
                if (tid < 32) {                                 // warp 0: scan 256 bins, 8 per lane
                    int loc[8], sum = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { loc[j] = s_hist[tid * 8 + j]; sum += loc[j]; }
                    int incl = sum;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) {
                        const int t = __shfl_up_sync(0xffffffffu, incl, o);
                        if (tid >= o) incl += t;
                    }
                    int cum = incl - sum;                       // members in lower bins
                    const bool here = cum < remaining && remaining <= incl;
                    if (here) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (remaining <= cum + loc[j]) {
                                s_sel[0] = prefix | ((unsigned)(tid * 8 + j) << shift);
                                s_sel[1] = (unsigned)(remaining - cum);
                                break;
                            }
                            cum += loc[j];
                        }
                    }
                }
                __syncthreads();
                prefix = s_sel[0];
                remaining = (int)s_sel[1];
                __syncthreads();
            }
            T = prefix;
            rem = remaining;
        }
        int my_eq = 0;
        if (!all)
            for (int c = c_lo; c < c_hi; ++c) my_eq += member(c) && s_key[c] == T;
        int dummy;
        int eq_rank = all ? 0 : block_excl_scan(my_eq, s_warp, dummy);
        int my_take = 0;
        unsigned take_bits = 0u;                                // per <= 32 candidates per thread
        for (int c = c_lo; c < c_hi; ++c) {
            bool take = false;
            if (member(c)) {
                if (all || s_key[c] < T) take = true;
                else if (s_key[c] == T) take = eq_rank++ < rem;
            }
            if (take) { take_bits |= 1u << (c - c_lo); ++my_take; }
        }
        int total;
        int pos = base + block_excl_scan(my_take, s_warp, total);
        for (int c = c_lo; c < c_hi; ++c) {
            if (!((take_bits >> (c - c_lo)) & 1u)) continue;
            const size_t r = (size_t)b * p.num + pos++;
            const float4 bx = c < G ? s_gt[c] : props[c - G];
            float* ro = p.rois + r * 5;
            ro[0] = (float)b; ro[1] = bx.x; ro[2] = bx.y; ro[3] = bx.z; ro[4] = bx.w;
            const int ind = s_ind[c];
            p.kind[r] = positives ? 1 : 0;
            p.row_gt_box[r] = positives ? s_gt[ind - 1] : make_float4(0.f, 0.f, 0.f, 0.f);
            p.row_gt_label[r] = positives ? p.gt_labels[(size_t)b * G + ind - 1] : 0ll;
            p.row_gt_index[r] = positives ? ind - 1 : -1;
            p.row_is_gt[r] = c < G ? 1 : 0;
            p.row_cand[r] = c < G ? c : (p.add_gt ? g : 0) + (c - G);     // index in cat([gt, props])
        }
        return total;
    };
    const int npos_sel = select_and_emit(true, npos_c, want_pos, 0);
    const int nneg_sel = select_and_emit(false, nneg_c, want_neg, npos_sel);
    for (int i = npos_sel + nneg_sel + tid; i < p.num; i += kAsThreads) {   // pad rows
        const size_t r = (size_t)b * p.num + i;
        float* ro = p.rois + r * 5;
        ro[0] = (float)b; ro[1] = 0.f; ro[2] = 0.f; ro[3] = 0.f; ro[4] = 0.f;
        p.kind[r] = 2;
        p.row_gt_box[r] = make_float4(0.f, 0.f, 0.f, 0.f);
        p.row_gt_label[r] = 0ll;
        p.row_gt_index[r] = -1;
        p.row_is_gt[r] = 0;
        p.row_cand[r] = -1;
    }
    if (tid == 0) {
        int* cn = p.counts + (size_t)b * 4;
        cn[0] = npos_sel; cn[1] = nneg_sel; cn[2] = npos_c; cn[3] = nneg_c;
    }
}

}  // namespace htd

using namespace htd;

extern "C" {

int htd_assign_sample(const float* props, const unsigned char* valid, int B, int N,
                      const float* gt_boxes, const long long* gt_labels, const int32_t* num_gt,
                      int G, const float* keys, float pos_iou_thr, float neg_iou_thr,
                      float min_pos_iou, int match_low_quality, int add_gt_as_proposals, int num,
                      int num_pos, float neg_pos_ub, float* rois, unsigned char* kind,
                      float* row_gt_boxes, long long* row_gt_labels, unsigned char* row_is_gt,
                      int32_t* row_cand, int32_t* row_gt_index, int32_t* counts, int32_t* gt_inds,
                      float* max_overlaps, htd_stream_t stream) {
    HTD_CHECK_ARG(B >= 0 && N >= 0 && G >= 0 && G <= HTD_MAX_GT && num >= 1 && num_pos >= 0 &&
                      num_pos <= num && G + N <= HTD_MAX_CANDIDATES,
                  "htd_assign_sample: bad sizes B=%d N=%d G=%d num=%d num_pos=%d (G <= %d, G+N <= %d)",
                  B, N, G, num, num_pos, HTD_MAX_GT, HTD_MAX_CANDIDATES);
    if (B == 0) return HTD_OK;
    HTD_CHECK_ARG((N == 0 || props) && (G == 0 || (gt_boxes && gt_labels)) && num_gt && keys && rois &&
                      kind && row_gt_boxes && row_gt_labels && row_is_gt && row_cand &&
                      row_gt_index && counts,
                  "htd_assign_sample: null pointer");
    AsParams p;
    p.props = reinterpret_cast<const float4*>(props);
    p.valid = valid;
    p.gt_boxes = reinterpret_cast<const float4*>(gt_boxes);
    p.gt_labels = gt_labels;
    p.num_gt = num_gt;
    p.keys = keys;
    p.B = B; p.N = N; p.G = G;
    p.pos_thr = pos_iou_thr; p.neg_thr = neg_iou_thr; p.min_pos_iou = min_pos_iou;
    p.neg_pos_ub = neg_pos_ub;
    p.match_low_quality = match_low_quality; p.add_gt = add_gt_as_proposals;
    p.num = num; p.num_pos = num_pos;
    p.rois = rois; p.kind = kind;
    p.row_gt_box = reinterpret_cast<float4*>(row_gt_boxes);
    p.row_gt_label = row_gt_labels;
    p.row_is_gt = row_is_gt; p.row_cand = row_cand; p.row_gt_index = row_gt_index;
    p.counts = counts; p.gt_inds = gt_inds; p.max_ov = max_overlaps;
    const size_t smem = (size_t)G * 16 + (size_t)(G + N) * 8 + (size_t)G * 4;
    HTD_SMEM_OPTIN(assign_sample_kernel, HTD_MAX_GT * 20 + HTD_MAX_CANDIDATES * 8,
                   "htd_assign_sample");
    assign_sample_kernel<<<B, kAsThreads, smem, (cudaStream_t)stream>>>(p);
    HTD_CHECK_LAUNCH("htd_assign_sample");
    return HTD_OK;
}

}  // extern "C"
