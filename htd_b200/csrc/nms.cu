// Multi-class NMS of the test branch - the step right after the RoI head at inference
// (SURVEY.md section 8 row f3): BBoxHead.get_bboxes -> multiclass_nms
// (bbox_heads/bbox_head.py:188-225, core/post_processing/bbox_nms.py:7-71) -> mmcv.ops.batched_nms /
// nms (mmcv-full 1.2.1, un-vendored; published algorithm restated in oracle/restate.py):
//   * candidates = all (RoI k, class c) with score[k,c] > score_thr, in row-major (k, c) order;
//   * boxes are shifted by c * (max coordinate over the candidates + 1) so that one class-agnostic
//     NMS separates the classes (the shift is done in fp32 and changes the IoU bits: kept);
//   * greedy NMS in descending score order, suppress when IoU > iou_thr, IoU = inter / (Sa + Sb - inter);
//   * the survivors in descending score order, first max_num of them.
// The reference does this with masked_select / nonzero / sort / a bit-matrix kernel and a host-side
// reduction.  Here: three small kernels, no host sync, static output shapes, ties in the score
// broken by the candidate index (k * C + c) - the order of a stable sort.
#include "common.cuh"

namespace htd {

constexpr int kNmsThreads = 256;

__device__ __forceinline__ void atomic_max_float(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned*>(addr), __float_as_uint(v));
}

// ws_f[0] = max coordinate, ws_n[c] = candidates of class c, ws_n[C + c] = survivors of class c
__global__ void nms_init_kernel(float* ws_f, int* ws_n, int C) {
    for (int i = threadIdx.x; i < 2 * C; i += blockDim.x) ws_n[i] = 0;
    if (threadIdx.x == 0) ws_f[0] = -INFINITY;
}

// one CTA per class: candidate RoIs in ascending k, global max coordinate
__global__ void __launch_bounds__(kNmsThreads) nms_collect_kernel(
    const float* __restrict__ boxes, int box_classes, const float* __restrict__ scores, int K, int C,
    float score_thr, int* __restrict__ cand_k, int* __restrict__ ws_n, float* __restrict__ ws_f) {
    __shared__ int s_warp[kNmsThreads / 32];
    __shared__ int s_base;
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_base = 0;
    float mx = -INFINITY;
    __syncthreads();
    for (int k0 = 0; k0 < K; k0 += kNmsThreads) {
        const int k = k0 + tid;
        const bool ok = k < K && scores[(size_t)k * (C + 1) + c] > score_thr;
        const unsigned m = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) s_warp[warp] = __popc(m);
        __syncthreads();
        int off = s_base;
        for (int w = 0; w < warp; ++w) off += s_warp[w];
        if (ok) {
            cand_k[(size_t)c * K + off + __popc(m & ((1u << lane) - 1u))] = k;
            const float4 b = *reinterpret_cast<const float4*>(
                boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
            mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
        }
        __syncthreads();
        if (tid == 0) {
            int t = 0;
            for (int w = 0; w < kNmsThreads / 32; ++w) t += s_warp[w];
            s_base += t;
        }
        __syncthreads();
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0 && mx > -INFINITY) atomic_max_float(ws_f, mx);
    if (tid == 0) ws_n[c] = s_base;
}

// mmcv nms IoU (offset 0) on the shifted boxes, plain IEEE fp32 operations in that order
__device__ __forceinline__ float nms_iou(const float4 a, const float4 b) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
    const float inter = __fmul_rn(w, h);
    const float sa = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float sb = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

// IoU(a, b) > thr with the division only for boxes that intersect: disjoint boxes have IoU 0 (or
// 0 / 0 = NaN for two empty boxes), which is never above a threshold >= 0 - same decision, and most
// pairs of RPN candidates are disjoint
__device__ __forceinline__ bool nms_over(const float4 a, const float4 b, float thr) {
    if (thr >= 0.f && (fminf(a.z, b.z) <= fmaxf(a.x, b.x) || fminf(a.w, b.w) <= fmaxf(a.y, b.y)))
        return false;
    return nms_iou(a, b) > thr;
}

// one CTA per class: order by (score desc, k asc) by rank counting, greedy suppression, survivors
// (still in that order) to kept_k / kept_score.  The greedy pass walks the sorted candidates 64 at a
// time: (A) the 64 x 64 overlap bits inside the chunk, all threads; (B) one thread resolves the chunk
// sequentially on those bits - exactly the reference's order of decisions; (C) all threads test
// every later candidate against the chunk's survivors.  Three barriers per 64 candidates instead
// of one per candidate (the RPN hands 2000 candidates per level to this kernel: 2.5 ms -> ~0.1 ms).
__global__ void __launch_bounds__(1024) nms_class_kernel(
    const float* __restrict__ boxes, int box_classes, const float* __restrict__ scores, int K, int C,
    float iou_thr, const int* __restrict__ cand_k, int* __restrict__ ws_n,
    const float* __restrict__ ws_f, int* __restrict__ kept_k, float* __restrict__ kept_score) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int c = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int n = ws_n[c];
    if (n == 0) return;
    float4* s_box = reinterpret_cast<float4*>(smem);                 // [n] sorted, shifted
    float* s_score = reinterpret_cast<float*>(s_box + K);            // [n] sorted
    int* s_k = reinterpret_cast<int*>(s_score + K);                  // [n] sorted
    float* u_score = reinterpret_cast<float*>(s_k + K);              // [n] unsorted
    unsigned char* s_supp = reinterpret_cast<unsigned char*>(u_score + K);
    __shared__ unsigned long long s_m[64];                           // chunk row a: later members it overlaps
    __shared__ unsigned long long s_keep;
    const float shift = __fmul_rn((float)c, __fadd_rn(ws_f[0], 1.f));       // idxs * (max + 1)
    const int* ck = cand_k + (size_t)c * K;
    for (int j = tid; j < n; j += nthr) u_score[j] = scores[(size_t)ck[j] * (C + 1) + c];
    __syncthreads();
    for (int j = tid; j < n; j += nthr) {
        const float sj = u_score[j];
        int rank = 0;
        for (int i = 0; i < n; ++i) {
            const float si = u_score[i];
            rank += (si > sj) || (si == sj && i < j);            // candidates are in ascending k
        }
        const int k = ck[j];
        const float4 b = *reinterpret_cast<const float4*>(
            boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
        s_box[rank] = make_float4(__fadd_rn(b.x, shift), __fadd_rn(b.y, shift),
                                  __fadd_rn(b.z, shift), __fadd_rn(b.w, shift));
        s_score[rank] = sj;
        s_k[rank] = k;
        s_supp[rank] = 0;
    }
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += 64) {
        const int m = min(64, n - i0);
        if (tid < 64) s_m[tid] = 0ull;
        __syncthreads();
        for (int p = tid; p < 64 * 64; p += nthr) {                // (A)
            const int a = p >> 6, b = p & 63;
            if (a < b && b < m && nms_over(s_box[i0 + a], s_box[i0 + b], iou_thr))
                atomicOr(&s_m[a], 1ull << b);
        }
        __syncthreads();
        if (tid == 0) {                                            // (B)
            unsigned long long supp = 0ull, keep = 0ull;
            for (int b = 0; b < m; ++b)
                if (s_supp[i0 + b]) supp |= 1ull << b;
            for (int b = 0; b < m; ++b)
                if (!((supp >> b) & 1ull)) {
                    keep |= 1ull << b;
                    supp |= s_m[b];
                }
            for (int b = 0; b < m; ++b) s_supp[i0 + b] = (unsigned char)((supp >> b) & 1ull) & (unsigned char)!((keep >> b) & 1ull);
            s_keep = keep;
        }
        __syncthreads();
        const unsigned long long keep = s_keep;
        for (int j = i0 + 64 + tid; j < n; j += nthr) {             // (C)
            if (s_supp[j]) continue;
            const float4 bj = s_box[j];
            unsigned long long kk = keep;
            while (kk) {
                const int b = __ffsll((long long)kk) - 1;
                kk &= kk - 1ull;
                if (nms_over(s_box[i0 + b], bj, iou_thr)) {
                    s_supp[j] = 1;
                    break;
                }
            }
        }
        __syncthreads();
    }
    // compact the survivors in order (single warp; n is small)
    if (tid < 32) {
        int base = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + tid;
            const bool keep = j < n && !s_supp[j];
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int r = base + __popc(m & ((1u << tid) - 1u));
                kept_k[(size_t)c * K + r] = s_k[j];
                kept_score[(size_t)c * K + r] = s_score[j];
            }
            base += __popc(m);
        }
        if (tid == 0) ws_n[C + c] = base;
    }
}

// ---- few classes, many candidates per class (the RPN: 5 levels x 2000 boxes) ----------------------
// One CTA per class walks its candidates one after the other however it is organised; here the
// pair tests go to the whole GPU instead: (1) nms_sort_kernel orders a class's candidates (rank
// counting) into the workspace, (2) nms_mask_kernel - a grid of 64 x 64 blocks per class - writes
// for every candidate the 64-bit words of the LATER candidates it overlaps, (3) nms_reduce_kernel
// resolves a class sequentially on those words, 64 candidates per trip from shared memory.  Same
// decisions in the same order as the greedy loop (mmcv's CUDA nms is organised the same way, with
// step 3 on the host).
constexpr int kMaskMaxClasses = 8;

__global__ void __launch_bounds__(1024) nms_sort_kernel(
    const float* __restrict__ boxes, int box_classes, const float* __restrict__ scores, int K, int C,
    const int* __restrict__ cand_k, const int* __restrict__ ws_n, const float* __restrict__ ws_f,
    float4* __restrict__ sbox, float* __restrict__ sscore, int* __restrict__ sk) {
    extern __shared__ __align__(16) unsigned char smem[];
    float* u_score = reinterpret_cast<float*>(smem);
    const int c = blockIdx.x, tid = threadIdx.x, nthr = blockDim.x;
    const int n = ws_n[c];
    if (n == 0) return;
    const float shift = __fmul_rn((float)c, __fadd_rn(ws_f[0], 1.f));       // idxs * (max + 1)
    const int* ck = cand_k + (size_t)c * K;
    for (int j = tid; j < n; j += nthr) u_score[j] = scores[(size_t)ck[j] * (C + 1) + c];
    __syncthreads();
    for (int j = tid; j < n; j += nthr) {
        const float sj = u_score[j];
        int rank = 0;
        for (int i = 0; i < n; ++i) {
            const float si = u_score[i];
            rank += (si > sj) || (si == sj && i < j);            // candidates are in ascending k
        }
        const int k = ck[j];
        const float4 b = *reinterpret_cast<const float4*>(
            boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
        sbox[(size_t)c * K + rank] = make_float4(__fadd_rn(b.x, shift), __fadd_rn(b.y, shift),
                                                 __fadd_rn(b.z, shift), __fadd_rn(b.w, shift));
        sscore[(size_t)c * K + rank] = sj;
        sk[(size_t)c * K + rank] = k;
    }
}

// block (bx, by, c): rows by*64 .. +63 against columns bx*64 .. +63 (bx >= by); mask[c][row][bx]
__global__ void __launch_bounds__(64) nms_mask_kernel(const float4* __restrict__ sbox,
                                                      const int* __restrict__ ws_n, int K, int words,
                                                      float iou_thr,
                                                      unsigned long long* __restrict__ mask) {
    const int c = blockIdx.z, bx = blockIdx.x, by = blockIdx.y, t = threadIdx.x;
    const int n = ws_n[c];
    if (bx < by || by * 64 >= n || bx * 64 >= n) return;
    __shared__ float4 s_col[64];
    const float4* b = sbox + (size_t)c * K;
    const int col = bx * 64 + t;
    if (col < n) s_col[t] = b[col];
    __syncthreads();
    const int row = by * 64 + t;
    if (row >= n) return;
    const float4 br = b[row];
    const int ncol = min(64, n - bx * 64);
    unsigned long long w = 0ull;
    for (int j = (bx == by ? t + 1 : 0); j < ncol; ++j)
        if (nms_over(br, s_col[j], iou_thr)) w |= 1ull << j;
    mask[((size_t)c * K + row) * words + bx] = w;
}

__global__ void __launch_bounds__(256) nms_reduce_kernel(const unsigned long long* __restrict__ mask,
                                                         const float* __restrict__ sscore,
                                                         const int* __restrict__ sk, int K, int C,
                                                         int words, int* __restrict__ ws_n,
                                                         int* __restrict__ kept_k,
                                                         float* __restrict__ kept_score) {
    extern __shared__ __align__(16) unsigned char smem[];
    unsigned long long* s_rows = reinterpret_cast<unsigned long long*>(smem);      // [64][words]
    __shared__ unsigned long long s_keep[64];                                      // per chunk (<= 4096 / 64)
    const int c = blockIdx.x, tid = threadIdx.x, lane = tid & 31;
    const int n = ws_n[c];
    if (n == 0) return;
    const int nw = (n + 63) >> 6;
    unsigned long long r0 = 0ull, r1 = 0ull;      // warp 0: removed words lane and lane + 32
    for (int w = 0; w < nw; ++w) {
        const int rows = min(64, n - w * 64);
        // rows of this chunk, words w .. nw-1 (the only ones the mask kernel wrote)
        for (int i = tid; i < rows * (nw - w); i += 256) {
            const int rr = i / (nw - w), ww = w + i - rr * (nw - w);
            s_rows[rr * words + ww] = mask[((size_t)c * K + w * 64 + rr) * words + ww];
        }
        __syncthreads();
        if (tid < 32) {
            unsigned long long cw = __shfl_sync(0xffffffffu, w < 32 ? r0 : r1, w & 31);
            unsigned long long keep = 0ull;
            for (int b = 0; b < rows; ++b) {
                if (!((cw >> b) & 1ull)) {
                    keep |= 1ull << b;
                    cw |= s_rows[b * words + w];
                    if (lane >= w && lane < nw) r0 |= s_rows[b * words + lane];
                    if (lane + 32 >= w && lane + 32 < nw) r1 |= s_rows[b * words + lane + 32];
                }
            }
            if (lane == 0) s_keep[w] = keep;
        }
        __syncthreads();
    }
    // compact the survivors in order
    if (tid < 32) {
        int base = 0;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + tid;
            const bool keep = j < n && ((s_keep[j >> 6] >> (j & 63)) & 1ull);
            const unsigned m = __ballot_sync(0xffffffffu, keep);
            if (keep) {
                const int r = base + __popc(m & ((1u << tid) - 1u));
                kept_k[(size_t)c * K + r] = sk[(size_t)c * K + j];
                kept_score[(size_t)c * K + r] = sscore[(size_t)c * K + j];
            }
            base += __popc(m);
        }
        if (tid == 0) ws_n[C + c] = base;
    }
}

// every survivor finds its global rank by (score desc, k * C + c asc); ranks < max_num are written
__global__ void __launch_bounds__(kNmsThreads) nms_merge_kernel(
    const float* __restrict__ boxes, int box_classes, int K, int C, const int* __restrict__ ws_n,
    const int* __restrict__ kept_k, const float* __restrict__ kept_score, int max_num,
    float* __restrict__ det, long long* __restrict__ labels, int* __restrict__ count) {
    const int c = blockIdx.y;
    const int n = ws_n[C + c];
    if (blockIdx.x == 0 && c == 0 && threadIdx.x == 0) {
        int t = 0;
        for (int i = 0; i < C; ++i) t += ws_n[C + i];
        count[0] = min(t, max_num);
    }
    for (int r = blockIdx.x * kNmsThreads + threadIdx.x; r < n; r += gridDim.x * kNmsThreads) {
        const float s = kept_score[(size_t)c * K + r];
        const int k = kept_k[(size_t)c * K + r];
        const long long flat = (long long)k * C + c;
        int rank = 0;
        for (int c2 = 0; c2 < C && rank < max_num; ++c2) {
            const int n2 = ws_n[C + c2];
            const float* ks = kept_score + (size_t)c2 * K;
            const int* kk = kept_k + (size_t)c2 * K;
            // a class's survivors are sorted by (score desc, k asc): the number of them that come
            // before this one is a lower bound in that order
            int lo = 0, hi = n2;
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                const float s2 = ks[mid];
                if (s2 > s || (s2 == s && (long long)kk[mid] * C + c2 < flat)) lo = mid + 1;
                else hi = mid;
            }
            rank += lo;
        }
        if (rank < max_num) {
            const float4 b = *reinterpret_cast<const float4*>(
                boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
            float* o = det + (size_t)rank * 5;
            o[0] = b.x; o[1] = b.y; o[2] = b.z; o[3] = b.w; o[4] = s;
            labels[rank] = c;
        }
    }
}


// ------------------------------------------------------------------------------------------
// Soft-NMS variant (configs/htd/htd_resnet101_2x.py:298: nms=dict(type='soft_nms', iou_thr=0.5,
// min_score=0.05); dispatch mmdet/core/post_processing/bbox_nms.py:61 -> mmcv.ops.batched_nms ->
// mmcv.ops.soft_nms, mmcv-full 1.2.1, un-vendored: published algorithm restated literally in
// oracle/soft_nms_ref.c).  The reference loop is sequential: pick the FIRST maximum of the current
// array, swap it to the front, decay every later score by its overlap with the pick, and drop a box
// that falls below min_score by overwriting it with the last box.  Which of two equal scores is
// picked first depends on that swap / overwrite history, so the whole process is reproduced as it
// is - by ONE CTA whose 1024 threads do each of its inner passes in parallel:
//   argmax   (score, lowest position) reduction over the live range;
//   decay    all later boxes at once (plain IEEE fp32 in the reference's operation order);
//   removal  the reference's "overwrite with the last box" scan equals an unstable compaction:
//            with m survivors behind the pick, the k-th dead slot among the first m (ascending) is
//            filled by the k-th live box counted from the end - block scan + one gather.
// Picks come out in descending score order, so the loop stops after max_num picks (the reference
// computes all of them and then slices).  No host sync, one launch.
// ------------------------------------------------------------------------------------------
constexpr int kSoftThreads = 1024;

struct SoftNmsWs {
    float4* box;      // shifted boxes (boxes_for_nms), current order
    float* score;
    float* area;
    int* id;          // k * C + c of the candidate
    int* tmp;         // filler positions
    int* roi_off;     // [K + 1] exclusive scan of candidates per RoI
};

__device__ __forceinline__ int block_exclusive_scan(int v, int* s_warp, int& total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();                                  // s_warp reuse
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        int w = s_warp[lane];
        int wi = w;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, wi, o);
            if (lane >= o) wi += t;
        }
        s_warp[lane] = wi - w;                        // exclusive warp offsets
        if (lane == 31) s_warp[32] = wi;
    }
    __syncthreads();
    total = s_warp[32];
    return s_warp[warp] + incl - v;
}

__global__ void __launch_bounds__(kSoftThreads) soft_nms_kernel(
    const float* __restrict__ boxes, int box_classes, const float* __restrict__ scores, int K, int C,
    float score_thr, float iou_thr, float min_score, int method, int max_num, SoftNmsWs ws,
    float* __restrict__ det, long long* __restrict__ labels, int* __restrict__ count) {
    __shared__ int s_warp[33];
    __shared__ float s_red_f[32];
    __shared__ int s_red_i[32];
    __shared__ float4 s_pick;
    __shared__ float s_pick_area;
    __shared__ int s_any;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int C1 = C + 1;

    // ---- candidates in row-major (k, c) order + the largest coordinate among them ------------
    const int per = (K + kSoftThreads - 1) / kSoftThreads;       // contiguous RoIs per thread
    const int k_lo = min(tid * per, K), k_hi = min(k_lo + per, K);
    int mine = 0;
    float mx = -INFINITY;
    for (int k = k_lo; k < k_hi; ++k) {
        int cnt = 0;
        for (int c = 0; c < C; ++c)
            if (scores[(size_t)k * C1 + c] > score_thr) {
                ++cnt;
                const float4 b = *reinterpret_cast<const float4*>(
                    boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
                mx = fmaxf(mx, fmaxf(fmaxf(b.x, b.y), fmaxf(b.z, b.w)));
            }
        mine += cnt;
    }
    int n = 0;
    int off = block_exclusive_scan(mine, s_warp, n);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if (lane == 0) s_red_f[warp] = mx;
    __syncthreads();
    mx = s_red_f[0];
    for (int w = 1; w < kSoftThreads / 32; ++w) mx = fmaxf(mx, s_red_f[w]);
    const float shift1 = __fadd_rn(mx, 1.f);                      // max_coordinate + 1
    for (int k = k_lo; k < k_hi; ++k)
        for (int c = 0; c < C; ++c) {
            const float sc = scores[(size_t)k * C1 + c];
            if (sc > score_thr) {
                const float4 b = *reinterpret_cast<const float4*>(
                    boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
                const float sh = __fmul_rn((float)c, shift1);     // idxs.to(boxes) * (max + 1)
                const float4 q = make_float4(__fadd_rn(b.x, sh), __fadd_rn(b.y, sh),
                                             __fadd_rn(b.z, sh), __fadd_rn(b.w, sh));
                ws.box[off] = q;
                ws.score[off] = sc;
                ws.area[off] = __fmul_rn(__fsub_rn(q.z, q.x), __fsub_rn(q.w, q.y));
                ws.id[off] = k * C + c;
                ++off;
            }
        }
    __syncthreads();

    // ---- the sequential selection loop ----------------------------------------------------------
    int i = 0;
    for (; i < n && i < max_num; ++i) {
        // first maximum of score[i..n)
        float best = -INFINITY;
        int bpos = 0x7fffffff;
        for (int p = i + tid; p < n; p += kSoftThreads) {
            const float v = ws.score[p];
            if (bpos == 0x7fffffff || v > best) { best = v; bpos = p; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best, o);
            const int op = __shfl_xor_sync(0xffffffffu, bpos, o);
            if (op != 0x7fffffff && (bpos == 0x7fffffff || ov > best || (ov == best && op < bpos))) {
                best = ov;
                bpos = op;
            }
        }
        if (lane == 0) { s_red_f[warp] = best; s_red_i[warp] = bpos; }
        __syncthreads();
        if (tid == 0) {
            for (int w = 1; w < kSoftThreads / 32; ++w) {
                const float ov = s_red_f[w];
                const int op = s_red_i[w];
                if (op != 0x7fffffff && (bpos == 0x7fffffff || ov > best || (ov == best && op < bpos))) {
                    best = ov;
                    bpos = op;
                }
            }
            // swap the pick to position i, emit it
            const float4 pb = ws.box[bpos];
            const float pa = ws.area[bpos];
            const int pid = ws.id[bpos];
            ws.box[bpos] = ws.box[i];
            ws.score[bpos] = ws.score[i];
            ws.area[bpos] = ws.area[i];
            ws.id[bpos] = ws.id[i];
            ws.box[i] = pb;
            ws.score[i] = best;
            ws.area[i] = pa;
            ws.id[i] = pid;
            s_pick = pb;
            s_pick_area = pa;
            s_any = 0;
            const int k = pid / C, c = pid - k * C;
            const float4 ob = *reinterpret_cast<const float4*>(
                boxes + ((size_t)k * box_classes + (box_classes > 1 ? c : 0)) * 4);
            float* o = det + (size_t)i * 5;
            o[0] = ob.x; o[1] = ob.y; o[2] = ob.z; o[3] = ob.w; o[4] = best;
            labels[i] = c;
        }
        __syncthreads();
        // decay the later boxes; each thread owns a contiguous segment (same split as the scan)
        const int rem = n - i - 1;
        const int seg = (rem + kSoftThreads - 1) / kSoftThreads;
        const int r_lo = min(tid * seg, rem), r_hi = min(r_lo + seg, rem);
        const float4 a = s_pick;
        const float iarea = s_pick_area;
        int live = 0;
        for (int r = r_lo; r < r_hi; ++r) {
            const int p = i + 1 + r;
            const float4 b = ws.box[p];
            const float w = fmaxf(0.f, __fsub_rn(fminf(a.z, b.z), fmaxf(a.x, b.x)));
            const float h = fmaxf(0.f, __fsub_rn(fminf(a.w, b.w), fmaxf(a.y, b.y)));
            const float inter = __fmul_rn(w, h);
            const float ovr = __fdiv_rn(inter, __fsub_rn(__fadd_rn(iarea, ws.area[p]), inter));
            float weight = 1.f;
            if (ovr >= iou_thr) weight = method == 0 ? 0.f : __fsub_rn(1.f, ovr);
            const float ns = __fmul_rn(ws.score[p], weight);
            ws.score[p] = ns;
            live += !(ns < min_score);
        }
        if (live != r_hi - r_lo) s_any = 1;                      // benign race: all write 1
        __syncthreads();
        if (!s_any) continue;                                    // uniform
        // ---- removals: unstable compaction of [i+1, n) ------------------------------------------
        int m = 0;
        const int before = block_exclusive_scan(live, s_warp, m);
        int lb = before;                                         // live boxes before r
        for (int r = r_lo; r < r_hi; ++r) {
            const bool lv = !(ws.score[i + 1 + r] < min_score);
            if (lv && r >= m) ws.tmp[m - lb - 1] = i + 1 + r;    // rank from the end = live after r
            lb += lv;
        }
        __syncthreads();
        lb = before;
        for (int r = r_lo; r < r_hi && r < m; ++r) {
            const int p = i + 1 + r;
            const bool lv = !(ws.score[p] < min_score);
            if (!lv) {
                const int src = ws.tmp[r - lb];                  // dead slots before r = r - lb
                ws.box[p] = ws.box[src];
                ws.score[p] = ws.score[src];
                ws.area[p] = ws.area[src];
                ws.id[p] = ws.id[src];
            }
            lb += lv;
        }
        n = i + 1 + m;
        __syncthreads();
    }
    if (tid == 0) count[0] = i;
}

}  // namespace htd

using namespace htd;

extern "C" {

static bool nms_mask_path(int K, int C) { return C <= kMaskMaxClasses && K > 512; }

long long htd_multiclass_nms_workspace_bytes(int K, int C) {
    // cand_k, kept_k (int), kept_score (float): [C, K] each; counters [2C] ints; 1 float (16 B slot)
    long long b = (long long)C * K * 12 + (long long)C * 8 + 16;
    b = (b + 15) / 16 * 16;
    if (nms_mask_path(K, C))   // sorted boxes / scores / k + the overlap words [C, K, ceil(K / 64)]
        b += (long long)C * K * 24 + (long long)C * K * ((K + 63) / 64) * 8;
    return b;
}

int htd_multiclass_nms(const float* boxes, int box_classes, const float* scores, int K, int C,
                       float score_thr, float iou_thr, int max_num, float* det, long long* labels,
                       int32_t* count, void* workspace, htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && K <= HTD_NMS_MAX_ROIS && C >= 1 && C <= HTD_NMS_MAX_CLASSES &&
                      (box_classes == 1 || box_classes == C) && max_num >= 1 && score_thr >= 0.f,
                  "htd_multiclass_nms: bad arguments K=%d (<= %d) C=%d (<= %d) box_classes=%d "
                  "max_num=%d score_thr=%g (>= 0)", K, HTD_NMS_MAX_ROIS, C, HTD_NMS_MAX_CLASSES,
                  box_classes, max_num, (double)score_thr);
    HTD_CHECK_ARG(det && labels && count && workspace && (K == 0 || (boxes && scores)),
                  "htd_multiclass_nms: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    char* w = static_cast<char*>(workspace);
    float* ws_f = reinterpret_cast<float*>(w);
    int* ws_n = reinterpret_cast<int*>(w + 16);
    int* cand_k = ws_n + 2 * C;
    int* kept_k = cand_k + (size_t)C * K;
    float* kept_score = reinterpret_cast<float*>(kept_k + (size_t)C * K);
    nms_init_kernel<<<1, 256, 0, st>>>(ws_f, ws_n, C);
    HTD_CHECK_LAUNCH("htd_multiclass_nms(init)");
    if (K > 0) {
        nms_collect_kernel<<<C, kNmsThreads, 0, st>>>(boxes, box_classes, scores, K, C, score_thr,
                                                      cand_k, ws_n, ws_f);
        HTD_CHECK_LAUNCH("htd_multiclass_nms(collect)");
        if (nms_mask_path(K, C)) {
            const int words = (K + 63) / 64;
            size_t off = ((size_t)C * K * 12 + (size_t)C * 8 + 16 + 15) / 16 * 16;
            float4* sbox = reinterpret_cast<float4*>(w + off);
            float* sscore = reinterpret_cast<float*>(sbox + (size_t)C * K);
            int* sk = reinterpret_cast<int*>(sscore + (size_t)C * K);
            unsigned long long* mask = reinterpret_cast<unsigned long long*>(sk + (size_t)C * K);
            HTD_SMEM_OPTIN(nms_sort_kernel, HTD_NMS_MAX_ROIS * 4, "htd_multiclass_nms");
            nms_sort_kernel<<<C, 1024, (size_t)K * 4, st>>>(boxes, box_classes, scores, K, C, cand_k, ws_n,
                                                            ws_f, sbox, sscore, sk);
            HTD_CHECK_LAUNCH("htd_multiclass_nms(sort)");
            nms_mask_kernel<<<dim3(words, words, C), 64, 0, st>>>(sbox, ws_n, K, words, iou_thr, mask);
            HTD_CHECK_LAUNCH("htd_multiclass_nms(mask)");
            HTD_SMEM_OPTIN(nms_reduce_kernel, 64 * 64 * 8, "htd_multiclass_nms");
            nms_reduce_kernel<<<C, 256, (size_t)64 * words * 8, st>>>(mask, sscore, sk, K, C, words, ws_n,
                                                                     kept_k, kept_score);
            HTD_CHECK_LAUNCH("htd_multiclass_nms(reduce)");
        } else {
        const size_t smem = (size_t)K * (16 + 4 + 4 + 4 + 1);
        HTD_SMEM_OPTIN(nms_class_kernel, HTD_NMS_MAX_ROIS * 29, "htd_multiclass_nms");
        nms_class_kernel<<<C, K > 256 ? 1024 : kNmsThreads, smem, st>>>(boxes, box_classes, scores, K, C,
                                                                        iou_thr, cand_k, ws_n, ws_f,
                                                                        kept_k, kept_score);
        HTD_CHECK_LAUNCH("htd_multiclass_nms(class)");
        }
    }
    const int bx = K > 0 ? (K + kNmsThreads - 1) / kNmsThreads : 1;
    nms_merge_kernel<<<dim3(bx, C), kNmsThreads, 0, st>>>(boxes, box_classes, K, C, ws_n, kept_k,
                                                          kept_score, max_num, det, labels, count);
    HTD_CHECK_LAUNCH("htd_multiclass_nms(merge)");
    return HTD_OK;
}


long long htd_multiclass_soft_nms_workspace_bytes(int K, int C) {
    // box (16) + score + area + id + tmp (4 each) per candidate slot, K*C slots; roi_off unused tail
    return (long long)K * C * 32 + 64;
}

int htd_multiclass_soft_nms(const float* boxes, int box_classes, const float* scores, int K, int C,
                            float score_thr, float iou_thr, float min_score, int method,
                            int max_num, float* det, long long* labels, int32_t* count,
                            void* workspace, htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && K <= HTD_NMS_MAX_ROIS && C >= 1 && C <= HTD_NMS_MAX_CLASSES &&
                      (box_classes == 1 || box_classes == C) && max_num >= 1 && score_thr >= 0.f,
                  "htd_multiclass_soft_nms: bad arguments K=%d (<= %d) C=%d (<= %d) box_classes=%d "
                  "max_num=%d score_thr=%g (>= 0)", K, HTD_NMS_MAX_ROIS, C, HTD_NMS_MAX_CLASSES,
                  box_classes, max_num, (double)score_thr);
    HTD_CHECK_ARG(method == 0 || method == 1,
                  "htd_multiclass_soft_nms: method must be 0 (naive) or 1 (linear, the mmcv "
                  "default the HTD configs use); gaussian is not provided");
    HTD_CHECK_ARG(det && labels && count && workspace && (K == 0 || (boxes && scores)),
                  "htd_multiclass_soft_nms: null pointer");
    HTD_CHECK_ARG(((uintptr_t)workspace & 15) == 0, "htd_multiclass_soft_nms: workspace must be 16-byte aligned");
    const size_t slots = (size_t)K * C;
    char* w = static_cast<char*>(workspace);
    SoftNmsWs ws;
    ws.box = reinterpret_cast<float4*>(w);
    ws.score = reinterpret_cast<float*>(w + slots * 16);
    ws.area = ws.score + slots;
    ws.id = reinterpret_cast<int*>(ws.area + slots);
    ws.tmp = ws.id + slots;
    ws.roi_off = nullptr;
    soft_nms_kernel<<<1, kSoftThreads, 0, (cudaStream_t)stream>>>(
        boxes, box_classes, scores, K, C, score_thr, iou_thr, min_score, method, max_num, ws, det,
        labels, count);
    HTD_CHECK_LAUNCH("htd_multiclass_soft_nms");
    return HTD_OK;
}

}  // extern "C"
