// Multi-level RoIAlign for the HTD RoI head on B200 (sm_100a): FPN level assignment, footprint
// planning, separable gather forward and an atomic-free pixel-tile gather backward.
//
// Reference path replaced: SingleRoIExtractor.forward (single_level_roi_extractor.py:53-99) and the
// five RoIAlign calls of AdptRoIExtractor.forward (adaptative_roi_extractor.py:71-74,87), i.e.
// mmcv's roi_align_forward/backward (one thread per output element, 4*gh*gw scalar NCHW loads
// per output, global atomicAdd in backward).
//
// Design (DESIGN.md "RoIAlign kernels"):
//  * avg RoIAlign is separable:  Y[ph][pw][c] = sum_r sum_x Wy[ph][r] Wx[pw][x] F[r][x][c];
//    every footprint pixel of a bin is read ONCE (128-bit channels-last loads) instead of up to
//    4*gh*gw times.  Axis weights are computed in fp64 (roi_axis.h) and rounded to fp32.
//  * plan (htd_roi_plan): footprint box of every (level, RoI), an exclusive scan of their
//    extents, and the separable axis-weight TABLES  wy[row][p], wx[col][p]  (fp64 arithmetic,
//    rounded once to fp32) plus the pixel range of every bin.  Built once per extractor call and
//    shared by forward and backward, so the hot kernels contain no coordinate math at all.
//  * forward: one CTA per (RoI, level, output row ph), one warp per output bin; a lane owns 8 of
//    the 256 channels of a chunk; weights come from the tables via warp shuffles.  A 200x200-pixel
//    BA footprint on P2 is spread over 7 CTAs x 7 warps.
//  * backward: one CTA per 8x8-pixel tile of dX; it scans the RoI footprint boxes of its image,
//    compacts the intersecting RoIs IN INDEX ORDER (ballot + prefix), and accumulates
//    Wy^T dY Wx for its 64 pixels x 256 channels in registers.  Every dX element is written
//    exactly once: no atomics, no memset, bit-reproducible.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"
#include "roi_axis.h"
#include "tc05.cuh"

namespace htd {

static thread_local char g_err[512] = "";
// Measurement hooks exist only in the -DHTD_DEBUG_HOOKS build (libhtd_b200_hooks.so): kernel-variant
// selection, the backward trace and the HTD_FWD_KERNEL / HTD_BWD_KERNEL environment switches.  The
// product library always runs the default kernels.
#ifdef HTD_DEBUG_HOOKS
static unsigned long long* g_bwd_trace = nullptr;    // see htd_debug_set_bwd_trace
static int g_bwd_variant = -1;                       // see htd_debug_set_bwd_variant
static const char* hook_env(const char* name) { return getenv(name); }
#else
static constexpr unsigned long long* g_bwd_trace = nullptr;
static constexpr int g_bwd_variant = -1;
static const char* hook_env(const char*) { return nullptr; }
#endif

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

struct LevelDev {
    void* data;
    int H, W;
    float scale;
};

// ------------------------------------------------------------------------------------------
// level assignment
// ------------------------------------------------------------------------------------------
__global__ void level_assign_kernel(const float* __restrict__ rois, int K, int num_levels,
                                    float finest, int* __restrict__ levels) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    const float* r = rois + (size_t)k * 5;
    levels[k] = roi_level(r[1], r[2], r[3], r[4], finest, num_levels);
}

// ------------------------------------------------------------------------------------------
// footprints
// ------------------------------------------------------------------------------------------
struct FootParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int L, B, K, P, sr;
    const float* rois;
    const int* roi_level;
    int4* boxes;
    unsigned long long* pixel_count;
};

__global__ void footprint_kernel(const FootParams p) {
    int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= p.L * p.K) return;
    int l = idx / p.K, k = idx % p.K;
    int4 box = make_int4(0, -1, 0, -1);
    const float* r = p.rois + (size_t)k * 5;
    int b = (int)r[0];
    bool on = (p.roi_level == nullptr) || (p.roi_level[k] == l);
    if (on && b >= 0 && b < p.B) {
        Axis ay = make_axis(r[2], r[4], (double)p.lv[l].scale, p.P, p.lv[l].H, p.sr, 1);
        Axis ax = make_axis(r[1], r[3], (double)p.lv[l].scale, p.P, p.lv[l].W, p.sr, 1);
        int r0, r1, c0, c1;
        roi_range(ay, p.P, r0, r1);
        roi_range(ax, p.P, c0, c1);
        if (r1 >= r0 && c1 >= c0) {
            box = make_int4(r0, r1, c0, c1);
            if (p.pixel_count)
                atomicAdd(p.pixel_count + l,
                          (unsigned long long)(r1 - r0 + 1) * (unsigned long long)(c1 - c0 + 1));
        }
    }
    p.boxes[idx] = box;
}

// ------------------------------------------------------------------------------------------
// plan: extents scan + axis-weight tables
// ------------------------------------------------------------------------------------------
constexpr int kTabW = HTD_MAX_POOLED;          // floats per table row (weights of the P bins)
constexpr int kRangeInts = 4 * HTD_MAX_POOLED; // per entry: y lo[8], y hi[8], x lo[8], x hi[8]

// offsets[e] = sum_{e' < e} (fh + fw) over all entries e = l*K + k; offsets[n] = total rows.
__global__ void __launch_bounds__(1024) extent_scan_kernel(const int4* __restrict__ boxes, int n,
                                                           int* __restrict__ offsets) {
    __shared__ int s_warp[32];
    __shared__ int s_carry;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) s_carry = 0;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int e = base + tid;
        int v = 0;
        if (e < n) {
            const int4 b = boxes[e];
            if (b.y >= b.x && b.w >= b.z) v = (b.y - b.x + 1) + (b.w - b.z + 1);
        }
        int inc = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(0xffffffffu, inc, o);
            if (lane >= o) inc += u;
        }
        if (lane == 31) s_warp[warp] = inc;
        __syncthreads();
        if (warp == 0) {
            int w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += u;
            }
            s_warp[lane] = w;
        }
        __syncthreads();
        const int carry = s_carry;
        const int excl = carry + (warp > 0 ? s_warp[warp - 1] : 0) + inc - v;
        if (e < n) offsets[e] = excl;
        __syncthreads();
        if (tid == 1023) s_carry = carry + s_warp[31];
        __syncthreads();
    }
    if (tid == 0) offsets[n] = s_carry;
}

struct TableParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int L, K, P, sr;
    const float* rois;
    const int4* boxes;
    const int* offsets;
    int* ranges;
    float* weights;
};

__global__ void __launch_bounds__(128) weight_table_kernel(const TableParams p) {
    __shared__ Axis s_axis[2];
    const int e = blockIdx.x;
    const int l = e / p.K, k = e % p.K;
    const int4 box = p.boxes[e];
    int* rg = p.ranges + (size_t)e * kRangeInts;
    const int tid = threadIdx.x;
    if (box.y < box.x || box.w < box.z) {           // RoI does not touch this level
        if (tid < kRangeInts) rg[tid] = (tid / HTD_MAX_POOLED) % 2 == 0 ? 0 : -1;
        return;
    }
    const float* r = p.rois + (size_t)k * 5;
    if (tid == 0) s_axis[0] = make_axis(r[2], r[4], (double)p.lv[l].scale, p.P, p.lv[l].H, p.sr, 1);
    if (tid == 32) s_axis[1] = make_axis(r[1], r[3], (double)p.lv[l].scale, p.P, p.lv[l].W, p.sr, 1);
    __syncthreads();
    if (tid < 2 * HTD_MAX_POOLED) {                 // per-bin pixel ranges
        const int axis = tid / HTD_MAX_POOLED, pp = tid % HTD_MAX_POOLED;
        int lo = 0, hi = -1;
        if (pp < p.P) bin_range(s_axis[axis], pp, lo, hi);
        rg[axis * 2 * HTD_MAX_POOLED + pp] = lo;
        rg[axis * 2 * HTD_MAX_POOLED + HTD_MAX_POOLED + pp] = hi;
    }
    const int fh = box.y - box.x + 1, fw = box.w - box.z + 1;
    float* tab = p.weights + (size_t)p.offsets[e] * kTabW;
    const int n = (fh + fw) * kTabW;
    for (int base = 0; base < n; base += blockDim.x) {      // uniform trip count: shuffles below
        const int i = base + tid;
        const int row = i / kTabW, pp = i % kTabW;
        float w = 0.f;
        if (i < n && pp < p.P)
            w = row < fh ? axis_weight(s_axis[0], pp, box.x + row)
                         : axis_weight(s_axis[1], pp, box.z + (row - fh));
        // the 8 entries of a table row sit in 8 consecutive lanes: with pooled < 8 the last entry
        // carries the row sum (the rank-1 add-vector term of the tensor-pipe backward reads it)
        float sum = w;
        sum += __shfl_xor_sync(0xffffffffu, sum, 1);
        sum += __shfl_xor_sync(0xffffffffu, sum, 2);
        sum += __shfl_xor_sync(0xffffffffu, sum, 4);
        if (p.P < kTabW && pp == kTabW - 1) w = sum;
        if (i < n) tab[i] = w;
    }
}

// The whole plan in ONE launch (htd_roi_plan when no pixel count is requested): a CTA per entry
// (level, RoI) computes the footprint box, the per-bin pixel ranges and both axis-weight tables.
// Table rows sit at a fixed stride per entry - offsets[e] = K * (rows of the lower levels) +
// k * (H_l + W_l), or k * max_l (H_l + W_l) for a level-assigned extraction - inside the capacity
// htd_roi_plan_rows_bound already reserves, so no prefix scan (and no second and third launch:
// 0.042 ms of 0.12 ms for the single-level extraction at the bench size) is needed.
struct PlanParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int L, B, K, P, sr;
    int base[HTD_MAX_LEVELS];     // table rows before level l's block (all-level mode)
    int ext[HTD_MAX_LEVELS];      // H_l + W_l
    int ext_max;
    const float* rois;
    const int* roi_level;
    int4* boxes;
    int* offsets;
    int* ranges;
    float* weights;
};

__global__ void __launch_bounds__(128) roi_plan_fused_kernel(const PlanParams p) {
    __shared__ Axis s_axis[2];
    __shared__ int s_box[4];
    __shared__ int s_rng[kRangeInts];               // y lo[8], y hi[8], x lo[8], x hi[8]
    const int tid = threadIdx.x;
    // level-assigned extraction: one CTA per RoI - the entries of the levels the RoI is NOT on are
    // empty boxes / ranges, written by the same CTA; all-level extraction: one CTA per (level, RoI)
    const int k = p.roi_level ? blockIdx.x : blockIdx.x % p.K;
    const float* r = p.rois + (size_t)k * 5;
    const int b = (int)r[0];
    int l = p.roi_level ? p.roi_level[k] : blockIdx.x / p.K;
    if (p.roi_level) {
        for (int lo = 0; lo < p.L; ++lo) {
            if (lo == l && b >= 0 && b < p.B) continue;
            const int eo = lo * p.K + k;
            if (tid == 0) {
                p.offsets[eo] = k * p.ext_max;
                p.boxes[eo] = make_int4(0, -1, 0, -1);
            }
            if (tid < kRangeInts)
                p.ranges[(size_t)eo * kRangeInts + tid] = (tid / HTD_MAX_POOLED) % 2 == 0 ? 0 : -1;
        }
        if (k == 0 && tid == 1) p.offsets[p.L * p.K] = p.K * p.ext_max;
        if (!(l >= 0 && l < p.L && b >= 0 && b < p.B)) return;
    }
    const int e = l * p.K + k;
    const bool on = b >= 0 && b < p.B;
    int* rg = p.ranges + (size_t)e * kRangeInts;
    const int off = p.roi_level ? k * p.ext_max : p.K * p.base[l] + k * p.ext[l];
    if (tid == 0) p.offsets[e] = off;
    if (!p.roi_level && e == 0 && tid == 1)
        p.offsets[p.L * p.K] = p.K * (p.base[p.L - 1] + p.ext[p.L - 1]);
    // the serial fp64 chains are what a CTA costs in time: the two axes are set up by two threads,
    // the 2 x P bin ranges by 2 x P threads, and the footprint box is the union of the bin ranges
    // (roi_range on one thread per axis ran the P bin ranges one after the other: 2/3 of the time)
    if (on) {
        if (tid == 0) s_axis[0] = make_axis(r[2], r[4], (double)p.lv[l].scale, p.P, p.lv[l].H, p.sr, 1);
        if (tid == 32) s_axis[1] = make_axis(r[1], r[3], (double)p.lv[l].scale, p.P, p.lv[l].W, p.sr, 1);
    }
    __syncthreads();
    if (tid < kRangeInts) {                         // per-bin pixel ranges (one warp: 2 x P lanes busy)
        const int axis = tid / (2 * HTD_MAX_POOLED), hi_half = (tid / HTD_MAX_POOLED) & 1;
        const int pp = tid % HTD_MAX_POOLED;
        int lo = 0, hi = -1;
        if (on && !hi_half && pp < p.P) bin_range(s_axis[axis], pp, lo, hi);
        // lane pp computed both ends; the "hi" lane (pp + 8) takes the second one from it
        const int hi_from = __shfl_sync(0xffffffffu, hi, (tid & 16) + pp);
        s_rng[tid] = hi_half ? hi_from : lo;
    }
    __syncthreads();
    if (tid < 2) {                                  // union of the non-empty bin ranges of an axis
        const int* lo = s_rng + tid * 2 * HTD_MAX_POOLED;
        const int* hi = lo + HTD_MAX_POOLED;
        int jlo = 0, jhi = -1;
        bool any = false;
        for (int pp = 0; pp < p.P; ++pp) {
            if (hi[pp] < lo[pp]) continue;
            if (!any) { jlo = lo[pp]; jhi = hi[pp]; any = true; }
            else { jlo = min(jlo, lo[pp]); jhi = max(jhi, hi[pp]); }
        }
        s_box[2 * tid] = jlo;
        s_box[2 * tid + 1] = jhi;
    }
    __syncthreads();
    int4 box = make_int4(0, -1, 0, -1);
    if (on && s_box[1] >= s_box[0] && s_box[3] >= s_box[2])
        box = make_int4(s_box[0], s_box[1], s_box[2], s_box[3]);
    if (tid == 0) p.boxes[e] = box;
    if (box.y < box.x || box.w < box.z) {           // RoI does not touch this level
        if (tid < kRangeInts) rg[tid] = (tid / HTD_MAX_POOLED) % 2 == 0 ? 0 : -1;
        return;
    }
    if (tid < kRangeInts) rg[tid] = s_rng[tid];
    // Axis-weight tables, 64 rows at a time.  A row (pixel) has non-zero weight only for the bins
    // whose pixel range contains it - usually one or two of the P - and axis_weight (an fp64 sample
    // search, ~600 instructions) is what this kernel costs.  A thread PAIR owns a row: the threads
    // first pick their bins (pair member 0 the 1st, 3rd, .. covering bin, member 1 the 2nd, 4th, ..),
    // then every lane of the warp makes ONE converged axis_weight call; looping over the bins with
    // the call inside ran it once per bin touched by any of the warp's rows.
    __shared__ __align__(16) float s_tab[64][kTabW];
    const int fh = box.y - box.x + 1, fw = box.w - box.z + 1;
    const int rows = fh + fw;
    float* tab = p.weights + (size_t)off * kTabW;
    const int rl = tid >> 1, slot = tid & 1;
    for (int base = 0; base < rows; base += 64) {
        for (int i = tid; i < 64 * kTabW; i += 128) (&s_tab[0][0])[i] = 0.f;
        __syncthreads();
        const int row = base + rl;
        if (row < rows) {
            const int axis = row < fh ? 0 : 1;
            const int j = axis == 0 ? box.x + row : box.z + (row - fh);
            const int* lo = s_rng + axis * 2 * HTD_MAX_POOLED;
            const int* hi = lo + HTD_MAX_POOLED;
            int mine = -1, more = 0, seen = 0;
            for (int pp = 0; pp < p.P; ++pp)
                if (j >= lo[pp] && j <= hi[pp]) {
                    if ((seen & 1) == slot) {
                        if (mine < 0) mine = pp;
                        else more = 1;
                    }
                    ++seen;
                }
            if (mine >= 0) s_tab[rl][mine] = axis_weight(s_axis[axis], mine, j);
            if (more) {                               // bins narrower than a pixel: rare
                seen = 0;
                for (int pp = 0; pp < p.P; ++pp)
                    if (j >= lo[pp] && j <= hi[pp]) {
                        if ((seen & 1) == slot && pp != mine)
                            s_tab[rl][pp] = axis_weight(s_axis[axis], pp, j);
                        ++seen;
                    }
            }
        }
        __syncthreads();
        if (row < rows) {
            const float* t = s_tab[rl];
            float4 v = *reinterpret_cast<const float4*>(t + slot * 4);
            if (slot == 1 && p.P < kTabW)             // entry 7: the row sum (entries P .. 6 are 0)
                v.w = ((t[6] + t[7]) + (t[4] + t[5])) + ((t[2] + t[3]) + (t[0] + t[1]));
            *reinterpret_cast<float4*>(tab + (size_t)row * kTabW + slot * 4) = v;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct FwdParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int L, B, C, K, P;
    const float* rois;
    const int* roi_level;
    const int4* boxes;
    const int* offsets;
    const int* ranges;
    const float* weights;
    const float* bias;
    void* out;
    int region;     // bytes of the staging region in dynamic shared memory (barriers follow it)
    int strip;      // 1: stage the whole bin-row strip of a small RoI at once (see the kernel)
};

// Shared-memory staged forward.  One CTA per (RoI, level, output row ph), one warp per output
// bin (ph, pw).  Every warp runs its OWN copy pipeline: lane 0 streams the bin's footprint -
// each row segment is one contiguous run in channels-last memory - into a private
// kFwdStages-deep ring with cp.async.bulk (UBLKCP) + mbarrier transaction counts, kFwdPx pixels
// per stage, while the warp reduces the previous stage from shared memory.  No cross-warp
// barrier, no idle spinning; bytes in flight are set by the rings, not by registers.
#ifndef HTD_FWD_STAGES
#define HTD_FWD_STAGES 2
#endif
#ifndef HTD_FWD_PX
#define HTD_FWD_PX 8
#endif
constexpr int kFwdStages = HTD_FWD_STAGES;
constexpr int kFwdPx = HTD_FWD_PX;
constexpr int kFwdStripBytes = 72 * 1024;      // strip-path staging budget (>= the bf16 rings)
inline int fwd_smem_max(int elem) {
    int region = HTD_MAX_POOLED * kFwdStages * kFwdPx * 256 * elem;
    if (region < kFwdStripBytes) region = kFwdStripBytes;
    return region + HTD_MAX_POOLED * kFwdStages * 8;
}

template <typename TIn>
__device__ __forceinline__ void ld_smem8(const TIn* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld_smem8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld_smem8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(HTD_MAX_POOLED * 32) roi_align_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(128) uint8_t fwd_smem[];
    const int pw = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int P = p.P;
    const int ph = (int)(blockIdx.x % P);
    const long long task = blockIdx.x / P;
    int l, k;
    if (p.roi_level) { k = (int)task; l = p.roi_level[k]; }
    else { l = (int)(task / p.K); k = (int)(task % p.K); }
    const int b = (int)p.rois[(size_t)k * 5];
    const bool bvalid = (b >= 0 && b < p.B);
    const bool valid = bvalid && (l >= 0 && l < p.L);
    const int PP = P * P;
    const int cw = min(p.C, 256);                       // channels staged per pixel
    const int stage_elems = kFwdPx * cw;
    TIn* ring = reinterpret_cast<TIn*>(fwd_smem) + (size_t)pw * kFwdStages * stage_elems;
    uint64_t* bars = reinterpret_cast<uint64_t*>(fwd_smem + p.region);
    uint64_t* full_bar = bars + pw * kFwdStages;

    int H = 1, W = 1, ry0 = 0, ry1 = -1, cx0 = 0, cx1 = -1, bz = 0, bw = -1;
    const TIn* feat = nullptr;
    const float* wyt = nullptr;
    const float* wxt = nullptr;
    if (valid) {
        const size_t e = (size_t)l * p.K + k;
        const int4 box = p.boxes[e];
        if (box.y >= box.x && box.w >= box.z) {
            const int* rg = p.ranges + e * kRangeInts;
            ry0 = rg[ph]; ry1 = rg[HTD_MAX_POOLED + ph];
            cx0 = rg[2 * HTD_MAX_POOLED + pw]; cx1 = rg[3 * HTD_MAX_POOLED + pw];
            const size_t off = (size_t)p.offsets[e];
            const int fh = box.y - box.x + 1;
            wyt = p.weights + (off + (ry0 - box.x)) * kTabW + ph;
            wxt = p.weights + (off + fh + (cx0 - box.z)) * kTabW + pw;
            H = p.lv[l].H; W = p.lv[l].W;
            feat = static_cast<const TIn*>(p.lv[l].data);
            bz = box.z; bw = box.w;
        }
    }
    const int ny = ry1 - ry0 + 1, nx = cx1 - cx0 + 1;
    TOut* orow = static_cast<TOut*>(p.out) + ((size_t)task * PP + ph * P + pw) * p.C;

    // ---- strip path (small RoIs: every single-level extraction, the coarse levels of BA): the
    // footprint rows of this bin row over the RoI's full width fit the staging region, so they are
    // fetched with ONE bulk copy per row, all in flight at once (one memory latency per CTA instead
    // of one per ring turn, and pixels shared by neighbouring bins are read once); the seven warps
    // then reduce their bins from the shared strip in the same order as the ring path (same bits).
    const int nxs = bw - bz + 1;
    if (p.strip && p.C <= 256 && ny > 0 && nxs > 0 &&
        (size_t)ny * nxs * p.C * sizeof(TIn) <= (size_t)p.region) {           // CTA-uniform
        TIn* sbuf = reinterpret_cast<TIn*>(fwd_smem);
        const uint32_t row_bytes = (uint32_t)(nxs * p.C * sizeof(TIn));
        if (threadIdx.x == 0) {
            mbar_init(bars, 1);
            fence_mbar_init();
        }
        __syncthreads();
        if (pw == 0) {
            if (lane == 0) mbar_expect_tx(bars, (uint32_t)ny * row_bytes);
            __syncwarp();
            for (int r = lane; r < ny; r += 32)
                bulk_g2s(sbuf + (size_t)r * nxs * p.C,
                         feat + (((size_t)b * H + ry0 + r) * W + bz) * p.C, row_bytes, bars);
        }
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const bool lane_on = lane * 8 < p.C;
        mbar_wait(bars, 0u);                              // every warp: no exit with copies in flight
        if (nx > 0) {
            float wy_l = 0.f, wx_l = 0.f;
            for (int y = 0; y < ny; ++y) {
                if ((y & 31) == 0) wy_l = (y + lane < ny) ? __ldg(wyt + (size_t)(y + lane) * kTabW) : 0.f;
                const float wyv = __shfl_sync(0xffffffffu, wy_l, y & 31);
                if (wyv == 0.f) continue;
                const TIn* src = sbuf + ((size_t)y * nxs + (cx0 - bz)) * p.C + lane * 8;
                for (int x = 0; x < nx; ++x) {
                    if ((x & 31) == 0) wx_l = (x + lane < nx) ? __ldg(wxt + (size_t)(x + lane) * kTabW) : 0.f;
                    const float w = wyv * __shfl_sync(0xffffffffu, wx_l, x & 31);
                    if (lane_on) {
                        float v[8];
                        ld_smem8<TIn>(src + (size_t)x * p.C, v);
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, v[e], acc[e]);
                    }
                }
            }
        }
        if (p.bias && bvalid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = lane * 8 + e;
                if (c < p.C) acc[e] += __ldg(p.bias + (size_t)b * p.C + c);
            }
        }
        Vec8<TOut, false>::store(orow, lane, p.C, acc);
        return;
    }

    const int nseg = (nx + kFwdPx - 1) / kFwdPx;        // segments per footprint row
    const int total = (ny > 0 && nx > 0) ? ny * nseg : 0;

    if (total > 0) {
        if (lane == 0) {
            for (int s = 0; s < kFwdStages; ++s) mbar_init(full_bar + s, 1);
            fence_mbar_init();
        }
        __syncwarp();
    }
    int tt = 0;                                         // stage uses so far (ring position / parity)
    for (int cb = 0; cb < p.C; cb += 256) {
        const int nch = min(256, p.C - cb);
        const bool lane_on = lane * 8 < nch;
        const uint32_t px_bytes = (uint32_t)(nch * sizeof(TIn));
        // lane 0: stream segment t of this channel block into ring slot (tt0 + t) % stages
        auto issue = [&](int t, int slot) {
            const int y = ry0 + t / nseg, xs = cx0 + (t % nseg) * kFwdPx;
            const int npx = min(kFwdPx, cx1 - xs + 1);
            const TIn* src = feat + (((size_t)b * H + y) * W + xs) * p.C + cb;
            TIn* dst = ring + (size_t)slot * stage_elems;
            mbar_expect_tx(full_bar + slot, npx * px_bytes);
            if (p.C <= 256) {
                bulk_g2s(dst, src, npx * px_bytes, full_bar + slot);
            } else {
                for (int x = 0; x < npx; ++x)
                    bulk_g2s(dst + (size_t)x * cw, src + (size_t)x * p.C, px_bytes, full_bar + slot);
            }
        };
        if (lane == 0)
            for (int t = 0; t < min(kFwdStages, total); ++t) issue(t, (tt + t) % kFwdStages);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        float wy_l = 0.f, wx_l = 0.f;
        int t = 0;
        for (int y = 0; y < ny && total > 0; ++y) {
            if ((y & 31) == 0) wy_l = (y + lane < ny) ? __ldg(wyt + (size_t)(y + lane) * kTabW) : 0.f;
            const float wyv = __shfl_sync(0xffffffffu, wy_l, y & 31);
            for (int j = 0; j < nseg; ++j, ++t, ++tt) {
                const int xo = j * kFwdPx;                 // offset of the segment inside the bin
                if ((xo & 31) == 0)
                    wx_l = (xo + lane < nx) ? __ldg(wxt + (size_t)(xo + lane) * kTabW) : 0.f;
                const int slot = tt % kFwdStages;
                mbar_wait(full_bar + slot, (uint32_t)(tt / kFwdStages) & 1u);
                if (wyv != 0.f) {
                    const int npx = min(kFwdPx, nx - xo);
                    const TIn* src = ring + (size_t)slot * stage_elems + lane * 8;
#pragma unroll
                    for (int x = 0; x < kFwdPx; ++x) {
                        if (x < npx) {
                            const float w = wyv * __shfl_sync(0xffffffffu, wx_l, (xo + x) & 31);
                            if (lane_on) {
                                float v[8];
                                ld_smem8<TIn>(src + (size_t)x * cw, v);
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, v[e], acc[e]);
                            }
                        }
                    }
                }
                __syncwarp();                            // all lanes done with the slot
                if (lane == 0 && t + kFwdStages < total) issue(t + kFwdStages, slot);
            }
        }
        if (p.bias && bvalid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = lane * 8 + e;
                if (c < nch) acc[e] += __ldg(p.bias + (size_t)b * p.C + cb + c);
            }
        }
        Vec8<TOut, false>::store(orow + cb, lane, nch, acc);
    }
}

// ------------------------------------------------------------------------------------------
// persistent forward (single-level extraction; every strip is small)
// ------------------------------------------------------------------------------------------
// The kernel above pays one exposed memory latency per CTA (a CTA = one bin row of one RoI) and
// only three CTAs fit an SM, so 7168 CTAs take ~16 rounds of a few microseconds: 0.089 ms and 34 %
// of the HBM roofline for SingleRoIExtractor at the bench size (ncu: DRAM at 12 % of peak).  Here one
// CTA per SM stays resident and walks the same work units (RoI, bin row) in order:
//   warp P      producer: reads the plan entries of a unit several units AHEAD of the consumers,
//               carves the strip's footprint out of a byte ring in shared memory (~200 KB: five to
//               seven strips in flight), starts one bulk copy per footprint row plus one for each
//               slice of the separable axis-weight tables (cp.async.bulk, completion counted on the
//               unit's mbarrier) and publishes a small descriptor;
//   warps 0..P-1  one output bin (ph, pw) each: wait for the unit, reduce the bin from the strip
//               with weights read from the staged tables (broadcast shared loads - no shuffles), add
//               the SFA bias, store 512 B, release the unit.
// The arithmetic (w = wy * wx, one fma per pixel in row-major order) is the one of the strip / ring
// paths above, so results are bit-identical.  A strip that does not fit the ring (never the case
// for a level-assigned RoI of an 800x1333 image) is reduced straight from global memory.
constexpr int kPfUnits = 8;                      // descriptor ring (units in flight)
constexpr int kPfGroups = 3;                     // consumer groups of P warps: units i, i+1, i+2 are reduced concurrently
                                                 // (one group alone leaves the SM's schedulers idle: 7 warps
                                                 // of dependent shared-load -> fma chains reached 0.095 ms)
constexpr int kPfRingBytes = 200 * 1024;         // staging ring
// One unit = one bin row of one RoI; 32 ints, so that the producer's lanes move it with one shared
// load and one shared store each.
struct PfUnit {
    const void* gsrc;            // first footprint pixel of the strip
    const float* wy_src;         // the strip's rows / the RoI's columns in the axis-weight tables
    const float* wx_src;
    long long out_off;           // output element offset of bin (ph, 0)
    int live, direct, b, ny, nxs, W, need;       // need: ring bytes (strip + both table slices)
    int dx0[HTD_MAX_POOLED], nx[HTD_MAX_POOLED - 1];
    int ph;
    int buf_off;                 // ring offset of the strip; the table slices follow it
};
static_assert(sizeof(PfUnit) == 128 && HTD_MAX_POOLED == 8, "PfUnit is moved as 32 ints");
constexpr int kPfBatch = 32;                     // units whose plan is fetched at once (one per lane)
constexpr int kPfSmem = kPfRingBytes + (kPfUnits + kPfBatch) * (int)sizeof(PfUnit) + 2 * kPfUnits * 8 +
                        kPfUnits * 4 + 16;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__((kPfGroups * (HTD_MAX_POOLED - 1) + 1) * 32, 1) roi_align_fwd_persist_kernel(const FwdParams p,
                                                                                         int units) {
    extern __shared__ __align__(128) uint8_t pf_smem[];
    uint8_t* ringb = pf_smem;
    PfUnit* desc = reinterpret_cast<PfUnit*>(pf_smem + kPfRingBytes);
    PfUnit* plan = desc + kPfUnits;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(plan + kPfBatch);
    uint64_t* empty_bar = full_bar + kPfUnits;
    int* used = reinterpret_cast<int*>(empty_bar + kPfUnits);   // ring bytes (strip + wrap waste) per slot
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int P = p.P, PP = P * P;
    const int grid = (int)gridDim.x;
    if (threadIdx.x == 0) {
        for (int q = 0; q < kPfUnits; ++q) {
            mbar_init(full_bar + q, 1);
            mbar_init(empty_bar + q, P);
        }
        fence_mbar_init();
    }
    __syncthreads();
    if (warp == kPfGroups * P) {
        // ===== producer =====
        // It is the only warp that starts copies, so everything it does per unit is serial time of
        // the whole CTA (ncu: the consumers sat on the full barrier 58 - 77 % of the time).  (1) The
        // plan entries of a unit sit behind two dependent global reads (level -> box / ranges /
        // offset): the lanes fetch the plans of kPfBatch units at once - one unit per lane, the
        // same two round trips per BATCH - and derive every per-unit quantity there, in parallel.
        // (2) The issue loop then only allocates ring space, moves the 128-byte record into the
        // descriptor ring (one int per lane) and starts the copies; indices are 32-bit.
        int head = 0, free_bytes = kPfRingBytes;
        int oldest = 0;                              // local index of the oldest unreleased unit
        int i = 0;
        for (int u0 = (int)blockIdx.x; u0 < units; u0 += kPfBatch * grid) {
            {
                const long long u = (long long)u0 + (long long)lane * grid;
                PfUnit pl;
                pl.gsrc = nullptr; pl.wy_src = nullptr; pl.wx_src = nullptr; pl.out_off = 0;
                pl.live = 0; pl.direct = 0; pl.b = -1; pl.ny = 0; pl.nxs = 0; pl.W = 1; pl.need = 0;
                pl.ph = 0; pl.buf_off = 0;
#pragma unroll
                for (int j = 0; j < HTD_MAX_POOLED; ++j) pl.dx0[j] = 0;
#pragma unroll
                for (int j = 0; j < HTD_MAX_POOLED - 1; ++j) pl.nx[j] = 0;
                if (u < units) {
                    const int k = (int)u / P;
                    const int ph = (int)u - k * P;
                    const int l = p.roi_level[k];
                    const int b = (int)p.rois[(size_t)k * 5];
                    pl.b = b;
                    pl.ph = ph;
                    pl.out_off = ((long long)k * PP + (long long)ph * P) * p.C;
                    if ((b >= 0 && b < p.B) && (l >= 0 && l < p.L)) {
                        const size_t e = (size_t)l * p.K + k;
                        const int4 box = p.boxes[e];
                        const int* rg = p.ranges + e * kRangeInts;
                        const int ry0 = rg[ph], ry1 = rg[HTD_MAX_POOLED + ph];
                        const int4* rx = reinterpret_cast<const int4*>(rg + 2 * HTD_MAX_POOLED);
                        const int4 lo0 = rx[0], lo1 = rx[1], hi0 = rx[2], hi1 = rx[3];
                        const size_t off = (size_t)p.offsets[e];
                        const int ny = ry1 - ry0 + 1, nxs = box.w - box.z + 1;
                        if (box.y >= box.x && box.w >= box.z && ny > 0 && nxs > 0) {
                            const int H = p.lv[l].H, W = p.lv[l].W;
                            const int cx0[8] = {lo0.x, lo0.y, lo0.z, lo0.w, lo1.x, lo1.y, lo1.z, lo1.w};
                            const int cx1[8] = {hi0.x, hi0.y, hi0.z, hi0.w, hi1.x, hi1.y, hi1.z, hi1.w};
#pragma unroll
                            for (int j = 0; j < HTD_MAX_POOLED; ++j) pl.dx0[j] = cx0[j] - box.z;
#pragma unroll
                            for (int j = 0; j < HTD_MAX_POOLED - 1; ++j)
                                pl.nx[j] = j < P ? cx1[j] - cx0[j] + 1 : 0;
                            const long long strip = (long long)ny * nxs * p.C * (long long)sizeof(TIn);
                            const int tab_bytes = (ny + nxs) * kTabW * 4;
                            pl.direct = strip + tab_bytes > kPfRingBytes;
                            pl.need = (pl.direct ? 0 : (int)strip) + tab_bytes;      // multiple of 16
                            pl.live = 1; pl.ny = ny; pl.nxs = nxs; pl.W = W;
                            pl.gsrc = static_cast<const TIn*>(p.lv[l].data) +
                                      (((size_t)b * H + ry0) * W + box.z) * p.C;
                            pl.wy_src = p.weights + (off + (size_t)(ry0 - box.x)) * kTabW;
                            pl.wx_src = p.weights + (off + (size_t)(box.y - box.x + 1)) * kTabW;
                        }
                    }
                }
                plan[lane] = pl;
            }
            __syncwarp();
            const int nb = min(kPfBatch, (units - u0 + grid - 1) / grid);
            for (int j = 0; j < nb; ++j, ++i) {
                const PfUnit* pl = plan + j;
                const int q = i % kPfUnits;
                const int need = pl->need;
                // ---- ring space: units are released in order
                int waste = 0;
                bool wrap = false;
                for (;;) {
                    wrap = head + need > kPfRingBytes;
                    waste = wrap ? kPfRingBytes - head : 0;
                    if (oldest + kPfUnits > i && free_bytes >= need + waste) break;
                    if (oldest == i) {           // ring drained: restart at its beginning (need <= ring)
                        head = 0;
                        continue;
                    }
                    const int oq = oldest % kPfUnits;
                    mbar_wait(empty_bar + oq, (uint32_t)(oldest / kPfUnits) & 1u);
                    free_bytes += used[oq];
                    ++oldest;
                }
                if (wrap) head = 0;
                const int buf = head;
                head += need;
                free_bytes -= need + waste;
                // ---- descriptor: one int per lane, the last one is the ring offset
                reinterpret_cast<int*>(desc + q)[lane] =
                    lane == 31 ? buf : reinterpret_cast<const int*>(pl)[lane];
                if (lane == 0) used[q] = need + waste;
                const int live = pl->live, ny = pl->ny, nxs = pl->nxs;
                __syncwarp();
                if (!live) {
                    if (lane == 0) mbar_arrive(full_bar + q);
                    continue;
                }
                if (lane == 0) mbar_expect_tx(full_bar + q, (uint32_t)need);
                __syncwarp();
                const int row_bytes = nxs * p.C * (int)sizeof(TIn);
                const int strip = pl->direct ? 0 : ny * row_bytes;
                if (lane == 31)
                    bulk_g2s(ringb + buf + strip, pl->wy_src, (uint32_t)(ny * kTabW * 4), full_bar + q);
                else if (lane == 30)
                    bulk_g2s(ringb + buf + strip + ny * kTabW * 4, pl->wx_src,
                             (uint32_t)(nxs * kTabW * 4), full_bar + q);
                if (strip) {
                    const TIn* src = static_cast<const TIn*>(pl->gsrc);
                    const size_t gstride = (size_t)pl->W * p.C;
                    for (int r = lane; r < ny; r += 32)
                        bulk_g2s(ringb + buf + (size_t)r * row_bytes, src + (size_t)r * gstride,
                                 (uint32_t)row_bytes, full_bar + q);
                }
            }
            __syncwarp();                            // the next batch overwrites the staged plans
        }
    } else if (warp < kPfGroups * P) {
        // ===== consumers: group g reduces the units i = g, g + G, ...; a warp = one bin (ph, pw) =====
        const int pw = warp % P, grp = warp / P;
        const bool lane_on = lane * 8 < p.C;
        int i = grp;
        for (int u = (int)blockIdx.x + grp * grid; u < units; u += kPfGroups * grid, i += kPfGroups) {
            const int q = i % kPfUnits;
            mbar_wait(full_bar + q, (uint32_t)(i / kPfUnits) & 1u);
            const PfUnit* d = desc + q;
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
            const int b = d->b;
            float bv[8];                             // SFA bias: in flight while the bin is reduced
#pragma unroll
            for (int e = 0; e < 8; ++e) bv[e] = 0.f;
            if (p.bias && b >= 0 && b < p.B && lane_on) {
#pragma unroll
                for (int e = 0; e < 8; ++e) bv[e] = __ldg(p.bias + (size_t)b * p.C + lane * 8 + e);
            }
            if (d->live) {
                const int ny = d->ny, nxs = d->nxs, nx = d->nx[pw], dx0 = d->dx0[pw];
                const int strip = d->direct ? 0 : ny * nxs * p.C * (int)sizeof(TIn);
                const float* ty = reinterpret_cast<const float*>(ringb + d->buf_off + strip) + d->ph;
                const float* tx = ty - d->ph + (size_t)(ny + dx0) * kTabW + pw;
                if (nx > 0 && lane_on) {
                    if (!d->direct) {
                        const TIn* sb = reinterpret_cast<const TIn*>(ringb + d->buf_off) + (size_t)dx0 * p.C + lane * 8;
                        for (int y = 0; y < ny; ++y) {
                            const float wyv = ty[(size_t)y * kTabW];
                            if (wyv == 0.f) continue;
                            const TIn* src = sb + (size_t)y * nxs * p.C;
#pragma unroll 4
                            for (int x = 0; x < nx; ++x) {
                                const float w = wyv * tx[(size_t)x * kTabW];
                                float v[8];
                                ld_smem8<TIn>(src + (size_t)x * p.C, v);
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, v[e], acc[e]);
                            }
                        }
                    } else {
                        const TIn* gb = static_cast<const TIn*>(d->gsrc) + (size_t)dx0 * p.C + lane * 8;
                        for (int y = 0; y < ny; ++y) {
                            const float wyv = ty[(size_t)y * kTabW];
                            if (wyv == 0.f) continue;
                            const TIn* src = gb + (size_t)y * d->W * p.C;
                            for (int x = 0; x < nx; ++x) {
                                const float w = wyv * tx[(size_t)x * kTabW];
                                float v[8];
                                ld_smem8<TIn>(src + (size_t)x * p.C, v);      // plain vector load
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, v[e], acc[e]);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += bv[e];
            TOut* orow = static_cast<TOut*>(p.out) + d->out_off + (size_t)pw * p.C;
            Vec8<TOut, false>::store(orow, lane, p.C, acc);
            __syncwarp();
            if (lane == 0) mbar_arrive(empty_bar + q);
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct BwdSource {                   // one extractor call whose gradient lands in dX
    const float* rois;
    const int4* boxes;
    const int* offsets;
    const int* ranges;
    const float* weights;
    const void* dy;
    const float* scale;
    const float* addvec;             // fp32, or bf16 when av_bf16 (tensor-pipe kernel only)
    int K, dy_per_level, ring_edge, av_bf16;
};

struct BwdParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int tile_start[HTD_MAX_LEVELS + 1];
    int L, B, C, P, nsrc, nchw;
    unsigned long long* trace;       // diagnostics: per-CTA {t0, t1, hits, blocks, smid, level} or null
    BwdSource src[HTD_MAX_BWD_SOURCES];
};

__device__ __forceinline__ void st_elem(float* p, float v) { *p = v; }
__device__ __forceinline__ void st_elem(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

constexpr int kTile = 8;  // 8x8 pixels per CTA
constexpr int kBwdWarps = 16;          // consumer warps: tile row (w & 7) x column half (w >> 3)
constexpr int kBwdThreads = (kBwdWarps + 1) * 32;   // + 1 producer warp
constexpr int kBwdChunk = kBwdWarps * 32;          // RoIs scanned per pass
constexpr int kHalf = kTile / 2;      // pixels per consumer warp

#ifndef HTD_BWD_STAGES_BF16
#define HTD_BWD_STAGES_BF16 4
#endif
#ifndef HTD_BWD_STAGES_F32
#define HTD_BWD_STAGES_F32 2
#endif
template <typename TDy>
struct BwdStages {                     // dY ring depth: 25 KB (bf16) / 50 KB (fp32) per slot at P = 7
    static constexpr int value = sizeof(TDy) == 2 ? HTD_BWD_STAGES_BF16 : HTD_BWD_STAGES_F32;
};

// Backward gather.  One CTA per 8x8-pixel tile of dX.  Per chunk of 512 RoIs the 16 consumer warps
// find the RoIs whose footprint box touches the tile and compact them in ascending index order.
// The PRODUCER warp (warp 8) then resolves, for all hits of the chunk in parallel, which output
// bins sample the tile (ranges table) and where the tile's rows/columns sit in the axis-weight
// tables; per hit it streams - with cp.async.bulk (UBLKCP) + mbarrier transaction counts, several
// hits ahead of the consumers - the tile's slices of the weight tables and the needed dY bins (per
// bin row one contiguous run) into a ring of shared-memory slots.  Each consumer warp owns half a
// tile COLUMN (4 pixels) and works separably: t[c] = sum_pw wx[pw][col] dY[ph][pw][c] once per bin
// row, then acc[r][c] += wy[ph][r] t[c] for its 4 rows - registers only, reading shared memory only.  Every dX element is written exactly once:
// no atomics, no memset, bit-reproducible.
template <typename TDy, typename TDx>
__global__ void __launch_bounds__(kBwdThreads, 1) roi_align_bwd_kernel(const BwdParams p) {
    constexpr int kS = BwdStages<TDy>::value;
    extern __shared__ __align__(128) uint8_t bwd_smem[];
    __shared__ __align__(16) float s_wy[kS][kTile][kTabW];      // [slot][tile row][bin]
    __shared__ __align__(16) float s_wx[kS][kTile][kTabW];      // [slot][tile col][bin]
    __shared__ int4 s_meta[kBwdChunk];                  // per hit: k, bins, tile-relative ranges, y table row
    __shared__ int s_xrow[kBwdChunk];                   // per hit: x table row of the first staged column
    __shared__ float s_scale[kBwdChunk];              // per hit: BA level weight (1 when unused)
    __shared__ __align__(16) float s_av[kS][256];     // per slot: per-channel add vector of the hit
    __shared__ int s_hits[kBwdChunk];
    __shared__ int s_wcnt[2][kBwdWarps];                  // double-buffered by chunk parity
    __shared__ uint64_t s_full[kS], s_empty[kS];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = (warp == kBwdWarps);
    // coarse levels first: their tiles collect the most RoIs and must not form the tail of the grid
    const int bid = (int)(gridDim.x - 1 - blockIdx.x);
    int l = 0;
    while (l + 1 < p.L && bid >= p.tile_start[l + 1]) ++l;
    const int H = p.lv[l].H, W = p.lv[l].W;
    const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile;
    int local = bid - p.tile_start[l];
    const int b = local / (tiles_x * tiles_y);
    local -= b * tiles_x * tiles_y;
    const int row0 = (local / tiles_x) * kTile, col0 = (local % tiles_x) * kTile;
    const int P = p.P, PP = P * P;
    const int tcol = warp & 7, rq = (warp >> 3) * kHalf;   // consumers: tile column, first of 4 rows
    const int col = col0 + tcol;
    const int cw = min(p.C, 256);
    TDx* dx = static_cast<TDx*>(p.lv[l].data);
    TDy* dybuf = reinterpret_cast<TDy*>(bwd_smem);          // [kS][PP * cw]
    const size_t buf_elems = (size_t)PP * cw;

    if (tid == 0) {
        for (int s = 0; s < kS; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kBwdWarps); }
        fence_mbar_init();
    }
    __syncthreads();
    unsigned g = 0;                               // hits processed so far (slot / parity counter)

    for (int cb = 0; cb < p.C; cb += 256) {
        const int nch = min(256, p.C - cb);
        const bool lane_on = lane * 8 < nch;
        float acc[kHalf][8];                      // [row of the strip][channel of the lane]
#pragma unroll
        for (int x = 0; x < kHalf; ++x)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[x][e] = 0.f;

        int pass = 0;
        for (int si = 0; si < p.nsrc; ++si) {
        const BwdSource& q = p.src[si];
        const TDy* dy = static_cast<const TDy*>(q.dy);
        const int4* boxes = q.boxes + (size_t)l * q.K;
        for (int k0 = 0; k0 < q.K; k0 += kBwdChunk, ++pass) {
            int* wcnt = s_wcnt[pass & 1];
            // ---- RoIs of this chunk that touch the tile, in ascending index order (consumers)
            if (!producer) {
                const int k = k0 + tid;
                bool hit = false;
                if (k < q.K) {
                    const int4 bx = __ldg(boxes + k);
                    hit = (bx.y >= bx.x) && bx.x <= row0 + kTile - 1 && bx.y >= row0 &&
                          bx.z <= col0 + kTile - 1 && bx.w >= col0 &&
                          ((int)__ldg(q.rois + (size_t)k * 5) == b);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) wcnt[warp] = __popc(bal);
                __syncthreads();
                int base = 0;
#pragma unroll
                for (int w = 0; w < kBwdWarps; ++w)
                    if (w < warp) base += wcnt[w];
                if (hit) s_hits[base + __popc(bal & ((1u << lane) - 1u))] = k;
            } else {
                __syncthreads();
            }
            __syncthreads();
            int nh = 0;
#pragma unroll
            for (int w = 0; w < kBwdWarps; ++w) nh += wcnt[w];
            if (nh == 0) continue;                // CTA-uniform

            if (producer) {
                // ---- metadata of all hits, lanes in parallel
                for (int h = lane; h < nh; h += 32) {
                    const int kk = s_hits[h];
                    const size_t e = (size_t)l * q.K + kk;
                    const int4 bx = __ldg(boxes + kk);
                    const int off = __ldg(q.offsets + e);
                    const int* rg = q.ranges + e * kRangeInts;
                    int pa = 0, pb = -1, qa = 0, qb = -1;
                    bool first = true;
                    for (int pp = 0; pp < P; ++pp) {
                        const int lo = __ldg(rg + pp), hi = __ldg(rg + HTD_MAX_POOLED + pp);
                        if (hi >= lo && lo <= row0 + kTile - 1 && hi >= row0) {
                            if (first) { pa = pp; first = false; }
                            pb = pp;
                        }
                    }
                    first = true;
                    for (int pp = 0; pp < P; ++pp) {
                        const int lo = __ldg(rg + 2 * HTD_MAX_POOLED + pp);
                        const int hi = __ldg(rg + 3 * HTD_MAX_POOLED + pp);
                        if (hi >= lo && lo <= col0 + kTile - 1 && hi >= col0) {
                            if (first) { qa = pp; first = false; }
                            qb = pp;
                        }
                    }
                    if (pb < pa || qb < qa) { pb = pa - 1; qb = qa - 1; }
                    const int r_lo = max(bx.x, row0), r_hi = min(bx.y, row0 + kTile - 1);
                    const int c_lo = max(bx.z, col0), c_hi = min(bx.w, col0 + kTile - 1);
                    s_meta[h] = make_int4(kk, (pa & 0xff) | ((pb & 0xff) << 8) | ((qa & 0xff) << 16) | ((qb & 0xff) << 24),
                                          (r_lo - row0) | ((r_hi - row0) << 8) | ((c_lo - col0) << 16) | ((c_hi - col0) << 24),
                                          off + (r_lo - bx.x));
                    s_xrow[h] = off + (bx.y - bx.x + 1) + (c_lo - bx.z);
                    s_scale[h] = q.scale ? __ldg(q.scale + e) : 1.f;
                }
                __syncwarp();
                // ---- stream hit h into slot (g + h) % kS
                if (lane == 0) {
                    for (int h = 0; h < nh; ++h) {
                        const unsigned gi = g + h;
                        const int slot = gi % kS;
                        const int4 m = s_meta[h];
                        const int kk = m.x;
                        const int pa = (int)(signed char)(m.y & 0xff), pb = (int)(signed char)((m.y >> 8) & 0xff);
                        const int qa = (int)(signed char)((m.y >> 16) & 0xff), qb = (int)(signed char)((m.y >> 24) & 0xff);
                        const int rl = m.z & 0xff, rh = (m.z >> 8) & 0xff, cl = (m.z >> 16) & 0xff, ch = (m.z >> 24) & 0xff;
                        const int nq = qb - qa + 1, np = pb - pa + 1;
                        const uint32_t bin_bytes = (uint32_t)(nch * sizeof(TDy));
                        const uint32_t wy_bytes = (uint32_t)(rh - rl + 1) * kTabW * 4u;
                        const uint32_t wx_bytes = (uint32_t)(ch - cl + 1) * kTabW * 4u;
                        const uint32_t dy_bytes = (np > 0 && nq > 0) ? (uint32_t)(np * nq) * bin_bytes : 0u;
                        mbar_wait(&s_empty[slot], ((gi / kS) & 1u) ^ 1u);       // consumers left the slot
                        const uint32_t av_bytes = q.addvec ? (uint32_t)nch * 4u : 0u;
                        mbar_expect_tx(&s_full[slot], wy_bytes + wx_bytes + dy_bytes + av_bytes);
                        if (av_bytes)
                            bulk_g2s(&s_av[slot][0], q.addvec + ((size_t)l * q.K + kk) * p.C + cb, av_bytes,
                                     &s_full[slot]);
                        bulk_g2s(&s_wy[slot][rl][0], q.weights + (size_t)m.w * kTabW, wy_bytes, &s_full[slot]);
                        bulk_g2s(&s_wx[slot][cl][0], q.weights + (size_t)s_xrow[h] * kTabW, wx_bytes, &s_full[slot]);
                        if (dy_bytes) {
                            const size_t dyk = (q.dy_per_level ? (size_t)l * q.K : 0) + kk;
                            TDy* dst = dybuf + (size_t)slot * buf_elems;
                            for (int ph = pa; ph <= pb; ++ph) {
                                const TDy* src = dy + ((dyk * PP + ph * P + qa) * p.C + cb);
                                if (p.C <= 256) {
                                    bulk_g2s(dst, src, (uint32_t)nq * bin_bytes, &s_full[slot]);
                                    dst += (size_t)nq * cw;
                                } else {
                                    for (int q = 0; q < nq; ++q, dst += cw)
                                        bulk_g2s(dst, src + (size_t)q * p.C, bin_bytes, &s_full[slot]);
                                }
                            }
                        }
                    }
                }
            } else {
                // ===== consumers =====
                for (int h = 0; h < nh; ++h) {
                    const unsigned gi = g + h;
                    const int slot = gi % kS;
                    mbar_wait(&s_full[slot], (gi / kS) & 1u);
                    const int4 m = s_meta[h];
                    const int kk = m.x;
                    const int pa = (int)(signed char)(m.y & 0xff), pb = (int)(signed char)((m.y >> 8) & 0xff);
                    const int qa = (int)(signed char)((m.y >> 16) & 0xff), qb = (int)(signed char)((m.y >> 24) & 0xff);
                    const int rl = m.z & 0xff, rh = (m.z >> 8) & 0xff, cl = (m.z >> 16) & 0xff, ch = (m.z >> 24) & 0xff;
                    if (col < W && pb >= pa && tcol >= cl && tcol <= ch && rq <= rh && rq + kHalf - 1 >= rl) {
                        // separable inside the tile: t[c] = sum_pw wx[pw][col] g(ph,pw)[c] once per bin
                        // row, then acc[r][c] += wy[ph][r] t[c] for the warp's 4 rows
                        const int nq = qb - qa + 1;
                        const float sbase = s_scale[h];
                        float av[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) av[e] = 0.f;
                        if (q.addvec && lane_on) ld_smem8<float>(&s_av[slot][lane * 8], av);
                        const int e0 = (l == 0) ? q.ring_edge : -1;
                        const TDy* src = dybuf + (size_t)slot * buf_elems + lane * 8;
                        for (int ph = pa; ph <= pb; ++ph) {
                            float wyr[kHalf];
                            bool anyr = false;
#pragma unroll
                            for (int r = 0; r < kHalf; ++r) {
                                wyr[r] = (rq + r >= rl && rq + r <= rh) ? s_wy[slot][rq + r][ph] : 0.f;
                                anyr |= (wyr[r] != 0.f);
                            }
                            if (!anyr) continue;
                            float t[8];
#pragma unroll
                            for (int e = 0; e < 8; ++e) t[e] = 0.f;
                            float wsum = 0.f;
                            for (int pw = qa; pw <= qb; ++pw) {
                                const float wxv = s_wx[slot][tcol][pw];
                                if (wxv == 0.f) continue;
                                float sv = sbase;
                                if (e0 >= 0) {
                                    const bool interior = (e0 > 0) && ph >= e0 && ph < P - e0 &&
                                                          pw >= e0 && pw < P - e0;
                                    if (!interior) sv += 1.f;
                                }
                                wsum += wxv;
                                const float w = wxv * sv;
                                if (lane_on) {
                                    float v[8];
                                    ld_smem8<TDy>(src + (size_t)((ph - pa) * nq + (pw - qa)) * cw, v);
#pragma unroll
                                    for (int e = 0; e < 8; ++e) t[e] = fmaf(w, v[e], t[e]);
                                }
                            }
                            if (wsum == 0.f) continue;
#pragma unroll
                            for (int e = 0; e < 8; ++e) t[e] = fmaf(wsum, av[e], t[e]);
#pragma unroll
                            for (int r = 0; r < kHalf; ++r) {
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[r][e] = fmaf(wyr[r], t[e], acc[r][e]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_empty[slot]);
                }
            }
            g += (unsigned)nh;
            __syncthreads();                      // s_hits / s_meta are rewritten by the next chunk
        }
        }
        if (!producer && col < W) {
            if (!p.nchw) {
#pragma unroll
                for (int r = 0; r < kHalf; ++r) {
                    const int row = row0 + rq + r;
                    if (row < H)
                        Vec8<TDx, false>::store(dx + (((size_t)b * H + row) * W + col) * p.C + cb,
                                                lane, nch, acc[r]);
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const int c = cb + lane * 8 + e;
                    if (lane * 8 + e < nch) {
#pragma unroll
                        for (int r = 0; r < kHalf; ++r) {
                            const int row = row0 + rq + r;
                            if (row < H)
                                st_elem(dx + (((size_t)b * p.C + c) * H + row) * W + col, acc[r][e]);
                        }
                    }
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// backward, bf16 dY: the per-hit contraction on the tensor pipe
// ------------------------------------------------------------------------------------------
// For one (tile, RoI) hit the scalar kernel above evaluates
//     dX[px][c] += sum_{bin} wy[row(px)][ph(bin)] * wx[col(px)][pw(bin)] * sv(bin) * dY[bin][c]
// with 16 warps that each re-read every staged bin.  The sum is a [64 px x bins] x [bins x 256 ch]
// product, so this kernel keeps the tile-gather structure - hits compacted in index order, a
// producer warp's bulk-copy ring, one write per dX element, no atomics - and does the per-hit
// contraction with warp-level bf16 MMAs (m16n8k16, fp32 accumulators):
//   * scan: ALL chunks of ALL sources are scanned in one batch (counts per (chunk, warp), one
//     exclusive scan, records written in ascending (source, index) order); the thread that found
//     a hit also resolves its bins / table rows, so the producer only streams;
//   * K index of a staged bin = ph_rel * 8 + pw_rel (8 slots per bin row, so a K step of 16 is two
//     bin rows and the A fragment of a lane needs just wx[col][qa + 2t .. +1] once per hit and
//     wy[row][ph] of two rows per step); unused slots carry weight 0;
//   * ring of K-step BLOCKS (16 bin rows = 8448 B): a hit takes ceil(np / 2) consecutive blocks,
//     ONE full barrier (its first block's) and one producer wait (on its last block); the producer
//     places every needed bin as its own 512 B row at a 528 B pitch (one bulk copy per bin, lanes
//     in parallel), which makes the transposing ldmatrix of the B fragments bank-conflict free;
//   * a consumer warp owns 2 tile rows x 8 columns (M = 16) x 256 / kCG channels;
//   * the BA add vector is one more K row (slot 7 of the hit's first block, bf16) with weight
//     rowsum(wy) * colsum(wx) - the sums are entry 7 of the plan's table rows; with pooled == 8
//     it stays the fp32 rank-1 FMA term;
//   * the interpolation weights are rounded to bf16 (2^-9 relative) - inside the 2e-2 bf16 gate;
//     fp32 dY keeps the exact scalar kernel.
// Measured cost model at the BASELINE sizes (tools/trace_bwd.py, 2848 tiles, 44.8 k hits, 65.4 k
// blocks): t_tile = 12 us + 0.7 us * hits + 0.6 us * blocks with two CTAs per SM.
constexpr int kMmaRowBytes = 256 * 2 + 16;               // one bin: 256 bf16 + 16 B pad
constexpr int kMmaBlockBytes = 16 * kMmaRowBytes;        // one K step: two bin rows x 8 slots (8448)
constexpr int kStagePitch = 256 + 8;                     // floats per pixel in the store staging

__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0,
                                               uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
        "{%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

constexpr int kHitCap = 1024;          // hit records of one batch (>= RoIs per chunk)

// kCG = channel groups: 4 * kCG consumer warps, each 2 tile rows x 8 columns x (256 / kCG) channels
template <typename TDx, int kR, int kMinCtas, int kCG>
__global__ void __launch_bounds__((4 * kCG + 1) * 32, kMinCtas) roi_align_bwd_mma_kernel(const BwdParams p) {
    typedef __nv_bfloat16 TDy;
    constexpr int kW = 4 * kCG;                // consumer warps
    constexpr int kChunk = kW * 32;            // RoIs scanned per chunk
    constexpr int kThreads = kChunk + 32;      // + 1 producer warp
    constexpr int kCW = 256 / kCG;             // channels per consumer warp
    constexpr int kNT = kCW / 8;               // n-tiles (8 channels) per consumer warp
    // chunks scanned per batch: 5 bits of rank each in a 64-bit word, 128 (chunk, warp) counts
    constexpr int kBatchChunks = (128 / kW) < 12 ? (128 / kW) : 12;
    // ring of kR K-step blocks (16 bin rows each); a hit takes ceil(np / 2) consecutive blocks, so
    // the ring holds many small hits or a few large ones.  Tables live in a ring of kR per-hit slots.
    extern __shared__ __align__(128) uint8_t bwd_smem[];
    __shared__ __align__(16) float s_wy[kR][kTile][kTabW];      // [table slot][tile row][bin]
    __shared__ __align__(16) float s_wx[kR][kTile][kTabW];      // [table slot][tile col][bin]
    __shared__ int4 s_meta[kHitCap];    // per hit: k | source << 28, bins, tile-relative ranges, y table row
    __shared__ int s_xrow[kHitCap];     // per hit: x table row of the first staged column
    __shared__ float s_scale[kHitCap];  // per hit: BA level weight (1 when unused)
    __shared__ __align__(16) float s_av[kR][256];
    __shared__ int s_cnt[kBatchChunks * kW + 1];       // per (chunk, warp) counts -> bases
    __shared__ uint64_t s_full[kR], s_empty[kR];        // per block

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = (warp == kW);
    // coarse levels first: their tiles collect the most RoIs and must not form the tail of the grid
    const int bid = (int)(gridDim.x - 1 - blockIdx.x);
    int l = 0;
    while (l + 1 < p.L && bid >= p.tile_start[l + 1]) ++l;
    const int H = p.lv[l].H, W = p.lv[l].W;
    const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile;
    int local = bid - p.tile_start[l];
    const int b = local / (tiles_x * tiles_y);
    local -= b * tiles_x * tiles_y;
    const int row0 = (local / tiles_x) * kTile, col0 = (local % tiles_x) * kTile;
    const int P = p.P, PP = P * P;
    const int C = p.C;                                     // <= 256, multiple of 64 (host check)
    TDx* dx = static_cast<TDx*>(p.lv[l].data);
    const uint32_t ring_u32 = smem_u32(bwd_smem);
    unsigned long long t_start = 0ull;
    if (p.trace && tid == 0) asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_start));

    // the ring is read beyond the copied rows (K slots of weight 0): make it finite once
    for (int i = tid; i < kR * kMmaBlockBytes / 16; i += kThreads)
        reinterpret_cast<uint4*>(bwd_smem)[i] = make_uint4(0u, 0u, 0u, 0u);
    if (tid == 0) {
        for (int s = 0; s < kR; ++s) { mbar_init(&s_full[s], 1); mbar_init(&s_empty[s], kW); }
        fence_mbar_init();
    }
    fence_proxy_async_smem();
    __syncthreads();
    unsigned g = 0, gb = 0;                       // hits / blocks streamed so far (ring positions, parities)
    unsigned fullph = 0;                          // consumers: phase bit of every full barrier

    // consumer roles: pixel group (tile rows 2*pg, 2*pg+1) x channel group (64 channels)
    const int pg = warp & 3, cg = warp >> 2;
    const int gid = lane >> 2, tq = lane & 3;      // mma fragment coordinates
    const int r0 = 2 * pg, r1 = r0 + 1;
    const bool cg_on = cg * kCW < C;
    // ldmatrix lane address inside a slot: matrix m = lane >> 3 -> K half (m & 1), n-tile (m >> 1)
    const uint32_t ld_off = (uint32_t)((((lane >> 3) & 1) * 8 + (lane & 7)) * kMmaRowBytes +
                                       (cg * kCW + (lane >> 4) * 8) * 2);
    float acc[kNT][4];
#pragma unroll
    for (int j = 0; j < kNT; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

    // chunks of all sources, numbered consecutively: chunk c -> (source, first RoI)
    int total_chunks = 0;
    for (int si = 0; si < p.nsrc; ++si) total_chunks += (p.src[si].K + kChunk - 1) / kChunk;

    for (int c_begin = 0; c_begin < total_chunks;) {
        // ---- scan: up to kBatchChunks chunks at once.  Pass 1 counts the RoIs whose footprint box
        // touches the tile per (chunk, warp); the records are then written in ascending (source,
        // index) order, so the accumulation order - and the result - is fixed.
        const int nb = min(kBatchChunks, total_chunks - c_begin);
        unsigned hitmask = 0;
        unsigned long long ranks = 0ull;          // 5 bits per chunk: rank of the lane's hit in its warp
        if (!producer) {
            int si = 0, c = c_begin;
            while (c >= (p.src[si].K + kChunk - 1) / kChunk) { c -= (p.src[si].K + kChunk - 1) / kChunk; ++si; }
            for (int j = 0; j < nb; ++j) {
                const BwdSource& q = p.src[si];
                const int k = c * kChunk + tid;
                bool hit = false;
                if (k < q.K) {
                    const int4 bx = __ldg(q.boxes + (size_t)l * q.K + k);
                    hit = (bx.y >= bx.x) && bx.x <= row0 + kTile - 1 && bx.y >= row0 &&
                          bx.z <= col0 + kTile - 1 && bx.w >= col0 &&
                          ((int)__ldg(q.rois + (size_t)k * 5) == b);
                }
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                if (lane == 0) s_cnt[j * kW + warp] = __popc(bal);
                if (hit) {
                    hitmask |= 1u << j;
                    ranks |= (unsigned long long)__popc(bal & ((1u << lane) - 1u)) << (5 * j);
                }
                if (++c >= (q.K + kChunk - 1) / kChunk) { c = 0; ++si; }
            }
        }
        __syncthreads();
        if (warp == 0) {                           // exclusive scan of the nb * 16 counts, in place
            const int n = nb * kW;
            int v[4], sum = 0;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = lane * 4 + i;
                v[i] = idx < n ? s_cnt[idx] : 0;
                sum += v[i];
            }
            int inc = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int u = __shfl_up_sync(0xffffffffu, inc, o);
                if (lane >= o) inc += u;
            }
            int run = inc - sum;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int idx = lane * 4 + i;
                if (idx < n) s_cnt[idx] = run;
                run += v[i];
            }
            if (lane == 31) s_cnt[n] = inc;
        }
        __syncthreads();
        // chunks that fit the record table (the first always does); the rest is rescanned
        int nbf = 1;
        while (nbf < nb && s_cnt[(nbf + 1) * kW] <= kHitCap) ++nbf;
        const int nh = s_cnt[nbf * kW];
        if (!producer && hitmask) {
            int si = 0, c = c_begin;
            while (c >= (p.src[si].K + kChunk - 1) / kChunk) { c -= (p.src[si].K + kChunk - 1) / kChunk; ++si; }
            for (int j = 0; j < nbf; ++j) {
                const BwdSource& q = p.src[si];
                if ((hitmask >> j) & 1u) {
                    const int kk = c * kChunk + tid;
                    const int pos = s_cnt[j * kW + warp] + (int)((ranks >> (5 * j)) & 31ull);
                    const size_t e = (size_t)l * q.K + kk;
                    const int4 bx = __ldg(q.boxes + e);
                    const int off = __ldg(q.offsets + e);
                    const float sc = q.scale ? __ldg(q.scale + e) : 1.f;
                    const int4* rg = reinterpret_cast<const int4*>(q.ranges + e * kRangeInts);
                    int lo[8], hi[8];
                    int pa = 0, pb = -1, qa = 0, qb = -1;
                    {
                        const int4 a0 = __ldg(rg), a1 = __ldg(rg + 1), h0 = __ldg(rg + 2), h1 = __ldg(rg + 3);
                        lo[0] = a0.x; lo[1] = a0.y; lo[2] = a0.z; lo[3] = a0.w; lo[4] = a1.x; lo[5] = a1.y; lo[6] = a1.z; lo[7] = a1.w;
                        hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
                        bool first = true;
#pragma unroll
                        for (int pp = 0; pp < HTD_MAX_POOLED; ++pp)
                            if (pp < P && hi[pp] >= lo[pp] && lo[pp] <= row0 + kTile - 1 && hi[pp] >= row0) {
                                if (first) { pa = pp; first = false; }
                                pb = pp;
                            }
                    }
                    {
                        const int4 a0 = __ldg(rg + 4), a1 = __ldg(rg + 5), h0 = __ldg(rg + 6), h1 = __ldg(rg + 7);
                        lo[0] = a0.x; lo[1] = a0.y; lo[2] = a0.z; lo[3] = a0.w; lo[4] = a1.x; lo[5] = a1.y; lo[6] = a1.z; lo[7] = a1.w;
                        hi[0] = h0.x; hi[1] = h0.y; hi[2] = h0.z; hi[3] = h0.w; hi[4] = h1.x; hi[5] = h1.y; hi[6] = h1.z; hi[7] = h1.w;
                        bool first = true;
#pragma unroll
                        for (int pp = 0; pp < HTD_MAX_POOLED; ++pp)
                            if (pp < P && hi[pp] >= lo[pp] && lo[pp] <= col0 + kTile - 1 && hi[pp] >= col0) {
                                if (first) { qa = pp; first = false; }
                                qb = pp;
                            }
                    }
                    if (pb < pa || qb < qa) { pb = pa - 1; qb = qa - 1; }
                    const int r_lo = max(bx.x, row0), r_hi = min(bx.y, row0 + kTile - 1);
                    const int c_lo = max(bx.z, col0), c_hi = min(bx.w, col0 + kTile - 1);
                    s_meta[pos] = make_int4(kk | (si << 28),
                                            (pa & 0xff) | ((pb & 0xff) << 8) | ((qa & 0xff) << 16) | ((qb & 0xff) << 24),
                                            (r_lo - row0) | ((r_hi - row0) << 8) | ((c_lo - col0) << 16) | ((c_hi - col0) << 24),
                                            off + (r_lo - bx.x));
                    s_xrow[pos] = off + (bx.y - bx.x + 1) + (c_lo - bx.z);
                    s_scale[pos] = sc;
                }
                if (++c >= (q.K + kChunk - 1) / kChunk) { c = 0; ++si; }
            }
        }
        c_begin += nbf;
        __syncthreads();
        if (nh == 0) continue;                    // CTA-uniform

        if (producer) {
            // ---- stream the hits block by block: lane 0 the tables (with the first block), all
            // lanes the bins of the block's two bin rows
            for (int h = 0; h < nh; ++h) {
                const int4 m = s_meta[h];
                const int pa = (int)(signed char)(m.y & 0xff), pb = (int)(signed char)((m.y >> 8) & 0xff);
                const int qa = (int)(signed char)((m.y >> 16) & 0xff), qb = (int)(signed char)((m.y >> 24) & 0xff);
                const int nq = qb - qa + 1, np = pb - pa + 1;
                if (np <= 0 || nq <= 0) continue;
                const int kk = m.x & 0x0fffffff;
                const BwdSource& q = p.src[(unsigned)m.x >> 28];
                const int ts = g % kR;
                const uint32_t bin_bytes = (uint32_t)C * 2u;
                const TDy* dy = static_cast<const TDy*>(q.dy);
                const size_t dyk = (q.dy_per_level ? (size_t)l * q.K : 0) + kk;
                const int nst = (np + 1) >> 1;
                // pooled < 8: K slot 7 of the first block is free and takes the add vector as a bf16
                // row (weight rowsum * colsum); otherwise it travels as fp32 for the FMA form
                const bool av_row = q.addvec != nullptr && P < kTabW;
                const bool av_cvt = av_row && !q.av_bf16;      // fp32 add vector: converted here
                float4 av0 = make_float4(0.f, 0.f, 0.f, 0.f), av1 = av0;
                if (av_cvt && lane * 8 < C) {
                    const float4* ap = reinterpret_cast<const float4*>(q.addvec + ((size_t)l * q.K + kk) * C + lane * 8);
                    av0 = __ldg(ap);
                    av1 = __ldg(ap + 1);
                }
                uint64_t* full0 = &s_full[gb % kR];        // ONE full barrier per hit: its first block's
                // consumers release the blocks hit by hit and in order: when the LAST block of the
                // range is free, the earlier ones are too - one wait per hit
                if (lane == 0) {
                    const unsigned last = gb + nst - 1;
                    mbar_wait(&s_empty[last % kR], ((last / kR) & 1u) ^ 1u);
                }
                if (av_cvt) {
                    __syncwarp();
                    if (lane * 8 < C)
                        *reinterpret_cast<uint4*>(bwd_smem + (size_t)(gb % kR) * kMmaBlockBytes + 7 * kMmaRowBytes + lane * 16) =
                            make_uint4(pack_bf16x2(av0.x, av0.y), pack_bf16x2(av0.z, av0.w),
                                       pack_bf16x2(av1.x, av1.y), pack_bf16x2(av1.z, av1.w));
                    __syncwarp();                          // ordered before lane 0's arrive below
                }
                if (lane == 0) {
                    const int rl = m.z & 0xff, rh = (m.z >> 8) & 0xff, cl = (m.z >> 16) & 0xff, ch = (m.z >> 24) & 0xff;
                    const uint32_t wy_bytes = (uint32_t)(rh - rl + 1) * kTabW * 4u;
                    const uint32_t wx_bytes = (uint32_t)(ch - cl + 1) * kTabW * 4u;
                    const uint32_t av_bytes = !q.addvec ? 0u : !av_row ? (uint32_t)C * 4u : q.av_bf16 ? bin_bytes : 0u;
                    mbar_expect_tx(full0, (uint32_t)(np * nq) * bin_bytes + wy_bytes + wx_bytes + av_bytes);
                    if (av_bytes && !av_row)
                        bulk_g2s(&s_av[ts][0], q.addvec + ((size_t)l * q.K + kk) * C, av_bytes, full0);
                    if (av_bytes && av_row)                  // bf16 add vector: straight into K slot 7
                        bulk_g2s(bwd_smem + (size_t)(gb % kR) * kMmaBlockBytes + 7 * kMmaRowBytes,
                                 reinterpret_cast<const TDy*>(q.addvec) + ((size_t)l * q.K + kk) * C,
                                 av_bytes, full0);
                    bulk_g2s(&s_wy[ts][rl][0], q.weights + (size_t)m.w * kTabW, wy_bytes, full0);
                    bulk_g2s(&s_wx[ts][cl][0], q.weights + (size_t)s_xrow[h] * kTabW, wx_bytes, full0);
                }
                __syncwarp();
                const TDy* src0 = dy + (dyk * PP + pa * P + qa) * C;
                for (int i = lane; i < np * nq; i += 32) {
                    const int pr = i / nq, qr = i - pr * nq;
                    const int blk = (gb + (pr >> 1)) % kR;
                    bulk_g2s(bwd_smem + (size_t)blk * kMmaBlockBytes + ((pr & 1) * 8 + qr) * kMmaRowBytes,
                             src0 + (size_t)(pr * P + qr) * C, bin_bytes, full0);
                }
                gb += (unsigned)nst;
                ++g;
            }
        } else {
            // ===== consumers =====
            for (int h = 0; h < nh; ++h) {
                const int4 m = s_meta[h];
                const int pa = (int)(signed char)(m.y & 0xff), pb = (int)(signed char)((m.y >> 8) & 0xff);
                const int qa = (int)(signed char)((m.y >> 16) & 0xff), qb = (int)(signed char)((m.y >> 24) & 0xff);
                if (pb < pa || qb < qa) continue;
                const int nst = (pb - pa + 2) >> 1;
                const int ts = g % kR;
                const int rl = m.z & 0xff, rh = (m.z >> 8) & 0xff;
                const bool work = cg_on && r0 <= rh && r1 >= rl;        // warp-uniform
                // one wait per hit (its first block's barrier counts all bytes of the hit).  A warp
                // without work waits all the same: the wait keeps its release from running a whole
                // ring cycle ahead of slower warps.  Only first blocks' barriers ever complete, so
                // their phase is tracked per barrier.
                const int blk0 = gb % kR;
                mbar_wait(&s_full[blk0], (fullph >> blk0) & 1u);
                fullph ^= 1u << blk0;
                if (work) {
                    const BwdSource& q = p.src[(unsigned)m.x >> 28];
                    const int e0 = (l == 0) ? q.ring_edge : -1;
                    const bool av_row = q.addvec != nullptr && P < kTabW;     // K slot 7 of block 0
                    const bool av_fma = q.addvec != nullptr && !av_row;
                    const int cl = (m.z >> 16) & 0xff, ch = (m.z >> 24) & 0xff;
                    const float sbase = s_scale[h];
                    const bool colok = gid >= cl && gid <= ch;
                    const bool rok0 = r0 >= rl && r0 <= rh, rok1 = r1 >= rl && r1 <= rh;
                    const int pw0 = qa + 2 * tq, pw1 = pw0 + 1;
                    // BA border ring on the finest level: bins outside the interior count twice
                    const bool pin0 = e0 > 0 && pw0 >= e0 && pw0 < P - e0;
                    const bool pin1 = e0 > 0 && pw1 >= e0 && pw1 < P - e0;
                    // column factors of the lane's two K slots: b* for a border bin row, i* inside
                    float wx0 = 0.f, wx1 = 0.f, b0 = 0.f, b1 = 0.f, i0 = 0.f, i1 = 0.f;
                    float w70 = 0.f, w71 = 0.f, rs0 = 0.f, rs1 = 0.f;
                    for (int st = 0; st < nst; ++st) {
                        const int blk = (gb + st) % kR;
                        const int ph = pa + 2 * st;
                        if (st == 0) {
                            wx0 = (colok && pw0 <= qb) ? s_wx[ts][gid][pw0 & 7] : 0.f;
                            wx1 = (colok && pw1 <= qb) ? s_wx[ts][gid][pw1 & 7] : 0.f;
                            const float u0 = wx0 * sbase, u1 = wx1 * sbase;
                            b0 = e0 >= 0 ? u0 + wx0 : u0;
                            b1 = e0 >= 0 ? u1 + wx1 : u1;
                            i0 = pin0 ? u0 : b0;
                            i1 = pin1 ? u1 : b1;
                            if (av_row && tq == 3) {         // the lane holding K slot 7: rowsum * colsum
                                const float cs = colok ? s_wx[ts][gid][kTabW - 1] : 0.f;
                                w70 = rok0 ? s_wy[ts][r0][kTabW - 1] * cs : 0.f;
                                w71 = rok1 ? s_wy[ts][r1][kTabW - 1] * cs : 0.f;
                            }
                        }
                        const bool hb = ph + 1 <= pb;
                        const int phb = hb ? ph + 1 : ph;
                        const float wya0 = rok0 ? s_wy[ts][r0][ph] : 0.f;
                        const float wya1 = rok1 ? s_wy[ts][r1][ph] : 0.f;
                        const float wyb0 = (hb && rok0) ? s_wy[ts][r0][phb] : 0.f;
                        const float wyb1 = (hb && rok1) ? s_wy[ts][r1][phb] : 0.f;
                        if (av_fma) {
                            rs0 += wya0 + wyb0;
                            rs1 += wya1 + wyb1;
                        }
                        const bool ina = e0 > 0 && ph >= e0 && ph < P - e0;
                        const bool inb = e0 > 0 && phb >= e0 && phb < P - e0;
                        const float xa0 = ina ? i0 : b0, xa1 = ina ? i1 : b1;
                        const float xb0 = inb ? i0 : b0, xb1 = inb ? i1 : b1;
                        float h0 = wya0 * xa1, h1 = wya1 * xa1;
                        if (st == 0 && av_row && tq == 3) { h0 = w70; h1 = w71; }
                        uint32_t a[4];
                        a[0] = pack_bf16x2(wya0 * xa0, h0);
                        a[1] = pack_bf16x2(wya1 * xa0, h1);
                        a[2] = pack_bf16x2(wyb0 * xb0, wyb0 * xb1);
                        a[3] = pack_bf16x2(wyb1 * xb0, wyb1 * xb1);
                        if (__any_sync(0xffffffffu, (a[0] | a[1] | a[2] | a[3]) != 0u)) {
                            const uint32_t base = ring_u32 + (uint32_t)blk * kMmaBlockBytes + ld_off;
#pragma unroll
                            for (int jj = 0; jj < kNT / 2; ++jj) {
                                uint32_t bf[4];
                                ldmatrix_x4_trans(base + jj * 32u, bf);
                                mma_bf16_16816(acc[2 * jj], a, bf[0], bf[1]);
                                mma_bf16_16816(acc[2 * jj + 1], a, bf[2], bf[3]);
                            }
                        }
                        if (st == nst - 1 && av_fma) {
                            float cs = wx0 + wx1;             // column sum over the staged bin columns
                            cs += __shfl_xor_sync(0xffffffffu, cs, 1);
                            cs += __shfl_xor_sync(0xffffffffu, cs, 2);
                            const float w0 = rs0 * cs, w1 = rs1 * cs;
                            const float* av = &s_av[ts][cg * kCW + 2 * tq];
#pragma unroll
                            for (int j = 0; j < kNT; ++j) {
                                const float2 v = *reinterpret_cast<const float2*>(av + j * 8);
                                acc[j][0] = fmaf(w0, v.x, acc[j][0]);
                                acc[j][1] = fmaf(w0, v.y, acc[j][1]);
                                acc[j][2] = fmaf(w1, v.x, acc[j][2]);
                                acc[j][3] = fmaf(w1, v.y, acc[j][3]);
                            }
                        }
                    }
                    __syncwarp();
                }
                if (lane == 0)
                    for (int st = 0; st < nst; ++st) mbar_arrive(&s_empty[(gb + st) % kR]);
                gb += (unsigned)nst;
                ++g;
            }
        }
        if (c_begin < total_chunks) __syncthreads();      // the record table is rewritten by the next batch
    }

    // ---- every dX element of the tile is written once: accumulators -> staging -> 128-bit stores
    __syncthreads();                              // all copies consumed: the ring is free
    float* stage = reinterpret_cast<float*>(bwd_smem);      // [64 px][kStagePitch]
    if (!producer && cg_on) {
#pragma unroll
        for (int j = 0; j < kNT; ++j) {
            const int c = cg * kCW + j * 8 + 2 * tq;
            *reinterpret_cast<float2*>(stage + (r0 * kTile + gid) * kStagePitch + c) =
                make_float2(acc[j][0], acc[j][1]);
            *reinterpret_cast<float2*>(stage + (r1 * kTile + gid) * kStagePitch + c) =
                make_float2(acc[j][2], acc[j][3]);
        }
    }
    __syncthreads();
    if (p.trace && tid == 0) {
        unsigned long long t_end;
        unsigned smid;
        asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t_end));
        asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
        unsigned long long* rec = p.trace + (size_t)blockIdx.x * 6;
        rec[0] = t_start; rec[1] = t_end; rec[2] = g; rec[3] = gb; rec[4] = smid; rec[5] = (unsigned)l;
    }
    if (!producer && lane * 8 < C) {
#pragma unroll
        for (int i = 0; i < 64 / kW; ++i) {
            const int px = i * kW + warp;
            const int row = row0 + (px >> 3), col = col0 + (px & 7);
            if (row < H && col < W) {
                float v[8];
                ld_smem8<float>(stage + px * kStagePitch + lane * 8, v);
                if (!p.nchw) {
                    Vec8<TDx, false>::store(dx + (((size_t)b * H + row) * W + col) * C, lane, C, v);
                } else {
#pragma unroll
                    for (int e = 0; e < 8; ++e)
                        st_elem(dx + (((size_t)b * C + lane * 8 + e) * H + row) * W + col, v[e]);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// tensor-pipe forward (level-assigned extraction, bf16 features)
// ------------------------------------------------------------------------------------------
// The scalar kernels above spend ~19 instructions per (pixel, 8 channels) visit and are bound by
// instruction issue and by the dependent shared-load -> fma chains of few warps (ncu, persistent
// kernel: consumers wait for data 54 % of the time although the copy warp is idle 2/3 of it,
// because every strip stays in the ring for the ~3000 clk its reduction takes).  Here the x
// reduction of a footprint row is a small matrix product on the tensor pipe,
//     T[c, pw] = sum_px F[y, px, c] * wx[px, pw]        (M = 16 channels, N = 8 bins, K = 16 px)
// and the y reduction out[ph][pw][c] += wy[y][ph] * T[c, pw] stays on the accumulator fragments.
//   * task = up to kMfSeg = 4 consecutive bin rows of one RoI (two tasks per RoI at pooled 7: the
//     RoI-wide work - descriptor, x-table, B fragments - is paid twice per RoI instead of once per
//     bin row), numbered segment-major so that every CTA sees both segment sizes; tasks alternate between kMfGroups
//     groups of 4 consumer + 4 producer warps; inside a group warp pair cc owns channels
//     [64 cc, 64 cc + 64) and streams them through its OWN ring of 4 KB tiles - no cross-pair barrier.
//   * a tile = 32 pixels of one footprint row x 64 channels, fetched by ONE cp.async.bulk.tensor
//     (UTMALDG; 4-D map (C, W, H, B) of the level, box 64 x 32 x 1 x 1, 128-byte swizzle, zero
//     fill right of the image) - the swizzle makes the transposing ldmatrix of the A fragments
//     bank-conflict free, which plain row copies (pixels 512 B apart) cannot be.
//   * producer 0 of a group fetches the plans of 32 tasks at once (one per lane), publishes each as a
//     128-byte descriptor a few tasks AHEAD of the tiles and bulk-copies the task's slices of the
//     axis-weight tables behind it.
//   * weights enter the MMA as bf16 (2^-9 relative, inside the 2e-2 bf16 gate, like the backward
//     gather); with fp32 OUTPUT they enter as bf16 hi + lo pairs (two MMAs, ~2^-17 relative).
//   * footprints wider than 64 px or tasks spanning more than 64 rows (never a level-assigned RoI
//     of an 800x1333 image) are reduced straight from global memory by the consumer warps.
#ifndef HTD_MF_TILES                             // tuning builds (tools/fwd_variants.py) override these
#define HTD_MF_TILES 5
#endif
#ifndef HTD_MF_DESC
#define HTD_MF_DESC 4
#endif
#ifndef HTD_MF_AHEAD
#define HTD_MF_AHEAD 2
#endif
#ifndef HTD_MF_SEG
#define HTD_MF_SEG 4
#endif
constexpr int kMfGroups = 2;
constexpr int kMfTiles = HTD_MF_TILES;           // 4 KB tiles in flight per (group, channel chunk)
constexpr int kMfDesc = HTD_MF_DESC;             // task descriptors in flight per group
constexpr int kMfAhead = HTD_MF_AHEAD;           // descriptors published ahead of the tile issue
constexpr int kMfSeg = HTD_MF_SEG;               // bin rows per task
constexpr int kMfMaxRows = 64, kMfMaxPx = 64;
constexpr int kMfTileBytes = 32 * 128;
struct MfUnit {                                  // 32 ints: moved by the producer's lanes, one int each
    const void* gsrc;                            // pixel (row y0, column x0) of the footprint (fallback path)
    const float* wy_src;                         // the task's rows / the RoI's columns in the tables
    const float* wx_src;
    long long out_off;                           // output element offset of bin (ph0, 0)
    int live, fallback, b, l, nrows, nxs, nchunk, ph0, nph, x0, y0, W;
    short yoff[4], nyb[4];                       // per bin row: first footprint row (from y0), rows
    short dx0[HTD_MAX_POOLED - 1], nx[HTD_MAX_POOLED - 1];    // per bin column (fallback path)
    int wcls;                                    // box width class of chunk 0 (bits 0-1) and 1 (bits 2-3)
};
static_assert(sizeof(MfUnit) == 128 && kMfSeg <= 4, "MfUnit is moved as 32 ints");
struct MfSlot {
    MfUnit u;
    float wy[kMfMaxRows][kTabW];
    float wx[kMfMaxPx][kTabW];
};
// box width classes of a tile: 16 / 24 / 32 pixels (footprints of level-assigned RoIs are 15 - 30
// px wide: always fetching 32 doubled the L2 -> SM traffic, which is what bounds this kernel)
constexpr int kMfWidths = 3;
struct FwdMaps {
    CUtensorMap m[HTD_MAX_LEVELS * kMfWidths];
};
constexpr int kMfSmem = kMfGroups * 4 * kMfTiles * kMfTileBytes + kMfGroups * kMfDesc * (int)sizeof(MfSlot) +
                        kMfGroups * 32 * (int)sizeof(MfUnit) +
                        (kMfGroups * 4 * kMfTiles * 2 + kMfGroups * kMfDesc * 2) * 8 + 1024;
static_assert(kMfSmem <= 232448, "shared memory per CTA");

template <typename TOut>
__global__ void __launch_bounds__(kMfGroups * 8 * 32, 1)
    roi_align_fwd_mma_kernel(const __grid_constant__ FwdMaps maps, const FwdParams p, int tasks) {
    constexpr bool kLo = sizeof(TOut) == 4;              // fp32 output: weights as bf16 hi + lo
    extern __shared__ uint8_t mf_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(mf_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int grp = warp >> 3, role = warp & 7;
    const int cc = role & 3;                             // channel chunk of the warp's pair
    const bool producer = role >= 4;
    const int nact = p.C >> 6;                           // active pairs per group (C % 64 == 0)
    uint8_t* tiles = smem + (size_t)((grp * 4 + cc) * kMfTiles) * kMfTileBytes;
    MfSlot* slots = reinterpret_cast<MfSlot*>(smem + kMfGroups * 4 * kMfTiles * kMfTileBytes) + grp * kMfDesc;
    MfUnit* plan = reinterpret_cast<MfUnit*>(smem + kMfGroups * 4 * kMfTiles * kMfTileBytes +
                                             kMfGroups * kMfDesc * sizeof(MfSlot)) + grp * 32;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kMfGroups * 4 * kMfTiles * kMfTileBytes +
                                                 kMfGroups * kMfDesc * sizeof(MfSlot) +
                                                 kMfGroups * 32 * sizeof(MfUnit));
    uint64_t* full_bar = bars + (size_t)((grp * 4 + cc) * 2) * kMfTiles;
    uint64_t* empty_bar = full_bar + kMfTiles;
    uint64_t* dfull = bars + kMfGroups * 4 * kMfTiles * 2 + grp * 2 * kMfDesc;
    uint64_t* dempty = dfull + kMfDesc;
    const int P = p.P, PP = P * P;
    const int nseg = (P + kMfSeg - 1) / kMfSeg;          // tasks per RoI
    const int grid = (int)gridDim.x;
    if (threadIdx.x == 0) {
        for (int i = 0; i < kMfGroups * 4 * kMfTiles * 2; ++i) mbar_init(bars + i, 1);
        for (int g = 0; g < kMfGroups; ++g)
            for (int q = 0; q < kMfDesc; ++q) {
                mbar_init(bars + kMfGroups * 4 * kMfTiles * 2 + g * 2 * kMfDesc + q, 1);
                mbar_init(bars + kMfGroups * 4 * kMfTiles * 2 + g * 2 * kMfDesc + kMfDesc + q,
                          2 * nact - 1);                 // consumers + the producers that only read
            }
        fence_mbar_init();
    }
    {   // narrow boxes leave the tail of a tile untouched: no stale NaN patterns under zero weights
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < kMfGroups * 4 * kMfTiles * kMfTileBytes / 16; i += blockDim.x)
            z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async_smem();
    }
    __syncthreads();
    if (cc >= nact) return;
    // tasks of this CTA: i = 0, 1, ... -> task blockIdx.x + i * grid; group g takes i = g (mod groups)
    const int total_i = (int)blockIdx.x < tasks ? (tasks - (int)blockIdx.x + grid - 1) / grid : 0;
    const int n_g = total_i > grp ? (total_i - grp + kMfGroups - 1) / kMfGroups : 0;
    int t = 0;                                           // tiles of the pair so far

    if (producer) {
        int pub = 0;                                     // descriptors published so far (producer 0)
        for (int i0 = 0; i0 < n_g; i0 += 32 - kMfAhead) {
            // producer 0 keeps a window of 32 staged plans starting at i0; the window moves by
            // 32 - kMfAhead so that the tasks published ahead are still staged
            if (cc == 0) {
                const int ii = i0 + lane;
                MfUnit pl;
                pl.gsrc = nullptr; pl.wy_src = nullptr; pl.wx_src = nullptr; pl.out_off = 0;
                pl.live = 0; pl.fallback = 0; pl.b = -1; pl.l = 0; pl.nrows = 0; pl.nxs = 0; pl.nchunk = 0;
                pl.ph0 = 0; pl.nph = 0; pl.x0 = 0; pl.y0 = 0; pl.W = 1; pl.wcls = 0;
#pragma unroll
                for (int j = 0; j < 4; ++j) { pl.yoff[j] = 0; pl.nyb[j] = 0; }
#pragma unroll
                for (int j = 0; j < HTD_MAX_POOLED - 1; ++j) { pl.dx0[j] = 0; pl.nx[j] = 0; }
                if (ii < n_g) {
                    // segment-major task numbers: a CTA's tasks tk = blockIdx.x + i * grid then run
                    // through all segments (RoI-major numbering gave CTA b the SAME segment b % nseg of
                    // every RoI whenever nseg divides the grid - and the last segment has one bin row)
                    const int tk = (int)blockIdx.x + (ii * kMfGroups + grp) * grid;
                    const int sg = tk / p.K;
                    const int k = tk - sg * p.K;
                    const int ph0 = sg * kMfSeg;
                    const int nph = min(kMfSeg, P - ph0);
                    const int l = p.roi_level[k];
                    const int b = (int)p.rois[(size_t)k * 5];
                    pl.b = b;
                    pl.ph0 = ph0;
                    pl.nph = nph;
                    pl.out_off = ((long long)k * PP + (long long)ph0 * P) * p.C;
                    if ((b >= 0 && b < p.B) && (l >= 0 && l < p.L)) {
                        const size_t e = (size_t)l * p.K + k;
                        const int4 box = p.boxes[e];
                        const int4* rq = reinterpret_cast<const int4*>(p.ranges + e * kRangeInts);
                        const int4 ya = rq[0], yb = rq[1], yc = rq[2], yd = rq[3];
                        const int4 lo0 = rq[4], lo1 = rq[5], hi0 = rq[6], hi1 = rq[7];
                        const size_t off = (size_t)p.offsets[e];
                        const int ylo[8] = {ya.x, ya.y, ya.z, ya.w, yb.x, yb.y, yb.z, yb.w};
                        const int yhi[8] = {yc.x, yc.y, yc.z, yc.w, yd.x, yd.y, yd.z, yd.w};
                        // rows of the task: from the first non-empty bin row's first row on
                        int y0 = 0x7fffffff, y1 = -1;
#pragma unroll
                        for (int j = 0; j < 8; ++j)
                            if (j >= ph0 && j < ph0 + nph && yhi[j] >= ylo[j]) {
                                y0 = min(y0, ylo[j]);
                                y1 = max(y1, yhi[j]);
                            }
                        const int nxs = box.w - box.z + 1;
                        if (box.y >= box.x && box.w >= box.z && y1 >= y0 && nxs > 0) {
                            const int H = p.lv[l].H, W = p.lv[l].W;
                            const int cx0[8] = {lo0.x, lo0.y, lo0.z, lo0.w, lo1.x, lo1.y, lo1.z, lo1.w};
                            const int cx1[8] = {hi0.x, hi0.y, hi0.z, hi0.w, hi1.x, hi1.y, hi1.z, hi1.w};
#pragma unroll
                            for (int j = 0; j < HTD_MAX_POOLED - 1; ++j) {
                                pl.dx0[j] = (short)(cx0[j] - box.z);
                                pl.nx[j] = (short)(j < P ? cx1[j] - cx0[j] + 1 : 0);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j)
                                if (j >= ph0 && j < ph0 + nph && yhi[j] >= ylo[j]) {
                                    pl.yoff[j - ph0] = (short)(ylo[j] - y0);
                                    pl.nyb[j - ph0] = (short)(yhi[j] - ylo[j] + 1);
                                }
                            pl.live = 1; pl.l = l; pl.nrows = y1 - y0 + 1; pl.nxs = nxs; pl.W = W;
                            pl.fallback = (nxs > kMfMaxPx || pl.nrows > kMfMaxRows) ? 1 : 0;
                            pl.nchunk = (nxs + 31) >> 5;
                            {
                                const int w0 = min(nxs, 32), w1 = nxs - 32;     // pixels of chunk 0 / 1
                                const int c0 = w0 <= 16 ? 0 : w0 <= 24 ? 1 : 2;
                                const int c1 = w1 <= 16 ? 0 : w1 <= 24 ? 1 : 2;
                                pl.wcls = c0 | (c1 << 2);
                            }
                            pl.x0 = box.z; pl.y0 = y0;
                            pl.gsrc = static_cast<const __nv_bfloat16*>(p.lv[l].data) +
                                      (((size_t)b * H + y0) * W + box.z) * p.C;
                            pl.wy_src = p.weights + (off + (size_t)(y0 - box.x)) * kTabW;
                            pl.wx_src = p.weights + (off + (size_t)(box.y - box.x + 1)) * kTabW;
                        }
                    }
                }
                __syncwarp();
                plan[lane] = pl;
                __syncwarp();
            }
            const int nb = min(32 - kMfAhead, n_g - i0);
            for (int j = 0; j < nb; ++j) {
                const int ii = i0 + j;
                const int q = ii % kMfDesc;
                MfSlot* sl = slots + q;
                const MfUnit* d;
                if (cc == 0) {
                    // publish the descriptors (and table slices) of the tasks up to ii + kMfAhead
                    const int upto = min(min(ii + kMfAhead, n_g - 1), i0 + 31);
                    for (; pub <= upto; ++pub) {
                        const int pq = pub % kMfDesc;
                        MfSlot* ps = slots + pq;
                        const MfUnit* pd = plan + (pub - i0);
                        mbar_wait(dempty + pq, (uint32_t)((pub / kMfDesc) & 1) ^ 1u);
                        reinterpret_cast<int*>(&ps->u)[lane] = reinterpret_cast<const int*>(pd)[lane];
                        __syncwarp();
                        if (lane == 0) {
                            if (pd->live && !pd->fallback) {
                                mbar_expect_tx(dfull + pq, (uint32_t)((pd->nrows + pd->nxs) * kTabW * 4));
                                bulk_g2s(&ps->wy[0][0], pd->wy_src, (uint32_t)(pd->nrows * kTabW * 4), dfull + pq);
                                bulk_g2s(&ps->wx[0][0], pd->wx_src, (uint32_t)(pd->nxs * kTabW * 4), dfull + pq);
                            } else {
                                mbar_arrive(dfull + pq);
                            }
                        }
                    }
                    d = plan + j;
                } else {
                    d = &sl->u;
                    mbar_wait(dfull + q, (uint32_t)((ii / kMfDesc) & 1));
                }
                if (d->live && !d->fallback) {
                    const int nph = d->nph, nchunk = d->nchunk, x0 = d->x0, y0 = d->y0, b = d->b;
                    const int wcls = d->wcls;
                    for (int pi = 0; pi < nph; ++pi) {
                        const int ya = y0 + d->yoff[pi], ny = d->nyb[pi];
                        for (int r = 0; r < ny; ++r)
                            for (int xc = 0; xc < nchunk; ++xc, ++t) {
                                const int s = t % kMfTiles;
                                mbar_wait(empty_bar + s, (uint32_t)((t / kMfTiles) & 1) ^ 1u);
                                if (lane == 0) {
                                    const int wc = (wcls >> (2 * xc)) & 3;
                                    mbar_expect_tx(full_bar + s, (uint32_t)((16 + 8 * wc) * 128));
                                    tc::tma_load_4d(&maps.m[d->l * kMfWidths + wc], full_bar + s,
                                                    tiles + s * kMfTileBytes, cc * 64, x0 + xc * 32, ya + r, b);
                                }
                            }
                    }
                }
                if (cc != 0) {
                    __syncwarp();
                    if (lane == 0) mbar_arrive(dempty + q);
                }
            }
        }
        return;
    }

    // ===== consumers =====
    const int g8 = lane >> 2, tq = lane & 3;             // mma fragment coordinates
    // ldmatrix lane address inside a tile: matrix m = lane >> 3 -> channel half (m & 1) of the
    // 16-channel block, pixel half (m >> 1); row i = lane & 7 is the pixel -> swizzle phase i
    const int lm = lane >> 3, li = lane & 7;
    const uint32_t ld_row = (uint32_t)(((lm >> 1) * 8 + li) * 128);
    uint32_t ld_unit[4];
#pragma unroll
    for (int mt = 0; mt < 4; ++mt) ld_unit[mt] = (uint32_t)(((2 * mt + (lm & 1)) ^ li) << 4);
    const uint32_t tiles_u32 = smem_u32(tiles);
    for (int ii = 0; ii < n_g; ++ii) {
        const int q = ii % kMfDesc;
        const MfSlot* sl = slots + q;
        mbar_wait(dfull + q, (uint32_t)((ii / kMfDesc) & 1));
        const MfUnit* d = &sl->u;
        const int b = d->b, ph0 = d->ph0, nph = d->nph;
        const bool bias_on = p.bias && b >= 0 && b < p.B;
        TOut* obase = static_cast<TOut*>(p.out) + d->out_off + cc * 64;
        if (d->live && d->fallback) {
            // scalar reduction straight from global memory: a lane = 2 channels of the chunk
            const __nv_bfloat16* gb = static_cast<const __nv_bfloat16*>(d->gsrc) + cc * 64 + lane * 2;
            const int W = d->W;
            for (int pi = 0; pi < nph; ++pi) {
                const int ph = ph0 + pi, ny = d->nyb[pi], yo = d->yoff[pi];
                for (int pw = 0; pw < P; ++pw) {
                    float a0 = 0.f, a1 = 0.f;
                    const int nx = d->nx[pw], dx0 = d->dx0[pw];
                    for (int y = yo; y < yo + ny; ++y) {
                        const float wyv = __ldg(d->wy_src + (size_t)y * kTabW + ph);
                        if (wyv == 0.f) continue;
                        for (int x = 0; x < nx; ++x) {
                            const float w = wyv * __ldg(d->wx_src + (size_t)(dx0 + x) * kTabW + pw);
                            const uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(
                                gb + ((size_t)y * W + dx0 + x) * p.C));
                            a0 = fmaf(w, __uint_as_float(v << 16), a0);
                            a1 = fmaf(w, __uint_as_float(v & 0xffff0000u), a1);
                        }
                    }
                    if (bias_on) {
                        a0 += __ldg(p.bias + (size_t)b * p.C + cc * 64 + lane * 2);
                        a1 += __ldg(p.bias + (size_t)b * p.C + cc * 64 + lane * 2 + 1);
                    }
                    st_elem(obase + (size_t)(pi * P + pw) * p.C + lane * 2, a0);
                    st_elem(obase + (size_t)(pi * P + pw) * p.C + lane * 2 + 1, a1);
                }
            }
        } else {
            float bv[8];                                 // SFA bias of the lane's 8 channels
#pragma unroll
            for (int e = 0; e < 8; ++e)
                bv[e] = bias_on ? __ldg(p.bias + (size_t)b * p.C + cc * 64 + (e >> 1) * 16 + (e & 1) * 8 + g8)
                                : 0.f;
            const int nxs = d->nxs, nchunk = d->live ? d->nchunk : 0;
            const int wcls = d->wcls;
            // B fragments = wx[px][bin g8] of the lane's k rows, as bf16 (hi + lo for fp32 output)
            uint32_t bhi[2][2][2], blo[2][2][2];
#pragma unroll
            for (int xc = 0; xc < 2; ++xc)
#pragma unroll
                for (int ks = 0; ks < 2; ++ks)
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        const int px = xc * 32 + ks * 16 + h * 8 + 2 * tq;
                        float w0 = 0.f, w1 = 0.f;
                        if (xc < nchunk && g8 < P) {
                            if (px < nxs) w0 = sl->wx[px][g8];
                            if (px + 1 < nxs) w1 = sl->wx[px + 1][g8];
                        }
                        const uint32_t hi = pack_bf16x2(w0, w1);
                        bhi[xc][ks][h] = hi;
                        blo[xc][ks][h] = kLo ? pack_bf16x2(w0 - __uint_as_float(hi << 16),
                                                            w1 - __uint_as_float(hi & 0xffff0000u))
                                             : 0u;
                    }
            for (int pi = 0; pi < nph; ++pi) {
                const int ph = ph0 + pi;
                const int ny = d->live ? d->nyb[pi] : 0, yo = d->yoff[pi];
                float out[4][4];
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) out[mt][e] = 0.f;
                for (int r = 0; r < ny; ++r) {
                    float tacc[4][4];
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) tacc[mt][e] = 0.f;
                    const float wyv = sl->wy[yo + r][ph];
#pragma unroll
                    for (int xc = 0; xc < 2; ++xc) {
                        if (xc < nchunk) {
                            const int s = t % kMfTiles;
                            mbar_wait(full_bar + s, (uint32_t)((t / kMfTiles) & 1));
                            const uint32_t tb = tiles_u32 + (uint32_t)(s * kMfTileBytes) + ld_row;
#pragma unroll
                            for (int ks = 0; ks < 2; ++ks) {
                                if (ks == 1 && ((wcls >> (2 * xc)) & 3) == 0) break;   // 16-px box
#pragma unroll
                                for (int mt = 0; mt < 4; ++mt) {
                                    uint32_t a[4];
                                    ldmatrix_x4_trans(tb + (uint32_t)(ks * 2048) + ld_unit[mt], a);
                                    mma_bf16_16816(tacc[mt], a, bhi[xc][ks][0], bhi[xc][ks][1]);
                                    if (kLo) mma_bf16_16816(tacc[mt], a, blo[xc][ks][0], blo[xc][ks][1]);
                                }
                            }
                            __syncwarp();
                            if (lane == 0) mbar_arrive(empty_bar + s);
                            ++t;
                        }
                    }
#pragma unroll
                    for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                        for (int e = 0; e < 4; ++e) out[mt][e] = fmaf(wyv, tacc[mt][e], out[mt][e]);
                }
                // fragment element e of block mt: channel mt*16 + g8 + 8*(e >> 1), bin 2*tq + (e & 1)
                TOut* orow = obase + (size_t)pi * P * p.C;
#pragma unroll
                for (int mt = 0; mt < 4; ++mt)
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const int bin = 2 * tq + (e & 1);
                        if (bin < P)
                            st_elem(orow + (size_t)bin * p.C + mt * 16 + (e >> 1) * 8 + g8,
                                    out[mt][e] + bv[mt * 2 + (e >> 1)]);
                    }
            }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(dempty + q);
    }
}

static int fill_levels(LevelDev* dst, const HtdLevel* src, int L, const char* who) {
    HTD_CHECK_ARG(src != nullptr && L >= 1 && L <= HTD_MAX_LEVELS, "%s: need 1..%d levels, got %d",
                  who, HTD_MAX_LEVELS, L);
    for (int i = 0; i < L; ++i) {
        HTD_CHECK_ARG(src[i].data != nullptr && src[i].H > 0 && src[i].W > 0 &&
                          src[i].spatial_scale > 0.f,
                      "%s: level %d is malformed (data=%p H=%d W=%d scale=%g)", who, i, src[i].data,
                      src[i].H, src[i].W, (double)src[i].spatial_scale);
        dst[i].data = src[i].data;
        dst[i].H = src[i].H;
        dst[i].W = src[i].W;
        dst[i].scale = src[i].spatial_scale;
    }
    return HTD_OK;
}

}  // namespace htd

using namespace htd;

extern "C" {

int htd_abi_version(void) { return HTD_ABI_VERSION; }
#ifdef HTD_DEBUG_HOOKS
void htd_debug_set_bwd_trace(unsigned long long* records) { g_bwd_trace = records; }
void htd_debug_set_bwd_variant(int variant) { g_bwd_variant = variant; }
#endif
const char* htd_last_error(void) { return g_err; }

int htd_level_assign(const float* rois, int K, int num_levels, float finest_scale,
                     int32_t* levels, htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && num_levels >= 1 && finest_scale > 0.f,
                  "htd_level_assign: bad sizes K=%d levels=%d", K, num_levels);
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(rois && levels, "htd_level_assign: null pointer");
    level_assign_kernel<<<(K + 255) / 256, 256, 0, (cudaStream_t)stream>>>(rois, K, num_levels,
                                                                            finest_scale, levels);
    HTD_CHECK_LAUNCH("htd_level_assign");
    return HTD_OK;
}

int htd_roi_footprints(const HtdLevel* levels, int L, int B, const float* rois, int K,
                       const int32_t* roi_level, int pooled, int sampling_ratio, int32_t* boxes,
                       unsigned long long* pixel_count, htd_stream_t stream) {
    FootParams p;
    int rc = fill_levels(p.lv, levels, L, "htd_roi_footprints");
    if (rc) return rc;
    HTD_CHECK_ARG(K >= 0 && pooled >= 1 && pooled <= HTD_MAX_POOLED && B >= 1,
                  "htd_roi_footprints: bad sizes K=%d pooled=%d B=%d", K, pooled, B);
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(rois && boxes, "htd_roi_footprints: null pointer");
    p.L = L; p.B = B; p.K = K; p.P = pooled; p.sr = sampling_ratio;
    p.rois = rois; p.roi_level = roi_level; p.boxes = reinterpret_cast<int4*>(boxes);
    p.pixel_count = pixel_count;
    const int n = L * K;
    footprint_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(p);
    HTD_CHECK_LAUNCH("htd_roi_footprints");
    return HTD_OK;
}

long long htd_roi_plan_rows_bound(const HtdLevel* levels, int L, int K, int single_level) {
    if (!levels || L < 1 || L > HTD_MAX_LEVELS || K < 0) return -1;
    long long sum = 0, mx = 0;
    for (int l = 0; l < L; ++l) {
        const long long e = (long long)levels[l].H + levels[l].W;
        sum += e;
        if (e > mx) mx = e;
    }
    return (long long)K * (single_level ? mx : sum) + 1;
}

int htd_roi_plan(const HtdLevel* levels, int L, int B, const float* rois, int K,
                 const int32_t* roi_level, int pooled, int sampling_ratio, int32_t* boxes,
                 int32_t* offsets, int32_t* ranges, float* weights, long long rows_cap,
                 unsigned long long* pixel_count, htd_stream_t stream) {
    int rc;
    if (pixel_count == nullptr && K > 0) {
        // one launch: boxes, offsets (fixed stride), ranges and tables
        HTD_CHECK_ARG(levels && L >= 1 && L <= HTD_MAX_LEVELS && B >= 1 && rois && boxes && offsets &&
                          ranges && weights && pooled >= 1 && pooled <= HTD_MAX_POOLED,
                      "htd_roi_plan: bad arguments");
        const long long need1 = htd_roi_plan_rows_bound(levels, L, K, roi_level != nullptr);
        HTD_CHECK_ARG(rows_cap >= need1, "htd_roi_plan: weight table capacity %lld rows < bound %lld",
                      rows_cap, need1);
        HTD_CHECK_ARG(need1 < 2147483647LL && (long long)L * K < 2147483647LL, "htd_roi_plan: table too large");
        PlanParams q;
        rc = fill_levels(q.lv, levels, L, "htd_roi_plan");
        if (rc) return rc;
        q.L = L; q.B = B; q.K = K; q.P = pooled; q.sr = sampling_ratio;
        int acc = 0, mx = 0;
        for (int l = 0; l < L; ++l) {
            q.base[l] = acc;
            q.ext[l] = levels[l].H + levels[l].W;
            acc += q.ext[l];
            if (q.ext[l] > mx) mx = q.ext[l];
        }
        q.ext_max = mx;
        q.rois = rois; q.roi_level = roi_level; q.boxes = reinterpret_cast<int4*>(boxes);
        q.offsets = offsets; q.ranges = ranges; q.weights = weights;
        roi_plan_fused_kernel<<<roi_level ? K : L * K, 128, 0, (cudaStream_t)stream>>>(q);
        HTD_CHECK_LAUNCH("htd_roi_plan(fused)");
        return HTD_OK;
    }
    rc = htd_roi_footprints(levels, L, B, rois, K, roi_level, pooled, sampling_ratio, boxes,
                            pixel_count, stream);
    if (rc) return rc;
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(offsets && ranges && weights, "htd_roi_plan: null pointer");
    const long long need = htd_roi_plan_rows_bound(levels, L, K, roi_level != nullptr);
    HTD_CHECK_ARG(rows_cap >= need, "htd_roi_plan: weight table capacity %lld rows < bound %lld",
                  rows_cap, need);
    HTD_CHECK_ARG(need < 2147483647LL, "htd_roi_plan: table too large");
    cudaStream_t st = (cudaStream_t)stream;
    const int n = L * K;
    extent_scan_kernel<<<1, 1024, 0, st>>>(reinterpret_cast<const int4*>(boxes), n, offsets);
    HTD_CHECK_LAUNCH("htd_roi_plan(scan)");
    TableParams p;
    rc = fill_levels(p.lv, levels, L, "htd_roi_plan");
    if (rc) return rc;
    p.L = L; p.K = K; p.P = pooled; p.sr = sampling_ratio;
    p.rois = rois; p.boxes = reinterpret_cast<const int4*>(boxes); p.offsets = offsets;
    p.ranges = ranges; p.weights = weights;
    weight_table_kernel<<<n, 128, 0, st>>>(p);
    HTD_CHECK_LAUNCH("htd_roi_plan(tables)");
    return HTD_OK;
}

int htd_roi_align_fwd(const HtdLevel* levels, int L, int B, int C, int in_dtype,
                      const float* rois, int K, const int32_t* roi_level, int pooled,
                      const int32_t* boxes, const int32_t* offsets, const int32_t* ranges,
                      const float* weights, const float* bias, void* out, int out_dtype,
                      htd_stream_t stream) {
    FwdParams p;
    int rc = fill_levels(p.lv, levels, L, "htd_roi_align_fwd");
    if (rc) return rc;
    HTD_CHECK_ARG(K >= 0 && B >= 1 && pooled >= 1 && pooled <= HTD_MAX_POOLED,
                  "htd_roi_align_fwd: bad sizes K=%d B=%d pooled=%d", K, B, pooled);
    HTD_CHECK_ARG(C >= 8 && C % 8 == 0, "htd_roi_align_fwd: C=%d must be a positive multiple of 8",
                  C);
    HTD_CHECK_ARG((in_dtype == HTD_F32 || in_dtype == HTD_BF16) &&
                      (out_dtype == HTD_F32 || out_dtype == HTD_BF16),
                  "htd_roi_align_fwd: unsupported dtype in=%d out=%d", in_dtype, out_dtype);
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(rois && out && boxes && offsets && ranges && weights,
                  "htd_roi_align_fwd: null pointer");
    p.L = L; p.B = B; p.C = C; p.K = K; p.P = pooled;
    p.rois = rois; p.roi_level = roi_level; p.bias = bias; p.out = out;
    p.boxes = reinterpret_cast<const int4*>(boxes); p.offsets = offsets; p.ranges = ranges;
    p.weights = weights;
    const long long tasks = roi_level ? (long long)K : (long long)K * L;
    const long long blocks = tasks * pooled;
    HTD_CHECK_ARG(blocks < 2147483647LL, "htd_roi_align_fwd: too many bins (%lld)", blocks * pooled);
    dim3 grid((unsigned)blocks), block(pooled * 32);
    cudaStream_t st = (cudaStream_t)stream;
    const int cw = C < 256 ? C : 256;
    // level-assigned extraction (SingleRoIExtractor): every strip is small - persistent CTAs with a
    // producer warp that prefetches several work units ahead (HTD_FWD_KERNEL=cta keeps the
    // CTA-per-strip kernel for measurements)
    static int persist_on = -1;
    if (persist_on < 0) {
        const char* ev = hook_env("HTD_FWD_KERNEL");
        persist_on = (ev && (!strcmp(ev, "cta") || !strcmp(ev, "ring"))) ? 0 : 1;
    }
    // bf16 features: the tensor-pipe kernel (HTD_FWD_KERNEL=persist keeps the scalar persistent one)
    static int mma_on = -1;
    if (mma_on < 0) {
        const char* ev = hook_env("HTD_FWD_KERNEL");
        mma_on = (ev && !strcmp(ev, "persist")) ? 0 : 1;
    }
    if (persist_on && mma_on && roi_level != nullptr && pooled < HTD_MAX_POOLED && in_dtype == HTD_BF16 &&
        C % 64 == 0 && C <= 256) {
        FwdMaps maps;
        memset(&maps, 0, sizeof(maps));
        for (int l = 0; l < L; ++l) {
            const cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)levels[l].W, (cuuint64_t)levels[l].H,
                                        (cuuint64_t)B};
            const cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)levels[l].W * C * 2,
                                           (cuuint64_t)levels[l].H * levels[l].W * C * 2};
            for (int wc = 0; wc < kMfWidths; ++wc) {
                const cuuint32_t box[4] = {64u, (cuuint32_t)(16 + 8 * wc), 1u, 1u};
                rc = tc::make_map(&maps.m[l * kMfWidths + wc], levels[l].data, 4, dims, strides, box,
                                  "htd_roi_align_fwd(map)");
                if (rc) return rc;
            }
        }
        const int sms = sm_count();
        const long long mtasks = (long long)K * ((pooled + kMfSeg - 1) / kMfSeg);
        const unsigned mgrid = (unsigned)(mtasks < sms ? mtasks : sms);
        p.region = 0;
        p.strip = 1;
        if (out_dtype == HTD_F32) {
            HTD_SMEM_OPTIN((roi_align_fwd_mma_kernel<float>), kMfSmem, "htd_roi_align_fwd");
            roi_align_fwd_mma_kernel<float><<<mgrid, kMfGroups * 8 * 32, kMfSmem, st>>>(maps, p, (int)mtasks);
        } else {
            HTD_SMEM_OPTIN((roi_align_fwd_mma_kernel<__nv_bfloat16>), kMfSmem, "htd_roi_align_fwd");
            roi_align_fwd_mma_kernel<__nv_bfloat16><<<mgrid, kMfGroups * 8 * 32, kMfSmem, st>>>(maps, p, (int)mtasks);
        }
        HTD_CHECK_LAUNCH("htd_roi_align_fwd(tensor pipe)");
        return HTD_OK;
    }
    if (persist_on && roi_level != nullptr && pooled < HTD_MAX_POOLED && C <= 256) {
        const int sms = sm_count();
        const unsigned pgrid = (unsigned)(blocks < sms ? blocks : sms);
        p.region = 0;
        p.strip = 1;
#define HTD_FWDP_LAUNCH(TI, TO)                                                                   \
    do {                                                                                          \
        HTD_SMEM_OPTIN((roi_align_fwd_persist_kernel<TI, TO>), kPfSmem, "htd_roi_align_fwd");     \
        roi_align_fwd_persist_kernel<TI, TO><<<pgrid, (kPfGroups * pooled + 1) * 32, kPfSmem, st>>>(p, (int)blocks); \
    } while (0)
        if (in_dtype == HTD_F32 && out_dtype == HTD_F32) HTD_FWDP_LAUNCH(float, float);
        else if (in_dtype == HTD_F32 && out_dtype == HTD_BF16) HTD_FWDP_LAUNCH(float, __nv_bfloat16);
        else if (in_dtype == HTD_BF16 && out_dtype == HTD_F32) HTD_FWDP_LAUNCH(__nv_bfloat16, float);
        else HTD_FWDP_LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef HTD_FWDP_LAUNCH
        HTD_CHECK_LAUNCH("htd_roi_align_fwd(persistent)");
        return HTD_OK;
    }
    // staging region: the per-warp rings, or at least kFwdStripBytes for the strip path (three CTAs
    // of 72 KB + barriers per SM); HTD_FWD_KERNEL=ring switches the strip path off (measurements)
    static int strip_on = -1;
    if (strip_on < 0) {
        const char* ev = hook_env("HTD_FWD_KERNEL");
        strip_on = (ev && !strcmp(ev, "ring")) ? 0 : 1;
    }
    size_t region = (size_t)pooled * kFwdStages * kFwdPx * cw * dtype_size(in_dtype);
    if (strip_on && region < (size_t)kFwdStripBytes) region = kFwdStripBytes;
    p.region = (int)region;
    p.strip = strip_on;
    const size_t smem = region + (size_t)pooled * kFwdStages * sizeof(uint64_t);
#define HTD_FWD_LAUNCH(TI, TO)                                                                    \
    do {                                                                                          \
        HTD_SMEM_OPTIN((roi_align_fwd_kernel<TI, TO>), fwd_smem_max((int)sizeof(TI)),             \
                       "htd_roi_align_fwd");                                                      \
        roi_align_fwd_kernel<TI, TO><<<grid, block, smem, st>>>(p);                               \
    } while (0)
    if (in_dtype == HTD_F32 && out_dtype == HTD_F32) HTD_FWD_LAUNCH(float, float);
    else if (in_dtype == HTD_F32 && out_dtype == HTD_BF16) HTD_FWD_LAUNCH(float, __nv_bfloat16);
    else if (in_dtype == HTD_BF16 && out_dtype == HTD_F32) HTD_FWD_LAUNCH(__nv_bfloat16, float);
    else HTD_FWD_LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef HTD_FWD_LAUNCH
    HTD_CHECK_LAUNCH("htd_roi_align_fwd");
    return HTD_OK;
}

// bf16 dY: tensor-pipe contraction per hit (roi_align_bwd_mma_kernel).  HTD_BWD_KERNEL (or
// htd_debug_set_bwd_variant) selects for measurements: "scalar" = the FFMA kernel, "mma1".."mma5" =
// the warp layouts / ring depths listed at the launch below; default mma3.
static int bwd_variant() {
    static int env_variant = -1;
    if (env_variant < 0) {
        const char* ev = hook_env("HTD_BWD_KERNEL");
        env_variant = !ev ? 3 : !strcmp(ev, "scalar") ? 0 : !strcmp(ev, "mma1") ? 1 : !strcmp(ev, "mma2") ? 2 :
                      !strcmp(ev, "mma3") ? 3 : !strcmp(ev, "mma4") ? 4 : !strcmp(ev, "mma5") ? 5 : 3;
    }
    return (g_bwd_variant >= 0 && g_bwd_variant <= 5) ? g_bwd_variant : env_variant;
}

int htd_roi_align_bwd_uses_tensor_pipe(int C, int pooled, int dy_dtype) {
    (void)pooled;
    return (dy_dtype == HTD_BF16 && bwd_variant() != 0 && C % 64 == 0 && C >= 64 && C <= 256) ? 1 : 0;
}

int htd_roi_align_bwd_multi(const HtdLevel* grad_levels, int L, int B, int C, int dx_dtype,
                            int dx_nchw, const HtdBwdSource* sources, int nsrc, int pooled,
                            int dy_dtype, htd_stream_t stream) {
    BwdParams p;
    int rc = fill_levels(p.lv, grad_levels, L, "htd_roi_align_bwd");
    if (rc) return rc;
    HTD_CHECK_ARG(B >= 1 && pooled >= 1 && pooled <= HTD_MAX_POOLED,
                  "htd_roi_align_bwd: bad sizes B=%d pooled=%d", B, pooled);
    HTD_CHECK_ARG(C >= 8 && C % 8 == 0, "htd_roi_align_bwd: C=%d must be a positive multiple of 8",
                  C);
    HTD_CHECK_ARG((dx_dtype == HTD_F32 || dx_dtype == HTD_BF16) &&
                      (dy_dtype == HTD_F32 || dy_dtype == HTD_BF16),
                  "htd_roi_align_bwd: unsupported dtype dx=%d dy=%d", dx_dtype, dy_dtype);
    HTD_CHECK_ARG(nsrc >= 0 && nsrc <= HTD_MAX_BWD_SOURCES && (nsrc == 0 || sources),
                  "htd_roi_align_bwd: need 0..%d sources, got %d", HTD_MAX_BWD_SOURCES, nsrc);
    p.nsrc = 0;
    int any_av_bf16 = 0;
    for (int i = 0; i < nsrc; ++i) {
        const HtdBwdSource& q = sources[i];
        HTD_CHECK_ARG(q.K >= 0, "htd_roi_align_bwd: source %d has K=%d", i, q.K);
        if (q.K == 0) continue;
        HTD_CHECK_ARG(q.rois && q.boxes && q.offsets && q.ranges && q.weights && q.dy,
                      "htd_roi_align_bwd: source %d has a null pointer", i);
        BwdSource& d = p.src[p.nsrc++];
        d.rois = q.rois; d.boxes = reinterpret_cast<const int4*>(q.boxes); d.offsets = q.offsets;
        d.ranges = q.ranges; d.weights = q.weights; d.dy = q.dy; d.scale = q.scale;
        d.addvec = static_cast<const float*>(q.addvec); d.K = q.K; d.dy_per_level = q.dy_per_level;
        d.ring_edge = q.ring_edge;
        d.av_bf16 = (q.addvec != nullptr && q.addvec_dtype == HTD_BF16) ? 1 : 0;
        HTD_CHECK_ARG(q.addvec == nullptr || q.addvec_dtype == HTD_F32 || q.addvec_dtype == HTD_BF16,
                      "htd_roi_align_bwd: source %d has addvec_dtype=%d", i, q.addvec_dtype);
        any_av_bf16 |= d.av_bf16;
    }
    p.L = L; p.B = B; p.C = C; p.P = pooled; p.nchw = dx_nchw ? 1 : 0;
    p.trace = g_bwd_trace;

    long long total = 0;
    for (int l = 0; l < L; ++l) {
        p.tile_start[l] = (int)total;
        total += (long long)B * ((p.lv[l].H + kTile - 1) / kTile) * ((p.lv[l].W + kTile - 1) / kTile);
    }
    for (int l = L; l <= HTD_MAX_LEVELS; ++l) p.tile_start[l] = (int)total;
    HTD_CHECK_ARG(total < 2147483647LL, "htd_roi_align_bwd: too many tiles (%lld)", total);
    dim3 grid((unsigned)total), block(kBwdThreads);
    cudaStream_t st = (cudaStream_t)stream;
    const int cw = C < 256 ? C : 256;
    const size_t smem = (size_t)(dy_dtype == HTD_BF16 ? HTD_BWD_STAGES_BF16 : HTD_BWD_STAGES_F32) *
                        pooled * pooled * cw * dtype_size(dy_dtype);
#define HTD_BWD_LAUNCH(TY, TX)                                                                    \
    do {                                                                                          \
        HTD_SMEM_OPTIN((roi_align_bwd_kernel<TY, TX>),                                            \
                       BwdStages<TY>::value * HTD_MAX_POOLED * HTD_MAX_POOLED * 256 *             \
                           (int)sizeof(TY), "htd_roi_align_bwd");                                 \
        roi_align_bwd_kernel<TY, TX><<<grid, block, smem, st>>>(p);                               \
    } while (0)
    const int variant = bwd_variant();
    const bool mma = htd_roi_align_bwd_uses_tensor_pipe(C, pooled, dy_dtype) != 0;
    HTD_CHECK_ARG(!any_av_bf16 || (mma && pooled < kTabW),
                  "htd_roi_align_bwd: a bf16 addvec needs bf16 dy, pooled < %d, C %% 64 == 0, C <= 256 "
                  "(and HTD_BWD_KERNEL != scalar)", kTabW);
    if (mma) {
#define HTD_BWD_MMA_LAUNCH(TX, S, N, G)                                                           \
    do {                                                                                          \
        HTD_SMEM_OPTIN((roi_align_bwd_mma_kernel<TX, S, N, G>), S * kMmaBlockBytes,               \
                       "htd_roi_align_bwd(mma)");                                                 \
        roi_align_bwd_mma_kernel<TX, S, N, G>                                                     \
            <<<grid, (4 * G + 1) * 32, S * kMmaBlockBytes, st>>>(p);                               \
    } while (0)
#define HTD_BWD_MMA_VARIANT(S, N, G)                                                              \
    do {                                                                                          \
        if (dx_dtype == HTD_F32) HTD_BWD_MMA_LAUNCH(float, S, N, G);                              \
        else HTD_BWD_MMA_LAUNCH(__nv_bfloat16, S, N, G);                                          \
    } while (0)
        switch (variant) {
            case 1: HTD_BWD_MMA_VARIANT(12, 1, 4); break;    // 16 warps x 64 ch, 12 blocks
            case 2: HTD_BWD_MMA_VARIANT(8, 2, 4); break;     // same, 8 blocks, 2 CTAs / SM
            case 4: HTD_BWD_MMA_VARIANT(8, 2, 1); break;     // 4 warps x 256 ch, 8 blocks, 2 CTAs / SM
            case 5: HTD_BWD_MMA_VARIANT(12, 1, 2); break;    // 8 warps x 128 ch, 12 blocks
            default: HTD_BWD_MMA_VARIANT(8, 2, 2); break;    // 8 warps x 128 ch, 8 blocks, 2 CTAs / SM
        }
#undef HTD_BWD_MMA_VARIANT
#undef HTD_BWD_MMA_LAUNCH
        HTD_CHECK_LAUNCH("htd_roi_align_bwd(mma)");
        return HTD_OK;
    }
    if (dy_dtype == HTD_F32 && dx_dtype == HTD_F32) HTD_BWD_LAUNCH(float, float);
    else if (dy_dtype == HTD_F32 && dx_dtype == HTD_BF16) HTD_BWD_LAUNCH(float, __nv_bfloat16);
    else if (dy_dtype == HTD_BF16 && dx_dtype == HTD_F32) HTD_BWD_LAUNCH(__nv_bfloat16, float);
    else HTD_BWD_LAUNCH(__nv_bfloat16, __nv_bfloat16);
#undef HTD_BWD_LAUNCH
    HTD_CHECK_LAUNCH("htd_roi_align_bwd");
    return HTD_OK;
}

int htd_roi_align_bwd(const HtdLevel* grad_levels, int L, int B, int C, int dx_dtype,
                      const float* rois, int K, const int32_t* boxes, const int32_t* offsets,
                      const int32_t* ranges, const float* weights, int pooled, const void* dy,
                      int dy_dtype, int dy_per_level, const float* scale, int ring_edge,
                      const float* addvec, htd_stream_t stream) {
    HtdBwdSource q;
    q.rois = rois; q.boxes = boxes; q.offsets = offsets; q.ranges = ranges; q.weights = weights;
    q.dy = dy; q.scale = scale; q.addvec = addvec; q.K = K; q.dy_per_level = dy_per_level;
    q.ring_edge = ring_edge; q.addvec_dtype = HTD_F32;
    return htd_roi_align_bwd_multi(grad_levels, L, B, C, dx_dtype, 0, &q, 1, pooled, dy_dtype,
                                   stream);
}

}  // extern "C"
