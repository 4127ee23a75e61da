// Ranking step of the RPN proposal path (reference
// mmdet/models/dense_heads/rpn_head.py:125-134: `scores.sort(descending=True)` then the first
// nms_pre entries): the k largest of n keys per row, in descending order, ties by ascending index
// (the order of a stable sort; the reference's sort is unstable, i.e. any tie order is "its").
// One CTA per row:
//   1. radix select of the k-th largest key: four 8-bit passes over order-preserving uint32 images
//      of the floats (shared-memory histograms, digits scanned from the top);
//   2. one pass collects every key above the threshold and the keys equal to it (when they fit;
//      massive ties take an index-ordered compaction instead);
//   3. bitonic sort of the <= 8192 collected (key, ~index) pairs in shared memory.
// The rows are read 5 times from L2 (n <= a few 100 k floats); nothing is written but the k results.
#include "common.cuh"

namespace htd {

constexpr int kTopkThreads = 1024;
constexpr int kTopkCap = 8192;           // collected candidates (power of two)

__device__ __forceinline__ uint32_t key_of(float x) {
    const uint32_t u = x == 0.f ? 0u : __float_as_uint(x);      // -0 and +0 are one key (they compare equal)
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);          // ascending uint = ascending float
}
__device__ __forceinline__ float float_of(uint32_t k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

__global__ void __launch_bounds__(kTopkThreads, 1)
    topk_sorted_kernel(const float* __restrict__ keys, long long row_stride, int n, int k,
                       float* __restrict__ out_keys, int* __restrict__ out_idx) {
    extern __shared__ unsigned long long s_pairs[];              // [kTopkCap]
    __shared__ unsigned s_hist[256];
    __shared__ unsigned s_prefix, s_remaining, s_count, s_eq_total, s_eq_taken;
    __shared__ unsigned s_scan[kTopkThreads / 32];
    const int tid = threadIdx.x;
    const float* row = keys + (long long)blockIdx.x * row_stride;
    const bool vec = (reinterpret_cast<uintptr_t>(row) & 15u) == 0;
    float* ok = out_keys + (long long)blockIdx.x * k;
    int* oi = out_idx + (long long)blockIdx.x * k;

    // ---- 1. radix select: after pass p the top 8 (p + 1) bits of the k-th largest key are known
    if (tid == 0) { s_prefix = 0u; s_remaining = (unsigned)k; }
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 24 - 8 * pass;
        if (tid < 256) s_hist[tid] = 0u;
        __syncthreads();
        const unsigned prefix = s_prefix;
        // four elements per thread and trip (16-byte loads when the row allows): the passes are bound
        // by the latency of the row reads, one CTA has only 32 warps to hide it
        for (int base = 0; base < n; base += 4 * kTopkThreads) {   // uniform trip count: ballots below
          float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
          const int i0 = base + 4 * tid;
          if (vec && i0 + 3 < n) v4 = *reinterpret_cast<const float4*>(row + i0);
          else {
              if (i0 < n) v4.x = row[i0];
              if (i0 + 1 < n) v4.y = row[i0 + 1];
              if (i0 + 2 < n) v4.z = row[i0 + 2];
              if (i0 + 3 < n) v4.w = row[i0 + 3];
          }
          const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int i = i0 + e;
            const uint32_t u = i < n ? key_of(vv[e]) : 0u;
            bool valid = i < n && (pass == 0 || (u >> (shift + 8)) == prefix);
            const unsigned digit = (u >> shift) & 255u;
            if (pass == 0) {
                // sign + top exponent bits: a handful of values for real logits - 32 lanes adding to
                // the same counter serialise, so a warp adds each of its first few digits once
                unsigned active = __ballot_sync(0xffffffffu, valid);
                for (int it = 0; it < 4 && active; ++it) {
                    const int leader = __ffs(active) - 1;
                    const unsigned d0 = __shfl_sync(0xffffffffu, digit, leader);
                    const bool mine = valid && digit == d0;
                    const unsigned grp = __ballot_sync(0xffffffffu, mine);
                    if ((tid & 31) == leader) atomicAdd(&s_hist[d0], (unsigned)__popc(grp));
                    if (mine) valid = false;
                    active &= ~grp;
                }
            }
            if (valid) atomicAdd(&s_hist[digit], 1u);
          }
        }
        __syncthreads();
        if (tid == 0) {
            unsigned rem = s_remaining, cum = 0u;
            int d = 255;
            for (; d > 0; --d) {
                if (cum + s_hist[d] >= rem) break;
                cum += s_hist[d];
            }
            s_remaining = rem - cum;                 // still to take among the keys with digit d
            s_prefix = (prefix << 8) | (unsigned)d;
            if (pass == 3) s_eq_total = s_hist[d];   // keys equal to the threshold
        }
        __syncthreads();
    }
    const uint32_t thr = s_prefix;
    const unsigned need_eq = s_remaining;            // threshold-valued keys inside the top k (>= 1)
    const unsigned n_gt = (unsigned)k - need_eq;
    const bool all_eq = n_gt + s_eq_total <= (unsigned)kTopkCap;   // every tie fits: sort decides
    if (tid == 0) { s_count = 0u; s_eq_taken = 0u; }
    __syncthreads();

    // ---- 2. collect (key, ~index): descending 64-bit order = key descending, index ascending
    if (all_eq) {
        for (int base = 0; base < n; base += 4 * kTopkThreads) {
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            const int i0 = base + 4 * tid;
            if (vec && i0 + 3 < n) v4 = *reinterpret_cast<const float4*>(row + i0);
            else {
                if (i0 < n) v4.x = row[i0];
                if (i0 + 1 < n) v4.y = row[i0 + 1];
                if (i0 + 2 < n) v4.z = row[i0 + 2];
                if (i0 + 3 < n) v4.w = row[i0 + 3];
            }
            const float vv[4] = {v4.x, v4.y, v4.z, v4.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int i = i0 + e;
                if (i >= n) continue;
                const uint32_t u = key_of(vv[e]);
                if (u >= thr) {
                    const unsigned at = atomicAdd(&s_count, 1u);
                    s_pairs[at] = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
                }
            }
        }
    } else {
        // massive ties (e.g. saturated logits): the first need_eq threshold-valued keys in index order
        const int lane = tid & 31, warp = tid >> 5;
        for (int base = 0; base < n; base += kTopkThreads) {
            const int i = base + tid;
            const uint32_t u = i < n ? key_of(row[i]) : 0u;
            const bool gt = i < n && u > thr, eq = i < n && u == thr;
            if (gt) {
                const unsigned at = atomicAdd(&s_count, 1u);
                s_pairs[at] = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
            }
            const unsigned m = __ballot_sync(0xffffffffu, eq);
            if (lane == 0) s_scan[warp] = __popc(m);
            __syncthreads();
            unsigned before = s_eq_taken;
            for (int w = 0; w < warp; ++w) before += s_scan[w];
            const unsigned rank = before + __popc(m & ((1u << lane) - 1u));
            if (eq && rank < need_eq) {
                const unsigned at = atomicAdd(&s_count, 1u);
                s_pairs[at] = ((unsigned long long)u << 32) | (unsigned long long)(0xffffffffu - (unsigned)i);
            }
            __syncthreads();
            if (tid == 0) {
                unsigned t = 0u;
                for (int w = 0; w < kTopkThreads / 32; ++w) t += s_scan[w];
                s_eq_taken += t;
            }
            __syncthreads();
        }
    }
    __syncthreads();
    const unsigned cnt = s_count;
    unsigned m = 1u;
    while (m < cnt) m <<= 1;
    for (unsigned i = cnt + tid; i < m; i += kTopkThreads) s_pairs[i] = 0ull;      // below every key
    __syncthreads();

    // ---- 3. bitonic sort, descending
    for (unsigned size = 2u; size <= m; size <<= 1) {
        for (unsigned stride = size >> 1; stride > 0u; stride >>= 1) {
            for (unsigned t = tid; t < (m >> 1); t += kTopkThreads) {
                const unsigned lo = 2u * t - (t & (stride - 1u));          // index with bit `stride` clear
                const unsigned hi = lo + stride;
                const bool desc = (lo & size) == 0u;
                const unsigned long long a = s_pairs[lo], b = s_pairs[hi];
                if ((a < b) == desc) { s_pairs[lo] = b; s_pairs[hi] = a; }
            }
            __syncthreads();
        }
    }
    for (int j = tid; j < k; j += kTopkThreads) {
        const unsigned long long p = s_pairs[j];
        ok[j] = float_of((uint32_t)(p >> 32));
        oi[j] = (int)(0xffffffffu - (unsigned)(p & 0xffffffffull));
    }
}

}  // namespace htd

using namespace htd;

extern "C" {

int htd_topk_sorted(const float* keys, long long row_stride, int rows, int n, int k, float* out_keys,
                    int32_t* out_idx, htd_stream_t stream) {
    HTD_CHECK_ARG(rows >= 0 && n >= 1 && k >= 1 && k <= n && k <= HTD_TOPK_MAX && row_stride >= n,
                  "htd_topk_sorted: bad sizes rows=%d n=%d k=%d (1 <= k <= min(n, %d)) stride=%lld", rows,
                  n, k, HTD_TOPK_MAX, row_stride);
    if (rows == 0) return HTD_OK;
    HTD_CHECK_ARG(keys && out_keys && out_idx, "htd_topk_sorted: null pointer");
    const int smem = kTopkCap * (int)sizeof(unsigned long long);
    HTD_SMEM_OPTIN(topk_sorted_kernel, smem, "htd_topk_sorted");
    topk_sorted_kernel<<<rows, kTopkThreads, smem, (cudaStream_t)stream>>>(keys, row_stride, n, k, out_keys,
                                                                          out_idx);
    HTD_CHECK_LAUNCH("htd_topk_sorted");
    return HTD_OK;
}

}  // extern "C"
