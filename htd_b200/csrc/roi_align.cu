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
//  * forward: one warp per output bin, a lane owns 8 of the 256 channels of a chunk; work is a
//    flat list of bins so a 200x200-pixel BA footprint on P2 is spread over 49 warps.
//  * backward: one CTA per 8x8-pixel tile of dX; it scans the RoI footprint boxes of its image,
//    compacts the intersecting RoIs IN INDEX ORDER (ballot + prefix), and accumulates
//    Wy^T dY Wx for its 64 pixels x 256 channels in registers.  Every dX element is written
//    exactly once: no atomics, no memset, bit-reproducible.
#include <stdarg.h>

#include "common.cuh"
#include "roi_axis.h"

namespace htd {

static thread_local char g_err[512] = "";

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
// forward
// ------------------------------------------------------------------------------------------
struct FwdParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int L, B, C, K, P, sr;
    const float* rois;
    const int* roi_level;
    const float* bias;
    void* out;
    long long total_bins;
};

constexpr int kFwdWarps = 8;
constexpr int kWChunk = 64;

template <typename TIn, typename TOut>
__global__ void __launch_bounds__(kFwdWarps * 32) roi_align_fwd_kernel(const FwdParams p) {
    constexpr bool kSplit = SplitMap<TIn, TOut>::value;
    __shared__ float s_w[kFwdWarps][2][kWChunk];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long widx = (long long)blockIdx.x * kFwdWarps + warp;
    if (widx >= p.total_bins) return;
    const int PP = p.P * p.P;
    const int bin = (int)(widx % PP);
    const long long task = widx / PP;
    int l, k;
    if (p.roi_level) { k = (int)task; l = p.roi_level[k]; }
    else { l = (int)(task / p.K); k = (int)(task % p.K); }
    const float* r = p.rois + (size_t)k * 5;
    const float x1 = r[1], y1 = r[2], x2 = r[3], y2 = r[4];
    const int b = (int)r[0];
    const int ph = bin / p.P, pw = bin % p.P;
    const bool bvalid = (b >= 0 && b < p.B);
    const bool valid = bvalid && (l >= 0 && l < p.L);

    int H = 1, W = 1, r0 = 0, r1 = -1, c0 = 0, c1 = -1;
    const TIn* feat = nullptr;
    Axis ay, ax;
    if (valid) {
        H = p.lv[l].H; W = p.lv[l].W;
        feat = static_cast<const TIn*>(p.lv[l].data);
        const double sc = (double)p.lv[l].scale;
        ay = make_axis(y1, y2, sc, p.P, H, p.sr, 1);
        ax = make_axis(x1, x2, sc, p.P, W, p.sr, 1);
        bin_range(ay, ph, r0, r1);
        bin_range(ax, pw, c0, c1);
        if (c1 < c0) r1 = r0 - 1;
    }
    float* wy_s = s_w[warp][0];
    float* wx_s = s_w[warp][1];
    TOut* orow = static_cast<TOut*>(p.out) + ((size_t)task * PP + bin) * p.C;

    for (int cb = 0; cb < p.C; cb += 256) {
        const int nch = min(256, p.C - cb);
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int rr = r0; rr <= r1; rr += kWChunk) {
            const int nr = min(kWChunk, r1 - rr + 1);
            __syncwarp();
            for (int j = lane; j < nr; j += 32) wy_s[j] = axis_weight(ay, ph, rr + j);
            for (int cc = c0; cc <= c1; cc += kWChunk) {
                const int nc = min(kWChunk, c1 - cc + 1);
                __syncwarp();
                for (int j = lane; j < nc; j += 32) wx_s[j] = axis_weight(ax, pw, cc + j);
                __syncwarp();
                for (int y = 0; y < nr; ++y) {
                    const float wyv = wy_s[y];
                    const TIn* rowp = feat + (((size_t)b * H + (rr + y)) * W + cc) * p.C + cb;
#pragma unroll 4
                    for (int x = 0; x < nc; ++x) {
                        float v[8];
                        Vec8<TIn, kSplit>::load(rowp + (size_t)x * p.C, lane, nch, v);
                        const float w = wyv * wx_s[x];
#pragma unroll
                        for (int e = 0; e < 8; ++e) acc[e] = fmaf(w, v[e], acc[e]);
                    }
                }
            }
        }
        if (p.bias && bvalid) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = lane_chan<kSplit>(lane, e);
                if (c < nch) acc[e] += __ldg(p.bias + (size_t)b * p.C + cb + c);
            }
        }
        Vec8<TOut, kSplit>::store(orow + cb, lane, nch, acc);
    }
}

// ------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------
struct BwdParams {
    LevelDev lv[HTD_MAX_LEVELS];
    int tile_start[HTD_MAX_LEVELS + 1];
    int L, B, C, K, P, sr;
    const float* rois;
    const int4* boxes;
    const void* dy;
    int dy_per_level;
    const float* scale;
    int ring_edge;
    const float* addvec;
};

constexpr int kTile = 8;  // 8x8 pixels per CTA, one tile row per warp

template <typename TDy, typename TDx>
__global__ void __launch_bounds__(256) roi_align_bwd_kernel(const BwdParams p) {
    constexpr bool kSplit = SplitMap<TDy, TDx>::value;
    __shared__ float s_wy[2][HTD_MAX_POOLED][kTile];
    __shared__ float s_wx[2][HTD_MAX_POOLED][kTile];
    __shared__ int s_hits[256];
    __shared__ int s_wcnt[8];

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    int l = 0;
    while (l + 1 < p.L && (int)blockIdx.x >= p.tile_start[l + 1]) ++l;
    const int H = p.lv[l].H, W = p.lv[l].W;
    const double sc = (double)p.lv[l].scale;
    const int tiles_x = (W + kTile - 1) / kTile, tiles_y = (H + kTile - 1) / kTile;
    int local = (int)blockIdx.x - p.tile_start[l];
    const int b = local / (tiles_x * tiles_y);
    local -= b * tiles_x * tiles_y;
    const int row0 = (local / tiles_x) * kTile, col0 = (local % tiles_x) * kTile;
    const int PP = p.P * p.P;
    const int row = row0 + warp;
    TDx* dx = static_cast<TDx*>(p.lv[l].data);
    const TDy* dy = static_cast<const TDy*>(p.dy);
    const int4* boxes = p.boxes + (size_t)l * p.K;

    for (int cb = 0; cb < p.C; cb += 256) {
        const int nch = min(256, p.C - cb);
        float acc[kTile][8];
#pragma unroll
        for (int x = 0; x < kTile; ++x)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[x][e] = 0.f;

        for (int k0 = 0; k0 < p.K; k0 += 256) {
            // ---- find the RoIs of this chunk that touch the tile, in ascending index order
            const int k = k0 + tid;
            bool hit = false;
            if (k < p.K) {
                const int4 bx = __ldg(boxes + k);
                hit = (bx.y >= bx.x) && bx.x <= row0 + kTile - 1 && bx.y >= row0 &&
                      bx.z <= col0 + kTile - 1 && bx.w >= col0 &&
                      ((int)__ldg(p.rois + (size_t)k * 5) == b);
            }
            const unsigned bal = __ballot_sync(0xffffffffu, hit);
            if (lane == 0) s_wcnt[warp] = __popc(bal);
            __syncthreads();
            int base = 0, nh = 0;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                const int c = s_wcnt[w];
                if (w < warp) base += c;
                nh += c;
            }
            if (hit) s_hits[base + __popc(bal & ((1u << lane) - 1u))] = k;
            __syncthreads();

            for (int h = 0; h < nh; ++h) {
                const int kk = s_hits[h];
                const int buf = h & 1;
                if (tid < 2 * HTD_MAX_POOLED * kTile) {
                    const int axis = tid / (HTD_MAX_POOLED * kTile);
                    const int pp = (tid / kTile) % HTD_MAX_POOLED;
                    const int j = tid % kTile;
                    float wv = 0.f;
                    if (pp < p.P) {
                        const float* r = p.rois + (size_t)kk * 5;
                        if (axis == 0) {
                            Axis a = make_axis(r[2], r[4], sc, p.P, H, p.sr, 1);
                            wv = axis_weight(a, pp, row0 + j);
                        } else {
                            Axis a = make_axis(r[1], r[3], sc, p.P, W, p.sr, 1);
                            wv = axis_weight(a, pp, col0 + j);
                        }
                    }
                    if (axis == 0) s_wy[buf][pp][j] = wv; else s_wx[buf][pp][j] = wv;
                }
                __syncthreads();
                if (row < H) {
                    const size_t dyk = (p.dy_per_level ? (size_t)l * p.K : 0) + kk;
                    const float sbase = p.scale ? __ldg(p.scale + (size_t)l * p.K + kk) : 1.f;
                    float av[8];
#pragma unroll
                    for (int e = 0; e < 8; ++e) av[e] = 0.f;
                    if (p.addvec) {
                        const float* a = p.addvec + ((size_t)l * p.K + kk) * p.C + cb;
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            const int c = lane_chan<kSplit>(lane, e);
                            if (c < nch) av[e] = __ldg(a + c);
                        }
                    }
                    const int e0 = p.ring_edge;
                    for (int ph = 0; ph < p.P; ++ph) {
                        const float wyv = s_wy[buf][ph][warp];
                        if (wyv == 0.f) continue;
                        for (int pw = 0; pw < p.P; ++pw) {
                            float cw[kTile];
                            bool any = false;
#pragma unroll
                            for (int x = 0; x < kTile; ++x) {
                                cw[x] = s_wx[buf][pw][x];
                                any |= (cw[x] != 0.f);
                            }
                            if (!any) continue;
                            float sv = sbase;
                            if (e0 >= 0 && l == 0) {
                                const bool interior = (e0 > 0) && ph >= e0 && ph < p.P - e0 &&
                                                      pw >= e0 && pw < p.P - e0;
                                if (!interior) sv += 1.f;
                            }
                            float v[8];
                            Vec8<TDy, kSplit>::load(dy + ((dyk * PP + ph * p.P + pw) * p.C + cb),
                                                    lane, nch, v);
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = fmaf(sv, v[e], av[e]);
#pragma unroll
                            for (int x = 0; x < kTile; ++x) {
                                const float wgt = wyv * cw[x];
#pragma unroll
                                for (int e = 0; e < 8; ++e) acc[x][e] = fmaf(wgt, v[e], acc[x][e]);
                            }
                        }
                    }
                }
            }
            __syncthreads();
        }
        if (row < H) {
#pragma unroll
            for (int x = 0; x < kTile; ++x) {
                const int col = col0 + x;
                if (col < W)
                    Vec8<TDx, kSplit>::store(dx + (((size_t)b * H + row) * W + col) * p.C + cb,
                                             lane, nch, acc[x]);
            }
        }
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

int htd_roi_align_fwd(const HtdLevel* levels, int L, int B, int C, int in_dtype,
                      const float* rois, int K, const int32_t* roi_level, int pooled,
                      int sampling_ratio, const float* bias, void* out, int out_dtype,
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
    HTD_CHECK_ARG(rois && out, "htd_roi_align_fwd: null pointer");
    p.L = L; p.B = B; p.C = C; p.K = K; p.P = pooled; p.sr = sampling_ratio;
    p.rois = rois; p.roi_level = roi_level; p.bias = bias; p.out = out;
    const long long tasks = roi_level ? (long long)K : (long long)K * L;
    p.total_bins = tasks * pooled * pooled;
    const long long blocks = (p.total_bins + kFwdWarps - 1) / kFwdWarps;
    HTD_CHECK_ARG(blocks < 2147483647LL, "htd_roi_align_fwd: too many bins (%lld)", p.total_bins);
    dim3 grid((unsigned)blocks), block(kFwdWarps * 32);
    cudaStream_t st = (cudaStream_t)stream;
    if (in_dtype == HTD_F32 && out_dtype == HTD_F32)
        roi_align_fwd_kernel<float, float><<<grid, block, 0, st>>>(p);
    else if (in_dtype == HTD_F32 && out_dtype == HTD_BF16)
        roi_align_fwd_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(p);
    else if (in_dtype == HTD_BF16 && out_dtype == HTD_F32)
        roi_align_fwd_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(p);
    else
        roi_align_fwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>(p);
    HTD_CHECK_LAUNCH("htd_roi_align_fwd");
    return HTD_OK;
}

int htd_roi_align_bwd(const HtdLevel* grad_levels, int L, int B, int C, int dx_dtype,
                      const float* rois, int K, const int32_t* boxes, int pooled,
                      int sampling_ratio, const void* dy, int dy_dtype, int dy_per_level,
                      const float* scale, int ring_edge, const float* addvec,
                      htd_stream_t stream) {
    BwdParams p;
    int rc = fill_levels(p.lv, grad_levels, L, "htd_roi_align_bwd");
    if (rc) return rc;
    HTD_CHECK_ARG(K >= 0 && B >= 1 && pooled >= 1 && pooled <= HTD_MAX_POOLED,
                  "htd_roi_align_bwd: bad sizes K=%d B=%d pooled=%d", K, B, pooled);
    HTD_CHECK_ARG(C >= 8 && C % 8 == 0, "htd_roi_align_bwd: C=%d must be a positive multiple of 8",
                  C);
    HTD_CHECK_ARG((dx_dtype == HTD_F32 || dx_dtype == HTD_BF16) &&
                      (dy_dtype == HTD_F32 || dy_dtype == HTD_BF16),
                  "htd_roi_align_bwd: unsupported dtype dx=%d dy=%d", dx_dtype, dy_dtype);
    HTD_CHECK_ARG(K == 0 || (rois && boxes && dy), "htd_roi_align_bwd: null pointer");
    p.L = L; p.B = B; p.C = C; p.K = K; p.P = pooled; p.sr = sampling_ratio;
    p.rois = rois; p.boxes = reinterpret_cast<const int4*>(boxes); p.dy = dy;
    p.dy_per_level = dy_per_level; p.scale = scale; p.ring_edge = ring_edge; p.addvec = addvec;
    long long total = 0;
    for (int l = 0; l < L; ++l) {
        p.tile_start[l] = (int)total;
        total += (long long)B * ((p.lv[l].H + kTile - 1) / kTile) * ((p.lv[l].W + kTile - 1) / kTile);
    }
    for (int l = L; l <= HTD_MAX_LEVELS; ++l) p.tile_start[l] = (int)total;
    HTD_CHECK_ARG(total < 2147483647LL, "htd_roi_align_bwd: too many tiles (%lld)", total);
    dim3 grid((unsigned)total), block(256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dy_dtype == HTD_F32 && dx_dtype == HTD_F32)
        roi_align_bwd_kernel<float, float><<<grid, block, 0, st>>>(p);
    else if (dy_dtype == HTD_F32 && dx_dtype == HTD_BF16)
        roi_align_bwd_kernel<float, __nv_bfloat16><<<grid, block, 0, st>>>(p);
    else if (dy_dtype == HTD_BF16 && dx_dtype == HTD_F32)
        roi_align_bwd_kernel<__nv_bfloat16, float><<<grid, block, 0, st>>>(p);
    else
        roi_align_bwd_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, block, 0, st>>>(p);
    HTD_CHECK_LAUNCH("htd_roi_align_bwd");
    return HTD_OK;
}

}  // extern "C"
