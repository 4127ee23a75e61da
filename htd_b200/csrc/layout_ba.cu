// Layout conversion and the Border-aware Adaptation (BA) fuse kernels.
//
// Reference path replaced: AdptRoIExtractor.forward, adaptative_roi_extractor.py:76-91 -
//   atts = cat(atts).softmax(0); roi_feats = (atts * roi_feat).sum(0);
//   roi_feats_enhance = RoIAlign_0(feats[0], rois); roi_feats_enhance[:, :, e:-e, e:-e] = 0
//   return roi_feats + roi_feats_enhance
// which materialises [4,P,256,7,7] twice (repeat + product) and runs RoIAlign on P2 twice; here
// the four level samples R[l] come from ONE multi-level htd_roi_align_fwd launch (P2 sampled
// once) and the softmax-weighted sum + border ring (+ the HTDBBoxHead residual adds,
// htd_bbox_head.py:161-184) are one elementwise pass.
#include "common.cuh"

namespace htd {

// ------------------------------------------------------------------------------------------
// [N,R,S] -> [N,S,R] tiled transpose with dtype conversion
// ------------------------------------------------------------------------------------------
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <typename T>
__device__ __forceinline__ T from_f(float v);
template <>
__device__ __forceinline__ float from_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float v) { return __float2bfloat16_rn(v); }

// Two forms, both reading and writing global memory in 16-byte vectors where the shape allows:
//  * whole-matrix form (R*S <= kFlatMax): one CTA owns one [R,S] matrix; the source is read as a
//    flat vector stream into a padded fp32 tile, re-read along R, packed into a flat image of
//    the destination in shared memory and written out as a flat vector stream. This is the
//    [P,49,C] <-> [P,C,49] flatten-order change of the pooled RoI maps (49 is odd, so neither
//    side has vectorisable rows; the flat image is what makes both sides 16-byte traffic).
//  * tiled form: 64 x 64 tiles, vector width chosen per side from the divisibility of S / R.
//    This is the NCHW <-> NHWC pyramid conversion (R = 256 channels, S = H*W).
constexpr int kFlatMax = 12800;   // elements: 49*256 = 12544 and a little slack

template <typename T, int V>
__device__ __forceinline__ void ldvec(const T* p, float (&v)[V]) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (V == 8) {
            const float4 a = *reinterpret_cast<const float4*>(p);
            const float4 b = *reinterpret_cast<const float4*>(p + 4);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
            v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
        } else if constexpr (V == 4) {
            const float4 a = *reinterpret_cast<const float4*>(p);
            v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        } else if constexpr (V == 2) {
            const float2 a = *reinterpret_cast<const float2*>(p);
            v[0] = a.x; v[1] = a.y;
        } else {
            v[0] = p[0];
        }
    } else {
        if constexpr (V == 8) {
            const uint4 a = *reinterpret_cast<const uint4*>(p);
            const uint32_t w[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                v[2 * i] = __uint_as_float(w[i] << 16);
                v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            }
        } else if constexpr (V == 4) {
            const uint2 a = *reinterpret_cast<const uint2*>(p);
            v[0] = __uint_as_float(a.x << 16); v[1] = __uint_as_float(a.x & 0xffff0000u);
            v[2] = __uint_as_float(a.y << 16); v[3] = __uint_as_float(a.y & 0xffff0000u);
        } else if constexpr (V == 2) {
            const uint32_t a = *reinterpret_cast<const uint32_t*>(p);
            v[0] = __uint_as_float(a << 16); v[1] = __uint_as_float(a & 0xffff0000u);
        } else {
            v[0] = to_f<T>(p[0]);
        }
    }
}

template <typename T, int V>
__device__ __forceinline__ void stvec(T* p, const float (&v)[V]) {
    if constexpr (sizeof(T) == 4) {
        if constexpr (V == 8) {
            *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else if constexpr (V == 4) {
            *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
        } else if constexpr (V == 2) {
            *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
        } else {
            p[0] = v[0];
        }
    } else {
        if constexpr (V == 1) {
            p[0] = from_f<T>(v[0]);
        } else {
            uint32_t w[V / 2];
#pragma unroll
            for (int i = 0; i < V / 2; ++i) {
                const __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
                w[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            if constexpr (V == 8) *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
            else if constexpr (V == 4) *reinterpret_cast<uint2*>(p) = make_uint2(w[0], w[1]);
            else *reinterpret_cast<uint32_t*>(p) = w[0];
        }
    }
}

// whole-matrix form; dynamic shared memory = R*pitch source elements (pitch odd: the column walk
// of the second phase then spreads over the banks) + R*S destination elements
template <typename TS, typename TD>
__global__ void __launch_bounds__(256) transpose_flat_kernel(const TS* __restrict__ src,
                                                             TD* __restrict__ dst, int R, int S,
                                                             int pitch, int image_off,
                                                             const float* __restrict__ bias,
                                                             const float* __restrict__ rois, int B) {
    extern __shared__ __align__(16) unsigned char tr_smem[];
    TS* tile = reinterpret_cast<TS*>(tr_smem);
    TD* image = reinterpret_cast<TD*>(tr_smem + image_off);
    const int E = R * S;                       // multiple of 8 (checked on the host)
    const TS* s = src + (size_t)blockIdx.x * E;
    TD* d = dst + (size_t)blockIdx.x * E;
    constexpr int VL = 16 / (int)sizeof(TS);
    // four 16-byte loads in flight per thread before the first of them is scattered into the tile
    for (int i0 = threadIdx.x * VL; i0 < E; i0 += 4 * 256 * VL) {
        uint4 raw[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * 256 * VL;
            if (i < E) raw[k] = *reinterpret_cast<const uint4*>(s + i);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int i = i0 + k * 256 * VL;
            if (i < E) {
                const TS* e = reinterpret_cast<const TS*>(&raw[k]);
                int r = i / S, c = i - r * S;
#pragma unroll
                for (int j = 0; j < VL; ++j) {
                    tile[r * pitch + c] = e[j];
                    if (++c == S) { c = 0; ++r; }
                }
            }
        }
    }
    __syncthreads();
    // destination flat index o = c*R + r: consecutive lanes walk r, i.e. tile rows `pitch` apart;
    // (c, r) advance by 256 flat positions per trip without a division
    {
        // optional per-(image, column) bias: matrix n belongs to image rois[5 n] (the SFA vector
        // added to a RoI's channels while its map changes to the FC flatten order)
        const float* bn = nullptr;
        if (bias != nullptr) {
            const int img = (int)rois[(size_t)blockIdx.x * 5];
            if (img >= 0 && img < B) bn = bias + (size_t)img * S;
        }
        const int dc = 256 / R, dr = 256 - dc * R;
        int c = threadIdx.x / R, r = threadIdx.x - c * R;
#pragma unroll 4
        for (int o = threadIdx.x; o < E; o += 256) {
            float v = to_f<TS>(tile[r * pitch + c]);
            if (bn != nullptr) v += bn[c];
            image[o] = from_f<TD>(v);
            c += dc;
            r += dr;
            if (r >= R) { r -= R; ++c; }
        }
    }
    __syncthreads();
    constexpr int VB = 16 / (int)sizeof(TD);
    for (int i = threadIdx.x * VB; i < E; i += 256 * VB)
        *reinterpret_cast<uint4*>(d + i) = *reinterpret_cast<const uint4*>(image + i);
}

// tiled form
template <typename TS, typename TD, int VL, int VS>
__global__ void __launch_bounds__(256) transpose_kernel(const TS* __restrict__ src,
                                                        TD* __restrict__ dst, long long N, int R,
                                                        int S, int tiles_r, int tiles_s) {
    __shared__ float tile[64][65];
    long long blk = blockIdx.x;
    const int ts = (int)(blk % tiles_s);
    blk /= tiles_s;
    const int tr = (int)(blk % tiles_r);
    const long long n = blk / tiles_r;
    const TS* s = src + (size_t)n * R * S;
    TD* d = dst + (size_t)n * R * S;
    const int r0 = tr * 64, s0 = ts * 64;
    constexpr int LV = 64 / VL;                // vectors per tile row on the load side
#pragma unroll
    for (int i = threadIdx.x; i < 64 * LV; i += 256) {
        const int rr = i / LV, cv = (i - rr * LV) * VL;
        const int r = r0 + rr, c = s0 + cv;
        if (r < R && c < S) {                  // S % VL == 0: a vector never straddles the edge
            float v[VL];
            ldvec<TS, VL>(s + (size_t)r * S + c, v);
#pragma unroll
            for (int j = 0; j < VL; ++j) tile[rr][cv + j] = v[j];
        }
    }
    __syncthreads();
    constexpr int SV = 64 / VS;
#pragma unroll
    for (int i = threadIdx.x; i < 64 * SV; i += 256) {
        const int cc = i / SV, rv = (i - cc * SV) * VS;
        const int c = s0 + cc, r = r0 + rv;
        if (c < S && r < R) {
            float v[VS];
#pragma unroll
            for (int j = 0; j < VS; ++j) v[j] = tile[rv + j][cc];
            stvec<TD, VS>(d + (size_t)c * R + r, v);
        }
    }
}

// ------------------------------------------------------------------------------------------
// BA kernels
// ------------------------------------------------------------------------------------------
// out[n, c] = scale * sum over the PP bins of x[n, bin, c]: one CTA per row n; a thread owns 8
// channels (one 16-byte load per bin) of every `groups`-th bin, the bin groups meet in shared
// memory in a fixed order (deterministic).  C % 8 == 0, C <= 2048.
template <typename T>
__global__ void __launch_bounds__(256) bin_sum_kernel(const T* __restrict__ x, long long N, int PP,
                                                      int C, float scale, float* __restrict__ out) {
    __shared__ float red[8][256];                        // [bin group][channel], C <= 256
    const long long n = blockIdx.x;
    const T* xn = x + (size_t)n * PP * C;
    const int vecs = C / 8;                              // <= 256
    const bool exchange = C <= 256;
    const int groups = exchange ? (256 / vecs < 8 ? 256 / vecs : 8) : 1;
    const int v = threadIdx.x % vecs, g = threadIdx.x / vecs;
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    if (g < groups) {
#pragma unroll 4
        for (int b = g; b < PP; b += groups) {
            float t[8];
            ldvec<T, 8>(xn + (size_t)b * C + v * 8, t);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += t[j];
        }
    }
    if (exchange) {
        if (g < groups) {
#pragma unroll
            for (int j = 0; j < 8; ++j) red[g][v * 8 + j] = acc[j];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < C; c += 256) {
            float t = 0.f;
            for (int k = 0; k < groups; ++k) t += red[k][c];
            out[(size_t)n * C + c] = t * scale;
        }
    } else if (g == 0) {                                 // wide C: one group, no exchange
#pragma unroll
        for (int j = 0; j < 8; ++j) out[(size_t)n * C + v * 8 + j] = acc[j] * scale;
    }
}

// generic form (any C)
template <typename T>
__global__ void __launch_bounds__(256) bin_mean_kernel(const T* __restrict__ x, long long N, int PP,
                                                       int C, float* __restrict__ mean) {
    const long long n = blockIdx.x;
    const T* xn = x + (size_t)n * PP * C;
    const float inv = 1.f / (float)PP;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < PP; ++b) s += to_f<T>(xn[(size_t)b * C + c]);
        mean[(size_t)n * C + c] = s * inv;
    }
}

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]) {
    Vec8<T, false>::load(p, 0, 8, v);   // lane 0 of the contiguous map = 8 channels at p
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[8]) {
    Vec8<T, false>::store(p, 0, 8, v);
}

struct FuseParams {
    const void* R;
    const float* logits;
    int L, K, P, C, ring_edge, B;
    const void* add;
    const float* bias;
    const float* rois;
    float* w;
    void* out;
};

template <typename TR, typename TA, typename TO>
__global__ void __launch_bounds__(256) ba_fuse_fwd_kernel(const FuseParams p) {
    const int PP = p.P * p.P, C8 = p.C / 8;
    const long long total = (long long)p.K * PP * C8;
    const TR* R = static_cast<const TR*>(p.R);
    const TA* add = static_cast<const TA*>(p.add);
    TO* out = static_cast<TO*>(p.out);
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c8 = (int)(i % C8);
        const long long kb = i / C8;
        const int bin = (int)(kb % PP);
        const int k = (int)(kb / PP);
        // softmax over levels (L <= 8), recomputed per thread: L exps vs 8*L loads of R
        float wl[HTD_MAX_LEVELS];
        float mx = -INFINITY;
        for (int l = 0; l < p.L; ++l) {
            wl[l] = __ldg(p.logits + (size_t)l * p.K + k);
            mx = fmaxf(mx, wl[l]);
        }
        float den = 0.f;
        for (int l = 0; l < p.L; ++l) { wl[l] = expf(wl[l] - mx); den += wl[l]; }
        const float inv = 1.f / den;
        for (int l = 0; l < p.L; ++l) wl[l] *= inv;
        if (bin == 0 && c8 == 0)
            for (int l = 0; l < p.L; ++l) p.w[(size_t)l * p.K + k] = wl[l];
        const int ph = bin / p.P, pw = bin % p.P, e0 = p.ring_edge;
        float ring = 0.f;
        if (e0 >= 0) {
            const bool interior = (e0 > 0) && ph >= e0 && ph < p.P - e0 && pw >= e0 && pw < p.P - e0;
            ring = interior ? 0.f : 1.f;
        }
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        const size_t off = ((size_t)k * PP + bin) * p.C + (size_t)c8 * 8;
        for (int l = 0; l < p.L; ++l) {
            float v[8];
            ld8<TR>(R + (size_t)l * p.K * PP * p.C + off, v);
            const float wv = wl[l] + (l == 0 ? ring : 0.f);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(wv, v[e], acc[e]);
        }
        if (add) {
            float v[8];
            ld8<TA>(add + off, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += v[e];
        }
        if (p.bias) {
            const int b = (int)__ldg(p.rois + (size_t)k * 5);
            if (b >= 0 && b < p.B) {
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[e] += __ldg(p.bias + (size_t)b * p.C + c8 * 8 + e);
            }
        }
        st8<TO>(out + off, acc);
    }
}

template <typename TR, typename TG>
__global__ void __launch_bounds__(256) ba_fuse_bwd_kernel(const TR* __restrict__ R,
                                                          const TG* __restrict__ dout,
                                                          const float* __restrict__ w, int L, int K,
                                                          int PP, int C, float* __restrict__ da) {
    __shared__ float s_part[HTD_MAX_LEVELS][8];
    const int k = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int n8 = PP * C / 8;
    float dw[HTD_MAX_LEVELS];
#pragma unroll
    for (int l = 0; l < HTD_MAX_LEVELS; ++l) dw[l] = 0.f;
    for (int i = tid; i < n8; i += 256) {
        float g[8];
        ld8<TG>(dout + ((size_t)k * n8 + i) * 8, g);
#pragma unroll
        for (int l = 0; l < HTD_MAX_LEVELS; ++l) {
            if (l < L) {
                float v[8];
                ld8<TR>(R + (((size_t)l * K + k) * n8 + i) * 8, v);
#pragma unroll
                for (int e = 0; e < 8; ++e) dw[l] = fmaf(g[e], v[e], dw[l]);
            }
        }
    }
#pragma unroll
    for (int l = 0; l < HTD_MAX_LEVELS; ++l) {
        const float s = warp_sum(dw[l]);
        if (lane == 0) s_part[l][warp] = s;
    }
    __syncthreads();
    if (tid == 0) {
        float tot[HTD_MAX_LEVELS], mean = 0.f;
        for (int l = 0; l < L; ++l) {
            float s = 0.f;
            for (int i = 0; i < 8; ++i) s += s_part[l][i];
            tot[l] = s;
            mean += w[(size_t)l * K + k] * s;
        }
        for (int l = 0; l < L; ++l) da[(size_t)l * K + k] = w[(size_t)l * K + k] * (tot[l] - mean);
    }
}

// dbias[b, c] = sum over the RoIs k of image b of rowsum[k, c] (rowsum = per-RoI bin sums from
// bin_sum_kernel).  CTA = (32 channels, image b): 32 row lanes x 32 channels, lane l adds rows
// l, l + 32, ... in ascending order, the lanes meet in shared memory in lane order (deterministic).
__global__ void __launch_bounds__(1024) rows_by_image_kernel(const float* __restrict__ rowsum,
                                                             const float* __restrict__ rois, int K,
                                                             int C, int B, float* __restrict__ dbias) {
    __shared__ float red[32][33];
    const int cl = threadIdx.x & 31, l = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cl, b = blockIdx.y;
    float s = 0.f;
    if (c < C) {
#pragma unroll 4
        for (int k = l; k < K; k += 32) {
            const float v = rowsum[(size_t)k * C + c];
            const int img = (int)__ldg(rois + (size_t)k * 5);
            s += img == b ? v : 0.f;
        }
    }
    red[l][cl] = s;
    __syncthreads();
    if (l == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) t += red[i][cl];
        dbias[(size_t)b * C + c] = t;
    }
}

// generic form (any C), phase 1: block = kBiasRois consecutive RoIs, thread = channel
constexpr int kBiasRois = 4;
template <typename T>
__global__ void __launch_bounds__(256) bias_grad_partial_kernel(const T* __restrict__ g,
                                                                const float* __restrict__ rois,
                                                                int K, int PP, int C, int B,
                                                                float* __restrict__ partial) {
    const int kb = blockIdx.x;
    float* part = partial + (size_t)kb * B * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        for (int b = 0; b < B; ++b) part[(size_t)b * C + c] = 0.f;
        const int kend = min(K, (kb + 1) * kBiasRois);
        for (int k = kb * kBiasRois; k < kend; ++k) {
            const int b = (int)__ldg(rois + (size_t)k * 5);
            if (b < 0 || b >= B) continue;
            float s = 0.f;
            for (int bin = 0; bin < PP; ++bin) s += to_f<T>(g[((size_t)k * PP + bin) * C + c]);
            part[(size_t)b * C + c] += s;
        }
    }
}

__global__ void bias_grad_final_kernel(const float* __restrict__ partial, int nblk, int BC,
                                       float* __restrict__ dbias) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BC) return;
    float s = 0.f;
    for (int kb = 0; kb < nblk; ++kb) s += partial[(size_t)kb * BC + i];
    dbias[i] = s;
}


// ------------------------------------------------------------------------------------------
// Attention MLP of the BA extractor on the bin means (adaptative_roi_extractor.py:60-74: 1x1 conv
// C->H, tanh, 1x1 conv H->1 on the pooled vectors): logits[r] = b2 + sum_j W2[j] tanh(b1[j] +
// sum_c W1[j,c] m[r,c]).  rows = levels * RoIs (1024 at the bench size), H = 128: 67 MFLOP - the
// cost is launches, so forward is ONE kernel and backward three (the row gradient dm, which the
// backward gather waits for, first; the parameter sums after it), instead of ~8 and ~25
// library / elementwise launches.
// ------------------------------------------------------------------------------------------
constexpr int kMlpH = 128, kMlpRows = 4, kMlpMaxC = 256, kMlpChunks = 32;

constexpr int kMlpFwdRows = 16;                       // rows of m per CTA (two halves of 8)

template <typename TP>
__global__ void __launch_bounds__(2 * kMlpH) ba_mlp_fwd_kernel(const float* __restrict__ m, int rows,
                                                               int C, const TP* __restrict__ w1,
                                                               const TP* __restrict__ b1,
                                                               const TP* __restrict__ w2,
                                                               const TP* __restrict__ b2,
                                                               float* __restrict__ h,
                                                               float* __restrict__ logits) {
    __shared__ __align__(16) float sm[kMlpFwdRows][kMlpMaxC];
    __shared__ float wt[32][kMlpH + 1];               // W1 chunk, transposed: [c][j]
    __shared__ float red[kMlpFwdRows][kMlpH / 32];
    const int j = threadIdx.x % kMlpH, half = threadIdx.x / kMlpH;   // thread = hidden unit x 8 rows
    const int r0 = blockIdx.x * kMlpFwdRows;
    for (int i = threadIdx.x; i < kMlpFwdRows * C; i += 2 * kMlpH) {
        const int r = i / C, c = i - r * C;
        sm[r][c] = (r0 + r < rows) ? m[(size_t)(r0 + r) * C + c] : 0.f;
    }
    float acc[8];
    const float bj = ldv<TP>(b1 + j);
#pragma unroll
    for (int r = 0; r < 8; ++r) acc[r] = bj;
    // W1 chunk [128 j][32 c] as 16-byte vectors: VW elements per vector, NV vectors per thread;
    // the next chunk's vectors are in flight while the current one is consumed
    constexpr int VW = 16 / (int)sizeof(TP), NV = kMlpH * 32 / VW / (2 * kMlpH);
    uint4 pre[NV];
    auto fetch = [&](int c0) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = i * 2 * kMlpH + threadIdx.x, jj = v / (32 / VW), cq = (v % (32 / VW)) * VW;
            pre[i] = *reinterpret_cast<const uint4*>(w1 + (size_t)jj * C + c0 + cq);
        }
    };
    fetch(0);
    for (int c0 = 0; c0 < C; c0 += 32) {
        __syncthreads();                              // previous chunk consumed (and sm filled)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            const int v = i * 2 * kMlpH + threadIdx.x, jj = v / (32 / VW), cq = (v % (32 / VW)) * VW;
            const TP* e = reinterpret_cast<const TP*>(&pre[i]);
#pragma unroll
            for (int k = 0; k < VW; ++k) wt[cq + k][jj] = ldv<TP>(e + k);
        }
        if (c0 + 32 < C) fetch(c0 + 32);
        __syncthreads();
#pragma unroll
        for (int cq = 0; cq < 32; cq += 4) {
            const float w0 = wt[cq][j], w1v = wt[cq + 1][j], w2v = wt[cq + 2][j], w3v = wt[cq + 3][j];
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                const float4 x = *reinterpret_cast<const float4*>(&sm[half * 8 + r][c0 + cq]);
                acc[r] = fmaf(x.x, w0, acc[r]);
                acc[r] = fmaf(x.y, w1v, acc[r]);
                acc[r] = fmaf(x.z, w2v, acc[r]);
                acc[r] = fmaf(x.w, w3v, acc[r]);
            }
        }
    }
    const float w2j = ldv<TP>(w2 + j);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int row = half * 8 + r;
        const float hv = tanhf(acc[r]);
        if (r0 + row < rows) h[(size_t)(r0 + row) * kMlpH + j] = hv;
        float v = hv * w2j;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((j & 31) == 0) red[row][j >> 5] = v;
    }
    __syncthreads();
    if (threadIdx.x < kMlpFwdRows && r0 + (int)threadIdx.x < rows) {
        float v = ldv<TP>(b2);
#pragma unroll
        for (int w = 0; w < kMlpH / 32; ++w) v += red[threadIdx.x][w];
        logits[r0 + threadIdx.x] = v;
    }
}

// dpre[r,j] = da[r] W2[j] (1 - h[r,j]^2)  (kept for the parameter sums);
// dm[r,c] = inv_pp * sum_j dpre[r,j] W1[j,c]
constexpr int kMlpDmRows = 8;                         // rows per CTA of the dm kernel

template <typename TP>
__global__ void __launch_bounds__(256) ba_mlp_dm_kernel(const float* __restrict__ da,
                                                        const float* __restrict__ h, int rows, int C,
                                                        const TP* __restrict__ w1,
                                                        const TP* __restrict__ w2, float inv_pp,
                                                        float* __restrict__ dpre,
                                                        float* __restrict__ dm) {
    __shared__ __align__(16) float dp[kMlpDmRows][kMlpH];
    const int r0 = blockIdx.x * kMlpDmRows;
    for (int i = threadIdx.x; i < kMlpDmRows * kMlpH; i += 256) {
        const int r = i / kMlpH, j = i - r * kMlpH;
        float v = 0.f;
        if (r0 + r < rows) {
            const float hv = h[(size_t)(r0 + r) * kMlpH + j];
            v = da[r0 + r] * ldv<TP>(w2 + j) * (1.f - hv * hv);
            dpre[(size_t)(r0 + r) * kMlpH + j] = v;
        }
        dp[r][j] = v;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += 256) {
        float acc[kMlpDmRows];
#pragma unroll
        for (int r = 0; r < kMlpDmRows; ++r) acc[r] = 0.f;
#pragma unroll 1
        for (int j0 = 0; j0 < kMlpH; j0 += 16) {      // sixteen independent W1 loads in flight
            float w[16];
#pragma unroll
            for (int k = 0; k < 16; ++k) w[k] = ldv<TP>(w1 + (size_t)(j0 + k) * C + c);
#pragma unroll
            for (int k = 0; k < 16; k += 4) {
#pragma unroll
                for (int r = 0; r < kMlpDmRows; ++r) {
                    const float4 d = *reinterpret_cast<const float4*>(&dp[r][j0 + k]);
                    acc[r] = fmaf(d.x, w[k], acc[r]);
                    acc[r] = fmaf(d.y, w[k + 1], acc[r]);
                    acc[r] = fmaf(d.z, w[k + 2], acc[r]);
                    acc[r] = fmaf(d.w, w[k + 3], acc[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kMlpDmRows; ++r)
            if (r0 + r < rows) dm[(size_t)(r0 + r) * C + c] = acc[r] * inv_pp;
    }
}

// partial[chunk][j][c] = sum over the chunk's rows of dpre[r,j] m[r,c]; grid (H/8, kMlpChunks)
__global__ void __launch_bounds__(256) ba_mlp_dw_partial_kernel(const float* __restrict__ dpre,
                                                                const float* __restrict__ m,
                                                                int rows, int C,
                                                                float* __restrict__ partial) {
    const int jg = blockIdx.x * 8, per = (rows + kMlpChunks - 1) / kMlpChunks;
    const int ra = blockIdx.y * per, rb = min(rows, ra + per);
    for (int c = threadIdx.x; c < C; c += 256) {
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
        for (int r = ra; r < rb; ++r) {
            const float mv = m[(size_t)r * C + c];
            const float4 a = *reinterpret_cast<const float4*>(dpre + (size_t)r * kMlpH + jg);
            const float4 b = *reinterpret_cast<const float4*>(dpre + (size_t)r * kMlpH + jg + 4);
            acc[0] = fmaf(a.x, mv, acc[0]); acc[1] = fmaf(a.y, mv, acc[1]);
            acc[2] = fmaf(a.z, mv, acc[2]); acc[3] = fmaf(a.w, mv, acc[3]);
            acc[4] = fmaf(b.x, mv, acc[4]); acc[5] = fmaf(b.y, mv, acc[5]);
            acc[6] = fmaf(b.z, mv, acc[6]); acc[7] = fmaf(b.w, mv, acc[7]);
        }
#pragma unroll
        for (int k = 0; k < 8; ++k)
            partial[((size_t)blockIdx.y * kMlpH + jg + k) * C + c] = acc[k];
    }
}

// one CTA per hidden unit j: dW1[j,:] from the partials; db1[j], dW2[j] (and db2 in CTA 0) by a
// block reduction over the rows, in a fixed order (deterministic)
template <typename TP>
__global__ void __launch_bounds__(256) ba_mlp_dw_final_kernel(const float* __restrict__ partial,
                                                              const float* __restrict__ dpre,
                                                              const float* __restrict__ da,
                                                              const float* __restrict__ h, int rows,
                                                              int C, TP* __restrict__ dw1,
                                                              TP* __restrict__ db1,
                                                              TP* __restrict__ dw2,
                                                              TP* __restrict__ db2) {
    __shared__ float red[3][8];
    const int j = blockIdx.x;
    for (int c = threadIdx.x; c < C; c += 256) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < kMlpChunks; ++k) s += partial[((size_t)k * kMlpH + j) * C + c];
        stv<TP>(dw1 + (size_t)j * C + c, s);
    }
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;
    for (int r = threadIdx.x; r < rows; r += 256) {
        const float d = da[r];
        s1 += dpre[(size_t)r * kMlpH + j];
        s2 = fmaf(d, h[(size_t)r * kMlpH + j], s2);
        s3 += d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        s3 += __shfl_xor_sync(0xffffffffu, s3, o);
    }
    if ((threadIdx.x & 31) == 0) {
        red[0][threadIdx.x >> 5] = s1;
        red[1][threadIdx.x >> 5] = s2;
        red[2][threadIdx.x >> 5] = s3;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        float t1 = 0.f, t2 = 0.f, t3 = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) { t1 += red[0][w]; t2 += red[1][w]; t3 += red[2][w]; }
        stv<TP>(db1 + j, t1);
        stv<TP>(dw2 + j, t2);
        if (j == 0) stv<TP>(db2, t3);
    }
}

}  // namespace htd

using namespace htd;

#define DISPATCH2(dtA, dtB, CALL)                                              \
    do {                                                                       \
        if ((dtA) == HTD_F32 && (dtB) == HTD_F32) { CALL(float, float); }      \
        else if ((dtA) == HTD_F32) { CALL(float, __nv_bfloat16); }             \
        else if ((dtB) == HTD_F32) { CALL(__nv_bfloat16, float); }             \
        else { CALL(__nv_bfloat16, __nv_bfloat16); }                           \
    } while (0)

static bool dt_ok(int d) { return d == HTD_F32 || d == HTD_BF16; }

extern "C" {

int htd_layout_convert(const void* src, int src_dtype, void* dst, int dst_dtype, long long N,
                       int R, int S, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(src_dtype) && dt_ok(dst_dtype), "htd_layout_convert: bad dtype");
    HTD_CHECK_ARG(N >= 0 && R >= 1 && S >= 1, "htd_layout_convert: bad sizes");
    if (N == 0) return HTD_OK;
    HTD_CHECK_ARG(src && dst, "htd_layout_convert: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const size_t es = src_dtype == HTD_F32 ? 4 : 2, ed = dst_dtype == HTD_F32 ? 4 : 2;
    const long long E = (long long)R * S;
    const bool aligned16 = ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0);
    // whole-matrix form: an odd row length keeps the column walk conflict free as it is
    const int pitch = (S & 1) ? S : S + 1;
    const size_t image_off = ((size_t)R * pitch * es + 15) & ~(size_t)15;
    const size_t smem = image_off + (size_t)E * ed;
    if (E <= kFlatMax && E % 8 == 0 && aligned16 && smem <= 100 * 1024 && N < 2147483647LL) {
#define CALL(TS, TD)                                                                          \
    do {                                                                                      \
        HTD_SMEM_OPTIN((transpose_flat_kernel<TS, TD>), 100 * 1024, "htd_layout_convert");    \
        transpose_flat_kernel<TS, TD><<<(unsigned)N, 256, smem, st>>>(                        \
            static_cast<const TS*>(src), static_cast<TD*>(dst), R, S, pitch, (int)image_off,  \
            nullptr, nullptr, 0);                                                             \
    } while (0)
        DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
        HTD_CHECK_LAUNCH("htd_layout_convert");
        return HTD_OK;
    }
    const int tr = (R + 63) / 64, ts = (S + 63) / 64;
    const long long blocks = N * tr * ts;
    HTD_CHECK_ARG(blocks < 2147483647LL, "htd_layout_convert: tensor too large");
    // widest vector each side allows: rows of the source are S long, rows of the destination R
    // long, and every matrix starts E elements after the previous one
    auto width = [&](int row, size_t esz, const void* base) {
        int v = (int)(16 / esz);
        while (v > 1 && (row % v != 0 || E % v != 0 || (uintptr_t)base % (v * esz) != 0)) v >>= 1;
        return v > 4 ? 4 : v;                  // 4 elements per access keeps the tile loops short
    };
    const int vl = width(S, es, src), vs = width(R, ed, dst);
#define LAUNCH(TS, TD, VL, VS)                                                                 \
    transpose_kernel<TS, TD, VL, VS><<<(unsigned)blocks, 256, 0, st>>>(                        \
        static_cast<const TS*>(src), static_cast<TD*>(dst), N, R, S, tr, ts)
#define CALL(TS, TD)                                                                           \
    do {                                                                                       \
        if (vl == 4 && vs == 4) LAUNCH(TS, TD, 4, 4);                                          \
        else if (vl == 4) LAUNCH(TS, TD, 4, 1);                                                \
        else if (vl == 2 && vs == 4) LAUNCH(TS, TD, 2, 4);                                     \
        else if (vs == 4) LAUNCH(TS, TD, 1, 4);                                                \
        else LAUNCH(TS, TD, 1, 1);                                                             \
    } while (0)
    DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
#undef LAUNCH
    HTD_CHECK_LAUNCH("htd_layout_convert");
    return HTD_OK;
}

int htd_roi_flatten(const void* src, int src_dtype, void* dst, int dst_dtype, int K, int PP, int C,
                    const float* bias, const float* rois, int B, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(src_dtype) && dt_ok(dst_dtype) && K >= 0 && PP >= 1 && C >= 1,
                  "htd_roi_flatten: bad arguments");
    HTD_CHECK_ARG(!bias || (rois && B >= 1), "htd_roi_flatten: bias needs rois and B");
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(src && dst, "htd_roi_flatten: null pointer");
    const size_t es = src_dtype == HTD_F32 ? 4 : 2, ed = dst_dtype == HTD_F32 ? 4 : 2;
    const long long E = (long long)PP * C;
    const int pitch = (C & 1) ? C : C + 1;
    const size_t image_off = ((size_t)PP * pitch * es + 15) & ~(size_t)15;
    const size_t smem = image_off + (size_t)E * ed;
    HTD_CHECK_ARG(E <= kFlatMax && E % 8 == 0 && smem <= 100 * 1024 &&
                      ((uintptr_t)src % 16 == 0) && ((uintptr_t)dst % 16 == 0),
                  "htd_roi_flatten: a RoI map of %d bins x %d channels does not fit the "
                  "whole-matrix kernel (use htd_layout_convert)", PP, C);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TS, TD)                                                                          \
    do {                                                                                      \
        HTD_SMEM_OPTIN((transpose_flat_kernel<TS, TD>), 100 * 1024, "htd_roi_flatten");       \
        transpose_flat_kernel<TS, TD><<<(unsigned)K, 256, smem, st>>>(                        \
            static_cast<const TS*>(src), static_cast<TD*>(dst), PP, C, pitch, (int)image_off, \
            bias, rois, B);                                                                   \
    } while (0)
    DISPATCH2(src_dtype, dst_dtype, CALL);
#undef CALL
    HTD_CHECK_LAUNCH("htd_roi_flatten");
    return HTD_OK;
}

int htd_ba_bin_mean(const void* x, int x_dtype, long long N, int PP, int C, float* mean,
                    htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(x_dtype) && N >= 0 && PP >= 1 && C >= 1, "htd_ba_bin_mean: bad arguments");
    if (N == 0) return HTD_OK;
    HTD_CHECK_ARG(x && mean && N < 2147483647LL, "htd_ba_bin_mean: null pointer / too large");
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = C % 8 == 0 && C <= 2048 && ((uintptr_t)x % 16 == 0);
    const float inv = 1.f / (float)PP;
    if (x_dtype == HTD_F32) {
        if (vec) bin_sum_kernel<float><<<(unsigned)N, 256, 0, st>>>(static_cast<const float*>(x), N, PP, C, inv, mean);
        else bin_mean_kernel<float><<<(unsigned)N, 256, 0, st>>>(static_cast<const float*>(x), N, PP, C, mean);
    } else {
        if (vec) bin_sum_kernel<__nv_bfloat16><<<(unsigned)N, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), N, PP, C, inv, mean);
        else bin_mean_kernel<__nv_bfloat16><<<(unsigned)N, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), N, PP, C, mean);
    }
    HTD_CHECK_LAUNCH("htd_ba_bin_mean");
    return HTD_OK;
}

int htd_ba_fuse_fwd(const void* R, int r_dtype, const float* logits, int L, int K, int P, int C,
                    int ring_edge, const void* add, int add_dtype, const float* bias,
                    const float* rois, int B, float* w, void* out, int out_dtype,
                    htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(r_dtype) && dt_ok(out_dtype) && (!add || dt_ok(add_dtype)),
                  "htd_ba_fuse_fwd: bad dtype");
    HTD_CHECK_ARG(L >= 1 && L <= HTD_MAX_LEVELS && K >= 0 && P >= 1 && P <= HTD_MAX_POOLED &&
                      C >= 8 && C % 8 == 0,
                  "htd_ba_fuse_fwd: bad sizes L=%d K=%d P=%d C=%d", L, K, P, C);
    HTD_CHECK_ARG(!bias || (rois && B >= 1), "htd_ba_fuse_fwd: bias needs rois and B");
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(R && logits && w && out, "htd_ba_fuse_fwd: null pointer");
    // the residual `add` shares the dtype of R in every caller; keep the dispatch 2-way
    HTD_CHECK_ARG(!add || add_dtype == r_dtype, "htd_ba_fuse_fwd: add must have the dtype of R");
    FuseParams p{R, logits, L, K, P, C, ring_edge, B, add, bias, rois, w, out};
    const long long total = (long long)K * P * P * (C / 8);
    const unsigned blocks = (unsigned)((total + 255) / 256 < 148LL * 16 ? (total + 255) / 256 : 148LL * 16);
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TR, TO) ba_fuse_fwd_kernel<TR, TR, TO><<<blocks, 256, 0, st>>>(p)
    DISPATCH2(r_dtype, out_dtype, CALL);
#undef CALL
    HTD_CHECK_LAUNCH("htd_ba_fuse_fwd");
    return HTD_OK;
}

int htd_ba_fuse_bwd(const void* R, int r_dtype, const void* dout, int dout_dtype, const float* w,
                    int L, int K, int PP, int C, float* da, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(r_dtype) && dt_ok(dout_dtype), "htd_ba_fuse_bwd: bad dtype");
    HTD_CHECK_ARG(L >= 1 && L <= HTD_MAX_LEVELS && K >= 0 && PP >= 1 && C >= 8 && C % 8 == 0,
                  "htd_ba_fuse_bwd: bad sizes");
    if (K == 0) return HTD_OK;
    HTD_CHECK_ARG(R && dout && w && da, "htd_ba_fuse_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TR, TG)                                                                          \
    ba_fuse_bwd_kernel<TR, TG><<<K, 256, 0, st>>>(static_cast<const TR*>(R),                  \
                                                  static_cast<const TG*>(dout), w, L, K, PP, C, da)
    DISPATCH2(r_dtype, dout_dtype, CALL);
#undef CALL
    HTD_CHECK_LAUNCH("htd_ba_fuse_bwd");
    return HTD_OK;
}

int htd_bias_grad(const void* g, int g_dtype, const float* rois, int K, int PP, int C, int B,
                  float* partial, float* dbias, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(g_dtype) && K >= 0 && PP >= 1 && C >= 1 && B >= 1,
                  "htd_bias_grad: bad arguments");
    HTD_CHECK_ARG(dbias && (K == 0 || (g && rois && partial)), "htd_bias_grad: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (C % 8 == 0 && C <= 2048 && ((uintptr_t)g % 16 == 0) && K > 0) {
        // per-RoI bin sums [K, C] into the workspace, then the rows of each image
        if (g_dtype == HTD_F32)
            bin_sum_kernel<float><<<(unsigned)K, 256, 0, st>>>(static_cast<const float*>(g), K, PP, C, 1.f, partial);
        else
            bin_sum_kernel<__nv_bfloat16><<<(unsigned)K, 256, 0, st>>>(
                static_cast<const __nv_bfloat16*>(g), K, PP, C, 1.f, partial);
        HTD_CHECK_LAUNCH("htd_bias_grad(bin sums)");
        rows_by_image_kernel<<<dim3((C + 31) / 32, B), 1024, 0, st>>>(partial, rois, K, C, B, dbias);
        HTD_CHECK_LAUNCH("htd_bias_grad(rows)");
        return HTD_OK;
    }
    const int nblk = (K + kBiasRois - 1) / kBiasRois;
    if (nblk > 0) {
        if (g_dtype == HTD_F32)
            bias_grad_partial_kernel<float><<<nblk, 256, 0, st>>>(static_cast<const float*>(g), rois,
                                                                  K, PP, C, B, partial);
        else
            bias_grad_partial_kernel<__nv_bfloat16><<<nblk, 256, 0, st>>>(
                static_cast<const __nv_bfloat16*>(g), rois, K, PP, C, B, partial);
        HTD_CHECK_LAUNCH("htd_bias_grad(partial)");
    }
    bias_grad_final_kernel<<<(B * C + 255) / 256, 256, 0, st>>>(partial, nblk, B * C, dbias);
    HTD_CHECK_LAUNCH("htd_bias_grad(final)");
    return HTD_OK;
}


int htd_ba_mlp_supported(int C, int H) {
    return H == kMlpH && C >= 32 && C <= kMlpMaxC && C % 32 == 0;
}

long long htd_ba_mlp_workspace_floats(long long rows, int C) {
    return rows * kMlpH + (long long)kMlpChunks * kMlpH * C;
}

int htd_ba_mlp_fwd(const float* m, long long rows, int C, int H, const void* w1, const void* b1,
                   const void* w2, const void* b2, int p_dtype, float* h, float* logits,
                   htd_stream_t stream) {
    HTD_CHECK_ARG(htd_ba_mlp_supported(C, H) && dt_ok(p_dtype) && rows >= 0 && rows < 2147483647LL,
                  "htd_ba_mlp_fwd: unsupported sizes C=%d H=%d", C, H);
    if (rows == 0) return HTD_OK;
    HTD_CHECK_ARG(m && w1 && b1 && w2 && b2 && h && logits, "htd_ba_mlp_fwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned grid = (unsigned)((rows + kMlpFwdRows - 1) / kMlpFwdRows);
    if (p_dtype == HTD_F32)
        ba_mlp_fwd_kernel<float><<<grid, 2 * kMlpH, 0, st>>>(
            m, (int)rows, C, static_cast<const float*>(w1), static_cast<const float*>(b1),
            static_cast<const float*>(w2), static_cast<const float*>(b2), h, logits);
    else
        ba_mlp_fwd_kernel<__nv_bfloat16><<<grid, 2 * kMlpH, 0, st>>>(
            m, (int)rows, C, static_cast<const __nv_bfloat16*>(w1),
            static_cast<const __nv_bfloat16*>(b1), static_cast<const __nv_bfloat16*>(w2),
            static_cast<const __nv_bfloat16*>(b2), h, logits);
    HTD_CHECK_LAUNCH("htd_ba_mlp_fwd");
    return HTD_OK;
}

int htd_ba_mlp_bwd(const float* da, const float* h, const float* m, long long rows, int C, int H,
                   const void* w1, const void* w2, int p_dtype, float inv_pp, float* dm,
                   float* workspace, void* dw1, void* db1, void* dw2, void* db2,
                   htd_stream_t stream) {
    HTD_CHECK_ARG(htd_ba_mlp_supported(C, H) && dt_ok(p_dtype) && rows >= 1 && rows < 2147483647LL,
                  "htd_ba_mlp_bwd: unsupported sizes C=%d H=%d rows=%lld", C, H, rows);
    HTD_CHECK_ARG(da && h && m && w1 && w2 && dm && workspace && dw1 && db1 && dw2 && db2,
                  "htd_ba_mlp_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    float* dpre = workspace;
    float* partial = workspace + rows * kMlpH;
    const unsigned grid = (unsigned)((rows + kMlpDmRows - 1) / kMlpDmRows);
#define CALL(TP)                                                                               \
    do {                                                                                       \
        ba_mlp_dm_kernel<TP><<<grid, 256, 0, st>>>(da, h, (int)rows, C,                        \
                                                   static_cast<const TP*>(w1),                 \
                                                   static_cast<const TP*>(w2), inv_pp, dpre, dm); \
        ba_mlp_dw_partial_kernel<<<dim3(kMlpH / 8, kMlpChunks), 256, 0, st>>>(dpre, m, (int)rows, \
                                                                             C, partial);      \
        ba_mlp_dw_final_kernel<TP><<<kMlpH, 256, 0, st>>>(                                     \
            partial, dpre, da, h, (int)rows, C, static_cast<TP*>(dw1), static_cast<TP*>(db1),  \
            static_cast<TP*>(dw2), static_cast<TP*>(db2));                                     \
    } while (0)
    if (p_dtype == HTD_F32) CALL(float);
    else CALL(__nv_bfloat16);
#undef CALL
    HTD_CHECK_LAUNCH("htd_ba_mlp_bwd");
    return HTD_OK;
}

}  // extern "C"
