// FPN glue on channels-last maps - the producer side of the RoI head's pyramid (SURVEY.md section 8
// row f4; mmdet/models/necks/fpn.py:165-216).  The 1x1 lateral convolutions are rows-by-channels
// products and run on the dense tcgen05 kernel (csrc/dense_gemm.cu, kind NT, bias epilogue); here:
//   * top-down merge  lat[i-1] += interpolate(lat[i], size=lat[i-1].shape, mode='nearest')
//     (fpn.py:187-190) and its backward (the gather form of the nearest-neighbour scatter);
//   * the extra level  max_pool2d(out, 1, stride=2)  (fpn.py:201) = every other pixel, and its
//     backward.
// All maps are [B, H, W, C] in memory (channels-last), fp32 or bf16, C a multiple of 8 (bf16) / 4.
#include "common.cuh"

namespace htd {

// ATen's nearest source index (UpSample.h nearest_neighbor_compute_source_index): scale =
// (float)in / out, src = min((int)floorf(dst * scale), in - 1)
__device__ __forceinline__ int nearest_src(int dst, float scale, int in) {
    return min((int)floorf(__fmul_rn((float)dst, scale)), in - 1);
}

template <typename T>
struct Pack;                      // 16 bytes of T
template <>
struct Pack<float> {
    static constexpr int n = 4;
    static __device__ __forceinline__ uint4 add(uint4 a, uint4 b) {
        float4 x = *reinterpret_cast<float4*>(&a), y = *reinterpret_cast<float4*>(&b);
        float4 r = make_float4(x.x + y.x, x.y + y.y, x.z + y.z, x.w + y.w);
        return *reinterpret_cast<uint4*>(&r);
    }
};
template <>
struct Pack<__nv_bfloat16> {
    static constexpr int n = 8;
    static __device__ __forceinline__ uint32_t add2(uint32_t a, uint32_t b) {
        // bf16 + bf16 in fp32, one rounding: what ATen's bf16 add does
        const float lo = __uint_as_float(a << 16) + __uint_as_float(b << 16);
        const float hi = __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u);
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
    static __device__ __forceinline__ uint4 add(uint4 a, uint4 b) {
        return make_uint4(add2(a.x, b.x), add2(a.y, b.y), add2(a.z, b.z), add2(a.w, b.w));
    }
};

// out[b, y, x, :] = fine[b, y, x, :] + coarse[b, sy(y), sx(x), :]; one thread = 16 bytes
template <typename T>
__global__ void fpn_topdown_fwd_kernel(const T* __restrict__ fine, const T* __restrict__ coarse,
                                       T* __restrict__ out, int B, int Hf, int Wf, int Hc, int Wc,
                                       int C, float sy, float sx) {
    const int cv = C / Pack<T>::n;
    const long long total = (long long)B * Hf * Wf * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        long long r = i / cv;
        const int x = (int)(r % Wf);
        r /= Wf;
        const int y = (int)(r % Hf), b = (int)(r / Hf);
        const int ys = nearest_src(y, sy, Hc), xs = nearest_src(x, sx, Wc);
        const uint4 f = reinterpret_cast<const uint4*>(fine)[i];
        const uint4 u = reinterpret_cast<const uint4*>(coarse)[(((long long)b * Hc + ys) * Wc + xs) * cv + c];
        reinterpret_cast<uint4*>(out)[i] = Pack<T>::add(f, u);
    }
}

// dcoarse[b, ys, xs, :] = sum of dout over the fine pixels whose nearest source is (ys, xs); the
// candidates are the fine pixels around (ys / sy, xs / sx), each checked with the forward formula,
// accumulated in fp32 in ascending (y, x) order (deterministic)
template <typename T>
__global__ void fpn_topdown_bwd_kernel(const T* __restrict__ dout, T* __restrict__ dcoarse, int B,
                                       int Hf, int Wf, int Hc, int Wc, int C, float sy, float sx,
                                       int ry, int rx) {
    constexpr int n = Pack<T>::n;
    const int cv = C / n;
    const long long total = (long long)B * Hc * Wc * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        long long r = i / cv;
        const int xs = (int)(r % Wc);
        r /= Wc;
        const int ys = (int)(r % Hc), b = (int)(r / Hc);
        float acc[n];
#pragma unroll
        for (int e = 0; e < n; ++e) acc[e] = 0.f;
        const int y0 = max((int)((float)ys / sy) - 1, 0), x0 = max((int)((float)xs / sx) - 1, 0);
        for (int y = y0; y < min(y0 + ry, Hf); ++y) {
            if (nearest_src(y, sy, Hc) != ys) continue;
            for (int x = x0; x < min(x0 + rx, Wf); ++x) {
                if (nearest_src(x, sx, Wc) != xs) continue;
                const uint4 v = reinterpret_cast<const uint4*>(dout)[(((long long)b * Hf + y) * Wf + x) * cv + c];
                if (n == 4) {
                    const float4 f = *reinterpret_cast<const float4*>(&v);
                    acc[0] += f.x; acc[1] += f.y; acc[2] += f.z; acc[3] += f.w;
                } else {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        acc[2 * k] += __uint_as_float(w[k] << 16);
                        acc[2 * k + 1] += __uint_as_float(w[k] & 0xffff0000u);
                    }
                }
            }
        }
        uint4 o;
        if (n == 4) {
            float4 f = make_float4(acc[0], acc[1], acc[2], acc[3]);
            o = *reinterpret_cast<uint4*>(&f);
        } else {
            uint32_t w[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                __nv_bfloat162 h = __floats2bfloat162_rn(acc[2 * k], acc[2 * k + 1]);
                w[k] = *reinterpret_cast<uint32_t*>(&h);
            }
            o = make_uint4(w[0], w[1], w[2], w[3]);
        }
        reinterpret_cast<uint4*>(dcoarse)[i] = o;
    }
}

// forward: out[b, y, x, :] = in[b, 2y, 2x, :]  (out is [B, Ho, Wo, C], Ho = (H - 1) / 2 + 1)
// backward (scatter != 0): `in` is the gradient of out, `out` the [B, H, W, C] gradient of the input:
// out[b, y, x, :] = (y, x both even) ? in[b, y/2, x/2, :] : 0
__global__ void fpn_subsample_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int B,
                                     int H, int W, int Ho, int Wo, int cv, int scatter) {
    const int oh = scatter ? H : Ho, ow = scatter ? W : Wo;
    const long long total = (long long)B * oh * ow * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int c = (int)(i % cv);
        long long r = i / cv;
        const int x = (int)(r % ow);
        r /= ow;
        const int y = (int)(r % oh), b = (int)(r / oh);
        if (!scatter) {
            out[i] = in[(((long long)b * H + 2 * y) * W + 2 * x) * cv + c];
        } else {
            uint4 v = make_uint4(0u, 0u, 0u, 0u);
            if (!(y & 1) && !(x & 1)) v = in[(((long long)b * Ho + y / 2) * Wo + x / 2) * cv + c];
            out[i] = v;
        }
    }
}

static unsigned grid_for(long long total, int threads) {
    const long long want = (total + threads - 1) / threads;
    const long long cap = 148LL * 16;
    return (unsigned)(want < 1 ? 1 : want < cap ? want : cap);
}

}  // namespace htd

using namespace htd;

extern "C" {

int htd_fpn_topdown_fwd(const void* fine, const void* coarse, void* out, int dtype, int B, int Hf,
                        int Wf, int Hc, int Wc, int C, htd_stream_t stream) {
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_fpn_topdown_fwd: bad dtype %d", dtype);
    HTD_CHECK_ARG(B >= 0 && Hf >= 1 && Wf >= 1 && Hc >= 1 && Wc >= 1 && C >= 1 &&
                      C % (dtype == HTD_BF16 ? 8 : 4) == 0,
                  "htd_fpn_topdown_fwd: bad sizes B=%d fine %dx%d coarse %dx%d C=%d", B, Hf, Wf, Hc, Wc, C);
    if (B == 0) return HTD_OK;
    HTD_CHECK_ARG(fine && coarse && out, "htd_fpn_topdown_fwd: null pointer");
    const float sy = (float)Hc / (float)Hf, sx = (float)Wc / (float)Wf;
    const long long total = (long long)B * Hf * Wf * (C / (dtype == HTD_BF16 ? 8 : 4));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HTD_F32)
        fpn_topdown_fwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(
            static_cast<const float*>(fine), static_cast<const float*>(coarse), static_cast<float*>(out),
            B, Hf, Wf, Hc, Wc, C, sy, sx);
    else
        fpn_topdown_fwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(fine), static_cast<const __nv_bfloat16*>(coarse),
            static_cast<__nv_bfloat16*>(out), B, Hf, Wf, Hc, Wc, C, sy, sx);
    HTD_CHECK_LAUNCH("htd_fpn_topdown_fwd");
    return HTD_OK;
}

int htd_fpn_topdown_bwd(const void* dout, void* dcoarse, int dtype, int B, int Hf, int Wf, int Hc,
                        int Wc, int C, htd_stream_t stream) {
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_fpn_topdown_bwd: bad dtype %d", dtype);
    HTD_CHECK_ARG(B >= 0 && Hf >= 1 && Wf >= 1 && Hc >= 1 && Wc >= 1 && C >= 1 &&
                      C % (dtype == HTD_BF16 ? 8 : 4) == 0,
                  "htd_fpn_topdown_bwd: bad sizes B=%d fine %dx%d coarse %dx%d C=%d", B, Hf, Wf, Hc, Wc, C);
    if (B == 0) return HTD_OK;
    HTD_CHECK_ARG(dout && dcoarse, "htd_fpn_topdown_bwd: null pointer");
    const float sy = (float)Hc / (float)Hf, sx = (float)Wc / (float)Wf;
    // fine pixels per coarse pixel along an axis: at most ceil(Hf / Hc) + 1; + 1 for the start guess
    const int ry = (Hf + Hc - 1) / Hc + 3, rx = (Wf + Wc - 1) / Wc + 3;
    const long long total = (long long)B * Hc * Wc * (C / (dtype == HTD_BF16 ? 8 : 4));
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HTD_F32)
        fpn_topdown_bwd_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(
            static_cast<const float*>(dout), static_cast<float*>(dcoarse), B, Hf, Wf, Hc, Wc, C, sy, sx,
            ry, rx);
    else
        fpn_topdown_bwd_kernel<__nv_bfloat16><<<grid_for(total, 256), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(dout), static_cast<__nv_bfloat16*>(dcoarse), B, Hf, Wf, Hc,
            Wc, C, sy, sx, ry, rx);
    HTD_CHECK_LAUNCH("htd_fpn_topdown_bwd");
    return HTD_OK;
}

int htd_fpn_subsample(const void* in, void* out, int dtype, int B, int H, int W, int C, int backward,
                      htd_stream_t stream) {
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_fpn_subsample: bad dtype %d", dtype);
    HTD_CHECK_ARG(B >= 0 && H >= 1 && W >= 1 && C >= 1 && C % (dtype == HTD_BF16 ? 8 : 4) == 0,
                  "htd_fpn_subsample: bad sizes B=%d %dx%d C=%d", B, H, W, C);
    if (B == 0) return HTD_OK;
    HTD_CHECK_ARG(in && out, "htd_fpn_subsample: null pointer");
    const int Ho = (H - 1) / 2 + 1, Wo = (W - 1) / 2 + 1;
    const int cv = C / (dtype == HTD_BF16 ? 8 : 4);
    const long long total = (long long)B * (backward ? H : Ho) * (backward ? W : Wo) * cv;
    fpn_subsample_kernel<<<grid_for(total, 256), 256, 0, (cudaStream_t)stream>>>(
        static_cast<const uint4*>(in), static_cast<uint4*>(out), B, H, W, Ho, Wo, cv, backward ? 1 : 0);
    HTD_CHECK_LAUNCH("htd_fpn_subsample");
    return HTD_OK;
}

}  // extern "C"
