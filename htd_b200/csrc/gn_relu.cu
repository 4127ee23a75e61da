// Fused GroupNorm + ReLU (forward and backward) for the regression conv tower of HTDBBoxHead.
//
// Reference path: mmcv ConvModule(conv -> GN(36 groups) -> ReLU) x3 inside HTDBBoxHead.convs
// (htd_bbox_head.py:75-113, applied at :186) on [P, 576, 7, 7] RoI tensors.  ATen runs this as
// RowwiseMoments + an elementwise affine pass + ReLU forward, and ComputeInternalGradients +
// GroupNormBackward + two parameter reductions + threshold_backward in backward (~0.75 ms per
// step in the round-1 profile, 14 MB per layer).  Here: ONE forward launch and TWO backward
// launches per layer, channels-last, two-pass variance; statistics in fp32 for bf16 tensors and in
// fp64 for fp32 tensors (the parity configuration).
//
// Layout: x, y, dy, dx are [N, HW, C] (a torch [N, C, H, W] tensor in channels_last memory
// format); group g owns channels [g*cpg, (g+1)*cpg), cpg = C / G a multiple of 8.
#include "common.cuh"

namespace htd {

template <typename T>
__device__ __forceinline__ void ld8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void ld8<float>(const float* p, float (&v)[8]) {
    const float4 a = *reinterpret_cast<const float4*>(p);
    const float4 b = *reinterpret_cast<const float4*>(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void ld8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 u = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        v[2 * i] = __uint_as_float(w[i] << 16);
        v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
    }
}
template <typename T>
__device__ __forceinline__ void st8(T* p, const float (&v)[8]);
template <>
__device__ __forceinline__ void st8<float>(float* p, const float (&v)[8]) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <>
__device__ __forceinline__ void st8<__nv_bfloat16>(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
        w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
}

constexpr int kGnWarps = 8;

// Accumulator type of the group statistics: the fp32 configuration is the 1e-5 parity
// configuration, and GroupNorm backward after an average pool is a small residual of large terms
//  - its sums and the final combination run in fp64 there; bf16 uses fp32.
template <typename T> struct GnAcc { typedef float type; };
template <> struct GnAcc<float> { typedef double type; };

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// one warp per (n, g); work items of a group: (pixel, 8-channel chunk), strided over the lanes
template <typename T>
__global__ void __launch_bounds__(kGnWarps * 32) gn_relu_fwd_kernel(
    const T* __restrict__ x, const T* __restrict__ gamma, const T* __restrict__ beta,
    int N, int HW, int C, int G, float eps, T* __restrict__ y,
    typename GnAcc<T>::type* __restrict__ mean_out, typename GnAcc<T>::type* __restrict__ rstd_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kGnWarps + warp;
    if (item >= (long long)N * G) return;
    const int n = (int)(item / G), g = (int)(item % G);
    const int cpg = C / G, chunks = cpg / 8, items = HW * chunks;
    const T* xg = x + (size_t)n * HW * C + (size_t)g * cpg;
    typedef typename GnAcc<T>::type acc_t;
    const acc_t inv_m = (acc_t)1 / (acc_t)(HW * cpg);
    acc_t s = 0;
    for (int i = lane; i < items; i += 32) {
        float v[8];
        ld8<T>(xg + (size_t)(i / chunks) * C + (i % chunks) * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) s += (acc_t)v[e];
    }
    const acc_t mean_a = warp_sum(s) * inv_m;
    acc_t q = 0;
    for (int i = lane; i < items; i += 32) {
        float v[8];
        ld8<T>(xg + (size_t)(i / chunks) * C + (i % chunks) * 8, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) { const acc_t d = (acc_t)v[e] - mean_a; q += d * d; }
    }
    const acc_t rstd_a = (acc_t)1 / sqrt(warp_sum(q) * inv_m + (acc_t)eps);
    if (lane == 0) { mean_out[item] = mean_a; rstd_out[item] = rstd_a; }
    T* yg = y + (size_t)n * HW * C + (size_t)g * cpg;
    for (int i = lane; i < items; i += 32) {
        const int pix = i / chunks, c0 = (i % chunks) * 8;
        float v[8];
        ld8<T>(xg + (size_t)pix * C + c0, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = g * cpg + c0 + e;
            const acc_t pre = ((acc_t)v[e] - mean_a) * rstd_a * (acc_t)ldv<T>(gamma + c) + (acc_t)ldv<T>(beta + c);
            v[e] = fmaxf((float)pre, 0.f);
        }
        st8<T>(yg + (size_t)pix * C + c0, v);
    }
}

// backward, phase 1: dx and the per-(n, c) partial sums of dgamma / dbeta.
// part: [2][N][C] fp32 (dgamma partials, then dbeta partials).
template <typename T>
__global__ void __launch_bounds__(kGnWarps * 32) gn_relu_bwd_kernel(
    const T* __restrict__ x, const T* __restrict__ dy,
    const typename GnAcc<T>::type* __restrict__ mean_in,
    const typename GnAcc<T>::type* __restrict__ rstd_in, const T* __restrict__ gamma,
    const T* __restrict__ beta, int N, int HW, int C, int G, T* __restrict__ dx,
    float* __restrict__ part) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long item = (long long)blockIdx.x * kGnWarps + warp;
    if (item >= (long long)N * G) return;
    const int n = (int)(item / G), g = (int)(item % G);
    const int cpg = C / G, chunks = cpg / 8, items = HW * chunks;
    const size_t base = (size_t)n * HW * C + (size_t)g * cpg;
    const T* xg = x + base;
    const T* dg = dy + base;
    typedef typename GnAcc<T>::type acc_t;
    const acc_t mean = mean_in[item], rstd = rstd_in[item];
    const acc_t inv_m = (acc_t)1 / (acc_t)(HW * cpg);
    // pass 1: s1 = sum dy_eff*gamma, s2 = sum dy_eff*gamma*xhat ; per-channel sums for the params.
    // A lane always sees the same channel chunk when 32 % chunks == 0 (cpg = 8, 16, 32): its
    // per-channel partials live in registers and are reduced across the lanes of equal chunk.
    acc_t s1 = 0, s2 = 0;
    float pg[8], pb[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { pg[e] = 0.f; pb[e] = 0.f; }
    const bool fixed_chunk = (32 % chunks) == 0;
    for (int i = lane; i < items; i += 32) {
        const int pix = i / chunks, c0 = (i % chunks) * 8;
        float v[8], d[8];
        ld8<T>(xg + (size_t)pix * C + c0, v);
        ld8<T>(dg + (size_t)pix * C + c0, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = g * cpg + c0 + e;
            const acc_t ga = (acc_t)ldv<T>(gamma + c);
            const acc_t xh = ((acc_t)v[e] - mean) * rstd;
            // same expression as the forward pass: the ReLU mask is reproduced exactly
            const float de = (float)(xh * ga + (acc_t)ldv<T>(beta + c)) > 0.f ? d[e] : 0.f;
            s1 += (acc_t)de * ga;
            s2 += (acc_t)de * ga * xh;
            if (fixed_chunk) { pg[e] += (float)((acc_t)de * xh); pb[e] += de; }
            else {
                atomicAdd(part + ((size_t)0 * N + n) * C + c, (float)((acc_t)de * xh));   // rare shapes only
                atomicAdd(part + ((size_t)1 * N + n) * C + c, de);
            }
        }
    }
    s1 = warp_sum(s1) * inv_m;
    s2 = warp_sum(s2) * inv_m;
    if (fixed_chunk) {
        // lanes with equal (lane % chunks) hold the same channels: butterfly over the other bits
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            for (int o = 16; o >= chunks; o >>= 1) {
                pg[e] += __shfl_xor_sync(0xffffffffu, pg[e], o);
                pb[e] += __shfl_xor_sync(0xffffffffu, pb[e], o);
            }
        }
        if (lane < chunks) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int c = g * cpg + lane * 8 + e;
                part[((size_t)0 * N + n) * C + c] = pg[e];
                part[((size_t)1 * N + n) * C + c] = pb[e];
            }
        }
    }
    // pass 2: dx = rstd * (dy_eff*gamma - s1 - xhat*s2)
    T* dxg = dx + base;
    for (int i = lane; i < items; i += 32) {
        const int pix = i / chunks, c0 = (i % chunks) * 8;
        float v[8], d[8];
        ld8<T>(xg + (size_t)pix * C + c0, v);
        ld8<T>(dg + (size_t)pix * C + c0, d);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = g * cpg + c0 + e;
            const acc_t ga = (acc_t)ldv<T>(gamma + c);
            const acc_t xh = ((acc_t)v[e] - mean) * rstd;
            const float de = (float)(xh * ga + (acc_t)ldv<T>(beta + c)) > 0.f ? d[e] : 0.f;
            v[e] = (float)(rstd * ((acc_t)de * ga - s1 - xh * s2));
        }
        st8<T>(dxg + (size_t)pix * C + c0, v);
    }
}

// backward, phase 2: dgamma[c] = sum_n part[0][n][c], dbeta[c] = sum_n part[1][n][c].
// block = 32 channels x 8 row slices (blockIdx.y selects dgamma / dbeta), fixed summation order.
template <typename T>
__global__ void __launch_bounds__(256) gn_param_reduce_kernel(const float* __restrict__ part, int N,
                                                              int C, T* __restrict__ dgamma,
                                                              T* __restrict__ dbeta) {
    __shared__ float s_part[8][33];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    const float* src = part + (size_t)blockIdx.y * N * C;
    float a = 0.f;
    if (c < C)
        for (int n = ty; n < N; n += 8) a += src[(size_t)n * C + c];
    s_part[ty][tx] = a;
    __syncthreads();
    if (ty == 0 && c < C) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_part[i][tx];
        stv<T>((blockIdx.y == 0 ? dgamma : dbeta) + c, t);
    }
}


// ------------------------------------------------------------------------------------------
// ReLU + global average pool of the last tower conv (htd_bbox_head.py:109-113 ConvModule without a
// norm layer, then avg_pool :188-189): y[n,c] = mean_hw relu(x[n,hw,c]) in one pass, and its
// backward dx[n,hw,c] = x > 0 ? g[n,c] / HW : 0 in one pass (ATen: relu, mean, expand/div,
// threshold_backward - four passes over the largest activation of the head).
// One thread per (n, 8-channel chunk); consecutive threads read consecutive 16 / 32 bytes.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) relu_mean_fwd_kernel(const T* __restrict__ x, int N, int HW,
                                                            int C, T* __restrict__ y) {
    const int chunks = C / 8;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)N * chunks) return;
    const int n = (int)(i / chunks), c0 = (int)(i % chunks) * 8;
    const T* px = x + (size_t)n * HW * C + c0;
    float acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int hw = 0; hw < HW; ++hw) {
        float v[8];
        ld8<T>(px + (size_t)hw * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += fmaxf(v[e], 0.f);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = acc[e] / (float)HW;
    st8<T>(y + (size_t)n * C + c0, acc);
}

template <typename T>
__global__ void __launch_bounds__(256) relu_mean_bwd_kernel(const T* __restrict__ x,
                                                            const T* __restrict__ g, int N, int HW,
                                                            int C, T* __restrict__ dx) {
    const int chunks = C / 8;
    const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
    if (i >= (long long)N * chunks) return;
    const int n = (int)(i / chunks), c0 = (int)(i % chunks) * 8;
    float gv[8];
    ld8<T>(g + (size_t)n * C + c0, gv);
#pragma unroll
    for (int e = 0; e < 8; ++e) gv[e] = gv[e] / (float)HW;
    if (sizeof(T) == 2) {               // ATen rounds the expanded gradient to the tensor dtype
        float t[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) t[e] = __bfloat162float(__float2bfloat16_rn(gv[e]));
#pragma unroll
        for (int e = 0; e < 8; ++e) gv[e] = t[e];
    }
    const T* px = x + (size_t)n * HW * C + c0;
    T* pd = dx + (size_t)n * HW * C + c0;
    for (int hw = 0; hw < HW; ++hw) {
        float v[8], o[8];
        ld8<T>(px + (size_t)hw * C, v);
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = v[e] > 0.f ? gv[e] : 0.f;
        st8<T>(pd + (size_t)hw * C, o);
    }
}

static bool gn_args_ok(int N, int HW, int C, int G) {
    return N >= 0 && HW >= 1 && C >= 8 && G >= 1 && C % G == 0 && (C / G) % 8 == 0;
}

}  // namespace htd

using namespace htd;

extern "C" {

int htd_gn_relu_fwd(const void* x, int dtype, int N, int HW, int C, int G, const void* gamma,
                    const void* beta, float eps, void* y, void* mean, void* rstd,
                    htd_stream_t stream) {
    HTD_CHECK_ARG(gn_args_ok(N, HW, C, G), "htd_gn_relu_fwd: need C %% G == 0 and (C/G) %% 8 == 0 "
                  "(N=%d HW=%d C=%d G=%d)", N, HW, C, G);
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_gn_relu_fwd: bad dtype");
    if (N == 0) return HTD_OK;
    HTD_CHECK_ARG(x && y && gamma && beta && mean && rstd, "htd_gn_relu_fwd: null pointer");
    const long long items = (long long)N * G;
    const unsigned blocks = (unsigned)((items + kGnWarps - 1) / kGnWarps);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HTD_F32)
        gn_relu_fwd_kernel<float><<<blocks, kGnWarps * 32, 0, st>>>(
            static_cast<const float*>(x), static_cast<const float*>(gamma),
            static_cast<const float*>(beta), N, HW, C, G, eps, static_cast<float*>(y),
            static_cast<double*>(mean), static_cast<double*>(rstd));
    else
        gn_relu_fwd_kernel<__nv_bfloat16><<<blocks, kGnWarps * 32, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(gamma),
            static_cast<const __nv_bfloat16*>(beta), N, HW, C, G, eps,
            static_cast<__nv_bfloat16*>(y), static_cast<float*>(mean), static_cast<float*>(rstd));
    HTD_CHECK_LAUNCH("htd_gn_relu_fwd");
    return HTD_OK;
}

int htd_gn_relu_bwd(const void* x, const void* dy, int dtype, const void* mean, const void* rstd,
                    const void* gamma, const void* beta, int N, int HW, int C, int G, void* dx,
                    float* part, void* dgamma, void* dbeta, htd_stream_t stream) {
    HTD_CHECK_ARG(gn_args_ok(N, HW, C, G), "htd_gn_relu_bwd: need C %% G == 0 and (C/G) %% 8 == 0 "
                  "(N=%d HW=%d C=%d G=%d)", N, HW, C, G);
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_gn_relu_bwd: bad dtype");
    HTD_CHECK_ARG(dgamma && dbeta, "htd_gn_relu_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (N > 0) {
        HTD_CHECK_ARG(x && dy && mean && rstd && gamma && beta && dx && part,
                      "htd_gn_relu_bwd: null pointer");
        const int chunks = (C / G) / 8;
        if (32 % chunks != 0) {       // partials are accumulated with atomics for such shapes
            cudaError_t e = cudaMemsetAsync(part, 0, (size_t)2 * N * C * sizeof(float), st);
            if (e != cudaSuccess) { set_error("htd_gn_relu_bwd: memset failed"); return HTD_ERR_CUDA; }
        }
        const long long items = (long long)N * G;
        const unsigned blocks = (unsigned)((items + kGnWarps - 1) / kGnWarps);
        if (dtype == HTD_F32)
            gn_relu_bwd_kernel<float><<<blocks, kGnWarps * 32, 0, st>>>(
                static_cast<const float*>(x), static_cast<const float*>(dy),
                static_cast<const double*>(mean), static_cast<const double*>(rstd),
                static_cast<const float*>(gamma), static_cast<const float*>(beta),
                N, HW, C, G, static_cast<float*>(dx), part);
        else
            gn_relu_bwd_kernel<__nv_bfloat16><<<blocks, kGnWarps * 32, 0, st>>>(
                static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(dy),
                static_cast<const float*>(mean), static_cast<const float*>(rstd),
                static_cast<const __nv_bfloat16*>(gamma), static_cast<const __nv_bfloat16*>(beta), N, HW,
                C, G, static_cast<__nv_bfloat16*>(dx), part);
        HTD_CHECK_LAUNCH("htd_gn_relu_bwd");
    }
    if (dtype == HTD_F32)
        gn_param_reduce_kernel<float><<<dim3((C + 31) / 32, 2), 256, 0, st>>>(
            part, N, C, static_cast<float*>(dgamma), static_cast<float*>(dbeta));
    else
        gn_param_reduce_kernel<__nv_bfloat16><<<dim3((C + 31) / 32, 2), 256, 0, st>>>(
            part, N, C, static_cast<__nv_bfloat16*>(dgamma), static_cast<__nv_bfloat16*>(dbeta));
    HTD_CHECK_LAUNCH("htd_gn_relu_bwd(params)");
    return HTD_OK;
}

int htd_relu_mean_fwd(const void* x, int dtype, int N, int HW, int C, void* y, htd_stream_t stream) {
    HTD_CHECK_ARG(N >= 0 && HW >= 1 && C >= 8 && C % 8 == 0, "htd_relu_mean_fwd: need C %% 8 == 0 "
                  "(N=%d HW=%d C=%d)", N, HW, C);
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_relu_mean_fwd: bad dtype");
    if (N == 0) return HTD_OK;
    HTD_CHECK_ARG(x && y, "htd_relu_mean_fwd: null pointer");
    const long long items = (long long)N * (C / 8);
    const unsigned blocks = (unsigned)((items + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HTD_F32)
        relu_mean_fwd_kernel<float><<<blocks, 256, 0, st>>>(static_cast<const float*>(x), N, HW, C,
                                                            static_cast<float*>(y));
    else
        relu_mean_fwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), N, HW, C, static_cast<__nv_bfloat16*>(y));
    HTD_CHECK_LAUNCH("htd_relu_mean_fwd");
    return HTD_OK;
}

int htd_relu_mean_bwd(const void* x, const void* g, int dtype, int N, int HW, int C, void* dx,
                      htd_stream_t stream) {
    HTD_CHECK_ARG(N >= 0 && HW >= 1 && C >= 8 && C % 8 == 0, "htd_relu_mean_bwd: need C %% 8 == 0 "
                  "(N=%d HW=%d C=%d)", N, HW, C);
    HTD_CHECK_ARG(dtype == HTD_F32 || dtype == HTD_BF16, "htd_relu_mean_bwd: bad dtype");
    if (N == 0) return HTD_OK;
    HTD_CHECK_ARG(x && g && dx, "htd_relu_mean_bwd: null pointer");
    const long long items = (long long)N * (C / 8);
    const unsigned blocks = (unsigned)((items + 255) / 256);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == HTD_F32)
        relu_mean_bwd_kernel<float><<<blocks, 256, 0, st>>>(
            static_cast<const float*>(x), static_cast<const float*>(g), N, HW, C,
            static_cast<float*>(dx));
    else
        relu_mean_bwd_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), static_cast<const __nv_bfloat16*>(g), N, HW, C,
            static_cast<__nv_bfloat16*>(dx));
    HTD_CHECK_LAUNCH("htd_relu_mean_bwd");
    return HTD_OK;
}

}  // extern "C"
