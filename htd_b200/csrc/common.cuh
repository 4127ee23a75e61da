// Shared helpers for the htd_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <atomic>

#include "../../include/htd_b200.h"

namespace htd {

void set_error(const char* fmt, ...);

// SM count of the CURRENT device, cached per device (a process may drive several GPUs)
inline int sm_count() {
    static std::atomic<int> cache[64];
    int dev = 0;
    cudaGetDevice(&dev);
    int v = cache[dev & 63].load(std::memory_order_relaxed);
    if (v <= 0) {
        if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0)
            v = 148;
        cache[dev & 63].store(v, std::memory_order_relaxed);
    }
    return v;
}

#define HTD_CHECK_ARG(cond, ...)                                   \
    do {                                                           \
        if (!(cond)) {                                             \
            htd::set_error(__VA_ARGS__);                           \
            return HTD_ERR_INVALID_ARGUMENT;                       \
        }                                                          \
    } while (0)

// Opt a kernel in to more than 48 KB of dynamic shared memory ONCE PER DEVICE (the attribute is
// per device: a process-wide flag would leave a second GPU driven by the same process without it)
// and thread-safely (autograd worker threads); a failure is reported, not ignored.  Pass a
// templated kernel in parentheses.
#define HTD_SMEM_OPTIN(func, bytes, who)                                                         \
    do {                                                                                         \
        static std::atomic<unsigned long long> done__{0ull};                                     \
        int dev__ = 0;                                                                           \
        cudaGetDevice(&dev__);                                                                   \
        const unsigned long long bit__ = 1ull << (dev__ & 63);                                   \
        if (!(done__.load(std::memory_order_acquire) & bit__)) {                                 \
            cudaError_t oe__ = cudaFuncSetAttribute(                                             \
                func, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(bytes));                \
            if (oe__ != cudaSuccess) {                                                           \
                htd::set_error("%s: cannot reserve %d B of shared memory: %s", who, (int)(bytes), \
                               cudaGetErrorString(oe__));                                        \
                return HTD_ERR_CUDA;                                                             \
            }                                                                                    \
            done__.fetch_or(bit__, std::memory_order_release);                                   \
        }                                                                                        \
    } while (0)

#define HTD_CHECK_LAUNCH(name)                                                     \
    do {                                                                           \
        cudaError_t e__ = cudaGetLastError();                                      \
        if (e__ != cudaSuccess) {                                                  \
            htd::set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
            return HTD_ERR_CUDA;                                                   \
        }                                                                          \
    } while (0)

constexpr int kWarp = 32;

// scalar load / store of a tensor element as float (fp32 or bf16 tensors)
template <typename T>
__device__ __forceinline__ float ldv(const T* p);
template <>
__device__ __forceinline__ float ldv<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldv<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__device__ __forceinline__ void stv(T* p, float v);
template <>
__device__ __forceinline__ void stv<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stv<__nv_bfloat16>(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }


// ---------------------------------------------------------------------------------------------
// A warp covers 256 channels of one pixel with 128-bit accesses; a lane owns 8 channels.  The
// lane->channel map is a property of the KERNEL (all tensors it touches must agree):
//   kSplit = true  (all-fp32 kernels): channels {4l..4l+3} and {128+4l..128+4l+3}, so each of
//                  the two float4 accesses of a warp is one contiguous 512 B run;
//   kSplit = false (any bf16 operand): channels {8l..8l+7}; bf16 moves one uint4 (512 B per
//                  warp), fp32 moves two float4 at 8l and 8l+4.
// `p` points at the 256-channel chunk, `nch` = channels left in it (multiple of 8) masks the
// tail when C is not a multiple of 256.
// ---------------------------------------------------------------------------------------------
template <bool kSplit>
__device__ __forceinline__ int lane_chan(int lane, int e) {
    return kSplit ? ((e < 4) ? lane * 4 + e : 128 + lane * 4 + (e - 4)) : lane * 8 + e;
}

template <typename T, bool kSplit>
struct Vec8;

template <bool kSplit>
struct Vec8<float, kSplit> {
    static __device__ __forceinline__ void load(const float* __restrict__ p, int lane, int nch,
                                                float (&v)[8]) {
        const int c0 = kSplit ? lane * 4 : lane * 8, c1 = kSplit ? 128 + lane * 4 : lane * 8 + 4;
        float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
        if (c0 < nch) a = __ldg(reinterpret_cast<const float4*>(p + c0));
        if (c1 < nch) b = __ldg(reinterpret_cast<const float4*>(p + c1));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
        v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
    }
    static __device__ __forceinline__ void store(float* __restrict__ p, int lane, int nch,
                                                 const float (&v)[8]) {
        const int c0 = kSplit ? lane * 4 : lane * 8, c1 = kSplit ? 128 + lane * 4 : lane * 8 + 4;
        if (c0 < nch) *reinterpret_cast<float4*>(p + c0) = make_float4(v[0], v[1], v[2], v[3]);
        if (c1 < nch) *reinterpret_cast<float4*>(p + c1) = make_float4(v[4], v[5], v[6], v[7]);
    }
};

template <bool kSplit>
struct Vec8<__nv_bfloat16, kSplit> {
    static_assert(!kSplit, "bf16 tensors use the contiguous lane->channel map");
    static __device__ __forceinline__ void load(const __nv_bfloat16* __restrict__ p, int lane,
                                                int nch, float (&v)[8]) {
        const int c0 = lane * 8;
        uint4 u = make_uint4(0u, 0u, 0u, 0u);
        if (c0 < nch) u = __ldg(reinterpret_cast<const uint4*>(p + c0));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            v[2 * i] = __uint_as_float(w[i] << 16);
            v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
        }
    }
    static __device__ __forceinline__ void store(__nv_bfloat16* __restrict__ p, int lane, int nch,
                                                 const float (&v)[8]) {
        const int c0 = lane * 8;
        if (c0 >= nch) return;
        uint32_t w[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
            w[i] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(p + c0) = make_uint4(w[0], w[1], w[2], w[3]);
    }
};

template <typename A, typename B>
struct SplitMap {
    static constexpr bool value = false;
};
template <>
struct SplitMap<float, float> {
    static constexpr bool value = true;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 1-D bulk async copy global -> shared (SASS: UBLKCP), completion counted in bytes on `bar`.
// src/dst 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

inline size_t dtype_size(int dt) { return dt == HTD_BF16 ? 2 : 4; }

}  // namespace htd
