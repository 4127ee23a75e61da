// PGraph support kernels: sorted-space plan, row packing, IoU adjacency, masked softmax and its
// backward, per-group transpose, segment column sums.  The contractions themselves are in
// pgraph_gemm.cu.
//
// Reference path replaced: the per-(image, level) Python loop of HTDBBoxHead.forward,
// htd_bbox_head.py:195-219 - for every group ~15 ATen launches (boolean-mask gathers,
// bbox_overlaps broadcast temporaries [n,n,2], fill_diagonal_, diag, softmax ...) and a host sync
// (`.any()`).  Here ALL groups of the batch are processed by each launch; the only host read is
// the small group table of the plan.
#include "common.cuh"

namespace htd {

template <typename T>
__device__ __forceinline__ float ldf(const T* p);
template <>
__device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float ldf<__nv_bfloat16>(const __nv_bfloat16* p) {
    return __bfloat162float(*p);
}
template <typename T>
__device__ __forceinline__ void stf(T* p, float v);
template <>
__device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <>
__device__ __forceinline__ void stf<__nv_bfloat16>(__nv_bfloat16* p, float v) {
    *p = __float2bfloat16_rn(v);
}

// ------------------------------------------------------------------------------------------
// plan: one CTA, stable counting sort by key = level * B + image
// ------------------------------------------------------------------------------------------
constexpr int kPlanThreads = 1024;

__global__ void __launch_bounds__(kPlanThreads) plan_kernel(
    const float* __restrict__ rois, const int* __restrict__ levels, int K, int B, int L, int align,
    int Ncap, int* __restrict__ perm, int* __restrict__ pos, int2* __restrict__ rowspan,
    float4* __restrict__ boxes, int* __restrict__ table) {
    __shared__ int s_cnt[HTD_MAX_GROUPS];
    __shared__ int s_off[HTD_MAX_GROUPS];
    __shared__ int s_run[HTD_MAX_GROUPS];
    __shared__ int s_wcnt[kPlanThreads / 32][HTD_MAX_GROUPS];
    __shared__ int s_npad;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int G = L * B;
    for (int g = tid; g < G; g += kPlanThreads) { s_cnt[g] = 0; s_run[g] = 0; }
    for (int p = tid; p < Ncap; p += kPlanThreads) {
        perm[p] = -1;
        rowspan[p] = make_int2(0, 0);
        boxes[p] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();
    auto key_of = [&](int k) -> int {
        const int lv = levels[k];
        const float bf = rois[(size_t)k * 5];
        const int b = (int)bf;
        if (lv < 0 || lv >= L || !(bf >= 0.f) || b >= B) return -1;
        return lv * B + b;
    };
    for (int k = tid; k < K; k += kPlanThreads) {
        const int g = key_of(k);
        if (g >= 0) atomicAdd(&s_cnt[g], 1);
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int l = 0; l < L; ++l) {
            o = (o + align - 1) / align * align;
            const int lo = o;
            for (int b = 0; b < B; ++b) {
                const int g = l * B + b;
                o = (o + HTD_GROUP_ALIGN - 1) / HTD_GROUP_ALIGN * HTD_GROUP_ALIGN;
                s_off[g] = o;
                table[2 * g] = o;
                table[2 * g + 1] = s_cnt[g];
                o += s_cnt[g];
            }
            table[2 * G + 2 * l] = lo;
            table[2 * G + 2 * l + 1] = o - lo;
        }
        s_npad = (o + align - 1) / align * align;
        table[2 * G + 2 * L] = s_npad;
    }
    __syncthreads();
    // stable placement, kPlanThreads RoIs per round in index order
    for (int k0 = 0; k0 < K; k0 += kPlanThreads) {
        const int k = k0 + tid;
        const int g = (k < K) ? key_of(k) : -1;
        for (int i = lane; i < HTD_MAX_GROUPS; i += 32) s_wcnt[warp][i] = 0;
        __syncwarp();
        const unsigned peers = __match_any_sync(0xffffffffu, g);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        if (g >= 0 && rank == 0) s_wcnt[warp][g] = __popc(peers);
        __syncthreads();
        if (tid < G) {                    // exclusive scan over warps for key tid
            int acc = s_run[tid];
            for (int w = 0; w < kPlanThreads / 32; ++w) {
                const int c = s_wcnt[w][tid];
                s_wcnt[w][tid] = acc;
                acc += c;
            }
            s_run[tid] = acc;
        }
        __syncthreads();
        if (k < K) {
            int p = -1;
            if (g >= 0) {
                p = s_off[g] + s_wcnt[warp][g] + rank;
                perm[p] = k;
                rowspan[p] = make_int2(s_off[g], s_cnt[g]);
                const float* r = rois + (size_t)k * 5;
                boxes[p] = make_float4(r[1], r[2], r[3], r[4]);
            }
            pos[k] = p;
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------
// pack: gather rows by perm, convert, optional relu gate, row-major and/or transposed output
// ------------------------------------------------------------------------------------------
template <typename TS, typename TG, typename TD>
__global__ void __launch_bounds__(256) pack_kernel(const TS* __restrict__ src, long long lds,
                                                   const TG* __restrict__ gate, long long ldg,
                                                   const int* __restrict__ perm, int Npad, int D,
                                                   TD* __restrict__ dst, long long ldd,
                                                   TD* __restrict__ dstT, long long ldt, int ctiles) {
    __shared__ float tile[32][33];
    const int ct = blockIdx.x % ctiles, rt = blockIdx.x / ctiles;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int p0 = rt * 32, c0 = ct * 32;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int p = p0 + ty + i * 8, c = c0 + tx;
        float v = 0.f;
        if (p < Npad && c < D) {
            const int k = perm[p];
            if (k >= 0) {
                v = ldf<TS>(src + (size_t)k * lds + c);
                if (gate != nullptr && !(ldf<TG>(gate + (size_t)k * ldg + c) > 0.f)) v = 0.f;
            }
        }
        tile[ty + i * 8][tx] = v;
        if (dst != nullptr && p < Npad && c < ldd) stf<TD>(dst + (size_t)p * ldd + c, v);
    }
    if (dstT == nullptr) return;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int c = c0 + ty + i * 8, p = p0 + tx;
        if (c < D && p < Npad) stf<TD>(dstT + (size_t)c * ldt + p, tile[tx][ty + i * 8]);
    }
}

// ------------------------------------------------------------------------------------------
// IoU adjacency
// ------------------------------------------------------------------------------------------
// bbox_overlaps(mode='iou', is_aligned=False), iou2d_calculator.py:129-150, fp32, same operation
// order; explicit _rn intrinsics keep ptxas from contracting or reassociating anything.
__device__ __forceinline__ bool iou_positive(const float4 a, const float4 b) {
    const float area1 = __fmul_rn(__fsub_rn(a.z, a.x), __fsub_rn(a.w, a.y));
    const float area2 = __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y));
    const float ltx = fmaxf(a.x, b.x), lty = fmaxf(a.y, b.y);
    const float rbx = fminf(a.z, b.z), rby = fminf(a.w, b.w);
    const float w = fmaxf(__fsub_rn(rbx, ltx), 0.f), h = fmaxf(__fsub_rn(rby, lty), 0.f);
    const float overlap = __fmul_rn(w, h);
    float uni = __fsub_rn(__fadd_rn(area1, area2), overlap);
    uni = fmaxf(uni, 1e-6f);
    const float iou = __fdiv_rn(overlap, uni);
    return iou > 0.f;
}

__global__ void __launch_bounds__(128) graph_bits_kernel(const float4* __restrict__ boxes,
                                                         const int2* __restrict__ rowspan,
                                                         uint32_t* __restrict__ bits, int ldb,
                                                         int* __restrict__ deg) {
    __shared__ int s_part[4];
    const int p = blockIdx.x;
    const int2 sp = rowspan[p];
    const float4 me = boxes[p];
    int cnt = 0;
    for (int w = threadIdx.x; w < ldb; w += blockDim.x) {
        uint32_t m = 0u;
        const int j0 = w * 32;
        if (j0 < sp.y) {
            const int jn = min(32, sp.y - j0);
            for (int j = 0; j < jn; ++j) {
                const int q = sp.x + j0 + j;
                const bool on = (q == p) || iou_positive(me, __ldg(boxes + q));
                m |= (on ? 1u : 0u) << j;
            }
        }
        bits[(size_t)p * ldb + w] = m;
        cnt += __popc(m);
    }
    cnt = (int)warp_sum((float)cnt);   // <= 2^24 columns: exact in fp32
    if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = cnt;
    __syncthreads();
    if (threadIdx.x == 0) deg[p] = s_part[0] + s_part[1] + s_part[2] + s_part[3];
}

template <typename T>
__global__ void __launch_bounds__(256) graph_adj_kernel(const uint32_t* __restrict__ bits, int ldb,
                                                        const int* __restrict__ deg,
                                                        const int2* __restrict__ rowspan,
                                                        T* __restrict__ adj, long long ldn) {
    const int p = blockIdx.x;
    const int2 sp = rowspan[p];
    const float ri = sp.y > 0 ? powf((float)deg[p], -0.5f) : 0.f;
    for (int j = threadIdx.x; j < ldn; j += blockDim.x) {
        float v = 0.f;
        if (j < sp.y && ((bits[(size_t)p * ldb + (j >> 5)] >> (j & 31)) & 1u))
            v = __fmul_rn(ri, powf((float)deg[sp.x + j], -0.5f));
        stf<T>(adj + (size_t)p * ldn + j, v);
    }
}

// ------------------------------------------------------------------------------------------
// masked softmax forward / backward (one CTA per sorted row)
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float block_reduce(float v, bool is_max, float* s_buf) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const float u = __shfl_xor_sync(0xffffffffu, v, o);
        v = is_max ? fmaxf(v, u) : v + u;
    }
    __syncthreads();
    if (lane == 0) s_buf[warp] = v;
    __syncthreads();
    float r = s_buf[0];
    for (int w = 1; w < nw; ++w) r = is_max ? fmaxf(r, s_buf[w]) : r + s_buf[w];
    return r;
}

template <typename T>
__global__ void __launch_bounds__(256) masked_softmax_kernel(const float* __restrict__ S,
                                                             long long lds,
                                                             const uint32_t* __restrict__ bits,
                                                             int ldb,
                                                             const int2* __restrict__ rowspan,
                                                             T* __restrict__ out, long long ldo) {
    __shared__ float s_buf[8];
    const int p = blockIdx.x;
    const int n = rowspan[p].y;
    const float* s = S + (size_t)p * lds;
    const uint32_t* b = bits + (size_t)p * ldb;
    auto logit = [&](int j) -> float {
        return ((b[j >> 5] >> (j & 31)) & 1u) ? 0.f : s[j];
    };
    float mx = -INFINITY;
    for (int j = threadIdx.x; j < n; j += blockDim.x) mx = fmaxf(mx, logit(j));
    mx = block_reduce(mx, true, s_buf);
    float sum = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) sum += expf(logit(j) - mx);
    sum = block_reduce(sum, false, s_buf);
    const float inv = 1.f / sum;
    for (int j = threadIdx.x; j < ldo; j += blockDim.x)
        stf<T>(out + (size_t)p * ldo + j, j < n ? expf(logit(j) - mx) * inv : 0.f);
}

template <typename T>
__global__ void __launch_bounds__(256) softmax_bwd_kernel(const T* __restrict__ A, long long lda,
                                                          const float* __restrict__ dA,
                                                          long long ldda,
                                                          const uint32_t* __restrict__ bits, int ldb,
                                                          const int2* __restrict__ rowspan,
                                                          float* __restrict__ dS, long long ldds) {
    __shared__ float s_buf[8];
    const int p = blockIdx.x;
    const int n = rowspan[p].y;
    const T* a = A + (size_t)p * lda;
    const float* da = dA + (size_t)p * ldda;
    const uint32_t* b = bits + (size_t)p * ldb;
    float t = 0.f;
    for (int j = threadIdx.x; j < n; j += blockDim.x) t = fmaf(ldf<T>(a + j), da[j], t);
    t = block_reduce(t, false, s_buf);
    for (int j = threadIdx.x; j < ldds; j += blockDim.x) {
        float v = 0.f;
        if (j < n && !((b[j >> 5] >> (j & 31)) & 1u)) v = ldf<T>(a + j) * (da[j] - t);
        dS[(size_t)p * ldds + j] = v;
    }
}

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) group_transpose_kernel(const TI* __restrict__ in,
                                                              long long ldi,
                                                              const int2* __restrict__ rowspan,
                                                              float alpha, float beta,
                                                              TO* __restrict__ out, long long ldo) {
    const int p = blockIdx.x;
    const int2 sp = rowspan[p];
    const int i = p - sp.x;
    for (int j = threadIdx.x; j < ldo; j += blockDim.x) {
        float v = 0.f;
        if (j < sp.y) {
            if (alpha != 0.f) v = alpha * ldf<TI>(in + (size_t)p * ldi + j);
            if (beta != 0.f) v = fmaf(beta, ldf<TI>(in + (size_t)(sp.x + j) * ldi + i), v);
        }
        stf<TO>(out + (size_t)p * ldo + j, v);
    }
}

// block = 32 columns x 8 row slices
template <typename T>
__global__ void __launch_bounds__(256) segment_colsum_kernel(const T* __restrict__ x, long long ldx,
                                                             const int* __restrict__ seg, int D,
                                                             float* __restrict__ out) {
    __shared__ float s_part[8][33];
    const int s = blockIdx.y;
    const int lo = seg[2 * s], n = seg[2 * s + 1];
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + tx;
    float acc = 0.f;
    if (c < D)
        for (int r = ty; r < n; r += 8) acc += ldf<T>(x + (size_t)(lo + r) * ldx + c);
    s_part[ty][tx] = acc;
    __syncthreads();
    if (ty == 0 && c < D) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += s_part[i][tx];
        out[(size_t)s * D + c] = t;
    }
}

static bool dt_ok(int d) { return d == HTD_F32 || d == HTD_BF16; }

}  // namespace htd

using namespace htd;

extern "C" {

int htd_pgraph_plan(const float* rois, const int32_t* levels, int K, int B, int L, int align,
                    int Ncap, int32_t* perm, int32_t* pos, int32_t* rowspan, float* boxes,
                    int32_t* table, htd_stream_t stream) {
    HTD_CHECK_ARG(K >= 0 && B >= 1 && L >= 1 && align >= 1 && L * B <= HTD_MAX_GROUPS,
                  "htd_pgraph_plan: bad sizes K=%d B=%d L=%d (L*B <= %d)", K, B, L, HTD_MAX_GROUPS);
    HTD_CHECK_ARG(align % HTD_GROUP_ALIGN == 0, "htd_pgraph_plan: align must be a multiple of %d",
                  HTD_GROUP_ALIGN);
    HTD_CHECK_ARG(Ncap >= HTD_PLAN_ROW_CAPACITY(K, L, B, align),
                  "htd_pgraph_plan: Ncap=%d < required capacity %d", Ncap,
                  HTD_PLAN_ROW_CAPACITY(K, L, B, align));
    HTD_CHECK_ARG(perm && pos && rowspan && boxes && table && (K == 0 || (rois && levels)),
                  "htd_pgraph_plan: null pointer");
    plan_kernel<<<1, kPlanThreads, 0, (cudaStream_t)stream>>>(
        rois, levels, K, B, L, align, Ncap, perm, pos, reinterpret_cast<int2*>(rowspan),
        reinterpret_cast<float4*>(boxes), table);
    HTD_CHECK_LAUNCH("htd_pgraph_plan");
    return HTD_OK;
}

int htd_pgraph_pack(const void* src, int src_dtype, long long lds, const void* gate,
                    int gate_dtype, long long ldg, const int32_t* perm, int Npad, int D, void* dst,
                    long long ldd, void* dstT, long long ldt, int dst_dtype, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(src_dtype) && dt_ok(dst_dtype) && (!gate || dt_ok(gate_dtype)),
                  "htd_pgraph_pack: bad dtype");
    HTD_CHECK_ARG(Npad >= 0 && D >= 1 && (!dst || ldd >= D) && (!dstT || ldt >= Npad),
                  "htd_pgraph_pack: bad sizes Npad=%d D=%d ldd=%lld ldt=%lld", Npad, D, ldd, ldt);
    if (Npad == 0) return HTD_OK;
    HTD_CHECK_ARG(src && perm && (dst || dstT), "htd_pgraph_pack: null pointer");
    HTD_CHECK_ARG(!gate || gate_dtype == src_dtype, "htd_pgraph_pack: gate must have src's dtype");
    const int cols = dst ? (int)ldd : D;
    const int ctiles = (cols + 31) / 32, rtiles = (Npad + 31) / 32;
    cudaStream_t st = (cudaStream_t)stream;
#define CALL(TS, TD)                                                                             \
    pack_kernel<TS, TS, TD><<<ctiles * rtiles, 256, 0, st>>>(                                    \
        static_cast<const TS*>(src), lds, static_cast<const TS*>(gate), ldg, perm, Npad, D,      \
        static_cast<TD*>(dst), ldd, static_cast<TD*>(dstT), ldt, ctiles)
    if (src_dtype == HTD_F32 && dst_dtype == HTD_F32) { CALL(float, float); }
    else if (src_dtype == HTD_F32) { CALL(float, __nv_bfloat16); }
    else if (dst_dtype == HTD_F32) { CALL(__nv_bfloat16, float); }
    else { CALL(__nv_bfloat16, __nv_bfloat16); }
#undef CALL
    HTD_CHECK_LAUNCH("htd_pgraph_pack");
    return HTD_OK;
}

int htd_iou_graph_build(const float* boxes, const int32_t* rowspan, int Npad, uint32_t* bits,
                        int ldb, int32_t* deg, void* adj, int adj_dtype, long long ldn,
                        htd_stream_t stream) {
    HTD_CHECK_ARG(Npad >= 0 && ldb >= 1 && (!adj || (dt_ok(adj_dtype) && ldn >= 1 && ldn <= 32LL * ldb)),
                  "htd_iou_graph_build: bad sizes Npad=%d ldb=%d ldn=%lld", Npad, ldb, ldn);
    if (Npad == 0) return HTD_OK;
    HTD_CHECK_ARG(boxes && rowspan && bits && deg, "htd_iou_graph_build: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int2* rs = reinterpret_cast<const int2*>(rowspan);
    graph_bits_kernel<<<Npad, 128, 0, st>>>(reinterpret_cast<const float4*>(boxes), rs, bits, ldb,
                                            deg);
    HTD_CHECK_LAUNCH("htd_iou_graph_build(bits)");
    if (adj) {
        if (adj_dtype == HTD_F32)
            graph_adj_kernel<float><<<Npad, 256, 0, st>>>(bits, ldb, deg, rs,
                                                          static_cast<float*>(adj), ldn);
        else
            graph_adj_kernel<__nv_bfloat16><<<Npad, 256, 0, st>>>(
                bits, ldb, deg, rs, static_cast<__nv_bfloat16*>(adj), ldn);
        HTD_CHECK_LAUNCH("htd_iou_graph_build(adj)");
    }
    return HTD_OK;
}

int htd_pgraph_masked_softmax(const float* S, long long lds, const uint32_t* bits, int ldb,
                              const int32_t* rowspan, int Npad, void* out, int out_dtype,
                              long long ldo, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(out_dtype) && Npad >= 0 && ldo >= 1 && lds >= 1 && ldb >= 1,
                  "htd_pgraph_masked_softmax: bad arguments");
    if (Npad == 0) return HTD_OK;
    HTD_CHECK_ARG(S && bits && rowspan && out, "htd_pgraph_masked_softmax: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int2* rs = reinterpret_cast<const int2*>(rowspan);
    if (out_dtype == HTD_F32)
        masked_softmax_kernel<float><<<Npad, 256, 0, st>>>(S, lds, bits, ldb, rs,
                                                           static_cast<float*>(out), ldo);
    else
        masked_softmax_kernel<__nv_bfloat16><<<Npad, 256, 0, st>>>(
            S, lds, bits, ldb, rs, static_cast<__nv_bfloat16*>(out), ldo);
    HTD_CHECK_LAUNCH("htd_pgraph_masked_softmax");
    return HTD_OK;
}

int htd_pgraph_softmax_bwd(const void* A, int a_dtype, long long lda, const float* dA,
                           long long ldda, const uint32_t* bits, int ldb, const int32_t* rowspan,
                           int Npad, float* dS, long long ldds, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(a_dtype) && Npad >= 0 && lda >= 1 && ldda >= 1 && ldds >= 1 && ldb >= 1,
                  "htd_pgraph_softmax_bwd: bad arguments");
    if (Npad == 0) return HTD_OK;
    HTD_CHECK_ARG(A && dA && bits && rowspan && dS, "htd_pgraph_softmax_bwd: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int2* rs = reinterpret_cast<const int2*>(rowspan);
    if (a_dtype == HTD_F32)
        softmax_bwd_kernel<float><<<Npad, 256, 0, st>>>(static_cast<const float*>(A), lda, dA, ldda,
                                                        bits, ldb, rs, dS, ldds);
    else
        softmax_bwd_kernel<__nv_bfloat16><<<Npad, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(A), lda, dA, ldda, bits, ldb, rs, dS, ldds);
    HTD_CHECK_LAUNCH("htd_pgraph_softmax_bwd");
    return HTD_OK;
}

int htd_pgraph_group_transpose(const void* in, int in_dtype, long long ldi, const int32_t* rowspan,
                               int Npad, float alpha, float beta, void* out, int out_dtype,
                               long long ldo, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(in_dtype) && dt_ok(out_dtype) && Npad >= 0 && ldi >= 1 && ldo >= 1,
                  "htd_pgraph_group_transpose: bad arguments");
    if (Npad == 0) return HTD_OK;
    HTD_CHECK_ARG(in && rowspan && out && in != out, "htd_pgraph_group_transpose: null/aliased pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const int2* rs = reinterpret_cast<const int2*>(rowspan);
#define CALL(TI, TO)                                                                            \
    group_transpose_kernel<TI, TO><<<Npad, 256, 0, st>>>(static_cast<const TI*>(in), ldi, rs,   \
                                                         alpha, beta, static_cast<TO*>(out), ldo)
    if (in_dtype == HTD_F32 && out_dtype == HTD_F32) { CALL(float, float); }
    else if (in_dtype == HTD_F32) { CALL(float, __nv_bfloat16); }
    else if (out_dtype == HTD_F32) { CALL(__nv_bfloat16, float); }
    else { CALL(__nv_bfloat16, __nv_bfloat16); }
#undef CALL
    HTD_CHECK_LAUNCH("htd_pgraph_group_transpose");
    return HTD_OK;
}

int htd_pgraph_segment_colsum(const void* x, int x_dtype, long long ldx, const int32_t* seg,
                              int num_seg, int D, float* out, htd_stream_t stream) {
    HTD_CHECK_ARG(dt_ok(x_dtype) && num_seg >= 0 && D >= 1 && ldx >= D,
                  "htd_pgraph_segment_colsum: bad arguments");
    if (num_seg == 0) return HTD_OK;
    HTD_CHECK_ARG(x && seg && out, "htd_pgraph_segment_colsum: null pointer");
    dim3 grid((D + 31) / 32, num_seg);
    cudaStream_t st = (cudaStream_t)stream;
    if (x_dtype == HTD_F32)
        segment_colsum_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), ldx, seg, D,
                                                           out);
    else
        segment_colsum_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(x), ldx, seg, D, out);
    HTD_CHECK_LAUNCH("htd_pgraph_segment_colsum");
    return HTD_OK;
}

}  // extern "C"
