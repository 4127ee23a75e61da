// Grouped bf16 GEMM on the 5th-generation tensor cores (tcgen05 + TMEM accumulators, operands
// staged by TMA) for the PGraph aggregation of the HTD head.
//
// Reference path replaced (htd_bbox_head.py:204-217, one (image, level) group at a time, each a
// separate cuBLAS call plus ~15 elementwise launches):
//     roi_feat_mixed = mm(A_local, roi_feat)            [n,n] x [n,1024]
//     sim            = mm(sam_, sam_.t())               [n,1025] x [1025,n]
//     ...              matmul(A_global, roi_feat_mixed)  [n,n] x [n,1024]
// and their backward contractions.  All of them are instances of
//     D[M,N] = A[M,K] * B[N,K]^T       (both operands K-major, bf16, fp32 accumulate)
// executed here for ALL groups of the batch in one launch.
//
// Kernel structure (persistent CTAs, one per SM, looping over 128x256 output tiles, 192 threads):
//   warp 0     TMA producer: cp.async.bulk.tensor 2D loads of a 128x64 A tile and a 256x64 B tile
//              (128B swizzle) into a 4-stage shared-memory ring that runs across tile boundaries,
//              completion on mbarriers;
//   warp 1     allocates all 512 TMEM columns (two 256-column accumulators), then one elected lane
//              issues tcgen05.mma (cta_group::1, kind::f16, M=128, N=256, K=16) four per k-block,
//              releases the smem slot with tcgen05.commit and signals "accumulator full";
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns at a time, bias / ReLU, store D row-major
//              (fp32 or bf16, optionally through a row scatter map) and/or D^T (the K-major
//              operand of the next contraction), then "accumulator empty" - so the epilogue of
//              tile i overlaps the MMAs of tile i+1.
// The 128x256 tile halves the A-operand traffic per flop of a 128x128 tile (48 KB per 4.2 MFLOP):
// with 128x128 tiles the L2 -> SM operand stream capped the kernel near 45% of the tensor peak.
// Rows/columns of a tile that fall outside the group's M x N are computed on whatever the TMA
// fetched (next group's rows or zero fill) and simply not stored; only the K extent must be
// zero padded, which the packing kernels guarantee.
#include <cuda.h>

#include "common.cuh"

namespace htd {

constexpr int kBM = 128, kBN = 256, kBK = 64, kStages = 4;
constexpr int kATileBytes = kBM * kBK * 2;           // 16 KiB
constexpr int kBTileBytes = kBN * kBK * 2;           // 32 KiB
constexpr int kStageBytes = kATileBytes + kBTileBytes;
constexpr int kGemmThreads = 192;
constexpr int kAccStages = 2;                        // TMEM accumulator double buffer
constexpr int kTmemCols = kAccStages * kBN;          // 512: the whole tensor memory of the SM
constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + 256 /*barriers*/;

struct GemmParams {
    HtdGemmGroup grp[HTD_MAX_GROUPS];
    int tile_start[HTD_MAX_GROUPS + 1];
    int G;
    void* D;
    int d_is_bf16;
    long long ldd;
    const int* d_rowmap;
    void* DT;
    int dt_is_bf16;
    long long ldt;
    const float* bias;
    int relu;
    const HtdGemmGroup* grp_dev;      // device-resident schedule (htd_pgraph_schedule); NULL: use grp[]
    const int* tile_start_dev;
};

// tile index -> (group, local tile); total_tiles() bounds the persistent loops
__device__ __forceinline__ int total_tiles(const GemmParams& p) {
    return p.grp_dev != nullptr ? p.tile_start_dev[HTD_MAX_GROUPS] : p.tile_start[HTD_MAX_GROUPS];
}
__device__ __forceinline__ bool decode_tile_at(const GemmParams& p, int t, HtdGemmGroup& grp, int& local) {
    if (p.grp_dev != nullptr) {
        if (t >= p.tile_start_dev[HTD_MAX_GROUPS]) return false;
        int g = 0;
        while (g + 1 < HTD_MAX_GROUPS && t >= p.tile_start_dev[g + 1]) ++g;
        grp = p.grp_dev[g];
        local = t - p.tile_start_dev[g];
    } else {
        if (t >= p.tile_start[HTD_MAX_GROUPS]) return false;
        int g = 0;
        while (g + 1 < p.G && t >= p.tile_start[g + 1]) ++g;
        grp = p.grp[g];
        local = t - p.tile_start[g];
    }
    return true;
}
__device__ __forceinline__ bool decode_tile(const GemmParams& p, HtdGemmGroup& grp, int& local) {
    return decode_tile_at(p, (int)blockIdx.x, grp, local);
}

// Shared epilogue: one thread owns output row m of its group and 32 consecutive columns n0.. of
// it (v[j] = fp32 accumulators).  Applies bias / relu, then writes D (optionally through the
// row scatter map) and the transposed copy DT.
__device__ __forceinline__ void epilogue_store32(const GemmParams& p, const HtdGemmGroup& grp, int m,
                                                 int n0, float (&v)[32]) {
    if (m >= grp.M) return;
    if (p.bias != nullptr) {
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (n0 + j < grp.N) v[j] += __ldg(p.bias + grp.bias_off + n0 + j);
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
    }
    if (p.D != nullptr) {
        long long row = (long long)grp.d_row + m;
        if (p.d_rowmap != nullptr) row = p.d_rowmap[row];
        if (row >= 0) {
            const long long base = row * p.ldd + grp.d_col + n0;
            if (p.d_is_bf16) {
                __nv_bfloat16* Db = static_cast<__nv_bfloat16*>(p.D) + base;
                if (n0 + 32 <= grp.N && ((base & 7) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        uint32_t w[4];
#pragma unroll
                        for (int t = 0; t < 4; ++t) {
                            __nv_bfloat162 h = __floats2bfloat162_rn(v[j + 2 * t], v[j + 2 * t + 1]);
                            w[t] = *reinterpret_cast<uint32_t*>(&h);
                        }
                        *reinterpret_cast<uint4*>(Db + j) = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j < grp.N) Db[j] = __float2bfloat16_rn(v[j]);
                }
            } else {
                float* Df = static_cast<float*>(p.D) + base;
                if (n0 + 32 <= grp.N && ((base & 3) == 0)) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(Df + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j)
                        if (n0 + j < grp.N) Df[j] = v[j];
                }
            }
        }
    }
    if (p.DT != nullptr) {
        const long long base = (long long)(grp.dt_row + n0) * p.ldt + grp.dt_col + m;
        if (p.dt_is_bf16) {
            __nv_bfloat16* T = static_cast<__nv_bfloat16*>(p.DT) + base;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < grp.N) T[(long long)j * p.ldt] = __float2bfloat16_rn(v[j]);
        } else {
            float* T = static_cast<float*>(p.DT) + base;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (n0 + j < grp.N) T[(long long)j * p.ldt] = v[j];
        }
    }
}

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0,
                                            int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0),
        "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];"
                 ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tcgen05_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]),
          "=r"(v[7]), "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]),
          "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]),
          "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]),
          "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, 128B-swizzled operand tile: rows are 128 B apart, 8-row groups 1024 B apart (SBO),
// LBO unused (1), descriptor version 1 (sm_100), layout type 2 (SWIZZLE_128B).
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// kind::f16 instruction descriptor: D=f32 (bits 4-5 = 1), A=B=bf16 (bits 7-9, 10-12 = 1),
// both K-major (bits 15,16 = 0), N>>3 at bits 17-22, M>>4 at bits 24-28.
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kBN >> 3) << 17) |
                            ((uint32_t)(kBM >> 4) << 24);

// Persistent: one CTA per SM loops over output tiles t = blockIdx.x, blockIdx.x + gridDim.x, ...
// Tile = 128 x 256 (one tcgen05.mma M128 N256 K16 per 16 K-elements); the accumulator is double
// buffered in TMEM (2 x 256 columns) so the epilogue of tile i overlaps the MMAs of tile i+1.
__global__ void __launch_bounds__(kGemmThreads, 1)
    pgraph_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a,
                       const __grid_constant__ CUtensorMap tmap_b, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + kStages * kATileBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * kStageBytes);
    uint64_t* empty_bar = full_bar + kStages;
    uint64_t* tmem_full_bar = empty_bar + kStages;        // [kAccStages]
    uint64_t* tmem_empty_bar = tmem_full_bar + kAccStages;  // [kAccStages]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tmem_empty_bar + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = total_tiles(p);
    if ((int)blockIdx.x >= ntiles) return;                // uniform: surplus CTAs leave before setup

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_a)) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_b)) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) {
                mbar_init(full_bar + s, 1);
                mbar_init(empty_bar + s, 1);
            }
            for (int a = 0; a < kAccStages; ++a) {
                mbar_init(tmem_full_bar + a, 1);
                mbar_init(tmem_empty_bar + a, 4);          // one arrival per epilogue warp
            }
            fence_mbar_init();
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;"
                     ::"r"(smem_u32(tmem_ptr_smem)), "r"((uint32_t)kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            unsigned it = 0;                              // k-blocks issued so far (ring position)
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
                HtdGemmGroup grp;
                int local;
                decode_tile_at(p, t, grp, local);
                const int tiles_n = (grp.N + kBN - 1) / kBN;
                const int mt = local / tiles_n, nt = local % tiles_n;
                const int kblocks = (grp.K + kBK - 1) / kBK;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(empty_bar + s, ((it / kStages) & 1u) ^ 1u);
                    mbar_expect_tx(full_bar + s, kStageBytes);
                    tma_load_2d(&tmap_a, full_bar + s, smem_a + s * kATileBytes, grp.a_k0 + kb * kBK,
                                grp.a_row + mt * kBM);
                    tma_load_2d(&tmap_b, full_bar + s, smem_b + s * kBTileBytes, grp.b_k0 + kb * kBK,
                                grp.b_row + nt * kBN);
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            unsigned it = 0, lt = 0;                      // k-blocks / tiles consumed so far
            for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
                HtdGemmGroup grp;
                int local;
                decode_tile_at(p, t, grp, local);
                const int kblocks = (grp.K + kBK - 1) / kBK;
                const unsigned a = lt % kAccStages;
                mbar_wait(tmem_empty_bar + a, ((lt / kAccStages) & 1u) ^ 1u);   // epilogue drained it
                tcgen05_fence_after();
                const uint32_t tmem_d = tmem_base + a * kBN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const int s = it % kStages;
                    mbar_wait(full_bar + s, (it / kStages) & 1u);
                    tcgen05_fence_after();
                    const uint64_t adesc = make_smem_desc(smem_u32(smem_a + s * kATileBytes));
                    const uint64_t bdesc = make_smem_desc(smem_u32(smem_b + s * kBTileBytes));
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k) {
                        // advance 16 bf16 = 32 B along K inside the swizzle atom: +2 in >>4 units
                        umma_bf16(tmem_d, adesc + (uint64_t)(2 * k), bdesc + (uint64_t)(2 * k), kIdesc,
                                  (kb | k) != 0 ? 1u : 0u);
                    }
                    tcgen05_commit(empty_bar + s);       // frees the smem slot when the MMAs retire
                }
                tcgen05_commit(tmem_full_bar + a);       // accumulator of this tile complete
            }
        }
    } else {
        // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        unsigned lt = 0;
        for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++lt) {
            HtdGemmGroup grp;
            int local;
            decode_tile_at(p, t, grp, local);
            const int tiles_n = (grp.N + kBN - 1) / kBN;
            const int mt = local / tiles_n, nt = local % tiles_n;
            const int kblocks = (grp.K + kBK - 1) / kBK;
            const unsigned a = lt % kAccStages;
            mbar_wait(tmem_full_bar + a, (lt / kAccStages) & 1u);
            tcgen05_fence_after();
            const int m = (kblocks > 0) ? mt * kBM + q * 32 + lane : grp.M;   // K == 0: nothing stored
            const int nvalid = grp.N - nt * kBN;                              // columns of this tile
#pragma unroll 1
            for (int ch = 0; ch < kBN / 32; ++ch) {
                if (ch * 32 >= nvalid) break;                                  // warp-uniform
                uint32_t v[32];
                __syncwarp();  // tcgen05.ld is .sync.aligned: reconverge after the predicated stores
                tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + a * kBN + (uint32_t)(ch * 32), v);
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                epilogue_store32(p, grp, m, nt * kBN + ch * 32, f);
            }
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tmem_empty_bar + a);
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                     "r"((uint32_t)kTmemCols)
                     : "memory");
    }
}

// ------------------------------------------------------------------------------------------
// exact-fp32 variant (parity configuration): 64x64x16 smem tiles, 4x4 outputs per thread, FFMA
// ------------------------------------------------------------------------------------------
constexpr int kSM = 64, kSN = 64, kSK = 16;

__global__ void __launch_bounds__(256) pgraph_gemm_f32_kernel(const float* __restrict__ A,
                                                              long long a_ld,
                                                              const float* __restrict__ B,
                                                              long long b_ld, const GemmParams p) {
    __shared__ float sA[kSK][kSM + 4];
    __shared__ float sB[kSK][kSN + 4];
    HtdGemmGroup grp;
    int local;
    if (!decode_tile(p, grp, local)) return;
    const int tiles_n = (grp.N + kSN - 1) / kSN;
    const int m0 = (local / tiles_n) * kSM, n0 = (local % tiles_n) * kSN;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const int lr = tid >> 2, lk = (tid & 3) * 4;      // 64 rows x 4 k-quads
    for (int k0 = 0; k0 < grp.K; k0 += kSK) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int k = k0 + lk + e;
            float va = 0.f, vb = 0.f;
            if (k < grp.K) {
                if (m0 + lr < grp.M) va = A[(long long)(grp.a_row + m0 + lr) * a_ld + grp.a_k0 + k];
                if (n0 + lr < grp.N) vb = B[(long long)(grp.b_row + n0 + lr) * b_ld + grp.b_k0 + k];
            }
            sA[lk + e][lr] = va;
            sB[lk + e][lr] = vb;
        }
        __syncthreads();
#pragma unroll
        for (int k = 0; k < kSK; ++k) {
            float a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = sA[k][ty * 4 + i];
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = sB[k][tx * 4 + j];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
    }
    // epilogue (scalar; this variant exists for exactness, not speed)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int m = m0 + ty * 4 + i;
        if (m >= grp.M) continue;
        long long row = (long long)grp.d_row + m;
        if (p.D != nullptr && p.d_rowmap != nullptr) row = p.d_rowmap[row];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int n = n0 + tx * 4 + j;
            if (n >= grp.N) continue;
            float v = acc[i][j];
            if (p.bias != nullptr) v += p.bias[grp.bias_off + n];
            if (p.relu) v = fmaxf(v, 0.f);
            if (p.D != nullptr && row >= 0) {
                const long long o = row * p.ldd + grp.d_col + n;
                if (p.d_is_bf16) static_cast<__nv_bfloat16*>(p.D)[o] = __float2bfloat16_rn(v);
                else static_cast<float*>(p.D)[o] = v;
            }
            if (p.DT != nullptr) {
                const long long o = (long long)(grp.dt_row + n) * p.ldt + grp.dt_col + m;
                if (p.dt_is_bf16) static_cast<__nv_bfloat16*>(p.DT)[o] = __float2bfloat16_rn(v);
                else static_cast<float*>(p.DT)[o] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
// device-side schedule of the six PGraph contraction shapes
// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(HTD_SCHED_SETS * HTD_MAX_GROUPS) schedule_kernel(
    const int* __restrict__ table, int B, int L, int d, int ds, int bm, int bn,
    HtdGemmGroup* __restrict__ groups, int* __restrict__ tile_start) {
    __shared__ int s_tiles[HTD_SCHED_SETS][HTD_MAX_GROUPS];
    const int set = threadIdx.x / HTD_MAX_GROUPS, g = threadIdx.x % HTD_MAX_GROUPS;
    const int G = L * B;
    HtdGemmGroup q;
    q.M = q.N = q.K = 0;
    q.a_row = q.a_k0 = q.b_row = q.b_k0 = q.d_row = q.d_col = q.dt_row = q.dt_col = q.bias_off = 0;
    if (set < HTD_SCHED_LEVEL_ND) {
        if (g < G) {
            const int off = table[2 * g], n = table[2 * g + 1];
            if (n > 0) {
                q.a_row = off; q.d_row = off;
                if (set == HTD_SCHED_GROUP_ND) { q.M = n; q.N = d; q.K = n; q.b_k0 = off; q.dt_col = off; }
                else if (set == HTD_SCHED_GROUP_NN_S) { q.M = n; q.N = n; q.K = ds; q.b_row = off; }
                else if (set == HTD_SCHED_GROUP_NN_D) { q.M = n; q.N = n; q.K = d; q.b_row = off; }
                else { q.M = n; q.N = ds; q.K = n; q.b_k0 = off; }
            }
        }
    } else if (g < L) {
        const int loff = table[2 * G + 2 * g], ln = table[2 * G + 2 * g + 1];
        if (ln > 0) {
            if (set == HTD_SCHED_LEVEL_ND) {
                q.M = ln; q.N = d; q.K = d; q.a_row = loff; q.b_row = g * d; q.d_row = loff;
                q.dt_col = loff; q.bias_off = g * d;
            } else {
                q.M = d; q.N = d; q.K = ln; q.a_k0 = loff; q.b_k0 = loff; q.d_row = g * d;
            }
        }
    }
    groups[set * HTD_MAX_GROUPS + g] = q;
    s_tiles[set][g] = ((q.M + bm - 1) / bm) * ((q.N + bn - 1) / bn);
    __syncthreads();
    if (g == 0) {
        int acc = 0;
        int* ts = tile_start + set * (HTD_MAX_GROUPS + 1);
        for (int i = 0; i < HTD_MAX_GROUPS; ++i) { ts[i] = acc; acc += s_tiles[set][i]; }
        ts[HTD_MAX_GROUPS] = acc;
    }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* sym = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &qres) ==
                cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(sym);
    }
    return fn;
}

static int make_map(CUtensorMap* map, const void* base, long long rows, long long ld, int box_rows,
                    const char* who) {
    EncodeTiledFn enc = get_encode_fn();
    if (!enc) {
        set_error("%s: cuTensorMapEncodeTiled is unavailable", who);
        return HTD_ERR_CUDA;
    }
    static thread_local bool ctx_bound = false;     // see tc05.cuh: driver call on a fresh thread
    if (!ctx_bound) {
        cudaFree(nullptr);
        ctx_bound = true;
    }
    cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 2};
    cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides,
                     box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("%s: cuTensorMapEncodeTiled failed (%d) base=%p rows=%lld ld=%lld", who, (int)r, base,
                  rows, ld);
        return HTD_ERR_CUDA;
    }
    return HTD_OK;
}

}  // namespace htd

using namespace htd;

static int launch_gemm(const void* A, long long a_rows, long long a_ld, const void* B,
                       long long b_rows, long long b_ld, int ab_dtype, GemmParams& p,
                       long long total, cudaStream_t st) {
    const bool tc = (ab_dtype == HTD_BF16);
    if (total <= 0) return HTD_OK;
    HTD_CHECK_ARG(total < 2147483647LL, "htd_pgraph_gemm: too many tiles");
    if (!tc) {
        pgraph_gemm_f32_kernel<<<(unsigned)total, 256, 0, st>>>(static_cast<const float*>(A), a_ld,
                                                                static_cast<const float*>(B), b_ld, p);
        HTD_CHECK_LAUNCH("htd_pgraph_gemm(f32)");
        return HTD_OK;
    }
    HTD_CHECK_ARG(a_ld % 8 == 0 && b_ld % 8 == 0,
                  "htd_pgraph_gemm: leading dimensions must be multiples of 8 (a_ld=%lld b_ld=%lld)",
                  a_ld, b_ld);
    HTD_CHECK_ARG(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0,
                  "htd_pgraph_gemm: operands must be 16-byte aligned");
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, a_rows, a_ld, kBM, "htd_pgraph_gemm(A)");
    if (rc) return rc;
    rc = make_map(&mb, B, b_rows, b_ld, kBN, "htd_pgraph_gemm(B)");
    if (rc) return rc;
    HTD_SMEM_OPTIN(pgraph_gemm_kernel, kSmemBytes, "htd_pgraph_gemm");
    const int num_sms = sm_count();
    const unsigned grid = (unsigned)(total < num_sms ? total : num_sms);   // persistent CTAs
    pgraph_gemm_kernel<<<grid, kGemmThreads, kSmemBytes, st>>>(ma, mb, p);
    HTD_CHECK_LAUNCH("htd_pgraph_gemm(bf16)");
    return HTD_OK;
}

static int fill_common(GemmParams& p, const void* A, const void* B, int ab_dtype, long long a_rows,
                       long long b_rows, long long a_ld, long long b_ld, void* D, int d_dtype,
                       long long ldd, const int32_t* d_rowmap, void* DT, int dt_dtype, long long ldt,
                       const float* bias, int relu) {
    HTD_CHECK_ARG(A && B && (D || DT), "htd_pgraph_gemm: null pointer");
    HTD_CHECK_ARG(ab_dtype == HTD_F32 || ab_dtype == HTD_BF16, "htd_pgraph_gemm: bad operand dtype");
    HTD_CHECK_ARG((!D || d_dtype == HTD_F32 || d_dtype == HTD_BF16) &&
                      (!DT || dt_dtype == HTD_F32 || dt_dtype == HTD_BF16),
                  "htd_pgraph_gemm: bad output dtype");
    HTD_CHECK_ARG(a_rows > 0 && b_rows > 0 && a_ld > 0 && b_ld > 0, "htd_pgraph_gemm: bad extents");
    p.D = D;
    p.d_is_bf16 = (d_dtype == HTD_BF16);
    p.ldd = ldd;
    p.d_rowmap = d_rowmap;
    p.DT = DT;
    p.dt_is_bf16 = (dt_dtype == HTD_BF16);
    p.ldt = ldt;
    p.bias = bias;
    p.relu = relu;
    p.grp_dev = nullptr;
    p.tile_start_dev = nullptr;
    p.G = 0;
    return HTD_OK;
}

extern "C" int htd_pgraph_gemm(const void* A, long long a_rows, long long a_ld, const void* B,
                               long long b_rows, long long b_ld, int ab_dtype,
                               const HtdGemmGroup* groups, int G, void* D, int d_dtype,
                               long long ldd, const int32_t* d_rowmap, void* DT, int dt_dtype,
                               long long ldt, const float* bias, int relu, htd_stream_t stream) {
    HTD_CHECK_ARG(G >= 0 && G <= HTD_MAX_GROUPS, "htd_pgraph_gemm: G=%d exceeds %d", G, HTD_MAX_GROUPS);
    if (G == 0) return HTD_OK;
    HTD_CHECK_ARG(groups != nullptr, "htd_pgraph_gemm: null group table");
    GemmParams p;
    int rc = fill_common(p, A, B, ab_dtype, a_rows, b_rows, a_ld, b_ld, D, d_dtype, ldd, d_rowmap,
                         DT, dt_dtype, ldt, bias, relu);
    if (rc) return rc;
    const bool tc = (ab_dtype == HTD_BF16);
    const int bm = tc ? kBM : kSM, bn = tc ? kBN : kSN;
    long long total = 0;
    for (int g = 0; g < G; ++g) {
        const HtdGemmGroup& q = groups[g];
        HTD_CHECK_ARG(q.M >= 0 && q.N >= 0 && q.K >= 0 && q.a_row >= 0 && q.b_row >= 0 &&
                          q.a_k0 >= 0 && q.b_k0 >= 0 && q.d_row >= 0 && q.d_col >= 0 &&
                          q.dt_row >= 0 && q.dt_col >= 0,
                      "htd_pgraph_gemm: malformed group %d", g);
        HTD_CHECK_ARG(q.a_row + (long long)q.M <= a_rows && q.b_row + (long long)q.N <= b_rows &&
                          q.a_k0 + (long long)q.K <= a_ld && q.b_k0 + (long long)q.K <= b_ld,
                      "htd_pgraph_gemm: group %d exceeds the operand extents", g);
        HTD_CHECK_ARG(!tc || (q.a_k0 % 8 == 0 && q.b_k0 % 8 == 0),
                      "htd_pgraph_gemm: group %d: K offsets must be multiples of 8 (TMA needs "
                      "16-byte aligned coordinates), got a_k0=%d b_k0=%d", g, q.a_k0, q.b_k0);
        p.grp[g] = q;
        p.tile_start[g] = (int)total;
        total += (long long)((q.M + bm - 1) / bm) * ((q.N + bn - 1) / bn);
    }
    for (int g = G; g <= HTD_MAX_GROUPS; ++g) p.tile_start[g] = (int)total;
    p.G = G;
    return launch_gemm(A, a_rows, a_ld, B, b_rows, b_ld, ab_dtype, p, total, (cudaStream_t)stream);
}

extern "C" int htd_pgraph_schedule(const int32_t* table, int B, int L, int d, int ds, int ab_dtype,
                                   void* sched, htd_stream_t stream) {
    HTD_CHECK_ARG(table && sched, "htd_pgraph_schedule: null pointer");
    HTD_CHECK_ARG(B >= 1 && L >= 1 && L * B <= HTD_MAX_GROUPS && d >= 1 && ds >= 1,
                  "htd_pgraph_schedule: bad sizes B=%d L=%d d=%d ds=%d", B, L, d, ds);
    HTD_CHECK_ARG(ab_dtype == HTD_F32 || ab_dtype == HTD_BF16, "htd_pgraph_schedule: bad dtype");
    const bool tc = (ab_dtype == HTD_BF16);
    HtdGemmGroup* groups = static_cast<HtdGemmGroup*>(sched);
    int* tile_start = reinterpret_cast<int*>(groups + HTD_SCHED_SETS * HTD_MAX_GROUPS);
    schedule_kernel<<<1, HTD_SCHED_SETS * HTD_MAX_GROUPS, 0, (cudaStream_t)stream>>>(
        table, B, L, d, ds, tc ? kBM : kSM, tc ? kBN : kSN, groups, tile_start);
    HTD_CHECK_LAUNCH("htd_pgraph_schedule");
    return HTD_OK;
}

extern "C" long long htd_pgraph_max_tiles(int set, int K, int B, int L, int max_group, int Ncap,
                                          int d, int ds, int ab_dtype) {
    const bool tc = (ab_dtype == HTD_BF16);
    const long long bm = tc ? kBM : kSM, bn = tc ? kBN : kSN;
    auto cd = [](long long a, long long b) { return (a + b - 1) / b; };
    const long long G = (long long)L * B;
    const long long mrows = cd(K, bm) + G;           // sum over groups of ceil(n / bm)
    switch (set) {
        case HTD_SCHED_GROUP_ND: return mrows * cd(d, bn);
        case HTD_SCHED_GROUP_NN_S:
        case HTD_SCHED_GROUP_NN_D: return mrows * cd(max_group, bn);
        case HTD_SCHED_GROUP_NS: return mrows * cd(ds, bn);
        case HTD_SCHED_LEVEL_ND: return (cd(Ncap, bm) + L) * cd(d, bn);
        case HTD_SCHED_LEVEL_DD: return (long long)L * cd(d, bm) * cd(d, bn);
        default: return -1;
    }
}

extern "C" int htd_pgraph_gemm_scheduled(const void* A, long long a_rows, long long a_ld,
                                         const void* B, long long b_rows, long long b_ld,
                                         int ab_dtype, const void* sched, int set,
                                         long long max_tiles, void* D, int d_dtype, long long ldd,
                                         const int32_t* d_rowmap, void* DT, int dt_dtype,
                                         long long ldt, const float* bias, int relu,
                                         htd_stream_t stream) {
    HTD_CHECK_ARG(sched && set >= 0 && set < HTD_SCHED_SETS && max_tiles >= 0,
                  "htd_pgraph_gemm_scheduled: bad schedule arguments");
    GemmParams p;
    int rc = fill_common(p, A, B, ab_dtype, a_rows, b_rows, a_ld, b_ld, D, d_dtype, ldd, d_rowmap,
                         DT, dt_dtype, ldt, bias, relu);
    if (rc) return rc;
    const HtdGemmGroup* groups = static_cast<const HtdGemmGroup*>(sched);
    const int* tile_start = reinterpret_cast<const int*>(groups + HTD_SCHED_SETS * HTD_MAX_GROUPS);
    p.grp_dev = groups + set * HTD_MAX_GROUPS;
    p.tile_start_dev = tile_start + set * (HTD_MAX_GROUPS + 1);
    return launch_gemm(A, a_rows, a_ld, B, b_rows, b_ld, ab_dtype, p, max_tiles, (cudaStream_t)stream);
}
