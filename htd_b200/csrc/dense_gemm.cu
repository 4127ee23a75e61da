// Dense bf16 contractions of the HTD head on the 5th-generation tensor cores: the FC stacks
// (convfc_bbox_head.py:141-148, htd_bbox_head.py:114-121,191-192,227-228) and the 3x3 regression
// conv tower on 7x7 RoI maps (htd_bbox_head.py:75-113,186) - forward, data gradient and weight
// gradient - which the reference leaves to cuBLAS / cuDNN.
//
// One persistent kernel (one CTA per SM, 192 threads) serves six problem kinds; they differ only
// in how the TMA producer addresses the two operands and in the major-ness bits of the UMMA
// descriptors - the pipeline is the one of csrc/pgraph_gemm.cu:
//   warp 0     TMA producer: cp.async.bulk.tensor (2-D or 4-D boxes, 128-byte swizzle) into a
//              3-stage shared-memory ring (2 x 16 KB A + 32 KB B per stage), mbarrier completion;
//   warp 1     allocates the 512 TMEM columns (two 128 x 256 accumulators = one 256-row super tile
//              whose halves share every B stage), one lane issues tcgen05.mma.cta_group::1.kind::f16
//              (M = 128, N = bn <= 256, K = 16), 4 per stage and accumulator,
//              tcgen05.commit releases the stage / publishes the accumulator;
//   warps 2-5  epilogue: tcgen05.ld 32 lanes x 32 columns, bias / per-row-class bias / ReLU / ReLU
//              gate of the backward pass, store row-major or transposed, bf16 or fp32, or fp32
//              split-K partials - overlapped with the MMAs of the next tile.
//
// Kinds (D is always [M rows = TMEM lanes, N columns]):
//   NT  D = A[M,K] . B[N,K]^T                  both K-major            FC forward
//   NN  D = A[M,K] . B[K,N]                    B MN-major              FC data gradient
//   TN  D = A[K,M]^T . B[K,N]                  both MN-major           FC weight gradient
//   CONV_FPROP  D[co, pix] = sum_{tap,ci} W[co,tap,ci] . X[pix (+) tap, ci]
//               A = weights [Cout, 9*Cin] K-major; B = activations [P,7,7,Cin] through a 4-D box
//               (64 ch, 7, 7, 5 RoIs) whose start is shifted by the tap: the halo is the TMA
//               unit's out-of-bounds zero fill; N tile = 5 RoIs = 245 pixels of UMMA N = 256;
//               stored transposed -> Y [P,7,7,Cout] channels-last.
//   CONV_DGRAD  D[ci, pix] = sum_{tap,co} W[co,tap,ci] . dY[pix (-) tap, co]
//               A = the same weight matrix read MN-major (columns tap*Cin + ci, rows co);
//               B = dY through the 4-D box with the mirrored shift; stored transposed -> dX.
//   CONV_WGRAD  D[co, (tap,ci)] = sum_pix dY[pix, co] . X[pix (+) tap, ci]
//               A = dY [P*49, Cout] MN-major, B = X through the shifted 4-D box, MN-major; one
//               k-block = two RoIs (98 k rows; rows 98..111 of the 112-row stage stay zero); split-K
//               over RoIs.
#include "tc05.cuh"

namespace htd {

// kDSuper = 1: 256-row super tiles (two accumulators sharing each B stage, 3 stages of 64 KB);
// kDSuper = 0: 128-row tiles, 4 stages of 48 KB, double-buffered accumulator.  Measured (round 2,
// profiles/r02_dense_notes.md): the super tile does NOT pay - the kernel is bound by the shared-
// memory port (TMA fill + UMMA operand reads of an SS-mode MMA), and both accumulators re-read B.
#ifndef HTD_DENSE_SUPER
#define HTD_DENSE_SUPER 0
#endif
constexpr int kDSuper = HTD_DENSE_SUPER;
static_assert(kDSuper == 0, "the 256-row super tile was measured slower and its role loops were removed");
constexpr int kDBM = 128, kDBK = 64, kDStages = kDSuper ? 3 : 4;
constexpr int kDTileM = (kDSuper ? 2 : 1) * kDBM;    // rows of a work item
constexpr int kDAHalf = kDBM * kDBK * 2;             // 16 KiB: one 128-row half of the A stage
constexpr int kDATile = (kDSuper ? 2 : 1) * kDAHalf;
constexpr int kDBTile = 256 * kDBK * 2;              // 32 KiB (bn <= 256)
constexpr int kDStage = kDATile + kDBTile;
constexpr int kDThreads = 192;
constexpr int kDMaxStages = 4;
// conv wgrad stages hold 2 RoIs = 98 k rows padded to 112 (7 K=16 steps): chunks of 14 KB
constexpr int kWgRois = 2, kWgRows = 112, kWgChunk = kWgRows * 128;
constexpr int kWgStage = (2 + 4) * kWgChunk;         // A: 2 chunks, B: up to 4 chunks (bn <= 256)
// ring: 4 x 48 KB for the GEMM / fprop / dgrad kinds, 3 x 70 KB (bn = 192) or 2 x 84 KB (bn = 256)
// for conv wgrad
constexpr int kDRingBytes = 3 * 5 * kWgChunk > kDStages * kDStage ? 3 * 5 * kWgChunk : kDStages * kDStage;
static_assert(2 * kWgStage <= kDRingBytes && kDRingBytes + 1280 <= 232448, "shared-memory ring");
constexpr int kConvStage = 2 * 32 * 128 * 2;         // conv epilogue: two [32 px][128 ch] bf16 blocks
// largest launch: the wgrad ring, or the 4 x 48 KB ring plus the conv epilogue staging
constexpr int kDSmem = (kDRingBytes > kDStages * kDStage + kConvStage ? kDRingBytes : kDStages * kDStage + kConvStage) +
                       1024 /*align*/ + 256 /*barriers*/;
static_assert(kDSmem <= 232448, "shared memory per CTA");
constexpr int kChunk = 64 * 128;                     // one 64-element MN chunk x 64 k rows
// conv fprop / dgrad N tile: 4 RoIs as two halves of 2 RoIs (98 rows) whose second half starts at
// row 104 (a multiple of 8: the 128-byte swizzle phase of a TMA destination) -> UMMA N = 208
constexpr int kPP = 49, kRoisPerTile = 5;

// experiment switches of the kernels (HTD_DENSE_DEBUG bits) exist in the -DHTD_DEBUG_HOOKS build only
#ifdef HTD_DEBUG_HOOKS
#define DENSE_DBG(p) ((p).debug)
#else
#define DENSE_DBG(p) 0
#endif

struct DenseParams {
    int kind;
    int M, N;                     // output extents (conv fprop/dgrad: N = P * 49 pixels)
    int kblocks, splits, kb_per_split;
    int tiles_m, tiles_n, bn;
    int a_mn, b_mn;
    unsigned a_half_tx, b_tx;     // bytes landing per stage: per 128-row half of A, for B
    int nstages, a_bytes, stage_bytes;   // ring geometry: stages, bytes of the A part, of a stage
    int chunk_bytes, ksteps;      // MN-major chunk pitch (k rows * 128 B), K=16 steps per stage
    int rois_per_kb;              // conv wgrad: RoIs (49 k rows each) per k-block
    int Cin;                      // conv: channels of one tap in the weight matrix columns
    int kc_per_tap;               // conv fprop/dgrad: 64-channel chunks per tap
    int nt_per_tap;               // conv wgrad: N tiles per tap
    int zero_fill;                // conv wgrad: stage tails must read as zero
    // epilogue
    int transposed;
    void* D;
    int d_bf16;
    long long ldd;
    void* D2;                     // second output: v + row_bias[row_class[m], n]  (same dtype / ld)
    const void* bias;             // [N] fp32, or bf16 when bias_bf16
    int bias_bf16;
    const float* row_bias;        // [R, N]
    const int* row_class;         // [M]
    long long ld_rb;
    int relu;
    const __nv_bfloat16* gate;    // [M, N] (row-major) or [N, M] (transposed): v *= gate > 0
    long long ldg;
    float* partial;               // [splits, M, N] fp32 (splits > 1)
    int pair;                     // CTA-pair form (dense_pair_kernel): tiles_m counts 256-row tiles
    int ring_bytes;               // nstages * stage_bytes: the barriers follow the ring
    int debug;                    // experiments (HTD_DENSE_DEBUG): 1 = no MMA issue, 2 = no TMA loads
};

__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// One work item: a 256-row super tile (two 128-row accumulators that share every B stage) or, at
// the ragged end of M, a 128-row tile.
struct DenseWork {
    int sp, nt, mt;
    bool two;        // rows m0 + 128 .. exist: second accumulator in use
};

__device__ __forceinline__ bool dense_work(const DenseParams& p, int item, DenseWork& wk) {
    const int per_split = p.tiles_m * p.tiles_n;
    if (item >= per_split * p.splits) return false;
    wk.sp = item / per_split;
    const int rem = item - wk.sp * per_split;
    wk.nt = rem / p.tiles_m;
    wk.mt = rem - wk.nt * p.tiles_m;
    wk.two = kDSuper && p.M - wk.mt * kDTileM > kDBM;
    return true;
}

// One 32-column chunk of one accumulator row: f[j] = column n0 + j of row m, nc columns count.
// conv_t tiles (D^T = [pixel, channel]): columns [jl, jh) of the chunk are padding between the two
// halves of a CTA-pair tile and are skipped, the columns after them move up by jh - jl pixels
// (jl = jh = 32: none).
__device__ __forceinline__ void dense_store_chunk(const DenseParams& p, bool conv_t, int sp, int m,
                                                  int rcls, int n0, int nc, int jl, int jh,
                                                  float (&f)[32]) {
    if (conv_t) {                                      // out[pixel, m]: lanes = channels
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.D) + (size_t)n0 * p.ldd + m;
        const __nv_bfloat16* g = p.gate ? p.gate + (size_t)n0 * p.ldg + m : nullptr;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nc && !(j >= jl && j < jh)) {
                const int r = j >= jh ? j - (jh - jl) : j;             // pixel row after the gap
                float x = f[j];
                if (p.relu) x = fmaxf(x, 0.f);
                if (g && !(__bfloat162float(g[(size_t)r * p.ldg]) > 0.f)) x = 0.f;
                o[(size_t)r * p.ldd] = __float2bfloat16_rn(x);
            }
        return;
    }
    if (p.splits > 1) {                                // fp32 partial, finished later
        float* o = p.partial + ((size_t)sp * p.M + m) * p.N + n0;
        if (nc == 32 && (((size_t)m * p.N + n0) & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nc) o[j] = f[j];
        }
        return;
    }
    // row-major: bias, second output with the per-row-class bias, relu, gate
    if (p.bias != nullptr) {
        if (p.bias_bf16) {
            const __nv_bfloat16* bb = static_cast<const __nv_bfloat16*>(p.bias) + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nc) f[j] += __bfloat162float(bb[j]);
        } else {
            const float* bb = static_cast<const float*>(p.bias) + n0;
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nc) f[j] += __ldg(bb + j);
        }
    }
    if (p.D2 != nullptr) {
        const float* rb = p.row_bias + (size_t)rcls * p.ld_rb + n0;
        __nv_bfloat16* o2 = static_cast<__nv_bfloat16*>(p.D2) + (size_t)m * p.ldd + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nc) {
                float x = f[j] + __ldg(rb + j);
                if (p.relu) x = fmaxf(x, 0.f);
                o2[j] = __float2bfloat16_rn(x);
            }
    }
    if (p.relu) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = fmaxf(f[j], 0.f);
    }
    if (p.gate != nullptr) {
        const __nv_bfloat16* g = p.gate + (size_t)m * p.ldg + n0;
#pragma unroll
        for (int j = 0; j < 32; ++j)
            if (j < nc && !(__bfloat162float(g[j]) > 0.f)) f[j] = 0.f;
    }
    const size_t base = (size_t)m * p.ldd + n0;
    if (p.d_bf16) {
        __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.D) + base;
        if (nc == 32 && (base & 7) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint32_t u[4];
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    __nv_bfloat162 h = __floats2bfloat162_rn(f[j + 2 * t], f[j + 2 * t + 1]);
                    u[t] = *reinterpret_cast<uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(o + j) = make_uint4(u[0], u[1], u[2], u[3]);
            }
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nc) o[j] = __float2bfloat16_rn(f[j]);
        }
    } else {
        float* o = static_cast<float*>(p.D) + base;
        if (nc == 32 && (base & 3) == 0) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        } else {
#pragma unroll
            for (int j = 0; j < 32; ++j)
                if (j < nc) o[j] = f[j];
        }
    }
}

// Why 256 x 256 per CTA: the kernel is bound by the L2 -> SM operand delivery (about 35-40 B per
// clock and SM with all SMs loading; ncu: tensor pipe 36 % busy with 128 x 256 tiles and 48 KB
// per k-block, profiles/r02_dense_notes.md).  Two accumulators of 128 x 256 (all 512 TMEM columns)
// that share one B stage need 64 KB per k-block for twice the flops - a third less traffic per
// flop, the same ratio a cta_group::2 pair has.  Ragged 128-row tiles keep the double-buffered
// accumulator (their epilogue overlaps the next tile's MMAs); a 256-row tile's epilogue does not.
__global__ void __launch_bounds__(kDThreads, 1)
    dense_gemm_kernel(const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_b, const DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    // stage s = [A part | B part] at smem + s * stage_bytes (sizes are multiples of 1024 B)
    const int nst = p.nstages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.ring_bytes);
    uint64_t* empty_bar = full_bar + kDMaxStages;
    uint64_t* tfull_bar = empty_bar + kDMaxStages;    // [2] accumulator slots
    uint64_t* tempty_bar = tfull_bar + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int first = (int)blockIdx.x, stride = (int)gridDim.x;

    if (p.zero_fill) {                                // uniform
        uint4* z = reinterpret_cast<uint4*>(smem);
        for (int i = threadIdx.x; i < p.ring_bytes / 16; i += kDThreads)
            z[i] = make_uint4(0u, 0u, 0u, 0u);
        fence_proxy_async();                          // generic-proxy zeros -> visible to UMMA / TMA
    }
    if (warp == 0 && lane == 0) {
        tc::prefetch_map(&map_a);
        tc::prefetch_map(&map_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kDMaxStages; ++s) {
                mbar_init(full_bar + s, 1);
                mbar_init(empty_bar + s, 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(tfull_bar + a, 1);
                mbar_init(tempty_bar + a, 4);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tc::tmem_alloc(tmem_slot, 512);
    }
    tc::fence_before();
    __syncthreads();
    tc::fence_after();
    const uint32_t tmem_base = *tmem_slot;
    const int nb_chunks = (p.bn + 63) / 64;           // MN-major B: 64-column chunks of the tile

    if (warp == 0) {
        // ===== TMA producer =====
        // The loops of the two single-thread roles are executed by the WHOLE warp (waits by all
        // lanes, the issue by lane 0), and their bodies are kept short: one warp retires a
        // dependent instruction every 4-6 clocks, so the ~150 instructions per k-block of the first
        // version (runtime divisions for stage / phase / filter tap, descriptors rebuilt from
        // scratch) were a floor of ~700 clk per k-block on their own - more than the MMAs (490)
        // and as much as the operand delivery (profiles/r02_dense_notes.md, section 6).
        unsigned s = 0, ph = 0;                       // ring stage and its phase bit
        DenseWork wk;
        const bool conv_t = p.kind == HTD_DENSE_CONV_FPROP || p.kind == HTD_DENSE_CONV_DGRAD;
        for (int item = first; dense_work(p, item, wk); item += stride) {
            const int nt = wk.nt, m0 = wk.mt * kDTileM;
            const int kb0 = wk.sp * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
            const unsigned tx = p.b_tx + p.a_half_tx;
            // filter tap (dx, dy) and 64-channel chunk of the k-block, advanced incrementally
            int tap = 0, kc = 0, dx = -1, dy = -1;
            if (conv_t) {
                tap = kb0 / p.kc_per_tap;
                kc = kb0 - tap * p.kc_per_tap;
                dy = tap / 3 - 1;
                dx = tap - (dy + 1) * 3 - 1;
            }
            const int wtap = p.kind == HTD_DENSE_CONV_WGRAD ? nt / p.nt_per_tap : 0;
            const int wn = nt - wtap * p.nt_per_tap, wdy = wtap / 3 - 1, wdx = wtap - (wdy + 1) * 3 - 1;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar + s, ph ^ 1u);
                if (lane == 0) {
                    uint64_t* bar = full_bar + s;
                    uint8_t* sa = smem + s * p.stage_bytes;
                    uint8_t* sb = sa + p.a_bytes;
                    if (DENSE_DBG(p) & 2) {                // experiment: barrier traffic only
                        mbar_arrive(bar);
                    } else {
                        mbar_expect_tx(bar, tx);
                        switch (p.kind) {
                            case HTD_DENSE_NT:
                                tc::tma_load_2d(&map_a, bar, sa, kb * kDBK, m0);
                                tc::tma_load_2d(&map_b, bar, sb, kb * kDBK, nt * p.bn);
                                break;
                            case HTD_DENSE_NN:
                                tc::tma_load_2d(&map_a, bar, sa, kb * kDBK, m0);
                                for (int c = 0; c < nb_chunks; ++c)
                                    tc::tma_load_2d(&map_b, bar, sb + c * kChunk, nt * p.bn + c * 64, kb * kDBK);
                                break;
                            case HTD_DENSE_TN:
                                for (int c = 0; c < 2; ++c)
                                    tc::tma_load_2d(&map_a, bar, sa + c * kChunk, m0 + c * 64, kb * kDBK);
                                for (int c = 0; c < nb_chunks; ++c)
                                    tc::tma_load_2d(&map_b, bar, sb + c * kChunk, nt * p.bn + c * 64, kb * kDBK);
                                break;
                            case HTD_DENSE_CONV_FPROP:
                                tc::tma_load_2d(&map_a, bar, sa, tap * p.Cin + kc * 64, m0);
                                tc::tma_load_4d(&map_b, bar, sb, kc * 64, dx, dy, nt * kRoisPerTile);
                                break;
                            case HTD_DENSE_CONV_DGRAD:
                                for (int c = 0; c < 2; ++c)
                                    tc::tma_load_2d(&map_a, bar, sa + c * kChunk, tap * p.Cin + m0 + c * 64,
                                                    kc * 64);
                                tc::tma_load_4d(&map_b, bar, sb, kc * 64, -dx, -dy, nt * kRoisPerTile);
                                break;
                            default:   // HTD_DENSE_CONV_WGRAD: k-block = rois_per_kb RoIs
                                for (int c = 0; c < 2; ++c)
                                    tc::tma_load_2d(&map_a, bar, sa + c * p.chunk_bytes, m0 + c * 64,
                                                    kb * p.rois_per_kb * kPP);
                                for (int c = 0; c < nb_chunks; ++c)
                                    tc::tma_load_4d(&map_b, bar, sb + c * p.chunk_bytes, wn * p.bn + c * 64,
                                                    wdx, wdy, kb * p.rois_per_kb);
                                break;
                        }
                    }
                }
                __syncwarp();
                if (++s == (unsigned)nst) { s = 0; ph ^= 1u; }
                if (conv_t && ++kc == p.kc_per_tap) {
                    kc = 0;
                    ++tap;
                    if (++dx == 2) { dx = -1; ++dy; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (whole warp in the loop, lane 0 issues) =====
        const uint32_t idesc = tc::make_idesc(kDBM, p.bn, p.a_mn, p.b_mn);
        const uint32_t chunk = (uint32_t)p.chunk_bytes;
        // descriptors of stage 0; a stage / a K=16 step further is an addition to the 14-bit
        // start-address field (16-byte units; the ring ends below 2^18 bytes, so no carry out)
        const uint32_t s0a = smem_u32(smem), s0b = s0a + (uint32_t)p.a_bytes;
        const uint64_t ad0 = p.a_mn ? tc::desc_mnmajor(s0a, chunk) : tc::desc_kmajor(s0a);
        const uint64_t bd0 = p.b_mn ? tc::desc_mnmajor(s0b, chunk) : tc::desc_kmajor(s0b);
        const uint64_t a_k = p.a_mn ? 128u : 2u, b_k = p.b_mn ? 128u : 2u;   // 2048 B / 32 B per step
        const uint64_t st_d = (uint64_t)(p.stage_bytes >> 4);
        unsigned s = 0, ph = 0, uses[2] = {0u, 0u}, nsingle = 0;
        DenseWork wk;
        for (int item = first; dense_work(p, item, wk); item += stride) {
            const int kb0 = wk.sp * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
            const unsigned slot = nsingle++ & 1u;     // accumulator slots alternate
            mbar_wait(tempty_bar + slot, (uses[slot] & 1u) ^ 1u);
            tc::fence_after();
            const uint32_t acc0 = tmem_base + slot * 256;
            uint32_t accum = 0u;
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(full_bar + s, ph);
                tc::fence_after();
                if (lane == 0) {
                    const uint64_t ad = ad0 + s * st_d, bd = bd0 + s * st_d;
                    if (!(DENSE_DBG(p) & 1)) {
                        if (p.ksteps == kDBK / 16) {      // the common stage: 64 k = 4 steps, unrolled
#pragma unroll
                            for (int k = 0; k < kDBK / 16; ++k) {
                                tc::umma_bf16(acc0, ad + k * a_k, bd + k * b_k, idesc, accum);
                                accum = 1u;
                            }
                        } else {
#pragma unroll 1
                            for (int k = 0; k < p.ksteps; ++k) {
                                tc::umma_bf16(acc0, ad + k * a_k, bd + k * b_k, idesc, accum);
                                accum = 1u;
                            }
                        }
                    }
                    if (DENSE_DBG(p) & 8) mbar_arrive(empty_bar + s);   // experiment (with 1): no commit
                    else tc::commit(empty_bar + s);
                }
                __syncwarp();
                if (++s == (unsigned)nst) { s = 0; ph ^= 1u; }
            }
            if (lane == 0) tc::commit(tfull_bar + slot);
            ++uses[slot];
            __syncwarp();
        }
    } else {
        // ===== epilogue (warps 2..5): TMEM lane quarter = warp % 4 =====
        const int q = warp & 3;
        unsigned uses[2] = {0u, 0u}, nsingle = 0;
        DenseWork wk;
        for (int item = first; dense_work(p, item, wk); item += stride) {
            const int sp = wk.sp, nt = wk.nt;
            const int kb0 = sp * p.kb_per_split;
            const bool has_k = kb0 < p.kblocks;
            const bool conv_t = p.kind == HTD_DENSE_CONV_FPROP || p.kind == HTD_DENSE_CONV_DGRAD;
            // columns of this tile and where they go
            int nvalid, col0;
            if (conv_t) {
                col0 = nt * kRoisPerTile * kPP;                  // first pixel of the tile
                nvalid = min(kRoisPerTile * kPP, p.N - col0);
            } else if (p.kind == HTD_DENSE_CONV_WGRAD) {
                const int tap = nt / p.nt_per_tap, nn = nt - tap * p.nt_per_tap;
                col0 = tap * p.Cin + nn * p.bn;
                nvalid = min(p.bn, p.Cin - nn * p.bn);
            } else {
                col0 = nt * p.bn;
                nvalid = min(p.bn, p.N - col0);
            }
            for (int hh = 0; hh < 1; ++hh) {
                const unsigned slot = nsingle++ & 1u;
                if (DENSE_DBG(p) & 32) {                          // experiment: one polling lane per warp
                    if (lane == 0) mbar_wait(tfull_bar + slot, uses[slot] & 1u);
                    __syncwarp();
                } else {
                    mbar_wait(tfull_bar + slot, uses[slot] & 1u);
                }
                ++uses[slot];
                tc::fence_after();
                const int m = wk.mt * kDTileM + hh * kDBM + q * 32 + lane;
                const bool row_ok = m < p.M && has_k;
                const int rcls = (row_ok && p.row_class != nullptr) ? p.row_class[m] : 0;
#pragma unroll 1
                for (int ch = 0; ch * 32 < nvalid; ++ch) {
                    uint32_t v[32];
                    __syncwarp();
                    tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + slot * 256 + (uint32_t)(ch * 32), v);
                    if (conv_t) {
                        // D^T = [pixel, channel]: a lane holds ONE channel of 32 pixels.  The four
                        // warps put their [32 px][32 ch] blocks side by side in shared memory and
                        // the 128 threads then write whole pixel rows (256 B of channels) with
                        // 16-byte stores; the ReLU gate of the backward is read the same way.
                        // (2-byte stores, 64 B per warp instruction, cost ~8 us per tile.)
                        __nv_bfloat16* stg = reinterpret_cast<__nv_bfloat16*>(smem + p.ring_bytes + 256) +
                                             (ch & 1) * (32 * 128);
                        if (!(DENSE_DBG(p) & 4)) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) {
                                float x = __uint_as_float(v[j]);
                                if (p.relu) x = fmaxf(x, 0.f);
                                stg[j * 128 + q * 32 + lane] = __float2bfloat16_rn(x);
                            }
                        }
                        asm volatile("bar.sync 1, 128;" ::: "memory");
                        if (has_k && !(DENSE_DBG(p) & 4)) {
                            const int nc = min(32, nvalid - ch * 32);
                            const int m0t = wk.mt * kDTileM;
                            const int valid_ch = min(128, p.M - m0t);      // multiple of 8
                            const int t128 = q * 32 + lane;
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const int vv = t128 + i * 128, px = vv >> 4, c8 = (vv & 15) * 8;
                                if (px < nc && c8 < valid_ch) {
                                    uint4 val = *reinterpret_cast<const uint4*>(stg + px * 128 + c8);
                                    const size_t row = (size_t)(col0 + ch * 32 + px);
                                    if (p.gate != nullptr) {
                                        const uint4 gv = *reinterpret_cast<const uint4*>(
                                            p.gate + row * p.ldg + m0t + c8);
                                        const uint32_t gw[4] = {gv.x, gv.y, gv.z, gv.w};
                                        uint32_t w[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            // gate > 0 per bf16 half: sign bit clear and not zero
                                            const uint32_t lo = gw[k] & 0xffffu, hi = gw[k] >> 16;
                                            const bool klo = lo != 0u && lo < 0x8000u;
                                            const bool khi = hi != 0u && hi < 0x8000u;
                                            w[k] = (klo ? (w[k] & 0xffffu) : 0u) | (khi ? (w[k] & 0xffff0000u) : 0u);
                                        }
                                        val = make_uint4(w[0], w[1], w[2], w[3]);
                                    }
                                    *reinterpret_cast<uint4*>(static_cast<__nv_bfloat16*>(p.D) + row * p.ldd + m0t + c8) = val;
                                }
                            }
                        }
                        continue;
                    }
                    if (!row_ok || (DENSE_DBG(p) & 4)) continue;
                    const int nc = min(32, nvalid - ch * 32);          // valid columns of this chunk
                    float f[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                    dense_store_chunk(p, conv_t, sp, m, rcls, col0 + ch * 32, nc, 32, 32, f);
                }
                tc::fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(tempty_bar + slot);
            }
        }
    }
    tc::fence_before();
    __syncthreads();
    if (warp == 1) {
        tc::fence_after();
        tc::tmem_dealloc(tmem_base, 512);
    }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair form (cta_group::2): a cluster of two CTAs works on one 256-row tile.  CTA r stages rows
// m0 + 128 r .. of A and columns r * bn/2 .. of B, the even CTA issues M = 256 UMMAs that read both
// CTAs' shared memory and write both CTAs' TMEM (its own 128 rows in each).  Per k-block and SM
// that is 16 KB of A + half a B tile for a 128 x bn accumulator - two thirds of the bytes of the
// single-CTA form, which is bound by L2 -> SM delivery (profiles/r02_dense_notes.md) - and stages
// of 30-32 KB make the ring six deep.  Conv tiles are 4 RoIs: each CTA's half holds 2 RoIs = 98
// pixel rows in a 104-row half (UMMA N = 208; columns 98..103 of each half are padding).
//   full[s]   lives in the leader: 2 arrivals (one expect_tx per CTA) + the bytes of both CTAs
//   empty[s]  in each CTA: the leader's commit arrives on both (multicast)
//   tfull[a]  in each CTA: same;  tempty[a] in the leader: 8 arrivals (4 epilogue warps x 2 CTAs)
// ------------------------------------------------------------------------------------------------
constexpr int kPairStages = 6;
constexpr int kPairRois = 2, kPairHalfRows = 104;        // conv: RoIs / rows of one CTA's B half
constexpr int kPairJunk0 = kPairRois * kPP;              // 98: first padding column of a half

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kDThreads, 1)
    dense_pair_kernel(const __grid_constant__ CUtensorMap map_a,
                      const __grid_constant__ CUtensorMap map_b, const DenseParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                               ~static_cast<uintptr_t>(1023));
    const int nst = p.nstages;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + p.ring_bytes);
    uint64_t* empty_bar = full_bar + 8;
    uint64_t* tfull_bar = empty_bar + 8;              // [2] accumulator slots
    uint64_t* tempty_bar = tfull_bar + 2;             // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = tc::cluster_ctarank();
    const int first = (int)(blockIdx.x >> 1), stride = (int)(gridDim.x >> 1);
    const bool conv_t = p.kind == HTD_DENSE_CONV_FPROP || p.kind == HTD_DENSE_CONV_DGRAD;
    const int half_n = p.bn / 2;                      // B columns (K-major: rows) per CTA

    if (warp == 0 && lane == 0) {
        tc::prefetch_map(&map_a);
        tc::prefetch_map(&map_b);
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < 8; ++s) {
                mbar_init(full_bar + s, 2);
                mbar_init(empty_bar + s, 1);
            }
            for (int a = 0; a < 2; ++a) {
                mbar_init(tfull_bar + a, 1);
                mbar_init(tempty_bar + a, 8);
            }
            fence_mbar_init();
        }
        __syncwarp();
        tc::tmem_alloc_pair(tmem_slot, 512);
    }
    tc::fence_before();
    __syncthreads();
    tc::cluster_sync_all();                           // the peer's barriers exist before any remote arrive
    tc::fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs: own half of A and of B, signalled to the leader) =====
        unsigned s = 0, ph = 0;
        DenseWork wk;
        const unsigned tx = p.a_half_tx + p.b_tx;            // bytes this CTA lands per stage
        const int nhalf_chunks = half_n / 64;
        uint32_t bar0 = 0;
        if (lane == 0) bar0 = tc::map_to_cta(full_bar, 0);   // the leader's full[0]; full[s] = +8 s
        for (int item = first; dense_work(p, item, wk); item += stride) {
            const int nt = wk.nt, mh = wk.mt * 256 + (int)rank * kDBM;
            const int kb0 = wk.sp * p.kb_per_split;
            const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
            const int nb = nt * p.bn + (int)rank * half_n;   // GEMM kinds: first B column of this CTA
            const int roi0 = (nt * 2 + (int)rank) * kPairRois;
            int tap = 0, kc = 0, dx = -1, dy = -1;
            if (conv_t) {
                tap = kb0 / p.kc_per_tap;
                kc = kb0 - tap * p.kc_per_tap;
                dy = tap / 3 - 1;
                dx = tap - (dy + 1) * 3 - 1;
            }
            for (int kb = kb0; kb < kb1; ++kb) {
                mbar_wait(empty_bar + s, ph ^ 1u);
                if (lane == 0) {
                    const uint32_t bar = bar0 + 8u * s;
                    uint8_t* sa = smem + s * p.stage_bytes;
                    uint8_t* sb = sa + p.a_bytes;
                    if (DENSE_DBG(p) & 2) {                       // experiment: barrier traffic only
                        tc::mbar_arrive_cluster(bar);
                    } else {
                        tc::mbar_expect_tx_cluster(bar, tx);
                        switch (p.kind) {
                            case HTD_DENSE_NT:
                                tc::tma_load_2d_pair(&map_a, bar, sa, kb * kDBK, mh);
                                tc::tma_load_2d_pair(&map_b, bar, sb, kb * kDBK, nb);
                                break;
                            case HTD_DENSE_NN:
                                tc::tma_load_2d_pair(&map_a, bar, sa, kb * kDBK, mh);
                                for (int c = 0; c < nhalf_chunks; ++c)
                                    tc::tma_load_2d_pair(&map_b, bar, sb + c * kChunk, nb + c * 64, kb * kDBK);
                                break;
                            case HTD_DENSE_TN:
                                for (int c = 0; c < 2; ++c)
                                    tc::tma_load_2d_pair(&map_a, bar, sa + c * kChunk, mh + c * 64, kb * kDBK);
                                for (int c = 0; c < nhalf_chunks; ++c)
                                    tc::tma_load_2d_pair(&map_b, bar, sb + c * kChunk, nb + c * 64, kb * kDBK);
                                break;
                            case HTD_DENSE_CONV_FPROP:
                                tc::tma_load_2d_pair(&map_a, bar, sa, tap * p.Cin + kc * 64, mh);
                                tc::tma_load_4d_pair(&map_b, bar, sb, kc * 64, dx, dy, roi0);
                                break;
                            default:   // HTD_DENSE_CONV_DGRAD
                                for (int c = 0; c < 2; ++c)
                                    tc::tma_load_2d_pair(&map_a, bar, sa + c * kChunk, tap * p.Cin + mh + c * 64,
                                                         kc * 64);
                                tc::tma_load_4d_pair(&map_b, bar, sb, kc * 64, -dx, -dy, roi0);
                                break;
                        }
                    }
                }
                __syncwarp();
                if (++s == (unsigned)nst) { s = 0; ph ^= 1u; }
                if (conv_t && ++kc == p.kc_per_tap) {
                    kc = 0;
                    ++tap;
                    if (++dx == 2) { dx = -1; ++dy; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only; whole warp in the loop, lane 0 issues) =====
        if (rank == 0) {
            const uint32_t idesc = tc::make_idesc(256, p.bn, p.a_mn, p.b_mn);
            const uint32_t s0a = smem_u32(smem), s0b = s0a + (uint32_t)p.a_bytes;
            const uint64_t ad0 = p.a_mn ? tc::desc_mnmajor(s0a, kChunk) : tc::desc_kmajor(s0a);
            const uint64_t bd0 = p.b_mn ? tc::desc_mnmajor(s0b, kChunk) : tc::desc_kmajor(s0b);
            const uint64_t a_k = p.a_mn ? 128u : 2u, b_k = p.b_mn ? 128u : 2u;
            const uint64_t st_d = (uint64_t)(p.stage_bytes >> 4);
            unsigned s = 0, ph = 0, uses[2] = {0u, 0u}, ntile = 0;
            DenseWork wk;
            for (int item = first; dense_work(p, item, wk); item += stride) {
                const int kb0 = wk.sp * p.kb_per_split;
                const int kb1 = min(kb0 + p.kb_per_split, p.kblocks);
                const unsigned slot = ntile++ & 1u;
                mbar_wait(tempty_bar + slot, (uses[slot] & 1u) ^ 1u);
                tc::fence_after();
                const uint32_t acc0 = tmem_base + slot * 256;
                uint32_t accum = 0u;
                for (int kb = kb0; kb < kb1; ++kb) {
                    mbar_wait(full_bar + s, ph);
                    tc::fence_after();
                    if (lane == 0) {
                        const uint64_t ad = ad0 + s * st_d, bd = bd0 + s * st_d;
                        if (!(DENSE_DBG(p) & 1)) {
#pragma unroll
                            for (int k = 0; k < kDBK / 16; ++k) {
                                tc::umma_bf16_pair(acc0, ad + k * a_k, bd + k * b_k, idesc, accum);
                                accum = 1u;
                            }
                        }
                        tc::commit_pair(empty_bar + s, 3);
                    }
                    __syncwarp();
                    if (++s == (unsigned)nst) { s = 0; ph ^= 1u; }
                }
                if (lane == 0) tc::commit_pair(tfull_bar + slot, 3);
                ++uses[slot];
                __syncwarp();
            }
        }
    } else {
        // ===== epilogue (warps 2..5 of both CTAs): rows m0 + 128 rank + .. of the pair tile =====
        const int q = warp & 3;
        unsigned uses[2] = {0u, 0u}, ntile = 0;
        DenseWork wk;
        for (int item = first; dense_work(p, item, wk); item += stride) {
            const int sp = wk.sp, nt = wk.nt;
            const bool has_k = sp * p.kb_per_split < p.kblocks;
            // columns of this tile: conv = 4 RoIs in two halves with padding, GEMM = bn columns
            const int col0 = conv_t ? nt * 2 * kPairRois * kPP : nt * p.bn;
            const int ncols = conv_t ? p.bn : min(p.bn, p.N - col0);
            const unsigned slot = ntile++ & 1u;
            mbar_wait(tfull_bar + slot, uses[slot] & 1u);
            ++uses[slot];
            tc::fence_after();
            const int m = wk.mt * 256 + (int)rank * kDBM + q * 32 + lane;
            const bool row_ok = m < p.M && has_k;
            const int rcls = (row_ok && p.row_class != nullptr) ? p.row_class[m] : 0;
#pragma unroll 1
            for (int ch = 0; ch * 32 < ncols; ++ch) {
                uint32_t v[32];
                __syncwarp();
                tc::tmem_ld_32x32(tmem_base + ((uint32_t)(q * 32) << 16) + slot * 256 + (uint32_t)(ch * 32), v);
                if (!row_ok) continue;
                float f[32];
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
                const int c0 = ch * 32;
                int nc = min(32, ncols - c0), n0 = col0 + c0, jl = 32, jh = 32;
                if (conv_t) {
                    // column c of the tile: half = c / 104, pixel = col0 + half * 98 + c % 104 (< 98)
                    const int half = c0 >= kPairHalfRows ? 1 : 0;
                    const int in_half = c0 - half * kPairHalfRows;
                    n0 = col0 + half * kPairJunk0 + in_half;
                    // the padding window(s) that intersect this chunk, relative to the chunk
                    int gl = kPairJunk0 - in_half, gh = kPairHalfRows - in_half;   // of this half
                    if (gh <= 0) { gl += kPairHalfRows; gh += kPairHalfRows; }     // next half's
                    jl = max(0, min(32, gl));
                    jh = max(0, min(32, gh));
                    const int room = p.N - n0;                   // pixels left in the output
                    const int lim = room <= jl ? room : room + (jh - jl);
                    nc = max(0, min(nc, lim));
                }
                dense_store_chunk(p, conv_t, sp, m, rcls, n0, nc, jl, jh, f);
            }
            tc::fence_before();
            __syncwarp();
            if (lane == 0) tc::mbar_arrive_cluster(tc::map_to_cta(tempty_bar + slot, 0));
        }
    }
    tc::fence_before();
    __syncthreads();
    tc::cluster_sync_all();                           // nobody leaves while the peer may still touch it
    if (warp == 1) {
        tc::fence_after();
        tc::tmem_dealloc_pair(tmem_base, 512);
    }
}

// split-K finish: sum the fp32 partials in split order (deterministic), then the same epilogue
// (bias, per-row-class second output, relu, gate), 4 columns per thread
__global__ void __launch_bounds__(256) dense_finish_kernel(const DenseParams p) {
    const long long quads = ((long long)p.N + 3) / 4;
    const long long total = (long long)p.M * quads;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int m = (int)(i / quads), n0 = (int)(i - (long long)m * quads) * 4;
        const int nc = min(4, p.N - n0);
        float f[4] = {0.f, 0.f, 0.f, 0.f};
        for (int s = 0; s < p.splits; ++s) {
            const float* src = p.partial + ((size_t)s * p.M + m) * p.N + n0;
            for (int j = 0; j < nc; ++j) f[j] += src[j];
        }
        if (p.bias != nullptr)
            for (int j = 0; j < nc; ++j)
                f[j] += p.bias_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(p.bias)[n0 + j])
                                    : __ldg(static_cast<const float*>(p.bias) + n0 + j);
        if (p.D2 != nullptr) {
            const float* rb = p.row_bias + (size_t)p.row_class[m] * p.ld_rb + n0;
            __nv_bfloat16* o2 = static_cast<__nv_bfloat16*>(p.D2) + (size_t)m * p.ldd + n0;
            for (int j = 0; j < nc; ++j) {
                float x = f[j] + __ldg(rb + j);
                if (p.relu) x = fmaxf(x, 0.f);
                o2[j] = __float2bfloat16_rn(x);
            }
        }
        if (p.relu)
            for (int j = 0; j < nc; ++j) f[j] = fmaxf(f[j], 0.f);
        if (p.gate != nullptr)
            for (int j = 0; j < nc; ++j)
                if (!(__bfloat162float(p.gate[(size_t)m * p.ldg + n0 + j]) > 0.f)) f[j] = 0.f;
        const size_t base = (size_t)m * p.ldd + n0;
        if (p.d_bf16) {
            __nv_bfloat16* o = static_cast<__nv_bfloat16*>(p.D) + base;
            for (int j = 0; j < nc; ++j) o[j] = __float2bfloat16_rn(f[j]);
        } else {
            float* o = static_cast<float*>(p.D) + base;
            for (int j = 0; j < nc; ++j) o[j] = f[j];
        }
    }
}

// dz = dy * [y > 0] (optional) and the column sums of dz (the bias gradient), deterministic:
// grid (column blocks of 64, row chunks) -> partial [chunks, N]; a second pass adds the chunks.
constexpr int kCsRows = 64;
__global__ void __launch_bounds__(256) gate_colsum_kernel(
    const __nv_bfloat16* __restrict__ dy, long long ld_dy, const __nv_bfloat16* __restrict__ y,
    long long ld_y, int rows, int N, __nv_bfloat16* __restrict__ dz, long long ld_dz,
    float* __restrict__ partial) {
    __shared__ float s_sum[4][64];
    const int c = blockIdx.x * 64 + (threadIdx.x & 63), rg = threadIdx.x >> 6;
    const int r0 = blockIdx.y * kCsRows;
    float acc = 0.f;
    if (c < N)
        for (int r = r0 + rg; r < min(r0 + kCsRows, rows); r += 4) {
            float v = __bfloat162float(dy[(size_t)r * ld_dy + c]);
            if (y != nullptr && !(__bfloat162float(y[(size_t)r * ld_y + c]) > 0.f)) v = 0.f;
            if (dz != nullptr) dz[(size_t)r * ld_dz + c] = __float2bfloat16_rn(v);
            acc += v;
        }
    s_sum[rg][threadIdx.x & 63] = acc;
    __syncthreads();
    if (rg == 0 && c < N)
        partial[(size_t)blockIdx.y * N + c] =
            (s_sum[0][threadIdx.x] + s_sum[1][threadIdx.x]) + (s_sum[2][threadIdx.x] + s_sum[3][threadIdx.x]);
}
// out[j][c] = sum over chunks of partial[chunk][j][c], j < nvec (fixed order), fp32 or bf16 output
__global__ void __launch_bounds__(256) colsum_final_kernel(const float* __restrict__ partial,
                                                           int chunks, int nvec, int N,
                                                           void* __restrict__ out, int out_bf16) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, j = blockIdx.y;
    if (c >= N) return;
    float acc = 0.f;
    for (int k = 0; k < chunks; ++k) acc += partial[((size_t)k * nvec + j) * N + c];
    if (out_bf16) static_cast<__nv_bfloat16*>(out)[(size_t)j * N + c] = __float2bfloat16_rn(acc);
    else static_cast<float*>(out)[(size_t)j * N + c] = acc;
}

// Backward glue of the dual-output FC (dense.linear_dual, htd_bbox_head.py:161-164,191-192): with
// H = [relu(v); relu(v + corr[cls])] and dH its gradient,
//   dz[m]   = dH[m] * [H[m] > 0] + dH[M + m] * [H[M + m] > 0]          (gradient of v)
//   partial[chunk][0][c]     = column sums of dz            -> bias gradient
//   partial[chunk][1 + r][c] = column sums of dH[M + m] * [H[M + m] > 0] over rows of class r
// One pass over the four [M, N] operands instead of ~14 elementwise / compare / one-hot launches.
__global__ void __launch_bounds__(256) dual_gate_kernel(
    const __nv_bfloat16* __restrict__ dH, const __nv_bfloat16* __restrict__ H,
    const int* __restrict__ cls, int M, int N, int R, __nv_bfloat16* __restrict__ dz,
    float* __restrict__ partial) {
    extern __shared__ float s_acc[];                 // [4 row groups][1 + R][64]
    const int cl = threadIdx.x & 63, rg = threadIdx.x >> 6;
    const int c = blockIdx.x * 64 + cl;
    const int r0 = blockIdx.y * kCsRows;
    float acc[1 + 8];
#pragma unroll
    for (int j = 0; j < 9; ++j) acc[j] = 0.f;
    if (c < N)
        for (int r = r0 + rg; r < min(r0 + kCsRows, M); r += 4) {
            const size_t ia = (size_t)r * N + c, ib = (size_t)(M + r) * N + c;
            const float a = __bfloat162float(H[ia]) > 0.f ? __bfloat162float(dH[ia]) : 0.f;
            const float b = __bfloat162float(H[ib]) > 0.f ? __bfloat162float(dH[ib]) : 0.f;
            const float z = a + b;
            dz[ia] = __float2bfloat16_rn(z);
            acc[0] += z;
            const int k = cls[r];
#pragma unroll
            for (int j = 0; j < 8; ++j)
                if (j == k) acc[1 + j] += b;
        }
    for (int j = 0; j <= R; ++j) s_acc[(rg * (1 + R) + j) * 64 + cl] = acc[j];
    __syncthreads();
    if (rg == 0 && c < N)
        for (int j = 0; j <= R; ++j) {
            float t = 0.f;
            for (int g = 0; g < 4; ++g) t += s_acc[(g * (1 + R) + j) * 64 + cl];
            partial[((size_t)blockIdx.y * (1 + R) + j) * N + c] = t;
        }
}

// out[p, px, c] = a[p, px, c] + alpha * b[p, px, c] + g[img[p], c]  (channels-last RoI maps): the
// regression-branch input of HTDBBoxHead (x_reg + global_feat + alpha * BA, htd_bbox_head.py:163,184)
__global__ void __launch_bounds__(256) add3_kernel(const __nv_bfloat16* __restrict__ a,
                                                   const __nv_bfloat16* __restrict__ b, float alpha,
                                                   const __nv_bfloat16* __restrict__ g,
                                                   const float* __restrict__ rois, int P, int PP, int C,
                                                   int B, __nv_bfloat16* __restrict__ out) {
    const long long n8 = (long long)P * PP * C / 8;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n8;
         i += (long long)gridDim.x * blockDim.x) {
        const long long e = i * 8;
        const int c = (int)(e % C);
        const int roi = (int)(e / ((long long)PP * C));
        const uint4 ua = *reinterpret_cast<const uint4*>(a + e);
        const uint4 ub = *reinterpret_cast<const uint4*>(b + e);
        uint4 ug = make_uint4(0u, 0u, 0u, 0u);
        if (g != nullptr) {
            const int img = (int)rois[(size_t)roi * 5];
            if (img >= 0 && img < B) ug = *reinterpret_cast<const uint4*>(g + (size_t)img * C + c);
        }
        const uint32_t wa[4] = {ua.x, ua.y, ua.z, ua.w}, wb[4] = {ub.x, ub.y, ub.z, ub.w},
                       wg[4] = {ug.x, ug.y, ug.z, ug.w};
        uint32_t wo[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const float lo = __uint_as_float(wa[t] << 16) + alpha * __uint_as_float(wb[t] << 16) +
                             __uint_as_float(wg[t] << 16);
            const float hi = __uint_as_float(wa[t] & 0xffff0000u) + alpha * __uint_as_float(wb[t] & 0xffff0000u) +
                             __uint_as_float(wg[t] & 0xffff0000u);
            __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
            wo[t] = *reinterpret_cast<uint32_t*>(&h);
        }
        *reinterpret_cast<uint4*>(out + e) = make_uint4(wo[0], wo[1], wo[2], wo[3]);
    }
}

}  // namespace htd

using namespace htd;

static int dense_launch(const CUtensorMap& ma, const CUtensorMap& mb, DenseParams& p, cudaStream_t st) {
    const long long per_split = (long long)p.tiles_m * p.tiles_n;
    if (per_split <= 0 || p.kblocks <= 0) return HTD_OK;
    const long long items = per_split * p.splits;
    HTD_CHECK_ARG(items < 2147483647LL, "htd_dense_gemm: too many tiles");
    const int sms = sm_count();
    // only the ring this problem needs (4 x 48 KB for the GEMM / fprop / dgrad kinds): the rest of
    // the SM's shared memory stays free for the small kernels of the step's other branches, which
    // otherwise wait for a whole persistent CTA to retire
    p.ring_bytes = p.nstages * p.stage_bytes;
    const size_t smem = (size_t)p.ring_bytes + 1024 /*align*/ + 256 /*barriers*/ +
                        (p.transposed && !p.pair ? kConvStage : 0);
    if (p.pair) {
        HTD_SMEM_OPTIN(dense_pair_kernel, kDSmem, "htd_dense_gemm(pair)");
        const long long clusters = items < sms / 2 ? items : sms / 2;
        dense_pair_kernel<<<(unsigned)(2 * clusters), kDThreads, smem, st>>>(ma, mb, p);
    } else {
        HTD_SMEM_OPTIN(dense_gemm_kernel, kDSmem, "htd_dense_gemm");
        const unsigned grid = (unsigned)(items < sms ? items : sms);
        dense_gemm_kernel<<<grid, kDThreads, smem, st>>>(ma, mb, p);
    }
    HTD_CHECK_LAUNCH("htd_dense_gemm");
    if (p.splits > 1) {
        const long long quads = (long long)p.M * ((p.N + 3) / 4);
        const unsigned blocks = (unsigned)((quads + 255) / 256 < 8LL * sms ? (quads + 255) / 256 : 8LL * sms);
        dense_finish_kernel<<<blocks, 256, 0, st>>>(p);
        HTD_CHECK_LAUNCH("htd_dense_gemm(finish)");
    }
    return HTD_OK;
}

// HTD_DENSE_PAIR (0 = off, 1 = conv fprop / dgrad, 2 = also the GEMM kinds), HTD_DENSE_DEBUG and
// HTD_PAIR_STAGES are read by the -DHTD_DEBUG_HOOKS build only (or set through
// htd_debug_set_option); the product library runs the single-CTA form, which measured faster at
// the head's shapes (profiles/r02_dense_notes.md, section 5).
#ifdef HTD_DEBUG_HOOKS
static int g_dense_opt[3] = {-1, -1, -1};             // pair mode, debug bits, pair stages
static int dense_option(int which, const char* env) {
    if (g_dense_opt[which] < 0) {
        const char* e = getenv(env);
        g_dense_opt[which] = e ? atoi(e) : 0;
    }
    return g_dense_opt[which];
}
static int dense_pair_mode() { return dense_option(0, "HTD_DENSE_PAIR"); }
#else
static int dense_pair_mode() { return 0; }
#endif

static int pick_splits(long long tiles, int kblocks, int want, int sms) {
    if (want > 0) return want < kblocks ? want : (kblocks > 0 ? kblocks : 1);
    // A CTA's k-loop is latency bound (about 0.45 us per k-block once the 4-stage ring is primed),
    // so a grid that leaves SMs idle is cut along K until every CTA has about one ring of k-blocks
    // - as long as that saves more than the finishing pass costs (about 4 us: one more launch).
    if (tiles >= sms) return 1;
    int s = (int)(sms / tiles);
    const int by_ring = (kblocks + kDStages - 1) / kDStages;
    if (s > by_ring) s = by_ring;
    if (s < 2) return 1;
    const float saved_us = 0.45f * (float)kblocks * (1.f - 1.f / (float)s);
    return saved_us > 4.f ? s : 1;
}

// fills the launch parameters (and, with `maps`, the tensor maps) of one problem
static int dense_setup(const HtdDenseGemm* g, DenseParams& p, CUtensorMap* ma, CUtensorMap* mb,
                       bool maps) {
    HTD_CHECK_ARG(g != nullptr, "htd_dense_gemm: null descriptor");
    HTD_CHECK_ARG(g->kind >= HTD_DENSE_NT && g->kind <= HTD_DENSE_CONV_WGRAD,
                  "htd_dense_gemm: bad kind %d", g->kind);
    HTD_CHECK_ARG(g->d_dtype == HTD_F32 || g->d_dtype == HTD_BF16, "htd_dense_gemm: bad output dtype");
    if (maps) {
        HTD_CHECK_ARG(g->A && g->B && g->D, "htd_dense_gemm: null pointer");
        HTD_CHECK_ARG((((uintptr_t)g->A | (uintptr_t)g->B | (uintptr_t)g->D) & 15) == 0,
                      "htd_dense_gemm: operands must be 16-byte aligned");
    }
    const int sms = sm_count();
    memset(&p, 0, sizeof(p));
    p.kind = g->kind;
    p.nstages = kDStages; p.a_bytes = kDATile; p.stage_bytes = kDStage;
    p.chunk_bytes = kChunk; p.ksteps = kDBK / 16; p.rois_per_kb = 1;
    p.D = g->D;
    p.d_bf16 = g->d_dtype == HTD_BF16;
    p.ldd = g->ldd;
    p.bias = g->bias;
    p.bias_bf16 = g->bias_dtype == HTD_BF16;
    p.relu = g->relu;
    p.gate = static_cast<const __nv_bfloat16*>(g->gate);
    p.ldg = g->ldg;
    p.D2 = g->D2;
    p.row_bias = g->row_bias;
    p.row_class = g->row_class;
    p.ld_rb = g->ld_row_bias;
    HTD_CHECK_ARG(!g->D2 || (g->row_bias && g->row_class && g->d_dtype == HTD_BF16),
                  "htd_dense_gemm: D2 needs row_bias, row_class and a bf16 output");
    int rc;
    if (g->kind <= HTD_DENSE_TN) {
        const long long M = g->M, N = g->N, K = g->K;
        HTD_CHECK_ARG(M > 0 && N > 0 && K > 0 && g->lda > 0 && g->ldb > 0 && g->ldd >= N,
                      "htd_dense_gemm: bad extents M=%lld N=%lld K=%lld", M, N, K);
        HTD_CHECK_ARG(g->lda % 8 == 0 && g->ldb % 8 == 0,
                      "htd_dense_gemm: operand row pitches must be multiples of 8 elements");
        p.M = (int)M;
        p.N = (int)N;
        p.bn = N >= 256 ? 256 : (int)((N + 15) / 16 * 16);
        p.tiles_m = (int)((M + kDTileM - 1) / kDTileM);
        p.tiles_n = (int)((N + p.bn - 1) / p.bn);
        p.kblocks = (int)((K + kDBK - 1) / kDBK);
        p.a_mn = g->kind == HTD_DENSE_TN;
        p.b_mn = g->kind != HTD_DENSE_NT;
        // K-major B: one box of bn rows; MN-major B: whole 64-column chunks (bn may end inside one)
        p.a_half_tx = (unsigned)kDAHalf;
        p.b_tx = (unsigned)(p.b_mn ? (p.bn + 63) / 64 * kChunk : p.bn * kDBK * 2);
        if (maps) {
            // every tile tail (rows beyond M / N, columns beyond K, k rows beyond K) is zero-filled
            // by the TMA unit: nothing to pad on the caller's side
            if (!p.a_mn) rc = tc::make_map_2d(ma, g->A, M, K, g->lda, kDBM, "htd_dense_gemm(A)");
            else rc = tc::make_map_2d(ma, g->A, K, M, g->lda, kDBK, "htd_dense_gemm(A^T)");
            if (rc) return rc;
            if (!p.b_mn) rc = tc::make_map_2d(mb, g->B, N, K, g->ldb, p.bn, "htd_dense_gemm(B)");
            else rc = tc::make_map_2d(mb, g->B, K, N, g->ldb, kDBK, "htd_dense_gemm(B^T)");
            if (rc) return rc;
        }
    } else {
        const long long P = g->P, Cin = g->Cin, Cout = g->Cout;
        HTD_CHECK_ARG(P > 0 && Cin > 0 && Cout > 0 && Cin % 64 == 0 && Cout % 64 == 0,
                      "htd_dense_gemm(conv): P=%lld Cin=%lld Cout=%lld (channels must be multiples of 64)",
                      P, Cin, Cout);
        HTD_CHECK_ARG(g->pooled == 7, "htd_dense_gemm(conv): 7x7 RoI maps only (pooled=%d)", g->pooled);
        HTD_CHECK_ARG(!g->bias && !g->D2, "htd_dense_gemm(conv): no bias / second output");
        if (g->kind == HTD_DENSE_CONV_FPROP) {
            // A = W [Cout, 9*Cin], B = X [P,7,7,Cin] -> D^T = Y [P*49, Cout]
            p.M = (int)Cout; p.N = (int)(P * kPP); p.bn = 256; p.Cin = (int)Cin;
            p.kc_per_tap = (int)(Cin / 64); p.kblocks = 9 * p.kc_per_tap;
            p.tiles_m = (int)((Cout + kDTileM - 1) / kDTileM);
            p.tiles_n = (int)((P + kRoisPerTile - 1) / kRoisPerTile);
            p.a_mn = 0; p.b_mn = 0; p.transposed = 1;
            p.a_half_tx = (unsigned)kDAHalf; p.b_tx = (unsigned)(kRoisPerTile * kPP * 128);
            HTD_CHECK_ARG(g->ldd >= Cout && g->d_dtype == HTD_BF16, "htd_dense_gemm(conv fprop): bad output");
            if (maps) {
                rc = tc::make_map_2d(ma, g->A, Cout, 9 * Cin, 9 * Cin, kDBM, "htd_dense_gemm(conv W)");
                if (rc) return rc;
                rc = tc::make_map_roi(mb, g->B, P, 7, Cin, kRoisPerTile, "htd_dense_gemm(conv X)");
                if (rc) return rc;
            }
        } else if (g->kind == HTD_DENSE_CONV_DGRAD) {
            // A = W [Cout rows, 9*Cin cols] read MN-major, B = dY [P,7,7,Cout] -> D^T = dX [P*49, Cin]
            p.M = (int)Cin; p.N = (int)(P * kPP); p.bn = 256; p.Cin = (int)Cin;
            p.kc_per_tap = (int)(Cout / 64); p.kblocks = 9 * p.kc_per_tap;
            p.tiles_m = (int)((Cin + kDTileM - 1) / kDTileM);
            p.tiles_n = (int)((P + kRoisPerTile - 1) / kRoisPerTile);
            p.a_mn = 1; p.b_mn = 0; p.transposed = 1;
            p.a_half_tx = (unsigned)kDAHalf; p.b_tx = (unsigned)(kRoisPerTile * kPP * 128);
            HTD_CHECK_ARG(g->ldd >= Cin && g->d_dtype == HTD_BF16, "htd_dense_gemm(conv dgrad): bad output");
            if (maps) {
                rc = tc::make_map_2d(ma, g->A, Cout, 9 * Cin, 9 * Cin, kDBK, "htd_dense_gemm(conv W^T)");
                if (rc) return rc;
                rc = tc::make_map_roi(mb, g->B, P, 7, Cout, kRoisPerTile, "htd_dense_gemm(conv dY)");
                if (rc) return rc;
            }
        } else {
            // A = dY [P*49, Cout] MN-major, B = X [P,7,7,Cin] MN-major -> D = dW [Cout, 9*Cin]
            p.M = (int)Cout; p.N = (int)(9 * Cin); p.Cin = (int)Cin;
            p.bn = Cin % 256 == 0 ? 256 : (Cin % 192 == 0 ? 192 : (Cin % 128 == 0 ? 128 : 64));
            p.nt_per_tap = (int)(Cin / p.bn);
            p.tiles_m = (int)((Cout + kDTileM - 1) / kDTileM);
            p.tiles_n = 9 * p.nt_per_tap;
            // k-block = 2 RoIs = 98 k rows in stages of 112 rows (7 K=16 steps; rows 98..111 stay zero)
            p.rois_per_kb = kWgRois;
            p.kblocks = (int)((P + kWgRois - 1) / kWgRois);
            p.a_mn = 1; p.b_mn = 1; p.zero_fill = 1;
            p.chunk_bytes = kWgChunk; p.ksteps = kWgRows / 16;
            p.a_bytes = 2 * kWgChunk;
            p.stage_bytes = (2 + p.bn / 64) * kWgChunk;
            p.nstages = kDRingBytes / p.stage_bytes < kDMaxStages ? kDRingBytes / p.stage_bytes : kDMaxStages;
            p.a_half_tx = (unsigned)(2 * kWgRois * kPP * 128);
            p.b_tx = (unsigned)(p.bn / 64 * kWgRois * kPP * 128);
            static_assert(!kDSuper, "conv wgrad stage geometry assumes 128-row tiles");
            HTD_CHECK_ARG(g->ldd >= 9 * Cin, "htd_dense_gemm(conv wgrad): bad output pitch");
            HTD_CHECK_ARG(!g->gate && !g->relu, "htd_dense_gemm(conv wgrad): plain output only");
            if (maps) {
                rc = tc::make_map_2d(ma, g->A, P * kPP, Cout, Cout, kWgRois * kPP, "htd_dense_gemm(conv dY^T)");
                if (rc) return rc;
                rc = tc::make_map_roi(mb, g->B, P, 7, Cin, kWgRois, "htd_dense_gemm(conv X^T)");
                if (rc) return rc;
            }
        }
    }
    // ---- CTA-pair form where it applies (HTD_DENSE_PAIR=0 turns it off, =2 extends it from the
    // conv fprop / dgrad kinds to the plain GEMM kinds)
    const int pair_mode = dense_pair_mode();
    bool pair = false;
    if (pair_mode >= 1 && (g->kind == HTD_DENSE_CONV_FPROP || g->kind == HTD_DENSE_CONV_DGRAD)) {
        pair = true;
        p.bn = 2 * kPairHalfRows;                                   // UMMA N = 208
        p.tiles_n = (int)((g->P + 2 * kPairRois - 1) / (2 * kPairRois));
        p.b_tx = (unsigned)(kPairRois * kPP * 128);
        p.stage_bytes = kDAHalf + ((kPairHalfRows * 128 + 1023) & ~1023);
        if (maps) {
            const int ch = (int)(g->kind == HTD_DENSE_CONV_FPROP ? g->Cin : g->Cout);
            rc = tc::make_map_roi(mb, g->B, g->P, 7, ch, kPairRois, "htd_dense_gemm(conv pair B)");
            if (rc) return rc;
        }
    } else if (pair_mode >= 2 && g->kind <= HTD_DENSE_TN && p.M > kDBM &&
               (p.b_mn ? p.bn % 128 == 0 : p.bn % 16 == 0) && p.bn >= 32) {
        pair = true;
        const int half_n = p.bn / 2;
        p.b_tx = (unsigned)(p.b_mn ? half_n / 64 * kChunk : half_n * kDBK * 2);
        p.stage_bytes = kDAHalf + (int)((p.b_tx + 1023u) & ~1023u);
        if (maps && !p.b_mn) {                                      // K-major B: box of bn/2 rows
            rc = tc::make_map_2d(mb, g->B, g->N, g->K, g->ldb, half_n, "htd_dense_gemm(pair B)");
            if (rc) return rc;
        }
    }
    int pair_stages = 0;
#ifdef HTD_DEBUG_HOOKS
    p.debug = dense_option(1, "HTD_DENSE_DEBUG");
    pair_stages = dense_option(2, "HTD_PAIR_STAGES");
#endif
    if (pair) {
        const int want_st = pair_stages > 0 ? pair_stages : kPairStages;
        p.pair = 1;
        p.tiles_m = (p.M + 255) / 256;
        p.a_bytes = kDAHalf;
        p.nstages = kDRingBytes / p.stage_bytes < want_st ? kDRingBytes / p.stage_bytes : want_st;
    }
    const long long tiles = (long long)p.tiles_m * p.tiles_n;
    const bool can_split = !p.transposed;
    p.splits = can_split ? pick_splits(tiles, p.kblocks, g->splits, pair ? sms / 2 : sms) : 1;
    p.kb_per_split = (p.kblocks + p.splits - 1) / p.splits;
    p.splits = (p.kblocks + p.kb_per_split - 1) / p.kb_per_split;      // no empty slice
    return HTD_OK;
}

extern "C" {

#ifdef HTD_DEBUG_HOOKS
void htd_debug_set_option(const char* name, int value) {
    if (!strcmp(name, "dense_pair")) g_dense_opt[0] = value;
    else if (!strcmp(name, "dense_debug")) g_dense_opt[1] = value;
    else if (!strcmp(name, "pair_stages")) g_dense_opt[2] = value;
}
#endif

long long htd_dense_gemm_workspace_bytes(const HtdDenseGemm* g) {
    DenseParams p;
    if (dense_setup(g, p, nullptr, nullptr, false) != HTD_OK) return -1;
    return p.splits > 1 ? (long long)p.splits * p.M * p.N * 4 : 0;
}

int htd_dense_gemm(const HtdDenseGemm* g, void* workspace, long long workspace_bytes,
                   htd_stream_t stream) {
    DenseParams p;
    CUtensorMap ma, mb;
    int rc = dense_setup(g, p, &ma, &mb, true);
    if (rc) return rc;
    if (p.splits > 1) {
        const long long need = (long long)p.splits * p.M * p.N * 4;
        HTD_CHECK_ARG(workspace && workspace_bytes >= need,
                      "htd_dense_gemm: split-K needs a workspace of %lld bytes (got %lld)", need,
                      workspace_bytes);
        p.partial = static_cast<float*>(workspace);
    }
    return dense_launch(ma, mb, p, (cudaStream_t)stream);
}

int htd_gate_colsum(const void* dy, long long ld_dy, const void* y, long long ld_y, int rows, int N,
                    void* dz, long long ld_dz, float* partial, void* out, int out_dtype,
                    htd_stream_t stream) {
    HTD_CHECK_ARG(rows >= 0 && N > 0 && dy && partial && out, "htd_gate_colsum: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (rows + kCsRows - 1) / kCsRows;
    if (chunks > 0) {
        gate_colsum_kernel<<<dim3((N + 63) / 64, chunks), 256, 0, st>>>(
            static_cast<const __nv_bfloat16*>(dy), ld_dy, static_cast<const __nv_bfloat16*>(y), ld_y,
            rows, N, static_cast<__nv_bfloat16*>(dz), ld_dz, partial);
        HTD_CHECK_LAUNCH("htd_gate_colsum");
    }
    colsum_final_kernel<<<(N + 255) / 256, 256, 0, st>>>(partial, chunks, 1, N, out,
                                                          out_dtype == HTD_BF16);
    HTD_CHECK_LAUNCH("htd_gate_colsum(final)");
    return HTD_OK;
}

int htd_dual_gate(const void* dH, const void* H, const int32_t* cls, int M, int N, int R, void* dz,
                  float* partial, void* out, int out_dtype, htd_stream_t stream) {
    HTD_CHECK_ARG(M >= 0 && N > 0 && R >= 1 && R <= 8 && dH && H && cls && dz && partial && out,
                  "htd_dual_gate: bad arguments (M=%d N=%d R=%d, R <= 8)", M, N, R);
    cudaStream_t st = (cudaStream_t)stream;
    const int chunks = (M + kCsRows - 1) / kCsRows;
    if (chunks > 0) {
        dual_gate_kernel<<<dim3((N + 63) / 64, chunks), 256, (size_t)4 * (1 + R) * 64 * sizeof(float), st>>>(
            static_cast<const __nv_bfloat16*>(dH), static_cast<const __nv_bfloat16*>(H), cls, M, N, R,
            static_cast<__nv_bfloat16*>(dz), partial);
        HTD_CHECK_LAUNCH("htd_dual_gate");
    }
    colsum_final_kernel<<<dim3((N + 255) / 256, 1 + R), 256, 0, st>>>(partial, chunks, 1 + R, N, out,
                                                                      out_dtype == HTD_BF16);
    HTD_CHECK_LAUNCH("htd_dual_gate(final)");
    return HTD_OK;
}

int htd_add3(const void* a, const void* b, float alpha, const void* g, const float* rois, int P,
             int PP, int C, int B, void* out, htd_stream_t stream) {
    HTD_CHECK_ARG(P >= 0 && PP > 0 && C > 0 && C % 8 == 0 && (g == nullptr || (rois != nullptr && B > 0)),
                  "htd_add3: bad arguments (C must be a multiple of 8)");
    if (P == 0) return HTD_OK;
    HTD_CHECK_ARG(a && b && out, "htd_add3: null pointer");
    const long long n8 = (long long)P * PP * C / 8;
    const int sms = sm_count();
    const unsigned blocks = (unsigned)((n8 + 255) / 256 < 8LL * sms ? (n8 + 255) / 256 : 8LL * sms);
    add3_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(
        static_cast<const __nv_bfloat16*>(a), static_cast<const __nv_bfloat16*>(b), alpha,
        static_cast<const __nv_bfloat16*>(g), rois, P, PP, C, B, static_cast<__nv_bfloat16*>(out));
    HTD_CHECK_LAUNCH("htd_add3");
    return HTD_OK;
}

}  // extern "C"
