// Persistent, warp-specialised tcgen05 GEMM / implicit-GEMM convolution for sm_100a.
//
//   warp 0 (1 lane)  : TMA producer  -- A tile (128 x 64 fp16) + B tile (bn x 64 fp16) per k-block
//   warp 1 (1 lane)  : MMA issuer    -- 4 x tcgen05.mma (M=128, N<=bn, K=16) per k-block, fp32 in TMEM
//   warps 2..9       : epilogue      -- two warps per TMEM lane quarter, alternating 64-column chunks:
//                                       tcgen05.ld 32x32b -> bias / row-group bias / SiLU / GEGLU / alpha /
//                                       scaled residuals in registers (row-owner layout, residual loads issued
//                                       as one batch) -> fp16 into a SWIZZLE_128B staging tile -> TMA store
//                                       (hardware clips ragged rows/columns; no per-thread global stores)
//   two TMEM accumulator stages so the epilogue of tile i overlaps the main loop of tile i+1.
//
// The A operand is always a K-major SWIZZLE_128B tile written by TMA; what changes between a Linear,
// a 3x3 convolution and a (3,1,1) temporal convolution is only the tensor map and the coordinates of
// the box fetched for each k-block (shifted boxes; the halo is zero-filled by TMA OOB handling).
// The last N tile of a row may be narrower: its MMA N shrinks (instruction descriptor per tile).
//
// Replaces (reference = diffusers 0.24 blocks instantiated at
// /root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py:169-233 and
// /root/reference/MOFA-Video-Traj/models/controlnet_sdv.py:270-309): cuDNN Conv2d/Conv3d and cuBLAS
// Linear inside ResnetBlock2D / TemporalResnetBlock / BasicTransformerBlock / FeedForward(GEGLU).
//
// Round-1 profile notes (profiles/): the first version stored straight from registers (16 B per thread per
// row: half-used sectors) and its unrolled epilogue was > 64 KB of SASS: ncu showed stall_no_inst and
// exposed residual-load latency dominating K=320 GEMMs (15 % tensor-pipe).  Hence the compact loops here.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

constexpr int BM = 128;
constexpr int BK = 64;
constexpr int kEpiWarps = 8;
constexpr int kGemmThreads = 64 + 32 * kEpiWarps;
constexpr uint32_t kABytes = BM * BK * 2;
constexpr uint32_t kTmemCols = 512;
constexpr uint32_t kStagingBytes = 4096;  // 32 rows x 128 B per epilogue warp
constexpr bool kGeglu16 = true;           // GEGLU (LINEAR mode) on the 16-epilogue-warp instantiation (A/B switch)
constexpr bool kGegluCompact = false;     // rolled 8-column GEGLU body in the 8-warp kernel too (measured: K >= 640 slower)
constexpr int kGnSlots = 2;               // statistics (frames / batch items) one M tile may span in shared memory
constexpr int kGnGroups = 32;             // groups one N tile may span in shared memory
constexpr int kGnBufs = 3;                // tiles in flight: an epilogue warp is at most 2 tiles ahead of another (2 TMEM stages)

struct GemmKernelParams {
    int mode, act;
    int m_tiles, n_tiles, num_kb, kb_per_tap, kb_split;
    int N;      // weight rows
    int N_out;  // valid output columns (N, or N/2 for GEGLU)
    int bn;     // weight rows per N tile
    int stages;
    int tma_out;  // epilogue through staging + TMA store (needs 16-byte aligned rows, N_out >= 64)
    long long M;
    int H, W, tiles_x, tiles_y, BH, BW, dil, ks;
    int BN, n_img;  // conv: images per M tile (> 1 when one image is smaller than 128 pixels or its rows do not fill tiles)
    int sBH, sBN;   // conv: rows / images covered by one warp's 32-row TMA store box
    int T, HW, tiles_p;
    __half* out;
    long long ldc;
    const __half* bias;
    const __half* rowbias;
    long long ld_rowbias;
    long long rows_per_group;
    long long rowbias_mod;  // > 0: group = row % rowbias_mod instead of row / rows_per_group
    int vec_ok;             // leading dimensions multiples of 8 -> 16-byte accesses allowed
    const __half* res1;
    long long ldr1;
    const __half* res2;
    long long ldr2;
    float alpha, beta1, beta2;
    float* gn_stats;            // GroupNorm statistics of the output (see mofa_gemm_args.gn_stats) or nullptr
    long long gn_rows_per_stat;
    int gn_groups, gn_cpg, gn_c_off;
    int gn_debug;               // MOFA_GN_DEBUG (profiling experiments): 1 = no shared-memory adds, 2 = no row loop either
};

struct TileCoord {
    int n_img, y0, x0;  // conv
    int frame, p0;      // temporal (frame = b*T + t)
    long long m0;       // linear
};

MOFA_DEVICE TileCoord tile_coord(const GemmKernelParams& p, int mt) {
    TileCoord c;
    c.n_img = c.y0 = c.x0 = c.frame = c.p0 = 0;
    c.m0 = 0;
    if (p.mode == MOFA_A_LINEAR) {
        c.m0 = static_cast<long long>(mt) * BM;
    } else if (p.mode == MOFA_A_CONV3X3) {
        int per_img = p.tiles_x * p.tiles_y;
        const int grp = mt / per_img;
        c.n_img = grp * p.BN;
        int r = mt - grp * per_img;
        int ty = r / p.tiles_x;
        c.y0 = ty * p.BH;
        c.x0 = (r - ty * p.tiles_x) * p.BW;
    } else {
        c.frame = mt / p.tiles_p;
        c.p0 = (mt - c.frame * p.tiles_p) * BM;
    }
    return c;
}

// fire-and-forget fp32 add to GLOBAL memory.  atomicAdd() on a pointer the compiler cannot prove global compiles to ATOM (value
// returned: the warp waits a full round trip per call) behind an address-space test; the statistics flush issues 128 of them.
MOFA_DEVICE void red_add_global(float* addr, float v) {
    asm volatile("red.global.add.f32 [%0], %1;" ::"l"(__cvta_generic_to_global(addr)), "f"(v) : "memory");
}

union H8 {
    uint4 u;
    __half2 h2[4];
    __half h[8];
};

// act: 0 none, 1 SiLU, (2 GEGLU handled separately), 3 ReLU, 4 sigmoid, (5 ReLU after the residual), 6 GELU (erf),
// 7 quick-GELU x * sigmoid(1.702 x)  -- 6 / 7 are the CLIP vision MLP activations (hidden_act "gelu" / "quick_gelu")
MOFA_DEVICE float apply_act(float v, int act) {
    if (act == 1) return silu_f(v);
    if (act == 3) return fmaxf(v, 0.f);
    if (act == 4) return sigmoid_f(v);
    if (act == 6) return gelu_erf_f(v);
    if (act == 7) return v * sigmoid_f(1.702f * v);
    return v;
}

// fallback for narrow / unaligned outputs (N = 3, 4, 16, ...): scalar or 16-byte stores straight from registers
__device__ __noinline__ void epilogue_direct8(const GemmKernelParams& p, const float* vin, long long row,
                                              long long group, int n_bias, int n_out, bool act_silu) {
    const bool full = p.vec_ok && (n_out + 8 <= p.N_out);
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = vin[j];
    if (full) {
        if (p.bias) {
            H8 b;
            b.u = *reinterpret_cast<const uint4*>(p.bias + n_bias);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __half2float(b.h[j]);
        }
        if (p.rowbias) {
            H8 b;
            b.u = *reinterpret_cast<const uint4*>(p.rowbias + group * p.ld_rowbias + n_bias);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += __half2float(b.h[j]);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = apply_act(v[j], p.act) * p.alpha;
        if (p.res1) {
            H8 r;
            r.u = *reinterpret_cast<const uint4*>(p.res1 + row * p.ldr1 + n_out);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += p.beta1 * __half2float(r.h[j]);
        }
        if (p.res2) {
            H8 r;
            r.u = *reinterpret_cast<const uint4*>(p.res2 + row * p.ldr2 + n_out);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += p.beta2 * __half2float(r.h[j]);
        }
        if (p.act == 5) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = fmaxf(v[j], 0.f);
        }
        H8 o;
#pragma unroll
        for (int j = 0; j < 4; ++j) o.h2[j] = __floats2half2_rn(v[2 * j], v[2 * j + 1]);
        *reinterpret_cast<uint4*>(p.out + row * p.ldc + n_out) = o.u;
    } else {
        for (int j = 0; j < 8; ++j) {
            if (n_out + j >= p.N_out) break;
            float x = v[j];
            if (p.bias) x += __half2float(p.bias[n_bias + j]);
            if (p.rowbias) x += __half2float(p.rowbias[group * p.ld_rowbias + n_bias + j]);
            x = apply_act(x, p.act);
            x *= p.alpha;
            if (p.res1) x += p.beta1 * __half2float(p.res1[row * p.ldr1 + n_out + j]);
            if (p.res2) x += p.beta2 * __half2float(p.res2[row * p.ldr2 + n_out + j]);
            if (p.act == 5) x = fmaxf(x, 0.f);
            p.out[row * p.ldc + n_out + j] = __float2half_rn(x);
        }
    }
}

// kEpi selects the epilogue body so that the two production kernels stay small (the epilogue is unrolled 32 wide and was
// fetch-limited: stall_no_instruction 0.45 per issue, profiles/r1_ncu_geglu320_v3.csv):
//   0 plain (bias / row-group bias / alpha / residuals)   -- every conv and linear of the denoise step except GEGLU
//   1 GEGLU                                               -- FeedForward first projections
//   2 generic activations (SiLU, ReLU, sigmoid, GELU, quick-GELU, ReLU-after-residual): conditioning convs, CMP, VAE, CLIP
// kStats (plain only): GroupNorm statistics of the output accumulated from the staging tile (mofa_gemm_args.gn_stats).
// kEW = epilogue warps: 8 everywhere; 16 for the GEGLU kernel (LINEAR mode only), whose epilogue is the bottleneck at
// K = 320 / 640 and latency-limited with two warps per sub-partition (issue-active 59 %): four warps per sub-partition, each
// owning one 32-column half-chunk, 32 x 64-byte SWIZZLE_64B staging tiles, one {32 cols, 32 rows} TMA store each.
// 168 registers is the ceiling for a 320-thread CTA: the register file is split per sub-partition (16384 each) and ten warps
// land 3 + 3 + 2 + 2, so 3 warps x 32 x R <= 16384 (a 176-register build fails to launch: "too many resources requested");
// 18 warps (kEW = 16) land 5 + 5 + 4 + 4: 96 registers.
// k2: CTA pairs (cluster of 2, cta_group::2 MMAs of M = 256): CTA `rank` of pair `i` owns M tile 2 i + rank and stages the
// B rows [rank * n/2, (rank + 1) * n/2) of the pair's N tile; the leader (rank 0) issues every MMA.  Everything after
// the accumulator (the epilogues) is per CTA and unchanged.
template <int kEpi, bool kStats, int kEW, bool k2>
__global__ void __launch_bounds__(64 + 32 * kEW, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmA2,
               const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmOut,
               const GemmKernelParams p) {
    constexpr bool kGeglu = kEpi == 1;
    extern __shared__ uint8_t smem_raw[];
    // SWIZZLE_128B tiles need 1024 B alignment
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

    const uint32_t rank = k2 ? cluster_ctarank() : 0u;
    const uint32_t b_bytes = static_cast<uint32_t>(k2 ? (p.bn >> 1) : p.bn) * (BK * 2);   // B rows staged by THIS CTA
    const uint32_t stage_bytes = kABytes + b_bytes;
    uint8_t* staging = smem + static_cast<size_t>(p.stages) * stage_bytes;  // kEpiWarps x 4 KB, 1024-aligned
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kEpiWarps * kStagingBytes);  // same 32 KB for kEW = 16 (2 KB tiles)
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + p.stages;
    uint64_t* tfull_bar = bars + 2 * p.stages;
    uint64_t* tempty_bar = tfull_bar + 2;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(tempty_bar + 2);
    // kStats: per-CTA GroupNorm partial sums [kGnSlots statistics][kGnGroups groups][sum, sum of squares] of the tile being
    // drained; shared-memory atomics per chunk, ONE flush of global atomics per tile.  (Global atomics per chunk -- 4.6 M per
    // level-0 launch onto 3200 addresses -- doubled the conv time: profiles/r2_step_detail_b_gnstats_global_atomics.txt.)
    float* s_gn = reinterpret_cast<float*>(tmem_ptr_smem + 4);          // [kGnBufs][kGnSlots][kGnGroups][2]
    int* s_gn_cnt = reinterpret_cast<int*>(s_gn + kGnBufs * kGnSlots * kGnGroups * 2);   // warps done per buffer

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    if constexpr (kStats) {
        for (int i = threadIdx.x; i < kGnBufs * kGnSlots * kGnGroups * 2; i += blockDim.x) s_gn[i] = 0.f;
        if (threadIdx.x < kGnBufs) s_gn_cnt[threadIdx.x] = 0;
    }

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmA2);
        tma_prefetch_desc(&tmB);
        tma_prefetch_desc(&tmOut);
        for (int i = 0; i < p.stages; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], k2 ? 2 * kEW : kEW);   // pair: both CTAs' epilogue warps arrive at the leader
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if constexpr (k2) tmem_alloc_2sm(tmem_ptr_smem, kTmemCols);
        else tmem_alloc(tmem_ptr_smem, kTmemCols);
    }
    tc_fence_before();
    if constexpr (k2) cluster_sync_all();   // the peer's barriers must exist before any remote arrive / multicast commit
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t acc_stride = (static_cast<uint32_t>(p.bn) + 31u) & ~31u;

    // Tile order.  Default: tile = cta + i * grid with N fastest, so the CTAs that run concurrently share A tiles in L2.
    // With several M tiles per CTA each CTA instead walks ALL N tiles of one M tile back to back: the A tile
    // (K = 320: 80 KB) is re-read from L2 right after its first use instead of racing a neighbour CTA to DRAM
    // (ncu: 838 MB read for 590 MB of algorithmic input on the 320x320 projection; +19 % on that GEMM).  The M tiles
    // that do not fill a whole round of CTAs (m_tiles mod grid) are dealt out N-fastest again to keep the tail short.
    // (pairs: the same walk over PAIRS of M tiles; both CTAs of a pair see the same sequence, each takes its own half)
    const int grid_i = static_cast<int>(gridDim.x) >> (k2 ? 1 : 0), cta_i = static_cast<int>(blockIdx.x) >> (k2 ? 1 : 0);
    const int m_sched = k2 ? (p.m_tiles + 1) >> 1 : p.m_tiles;   // odd count: the last pair's second CTA runs an empty tile
    const int full_groups = (p.n_tiles > 1 && m_sched / grid_i >= 2) ? m_sched / grid_i : 0;
    const int grouped_tiles = full_groups * p.n_tiles;           // per CTA
    const int tail_tiles = (m_sched - full_groups * grid_i) * p.n_tiles;  // whole grid
    auto tile_at = [&](int i, int& mt, int& nt) {  // i-th tile of this CTA; false past the end
        bool ok = true;
        if (i < grouped_tiles) {
            const int g = i / p.n_tiles;
            mt = cta_i + g * grid_i;
            nt = i - g * p.n_tiles;
        } else {
            const int j = cta_i + (i - grouped_tiles) * grid_i;
            mt = full_groups * grid_i + j / p.n_tiles;
            nt = j % p.n_tiles;
            ok = j < tail_tiles;
        }
        if constexpr (k2) mt = 2 * mt + static_cast<int>(rank);
        return ok;
    };

    if (warp == 0) {
        // ===================== TMA producer =====================
        // Warp-uniform like the MMA issuer (one elected lane issues), and the k-block -> (tap, channel block) walk is
        // kept in counters: as a single-thread branch with a division per k-block this loop took ~225 instructions
        // (~1000 cycles) per stage against 400..600 cycles of MMA work per stage -- the 3x3 convolutions were
        // producer-bound (ncu, profiles/r2_ncu_conv320_producer_bound.txt: MMA warp 45 % of its time waiting for data
        // with the memory system at half load).
        const bool leader = elect_one();
        int stage = 0;
        uint32_t phase = 0;
        uint32_t s_off = 0;   // stage * stage_bytes
        int mt, nt;
        for (int ti = 0; tile_at(ti, mt, nt); ++ti) {
            const TileCoord tc = tile_coord(p, mt);
            int b_row0 = nt * p.bn;
            if constexpr (k2) {   // this CTA's half of the (possibly narrower, last) N tile
                int n_this = p.N - nt * p.bn;
                n_this = n_this >= p.bn ? p.bn : ((n_this + 31) & ~31);
                b_row0 += static_cast<int>(rank) * (n_this >> 1);
            }
            const int half_k = p.ks >> 1;  // conv "same" padding: taps centred on the output pixel
            int tb = 0, tt = 0;            // temporal: batch item and frame of this tile
            if (p.mode == MOFA_A_TEMPORAL3) {
                tb = tc.frame / p.T;
                tt = tc.frame - tb * p.T;
            }
            int tap = 0, kbt = 0, ky = 0, kx = 0;   // k-block = (tap, channel block kbt); conv tap = (ky, kx)
            for (int kb = 0; kb < p.num_kb; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                if (leader) {
                    uint8_t* sa = smem + s_off;
                    uint8_t* sb = sa + kABytes;
                    // pair: the leader's barrier counts the bytes of BOTH CTAs' loads (one arrival, the leader's)
                    if (!k2 || rank == 0) mbar_arrive_expect_tx(&full_bar[stage], k2 ? 2 * stage_bytes : stage_bytes);
                    auto load2 = [&](const CUtensorMap* m, void* dst, int c0, int c1) {
                        if constexpr (k2) tma_load_2d_2sm(m, &full_bar[stage], dst, c0, c1);
                        else tma_load_2d(m, &full_bar[stage], dst, c0, c1);
                    };
                    auto load4 = [&](const CUtensorMap* m, void* dst, int c0, int c1, int c2, int c3) {
                        if constexpr (k2) tma_load_4d_2sm(m, &full_bar[stage], dst, c0, c1, c2, c3);
                        else tma_load_4d(m, &full_bar[stage], dst, c0, c1, c2, c3);
                    };
                    if (p.mode == MOFA_A_LINEAR) {
                        if (kb < p.kb_split)
                            load2(&tmA, sa, kb * BK, static_cast<int>(tc.m0));
                        else
                            load2(&tmA2, sa, (kb - p.kb_split) * BK, static_cast<int>(tc.m0));
                    } else if (p.mode == MOFA_A_CONV3X3) {
                        load4(&tmA, sa, kbt * BK, tc.x0 + (kx - half_k) * p.dil, tc.y0 + (ky - half_k) * p.dil, tc.n_img);
                    } else {
                        load4(&tmA, sa, kbt * BK, tc.p0, tt + tap - 1, tb);
                    }
                    load2(&tmB, sb, kb * BK, b_row0);
                }
                if (++kbt == p.kb_per_tap) {
                    kbt = 0;
                    ++tap;
                    if (++kx == p.ks) {
                        kx = 0;
                        ++ky;
                    }
                }
                s_off += stage_bytes;
                if (++stage == p.stages) {
                    stage = 0;
                    s_off = 0;
                    phase ^= 1;
                }
            }
        }
        __syncwarp();
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the (warp-uniform) schedule and one elected lane issues, so descriptors, barrier addresses
        // and loop state live in uniform registers (inside a single-thread branch the compiler re-elects a lane and
        // broadcasts every operand before each tcgen05 instruction).  Pair: only the leader CTA issues.
        const bool leader = elect_one();
        if (!k2 || rank == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t iter = 0;
            int mt, nt;
            for (; tile_at(static_cast<int>(iter), mt, nt); ++iter) {
                int n_this = p.N - nt * p.bn;             // the last N tile of a row may be narrower
                n_this = n_this >= p.bn ? p.bn : ((n_this + (k2 ? 31 : 15)) & ~(k2 ? 31 : 15));
                const uint32_t idesc = umma_idesc_f16_m(static_cast<uint32_t>(n_this), k2 ? 256u : 128u);
                const uint32_t as = iter & 1u;
                const uint32_t aphase = (iter >> 1) & 1u;
                mbar_wait(&tempty_bar[as], aphase ^ 1);
                tc_fence_after();
                const uint32_t tmem_d = tmem_base + as * acc_stride;
                for (int kb = 0; kb < p.num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (leader) {
                        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
                        const uint64_t da = umma_desc_sw128_kmajor(sa);
                        const uint64_t db = umma_desc_sw128_kmajor(sa + kABytes);
                        // +32 B per K=16 step inside the 128 B swizzle atom: +2 in 16 B address units
                        if constexpr (k2) {
                            umma_f16_ss_2sm(tmem_d, da, db, idesc, kb != 0 ? 1u : 0u);
                            umma_f16_ss_2sm(tmem_d, da + 2, db + 2, idesc, 1u);
                            umma_f16_ss_2sm(tmem_d, da + 4, db + 4, idesc, 1u);
                            umma_f16_ss_2sm(tmem_d, da + 6, db + 6, idesc, 1u);
                            umma_commit_2sm(&empty_bar[stage]);
                        } else {
                            umma_f16_ss(tmem_d, da, db, idesc, kb != 0 ? 1u : 0u);
                            umma_f16_ss(tmem_d, da + 2, db + 2, idesc, 1u);
                            umma_f16_ss(tmem_d, da + 4, db + 4, idesc, 1u);
                            umma_f16_ss(tmem_d, da + 6, db + 6, idesc, 1u);
                            umma_commit(&empty_bar[stage]);
                        }
                    }
                    if (++stage == p.stages) {
                        stage = 0;
                        phase ^= 1;
                    }
                }
                if (leader) {
                    if constexpr (k2) umma_commit_2sm(&tfull_bar[as]);
                    else umma_commit(&tfull_bar[as]);
                }
            }
        }
        __syncwarp();
    } else if (warp >= 2) {
        // ===================== epilogue =====================
        const int e = warp - 2;
        const int q = warp & 3;   // TMEM lane quarter this warp may touch
        const int half = e >> 2;  // which of the two warps sharing that quarter
        const int r = q * 32 + lane;
        uint8_t* stg = staging + e * (kEW == 16 ? kStagingBytes / 2 : kStagingBytes);
        const int half_bn = p.bn >> 1;
        uint32_t iter = 0;
        int mt, nt;
        for (; tile_at(static_cast<int>(iter), mt, nt); ++iter) {
            const TileCoord tc = tile_coord(p, mt);
            const uint32_t as = iter & 1u;
            const uint32_t aphase = (iter >> 1) & 1u;

            bool valid;
            long long row;
            if (p.mode == MOFA_A_LINEAR) {
                row = tc.m0 + r;
                valid = row < p.M;
            } else if (p.mode == MOFA_A_CONV3X3) {
                // tile rows are ordered (image, y, x), x fastest -- the order the 4-D TMA box lands in shared memory
                const int x = tc.x0 + r % p.BW;
                const int ry = r / p.BW;
                const int y = tc.y0 + ry % p.BH;
                const int ni = tc.n_img + ry / p.BH;
                valid = (y < p.H) && (x < p.W) && (ni < p.n_img);
                row = (static_cast<long long>(ni) * p.H + y) * p.W + x;
            } else {
                const int pp = tc.p0 + r;
                valid = pp < p.HW && tc.frame < p.n_img;   // (n_img = B * T here; beyond it only for a pair's empty tile)
                row = static_cast<long long>(tc.frame) * p.HW + pp;
            }
            const long long group =
                (p.rowbias && valid) ? (p.rowbias_mod > 0 ? row % p.rowbias_mod : row / p.rows_per_group) : 0;
            const int out_cols_tile = kGeglu ? half_bn : p.bn;          // output columns per full N tile
            int out_cols = p.N_out - nt * out_cols_tile;                // ... of this tile
            out_cols = out_cols > out_cols_tile ? out_cols_tile : out_cols;

            // kStats: statistic indices this tile touches (rows are ordered, so [first valid row, last valid row] bounds them)
            long long st_first = 0;
            int gn_g_first = 0, gn_st = -1;
            uint32_t gn_vmask = 0, gn_bmask = 0;
            bool gn_smem = false;
            if constexpr (kStats) {
                long long r_first, r_last;
                if (p.mode == MOFA_A_LINEAR) {
                    r_first = tc.m0;
                    r_last = tc.m0 + BM - 1 < p.M - 1 ? tc.m0 + BM - 1 : p.M - 1;
                } else if (p.mode == MOFA_A_CONV3X3) {
                    r_first = (static_cast<long long>(tc.n_img) * p.H + tc.y0) * p.W + tc.x0;
                    const int nl = tc.n_img + p.BN - 1 < p.n_img - 1 ? tc.n_img + p.BN - 1 : p.n_img - 1;
                    const int yl = tc.y0 + p.BH - 1 < p.H - 1 ? tc.y0 + p.BH - 1 : p.H - 1;
                    const int xl = tc.x0 + p.BW - 1 < p.W - 1 ? tc.x0 + p.BW - 1 : p.W - 1;
                    r_last = (static_cast<long long>(nl) * p.H + yl) * p.W + xl;
                } else {
                    r_first = static_cast<long long>(tc.frame) * p.HW + tc.p0;
                    r_last = static_cast<long long>(tc.frame) * p.HW + (tc.p0 + BM - 1 < p.HW - 1 ? tc.p0 + BM - 1 : p.HW - 1);
                }
                st_first = r_first / p.gn_rows_per_stat;
                gn_smem = (r_last / p.gn_rows_per_stat - st_first) < kGnSlots;
                gn_g_first = (p.gn_c_off + nt * out_cols_tile) / p.gn_cpg;
                // this thread's row: statistic index relative to the tile's first, validity and segment-boundary masks
                gn_st = valid ? static_cast<int>(row / p.gn_rows_per_stat - st_first) : -1;
                const int st_prev = __shfl_up_sync(0xffffffffu, gn_st, 1);
                gn_vmask = __ballot_sync(0xffffffffu, valid);
                gn_bmask = __ballot_sync(0xffffffffu, lane > 0 && valid && gn_st != st_prev);
            }

            // Residual rows are this thread's own (one row per thread, 64 bytes per 32-column half-chunk): a load issued
            // where it is consumed exposes a full DRAM round trip per half-chunk with only two warps per sub-partition to
            // hide it (round 1 ncu on the K = 320 projections: long_scoreboard 3.97 per issue, issue-active 26 %).  So
            // the first half-chunk's residual is requested BEFORE the accumulator wait and each later one while the
            // previous half-chunk is being computed (one extra 4 x 16-byte register buffer).
            const bool has1 = p.res1 != nullptr, has2 = p.res2 != nullptr;  // uniform
            uint4 r1n[4];
            auto load_r1 = [&](int cc, int hh) {
                const int col0n = cc * 64 + hh * 32;
                const int n0n = nt * out_cols_tile + col0n;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    r1n[g] = make_uint4(0, 0, 0, 0);
                    if (valid && col0n < out_cols && n0n + g * 8 < p.N_out)
                        r1n[g] = __ldg(reinterpret_cast<const uint4*>(p.res1 + row * p.ldr1 + n0n) + g);
                }
            };
            // (not in the statistics kernel: its extra state would spill; its GEMMs are main-loop bound convolutions)
            constexpr bool kPrefetchRes = !kStats;
            if (kEW == 8 && kPrefetchRes && p.tma_out && has1 && half * 64 < out_cols) load_r1(half, 0);

            mbar_wait(&tfull_bar[as], aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q * 32) << 16) + as * acc_stride;

            if constexpr (kEW == 16 && kEpi == 0) {
                // plain epilogue, LINEAR mode, short K (the q|k|v / out / in projections at K = 320 are epilogue- and
                // latency-bound): warp (q, sub) drains the 32-column units sub, sub + 4, ... in two 16-column passes
                const int sub = e >> 2;
                const int n_units = (out_cols + 31) >> 5;
#pragma unroll 1
                for (int u = sub; u < n_units; u += 4) {
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll 1
                    for (int hh = 0; hh < 2; ++hh) {
                        const int col0 = u * 32 + hh * 16;
                        const int n0 = nt * p.bn + col0;       // global output column (= bias column: plain epilogue)
                        uint4 r1v[2], r2v[2], rbv[2], bvv[2];
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            const bool in = n0 + g * 8 < p.N_out;
                            r1v[g] = r2v[g] = rbv[g] = bvv[g] = make_uint4(0, 0, 0, 0);
                            if (has1 && valid && in)
                                r1v[g] = __ldg(reinterpret_cast<const uint4*>(p.res1 + row * p.ldr1 + n0) + g);
                            if (has2 && valid && in)
                                r2v[g] = __ldg(reinterpret_cast<const uint4*>(p.res2 + row * p.ldr2 + n0) + g);
                            if (p.rowbias && in)
                                rbv[g] = __ldg(reinterpret_cast<const uint4*>(p.rowbias + group * p.ld_rowbias + n0) + g);
                            if (p.bias && in) bvv[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + n0) + g);
                        }
                        uint32_t acc[16];
                        tmem_ld_32x16(taddr + col0, acc);
                        tmem_ld_wait();
#pragma unroll
                        for (int g = 0; g < 2; ++g) {
                            H8 b, rb, a1, a2, o;
                            b.u = bvv[g];
                            rb.u = rbv[g];
                            a1.u = r1v[g];
                            a2.u = r2v[g];
#pragma unroll
                            for (int j = 0; j < 8; ++j) {
                                float x = __uint_as_float(acc[g * 8 + j]) + __half2float(b.h[j]) + __half2float(rb.h[j]);
                                x *= p.alpha;
                                x = fmaf(p.beta1, __half2float(a1.h[j]), x);
                                x = fmaf(p.beta2, __half2float(a2.h[j]), x);
                                o.h[j] = __float2half_rn(x);
                            }
                            const int ci = hh * 2 + g;   // 16-byte chunk of the 64-byte staging row (SWIZZLE_64B)
                            *reinterpret_cast<uint4*>(stg + lane * 64 + ((ci ^ ((lane >> 1) & 3)) << 4)) = o.u;
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, stg, nt * p.bn + u * 32, static_cast<int>(tc.m0) + q * 32);
                        tma_store_commit();
                    }
                }
            } else if constexpr (kEW == 16) {
                // GEGLU, LINEAR mode: warp (q, sub) drains the 32-column half-chunks sub, sub + 4, ... of its lane quarter
                const int sub = e >> 2;
                const int n_units = (out_cols + 31) >> 5;
#pragma unroll 1
                for (int u = sub; u < n_units; u += 4) {
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
                    const int col0 = u * 32;
                    uint32_t a8[8], g8[8];
                    tmem_ld_32x8(taddr + col0, a8);
                    tmem_ld_32x8(taddr + half_bn + col0, g8);
#pragma unroll 1
                    for (int g = 0; g < 4; ++g) {
                        H8 bv, bg;
                        bv.u = make_uint4(0, 0, 0, 0);
                        bg.u = make_uint4(0, 0, 0, 0);
                        if (p.bias) {
                            const int nb = nt * p.bn + col0 + g * 8;
                            bv.u = __ldg(reinterpret_cast<const uint4*>(p.bias + nb));
                            bg.u = __ldg(reinterpret_cast<const uint4*>(p.bias + nb + half_bn));
                        }
                        tmem_ld_wait();
                        float va[8], ga[8];
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            va[j] = __uint_as_float(a8[j]);
                            ga[j] = __uint_as_float(g8[j]);
                        }
                        if (g < 3) {
                            tmem_ld_32x8(taddr + col0 + (g + 1) * 8, a8);
                            tmem_ld_32x8(taddr + half_bn + col0 + (g + 1) * 8, g8);
                        }
                        H8 o;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 y = geglu2(make_float2(va[2 * j], va[2 * j + 1]),
                                                    make_float2(ga[2 * j], ga[2 * j + 1]), bv.h2[j], bg.h2[j]);
                            o.h2[j] = __floats2half2_rn(y.x, y.y);
                        }
                        // 64-byte rows, SWIZZLE_64B: 16-byte chunk g of row r lives at chunk g ^ ((r >> 1) & 3)
                        *reinterpret_cast<uint4*>(stg + lane * 64 + ((g ^ ((lane >> 1) & 3)) << 4)) = o.u;
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if (lane == 0) {
                        tma_store_2d(&tmOut, stg, nt * out_cols_tile + col0, static_cast<int>(tc.m0) + q * 32);
                        tma_store_commit();
                    }
                }
            } else if (p.tma_out) {
                const int n_chunks = (out_cols + 63) >> 6;
#pragma unroll 1
                for (int c = half; c < n_chunks; c += 2) {
                    // the staging tile is free once the previous TMA store of this warp has read it
                    if (lane == 0) tma_store_wait_read<0>();
                    __syncwarp();
#pragma unroll 1
                    for (int hlf = 0; hlf < 2; ++hlf) {
                        const int col0 = c * 64 + hlf * 32;           // column inside the tile's output
                        if (col0 >= out_cols) break;
                        const int n_out0 = nt * out_cols_tile + col0; // global output column
                        // Every optional step below sits behind a WARP-UNIFORM branch around a whole 32-element
                        // loop, so the common case issues only: residual loads, tcgen05.ld, bias add, convert, store.
                        // (ncu on the first version: predicated-off activation/residual code still cost issue slots,
                        //  ~15 instructions per output element.)
                        if constexpr (kGeglu && kGegluCompact) {
                            // GEGLU half-chunk as a ROLLED loop over four 8-column groups (x8 TMEM loads, the next group's
                            // loads in flight while this group's GELUs run): ~200 instructions of loop body instead of a
                            // 32-wide unrolled ~700 (the epilogue was fetch-limited: stall_no_instruction 0.45 per issue).
                            // GEGLU launches carry no residual / row bias / alpha (mofa_gemm checks).
                            uint32_t a8[8], g8[8];
                            tmem_ld_32x8(taddr + col0, a8);
                            tmem_ld_32x8(taddr + half_bn + col0, g8);
#pragma unroll 1
                            for (int g = 0; g < 4; ++g) {
                                H8 bv, bg;
                                bv.u = make_uint4(0, 0, 0, 0);
                                bg.u = make_uint4(0, 0, 0, 0);
                                if (p.bias) {
                                    const int nb = nt * p.bn + col0 + g * 8;
                                    bv.u = __ldg(reinterpret_cast<const uint4*>(p.bias + nb));
                                    bg.u = __ldg(reinterpret_cast<const uint4*>(p.bias + nb + half_bn));
                                }
                                tmem_ld_wait();
                                float va[8], ga[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    va[j] = __uint_as_float(a8[j]);
                                    ga[j] = __uint_as_float(g8[j]);
                                }
                                if (g < 3) {
                                    tmem_ld_32x8(taddr + col0 + (g + 1) * 8, a8);
                                    tmem_ld_32x8(taddr + half_bn + col0 + (g + 1) * 8, g8);
                                }
                                H8 o;
#pragma unroll
                                for (int j = 0; j < 4; ++j) {
                                    const float2 y = geglu2(make_float2(va[2 * j], va[2 * j + 1]),
                                                            make_float2(ga[2 * j], ga[2 * j + 1]), bv.h2[j], bg.h2[j]);
                                    o.h2[j] = __floats2half2_rn(y.x, y.y);
                                }
                                const int ci = hlf * 4 + g;
                                *reinterpret_cast<uint4*>(stg + lane * 128 + ((ci ^ (lane & 7)) << 4)) = o.u;
                            }
                            continue;
                        }
                        uint4 r1[4], r2[4], rbv[4];
                        if (!kGeglu && p.rowbias) {  // row-group bias: requested with the residuals, not after the
#pragma unroll                                       // accumulator wait (its L2 latency sat on the critical path)
                            for (int g = 0; g < 4; ++g) {
                                const int nb = nt * p.bn + col0 + g * 8;
                                rbv[g] = make_uint4(0, 0, 0, 0);
                                if (nb < p.N_out)
                                    rbv[g] = __ldg(reinterpret_cast<const uint4*>(p.rowbias + group * p.ld_rowbias + nb));
                            }
                        }
                        if (has1) {
                            if constexpr (kPrefetchRes) {  // requested one half-chunk ago; now request this warp's next one
#pragma unroll
                                for (int g = 0; g < 4; ++g) r1[g] = r1n[g];
                                const int cn = hlf == 0 ? c : c + 2;
                                if (cn * 64 + (hlf ^ 1) * 32 < out_cols && cn < n_chunks) load_r1(cn, hlf ^ 1);
                            } else {
                                load_r1(c, hlf);
#pragma unroll
                                for (int g = 0; g < 4; ++g) r1[g] = r1n[g];
                            }
                        }
                        if (has2) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                r2[g] = make_uint4(0, 0, 0, 0);
                                if (valid && n_out0 + g * 8 < p.N_out)
                                    r2[g] = __ldg(reinterpret_cast<const uint4*>(p.res2 + row * p.ldr2 + n_out0) + g);
                            }
                        }
                        float v[32];
                        {
                            uint32_t acc[32];
                            tmem_ld_32x32(taddr + col0, acc);
                            if constexpr (kGeglu) {
                                uint32_t ag[32];
                                tmem_ld_32x32(taddr + half_bn + col0, ag);
                                // bias requested while the two TMEM loads are in flight (it tested slower to pre-load the
                                // bias into the accumulator with tcgen05.st before the MMAs: 525 vs 678 TFLOP/s at K = 320,
                                // profiles/r2_step_detail_a.txt -- the pre-load sits on the epilogue's critical path)
                                uint4 bvv[4], bgv[4];
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    const int nb = nt * p.bn + col0 + g * 8;
                                    bvv[g] = make_uint4(0, 0, 0, 0);
                                    bgv[g] = make_uint4(0, 0, 0, 0);
                                    if (p.bias) {
                                        bvv[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + nb));
                                        bgv[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + nb + half_bn));
                                    }
                                }
                                tmem_ld_wait();
#pragma unroll
                                for (int g = 0; g < 4; ++g) {
                                    H8 bv, bg;
                                    bv.u = bvv[g];
                                    bg.u = bgv[g];
#pragma unroll
                                    for (int j = 0; j < 4; ++j) {
                                        const float2 y = geglu2(
                                            make_float2(__uint_as_float(acc[g * 8 + 2 * j]), __uint_as_float(acc[g * 8 + 2 * j + 1])),
                                            make_float2(__uint_as_float(ag[g * 8 + 2 * j]), __uint_as_float(ag[g * 8 + 2 * j + 1])),
                                            bv.h2[j], bg.h2[j]);
                                        v[g * 8 + 2 * j] = y.x;
                                        v[g * 8 + 2 * j + 1] = y.y;
                                    }
                                }
                            } else {
                                uint4 bvv[4];  // bias: requested while the TMEM load is in flight
                                if (p.bias) {
#pragma unroll
                                    for (int g = 0; g < 4; ++g) {
                                        const int nb = nt * p.bn + col0 + g * 8;
                                        bvv[g] = make_uint4(0, 0, 0, 0);
                                        if (nb < p.N_out) bvv[g] = __ldg(reinterpret_cast<const uint4*>(p.bias + nb));
                                    }
                                }
                                tmem_ld_wait();
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(acc[j]);
                                if (p.bias) {
#pragma unroll
                                    for (int g = 0; g < 4; ++g) {
                                        H8 b;
                                        b.u = bvv[g];
#pragma unroll
                                        for (int j = 0; j < 8; ++j) v[g * 8 + j] += __half2float(b.h[j]);
                                    }
                                }
                                if (p.rowbias) {
#pragma unroll
                                    for (int g = 0; g < 4; ++g) {
                                        const int nb = nt * p.bn + col0 + g * 8;
                                        if (nb < p.N_out) {
                                            H8 b;
                                            b.u = rbv[g];
#pragma unroll
                                            for (int j = 0; j < 8; ++j) v[g * 8 + j] += __half2float(b.h[j]);
                                        }
                                    }
                                }
                                if constexpr (kEpi == 2) {
                                if (p.act == 1) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = silu_f(v[j]);
                                } else if (p.act == 3) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                                } else if (p.act == 4) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = sigmoid_f(v[j]);
                                } else if (p.act == 6) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = gelu_erf_f(v[j]);
                                } else if (p.act == 7) {
#pragma unroll
                                    for (int j = 0; j < 32; ++j) v[j] = v[j] * sigmoid_f(1.702f * v[j]);
                                }
                                }
                            }
                        }
                        if (p.alpha != 1.0f) {
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
                        }
                        if (has1) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                H8 a;
                                a.u = r1[g];
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[g * 8 + j] = fmaf(p.beta1, __half2float(a.h[j]), v[g * 8 + j]);
                            }
                        }
                        if (has2) {
#pragma unroll
                            for (int g = 0; g < 4; ++g) {
                                H8 a;
                                a.u = r2[g];
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[g * 8 + j] = fmaf(p.beta2, __half2float(a.h[j]), v[g * 8 + j]);
                            }
                        }
                        if constexpr (kEpi == 2) {
                            if (p.act == 5) {  // ReLU after the residual add (ResNet bottleneck)
#pragma unroll
                                for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
                            }
                        }
#pragma unroll
                        for (int g = 0; g < 4; ++g) {
                            H8 o;
#pragma unroll
                            for (int j = 0; j < 4; ++j) o.h2[j] = __floats2half2_rn(v[g * 8 + 2 * j], v[g * 8 + 2 * j + 1]);
                            const int ci = hlf * 4 + g;  // 16-byte chunk inside the 128-byte staging row
                            *reinterpret_cast<uint4*>(stg + lane * 128 + ((ci ^ (lane & 7)) << 4)) = o.u;
                        }
                    }
                    fence_proxy_async_smem();
                    __syncwarp();
                    if constexpr (kStats) if (p.gn_debug < 3) {
                        // GroupNorm statistics of what was just staged (the fp16-rounded outputs the consumer will read):
                        // lane l owns the column pair (2l, 2l+1) of this 64-column chunk and walks the warp's 32 rows of
                        // the swizzled staging tile (one 128-byte row per step: conflict-free), fp32 sums; a group never
                        // splits a pair (channels per group is even).  Common case: all 32 rows valid and inside one
                        // statistic -> a rolled 32-step loop and two atomics per lane; otherwise (ragged tiles, several
                        // small images per warp) a segment loop driven by two ballots.
                        const uint32_t vmask = gn_vmask, bmask = gn_bmask;
                        const int colp = nt * out_cols_tile + c * 64 + 2 * lane;     // global output column of the pair
                        const bool col_ok = colp < p.N_out;
                        const int grp = (p.gn_c_off + colp) / p.gn_cpg;
                        const int g_loc = grp - gn_g_first;
                        const bool to_smem = gn_smem && g_loc < kGnGroups;           // else: straight to global memory
                        float s0 = 0.f, s1 = 0.f, q0 = 0.f, q1 = 0.f;
                        // Lanes of one group are neighbours (cpg / 2 of them).  fp32 atomicAdd on shared memory is a
                        // compare-and-swap loop (ATOMS.CAST.SPIN): ten lanes on one address = ten rounds, and it made the
                        // temporal-conv GEMMs 2.2x slower (profiles/r2_step_detail_f/g).  So: inclusive prefix sums over the
                        // lanes, each group's total = difference of two prefixes, and only the LAST lane of a group touches
                        // memory (distinct addresses within the warp).
                        auto add_stat = [&](int st_seg) {   // st_seg: statistic index relative to st_first
                            const int gkey = col_ok ? grp : -1;
                            float ps = col_ok ? s0 + s1 : 0.f, pq = col_ok ? q0 + q1 : 0.f;
#pragma unroll
                            for (int o = 1; o < 32; o <<= 1) {
                                const float ts = __shfl_up_sync(0xffffffffu, ps, o), tq = __shfl_up_sync(0xffffffffu, pq, o);
                                if (lane >= o) {
                                    ps += ts;
                                    pq += tq;
                                }
                            }
                            const int g_next = __shfl_down_sync(0xffffffffu, gkey, 1);
                            const int g_prev = __shfl_up_sync(0xffffffffu, gkey, 1);
                            const bool is_last = lane == 31 || g_next != gkey;
                            const bool is_first = lane == 0 || g_prev != gkey;
                            // prefix just before my group's first lane: broadcast from the group's first lane
                            const uint32_t firsts = __ballot_sync(0xffffffffu, is_first);
                            const int first_lane = 31 - __clz(firsts & (0xffffffffu >> (31 - lane)));
                            const float bs = __shfl_sync(0xffffffffu, ps, (first_lane + 31) & 31);
                            const float bq = __shfl_sync(0xffffffffu, pq, (first_lane + 31) & 31);
                            if (is_last && gkey >= 0 && p.gn_debug < 1) {
                                const float ts = first_lane > 0 ? ps - bs : ps, tq = first_lane > 0 ? pq - bq : pq;
                                if (to_smem) {
                                    const uint32_t d = smem_u32(s_gn) + static_cast<uint32_t>(
                                        (((static_cast<int>(iter % kGnBufs) * kGnSlots + st_seg) * kGnGroups + g_loc) * 2) * 4);
                                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(d), "f"(ts) : "memory");
                                    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(d + 4), "f"(tq) : "memory");
                                } else {
                                    float* d = p.gn_stats + ((st_first + st_seg) * p.gn_groups + grp) * 2;
                                    red_add_global(d, ts);
                                    red_add_global(d + 1, tq);
                                }
                            }
                        };
                        const uint8_t* lane_ptr = stg + ((lane & 3) << 2);
                        const uint32_t chunk = lane >> 2;
                        if (vmask == 0xffffffffu && bmask == 0u && p.gn_debug < 2) {
                            // common case -- 32 valid rows of one statistic: straight-line, all 32 loads independent
                            // (a rolled loop with one load per trip cost 0.21 ms per level-0 temporal-conv launch: its 170
                            // clocks per trip were the load-to-use latency, serialised; profiles/r2_gnstats_experiments.txt)
                            uint32_t hv[32];
#pragma unroll
                            for (int rr = 0; rr < 32; ++rr)
                                hv[rr] = *reinterpret_cast<const uint32_t*>(lane_ptr + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                            float t0 = 0.f, t1 = 0.f, u0 = 0.f, u1 = 0.f;   // second accumulator set: shorter add chains
#pragma unroll
                            for (int rr = 0; rr < 32; rr += 2) {
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hv[rr]));
                                const float2 g2 = __half22float2(*reinterpret_cast<const __half2*>(&hv[rr + 1]));
                                s0 += f.x;
                                s1 += f.y;
                                q0 = fmaf(f.x, f.x, q0);
                                q1 = fmaf(f.y, f.y, q1);
                                t0 += g2.x;
                                t1 += g2.y;
                                u0 = fmaf(g2.x, g2.x, u0);
                                u1 = fmaf(g2.y, g2.y, u1);
                            }
                            s0 += t0;
                            s1 += t1;
                            q0 += u0;
                            q1 += u1;
                            add_stat(__shfl_sync(0xffffffffu, gn_st, 0));
                        } else {
                            int seg_row = -1;
#pragma unroll 1
                            for (int rr = 0; rr < (p.gn_debug >= 2 ? 1 : 32); ++rr) {
                                if (!((vmask >> rr) & 1u)) continue;                      // warp-uniform
                                if (seg_row >= 0 && ((bmask >> rr) & 1u)) {               // statistic index changes here
                                    add_stat(__shfl_sync(0xffffffffu, gn_st, seg_row));
                                    s0 = s1 = q0 = q1 = 0.f;
                                    seg_row = rr;
                                }
                                if (seg_row < 0) seg_row = rr;
                                const uint32_t hv =
                                    *reinterpret_cast<const uint32_t*>(lane_ptr + rr * 128 + ((chunk ^ (rr & 7)) << 4));
                                const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&hv));
                                s0 += f.x;
                                s1 += f.y;
                                q0 = fmaf(f.x, f.x, q0);
                                q1 = fmaf(f.y, f.y, q1);
                            }
                            if (seg_row >= 0) add_stat(__shfl_sync(0xffffffffu, gn_st, seg_row));
                        }
                    }
                    if (lane == 0) {
                        const int n0 = nt * out_cols_tile + c * 64;
                        if (p.mode == MOFA_A_LINEAR) {
                            tma_store_2d(&tmOut, stg, n0, static_cast<int>(tc.m0) + q * 32);
                        } else if (p.mode == MOFA_A_CONV3X3) {
                            // this warp's 32 rows = sBN images x sBH rows x BW pixels of the tile
                            const int ry0 = (q * 32) / p.BW;
                            tma_store_4d(&tmOut, stg, n0, tc.x0, tc.y0 + ry0 % p.BH, tc.n_img + ry0 / p.BH);
                        } else {
                            const int b = tc.frame / p.T;
                            tma_store_4d(&tmOut, stg, n0, tc.p0 + q * 32, tc.frame - b * p.T, b);
                        }
                        tma_store_commit();
                    }
                }
            } else {
                // narrow / unaligned outputs: 32-column chunks straight from registers
                const int chunks = (p.bn + 31) / 32;
#pragma unroll 1
                for (int c = half; c < chunks; c += 2) {
                    uint32_t acc[32];
                    tmem_ld_32x32(taddr + c * 32, acc);
                    tmem_ld_wait();
                    if (valid) {
#pragma unroll 1
                        for (int g = 0; g < 4; ++g) {
                            const int ncol = c * 32 + g * 8;
                            const int n = nt * p.bn + ncol;
                            if (ncol < p.bn && n < p.N_out) {
                                float v[8];
#pragma unroll
                                for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(acc[g * 8 + j]);
                                epilogue_direct8(p, v, row, group, n, n, p.act == 1);
                            }
                        }
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if constexpr (k2) mbar_arrive_leader(&tempty_bar[as]);
                else mbar_arrive(&tempty_bar[as]);
            }
            if constexpr (kStats) {
                if (gn_smem) {
                    // No barrier: every warp counts itself done with this tile's buffer and the LAST one moves the partial
                    // sums to global memory.  (Two bar.sync per tile lock-stepped the eight warps and doubled the
                    // temporal-conv GEMMs: profiles/r2_step_detail_f_gnstats_smem_barriers.txt.)  Three buffers: a warp
                    // can be at most two tiles ahead of the slowest one (two TMEM stages), so the buffer of tile i is
                    // flushed and zeroed long before tile i + 3 touches it.
                    const int buf = static_cast<int>(iter % kGnBufs);
                    __syncwarp();
                    int done = 0;
                    if (lane == 0) {
                        __threadfence_block();
                        done = atomicAdd(&s_gn_cnt[buf], 1);
                    }
                    done = __shfl_sync(0xffffffffu, done, 0);
                    if (done == kEW - 1) {
                        __threadfence_block();
                        float* sb = s_gn + buf * kGnSlots * kGnGroups * 2;
                        for (int i = lane; i < kGnSlots * kGnGroups; i += 32) {
                            const float a = sb[2 * i], b = sb[2 * i + 1];   // all eight warps are done with this buffer
                            sb[2 * i] = 0.f;
                            sb[2 * i + 1] = 0.f;
                            if (a != 0.f || b != 0.f) {
                                const int g = gn_g_first + (i % kGnGroups);
                                float* d = p.gn_stats + ((st_first + i / kGnGroups) * p.gn_groups + g) * 2;
                                red_add_global(d, a);
                                red_add_global(d + 1, b);
                            }
                        }
                        __syncwarp();
                        if (lane == 0) s_gn_cnt[buf] = 0;
                    }
                }
            }
        }
        if (lane == 0) tma_store_wait_all<0>();  // smem must outlive the last bulk store
    }

    tc_fence_before();
    if constexpr (k2) cluster_sync_all();   // the peer may still be reading this CTA's shared / tensor memory
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if constexpr (k2) tmem_dealloc_2sm(tmem_base, kTmemCols);
        else tmem_dealloc(tmem_base, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode_fn() {
    static PFN_tmapEncodeTiled fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres);
        if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !ptr) {
            set_last_error("cuTensorMapEncodeTiled entry point not available (%s)", cudaGetErrorString(e));
            return nullptr;
        }
        fn = reinterpret_cast<PFN_tmapEncodeTiled>(ptr);
    }
    return fn;
}

// fp16 tensor map, SWIZZLE_128B, inner box = 64 elements (128 B); dims innermost first
int make_tmap_f16_sw(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, CUtensorMapSwizzle swizzle);

int make_tmap_f16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
    return make_tmap_f16_sw(m, ptr, rank, dims, strides_bytes, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

int make_tmap_f16_sw(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, CUtensorMapSwizzle swizzle) {
    PFN_tmapEncodeTiled fn = get_encode_fn();
    if (!fn) return MOFA_ERR_CUDA;
    cuuint64_t gdims[5];
    cuuint64_t gstrides[4];
    cuuint32_t gbox[5];
    cuuint32_t estr[5];
    for (int i = 0; i < rank; ++i) {
        gdims[i] = dims[i];
        gbox[i] = box[i];
        estr[i] = 1;
        if (i > 0) gstrides[i - 1] = strides_bytes[i - 1];
    }
    if ((reinterpret_cast<uintptr_t>(ptr) & 15) != 0) {
        set_last_error("tensor map base %p not 16-byte aligned", ptr);
        return MOFA_ERR_ARG;
    }
    for (int i = 0; i + 1 < rank; ++i) {
        if (gstrides[i] % 16 != 0) {
            set_last_error("tensor map stride %d (= %llu bytes) not a multiple of 16", i,
                           static_cast<unsigned long long>(gstrides[i]));
            return MOFA_ERR_ARG;
        }
    }
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(ptr), gdims,
                    gstrides, gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_last_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dims %llu %llu %llu %llu)", (int)r,
                       rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                       (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0));
        return MOFA_ERR_CUDA;
    }
    return MOFA_OK;
}

int num_sms() {
    static int n = 0;
    if (!n) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

static int floor_pow2(int x) {
    int p = 1;
    while (p * 2 <= x) p *= 2;
    return p;
}

}  // namespace mofa

using namespace mofa;

extern "C" int mofa_gemm(const mofa_gemm_args* a, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!a || !a->a || !a->w || !a->out) {
        set_last_error("mofa_gemm: null argument");
        return MOFA_ERR_ARG;
    }
    const bool geglu = a->act == 2;
    const int bn = a->bn;
    // 16 epilogue warps (LINEAR mode, K <= 384: five or six k-blocks per tile, the epilogue is the bottleneck): GEGLU and the
    // plain epilogue without GroupNorm statistics
    const bool short_k = a->mode == MOFA_A_LINEAR && a->K <= 384 && !a->a2;
    if (bn < 16 || bn > 256 || (bn % 16) != 0 || (geglu && (bn % 128) != 0)) {
        set_last_error("mofa_gemm: unsupported bn=%d (act=%d)", bn, a->act);
        return MOFA_ERR_ARG;
    }
    if (a->N <= 0 || (geglu && (a->N % bn) != 0)) {
        set_last_error("mofa_gemm: bad N=%d for bn=%d", a->N, bn);
        return MOFA_ERR_ARG;
    }
    const int n_out_cols = geglu ? a->N / 2 : a->N;
    const bool vec_ok = (n_out_cols % 8) == 0 && (a->ldc % 8) == 0 && !(a->res1 && (a->ldr1 % 8)) &&
                        !(a->res2 && (a->ldr2 % 8)) && !(a->rowbias && (a->ld_rowbias % 8)) &&
                        (reinterpret_cast<uintptr_t>(a->out) % 16) == 0;
    // TMA-store epilogue: 64-column chunks must not straddle N tiles (bn, or bn/2 for GEGLU, multiple of 64)
    const bool tma_out = vec_ok && n_out_cols >= 64 && ((geglu ? bn / 2 : bn) % 64) == 0;
    if (geglu && !tma_out) {
        set_last_error("mofa_gemm: GEGLU needs 16-byte aligned rows and bn %% 128 == 0");
        return MOFA_ERR_ARG;
    }
    if (geglu && (a->rowbias || a->res1 || a->res2 || a->alpha != 1.0f)) {
        set_last_error("mofa_gemm: GEGLU takes a bias only (no row-group bias, residual or alpha)");
        return MOFA_ERR_ARG;
    }

    // (plain: only with a residual -- its per-row loads are what the extra warps hide: out-projection with row bias + residual
    //  295 -> 360 TFLOP/s; the residual-free q|k|v projection is faster on 8 warps with 64-column stores: 786 vs 630)
    const bool use_ew16 = kGeglu16 && short_k && tma_out &&
                          (geglu || (a->act == 0 && !a->gn_stats && a->res1 != nullptr)) &&
                          ((geglu ? bn / 2 : bn) % 32) == 0;
    GemmKernelParams p;
    memset(&p, 0, sizeof(p));
    p.mode = a->mode;
    p.act = a->act;
    p.bn = bn;
    p.N = a->N;
    p.n_tiles = (a->N + bn - 1) / bn;
    p.N_out = n_out_cols;
    p.tma_out = tma_out ? 1 : 0;
    p.out = static_cast<__half*>(a->out);
    p.ldc = a->ldc;
    p.bias = static_cast<const __half*>(a->bias);
    p.rowbias = static_cast<const __half*>(a->rowbias);
    p.ld_rowbias = a->ld_rowbias;
    p.rows_per_group = a->rows_per_group > 0 ? a->rows_per_group : 1;
    p.rowbias_mod = a->rowbias_mod;
    p.vec_ok = vec_ok ? 1 : 0;
    p.res1 = static_cast<const __half*>(a->res1);
    p.ldr1 = a->ldr1;
    p.res2 = static_cast<const __half*>(a->res2);
    p.ldr2 = a->ldr2;
    p.alpha = a->alpha;
    p.beta1 = a->beta1;
    p.beta2 = a->beta2;
    p.gn_stats = nullptr;
    if (a->gn_stats) {
        if (!tma_out || a->gn_rows_per_stat <= 0 || a->gn_groups <= 0 || a->gn_cpg <= 0 || (a->gn_cpg & 1) ||
            a->gn_c_off < 0 || (a->gn_c_off & 1)) {
            set_last_error("mofa_gemm: gn_stats needs the TMA-store epilogue (16-byte rows, N_out >= 64) and an even "
                           "gn_cpg / gn_c_off (cpg=%d c_off=%d)", a->gn_cpg, a->gn_c_off);
            return MOFA_ERR_ARG;
        }
        p.gn_stats = a->gn_stats;
        {
            static int dbg = -1;
            if (dbg < 0) {
                const char* e = getenv("MOFA_GN_DEBUG");
                dbg = e ? atoi(e) : 0;
            }
            p.gn_debug = dbg;
        }
        p.gn_rows_per_stat = a->gn_rows_per_stat;
        p.gn_groups = a->gn_groups;
        p.gn_cpg = a->gn_cpg;
        p.gn_c_off = a->gn_c_off;
    }
    p.BW = 1;
    p.BH = 1;
    p.BN = 1;
    p.sBH = p.sBN = 1;
    p.tiles_p = 1;
    p.tiles_x = p.tiles_y = 1;
    p.T = 1;
    p.HW = 1;

    CUtensorMap tmA, tmA2, tmB, tmOut;
    long long Ktot = 0;
    int rc;
    const uint64_t ldc_b = static_cast<uint64_t>(a->ldc) * 2;
    if (a->mode == MOFA_A_LINEAR) {
        if (a->M <= 0 || a->K <= 0 || (a->K % 8) != 0 || (a->lda % 8) != 0) {
            set_last_error("mofa_gemm linear: bad M=%lld K=%lld lda=%lld", (long long)a->M, (long long)a->K,
                           (long long)a->lda);
            return MOFA_ERR_ARG;
        }
        Ktot = a->K;
        long long K1 = a->a2 ? a->K1 : a->K;
        if (a->a2 && ((K1 % BK) != 0 || K1 <= 0 || K1 >= a->K || (a->lda2 % 8) != 0)) {
            set_last_error("mofa_gemm linear split: K1=%lld must be a multiple of 64 inside (0,K)", (long long)K1);
            return MOFA_ERR_ARG;
        }
        p.M = a->M;
        p.m_tiles = static_cast<int>((a->M + BM - 1) / BM);
        p.num_kb = static_cast<int>((Ktot + BK - 1) / BK);
        p.kb_split = a->a2 ? static_cast<int>(K1 / BK) : p.num_kb;
        p.kb_per_tap = p.num_kb;
        uint64_t dims[2] = {static_cast<uint64_t>(K1), static_cast<uint64_t>(a->M)};
        uint64_t strides[1] = {static_cast<uint64_t>(a->lda) * 2};
        uint32_t box[2] = {BK, BM};
        if ((rc = make_tmap_f16(&tmA, a->a, 2, dims, strides, box)) != MOFA_OK) return rc;
        if (a->a2) {
            uint64_t dims2[2] = {static_cast<uint64_t>(a->K - K1), static_cast<uint64_t>(a->M)};
            uint64_t strides2[1] = {static_cast<uint64_t>(a->lda2) * 2};
            if ((rc = make_tmap_f16(&tmA2, a->a2, 2, dims2, strides2, box)) != MOFA_OK) return rc;
        } else {
            tmA2 = tmA;
        }
        if (tma_out) {
            uint64_t od[2] = {static_cast<uint64_t>(n_out_cols), static_cast<uint64_t>(a->M)};
            uint64_t os[1] = {ldc_b};
            if (use_ew16) {   // 16 epilogue warps: {32 columns, 32 rows} stores out of 64-byte-row staging tiles
                uint32_t ob[2] = {32, 32};
                if ((rc = make_tmap_f16_sw(&tmOut, a->out, 2, od, os, ob, CU_TENSOR_MAP_SWIZZLE_64B)) != MOFA_OK) return rc;
            } else {
                uint32_t ob[2] = {64, 32};
                if ((rc = make_tmap_f16(&tmOut, a->out, 2, od, os, ob)) != MOFA_OK) return rc;
            }
        }
    } else if (a->mode == MOFA_A_CONV3X3) {
        if (a->n_img <= 0 || a->H <= 0 || a->W <= 0 || a->C <= 0 || (a->C % BK) != 0) {
            set_last_error("mofa_gemm conv3x3: needs C %% 64 == 0 (C=%d)", a->C);
            return MOFA_ERR_ARG;
        }
        p.ks = a->ksize > 0 ? a->ksize : 3;
        if ((p.ks & 1) == 0 || p.ks > 7) {
            set_last_error("mofa_gemm conv: odd kernel size <= 7 expected (ksize=%d)", p.ks);
            return MOFA_ERR_ARG;
        }
        Ktot = static_cast<long long>(p.ks) * p.ks * a->C;
        p.H = a->H;
        p.W = a->W;
        p.dil = a->dilation > 0 ? a->dilation : 1;
        // M tile = BN images x BH rows x BW pixels (powers of two, product 128): the split with the fewest tiles.  A 9 x 16
        // feature map (level 3 of 576 x 1024) as 8-row tiles wastes 7/16 of every second tile (round 1: 681 vs 1209 TFLOP/s
        // for the same conv at 18 x 32); as {16 px, 1 row, 8 images} it wastes 12 %.
        p.BW = floor_pow2(a->W < 32 ? a->W : 32);
        {
            long long best = -1;
            for (int bh = BM / p.BW; bh >= 1; bh >>= 1) {
                const int bnn = BM / (p.BW * bh);
                const long long tiles = static_cast<long long>((a->W + p.BW - 1) / p.BW) * ((a->H + bh - 1) / bh) *
                                        ((a->n_img + bnn - 1) / bnn);
                if (best < 0 || tiles < best) {
                    best = tiles;
                    p.BH = bh;
                    p.BN = bnn;
                }
            }
        }
        p.n_img = a->n_img;
        p.tiles_x = (a->W + p.BW - 1) / p.BW;
        p.tiles_y = (a->H + p.BH - 1) / p.BH;
        p.m_tiles = ((a->n_img + p.BN - 1) / p.BN) * p.tiles_x * p.tiles_y;
        p.sBH = (32 / p.BW) < p.BH ? (32 / p.BW) : p.BH;
        if (p.sBH < 1) p.sBH = 1;
        p.sBN = 32 / (p.BW * p.sBH);
        if (p.sBN < 1) p.sBN = 1;
        p.kb_per_tap = a->C / BK;
        p.num_kb = p.ks * p.ks * p.kb_per_tap;
        p.kb_split = p.num_kb;
        uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
        uint64_t strides[3] = {(uint64_t)a->C * 2, (uint64_t)a->W * a->C * 2, (uint64_t)a->H * a->W * a->C * 2};
        uint32_t box[4] = {BK, (uint32_t)p.BW, (uint32_t)p.BH, (uint32_t)p.BN};
        if ((rc = make_tmap_f16(&tmA, a->a, 4, dims, strides, box)) != MOFA_OK) return rc;
        tmA2 = tmA;
        if (tma_out) {
            uint64_t od[4] = {(uint64_t)n_out_cols, (uint64_t)a->W, (uint64_t)a->H, (uint64_t)a->n_img};
            uint64_t os[3] = {ldc_b, (uint64_t)a->W * ldc_b, (uint64_t)a->H * a->W * ldc_b};
            uint32_t ob[4] = {64, (uint32_t)p.BW, (uint32_t)p.sBH, (uint32_t)p.sBN};
            if ((rc = make_tmap_f16(&tmOut, a->out, 4, od, os, ob)) != MOFA_OK) return rc;
        }
    } else if (a->mode == MOFA_A_TEMPORAL3) {
        if (a->B <= 0 || a->T <= 0 || a->HW <= 0 || a->C <= 0 || (a->C % BK) != 0) {
            set_last_error("mofa_gemm temporal3: needs C %% 64 == 0 (C=%d)", a->C);
            return MOFA_ERR_ARG;
        }
        Ktot = 3LL * a->C;
        p.T = a->T;
        p.HW = a->HW;
        p.n_img = a->B * a->T;
        p.tiles_p = (a->HW + BM - 1) / BM;
        p.m_tiles = a->B * a->T * p.tiles_p;
        p.kb_per_tap = a->C / BK;
        p.num_kb = 3 * p.kb_per_tap;
        p.kb_split = p.num_kb;
        uint64_t dims[4] = {(uint64_t)a->C, (uint64_t)a->HW, (uint64_t)a->T, (uint64_t)a->B};
        uint64_t strides[3] = {(uint64_t)a->C * 2, (uint64_t)a->HW * a->C * 2, (uint64_t)a->T * a->HW * a->C * 2};
        uint32_t box[4] = {BK, BM, 1, 1};
        if ((rc = make_tmap_f16(&tmA, a->a, 4, dims, strides, box)) != MOFA_OK) return rc;
        tmA2 = tmA;
        if (tma_out) {
            uint64_t od[4] = {(uint64_t)n_out_cols, (uint64_t)a->HW, (uint64_t)a->T, (uint64_t)a->B};
            uint64_t os[3] = {ldc_b, (uint64_t)a->HW * ldc_b, (uint64_t)a->T * a->HW * ldc_b};
            uint32_t ob[4] = {64, 32, 1, 1};
            if ((rc = make_tmap_f16(&tmOut, a->out, 4, od, os, ob)) != MOFA_OK) return rc;
        }
    } else {
        set_last_error("mofa_gemm: unknown mode %d", a->mode);
        return MOFA_ERR_ARG;
    }
    if (!tma_out) tmOut = tmA;
    // CTA pairs (cta_group::2, see the kernel): an even number of M tiles so both CTAs of a pair always have a tile, every
    // N tile a multiple of 32 wide so each half is a legal MMA N.  MOFA_GEMM_2CTA=0 switches them off (A/B measurements).
    static int pair_mode = -1;
    if (pair_mode < 0) {
        const char* e = getenv("MOFA_GEMM_2CTA");
        pair_mode = e ? atoi(e) : 1;
    }
    // (measured, level-0 shapes: 3x3 conv 1150 -> 1244 TFLOP/s, temporal conv 841 -> 884, K = 1280 projection 839 -> 904,
    //  q|k|v 848 -> 886; with the GroupNorm-statistics epilogue on a short main loop the lock step of the two CTAs'
    //  epilogues costs more than the pair saves: temporal conv 761 -> 731, so those stay single)
    // An odd number of M tiles leaves the last pair's second CTA an empty tile (all loads out of bounds = zeros, every
    // store clipped / predicated off).
    // 16-epilogue-warp kernels: GEGLU gains as pairs (K = 320: 963 -> 1005 TFLOP/s), the memory-bound plain projections
    // lose 5 % (pair_mode 2 forces them on for measurements).  Level-2 shapes with 225 M tiles: GEGLU K = 1280
    // 1343 -> 1478, K = 5120 projection 1301 -> 1354 (profiles/r2_gemm_pairs_ab.txt).
    const bool k2 = pair_mode != 0 && (!use_ew16 || geglu || pair_mode == 2) && p.m_tiles >= 2 && (bn % 32) == 0 &&
                    (a->N % 32) == 0 && num_sms() >= 2 && (a->max_ctas <= 0 || a->max_ctas >= 2) &&
                    !(p.gn_stats && a->mode != MOFA_A_CONV3X3);
    {
        uint64_t dims[2] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(a->N)};
        uint64_t strides[1] = {static_cast<uint64_t>(Ktot) * 2};
        uint32_t box[2] = {BK, static_cast<uint32_t>(k2 ? bn / 2 : bn)};
        if ((rc = make_tmap_f16(&tmB, a->w, 2, dims, strides, box)) != MOFA_OK) return rc;
    }

    const uint32_t stage_bytes = kABytes + static_cast<uint32_t>(k2 ? bn / 2 : bn) * BK * 2;
    const size_t fixed = kEpiWarps * kStagingBytes + 1024 + 256;
    int stages = static_cast<int>((227u * 1024u - fixed) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) stages = 2;
    p.stages = stages;
    const size_t smem_bytes = static_cast<size_t>(stages) * stage_bytes + kEpiWarps * kStagingBytes +
                              (2 * stages + 4) * 8 + 16 + 1024 + (p.gn_stats ? kGnBufs * kGnSlots * kGnGroups * 8 + 32 : 0);

    if (p.gn_stats) {
        const long long rows_total = a->mode == MOFA_A_LINEAR ? a->M
                                     : a->mode == MOFA_A_CONV3X3 ? static_cast<long long>(a->n_img) * a->H * a->W
                                                                 : static_cast<long long>(a->B) * a->T * a->HW;
        const long long n_stat = (rows_total + p.gn_rows_per_stat - 1) / p.gn_rows_per_stat;
        cudaMemsetAsync(p.gn_stats, 0, sizeof(float) * 2 * p.gn_groups * n_stat, stream);
    }
    const long long total = static_cast<long long>(k2 ? ((p.m_tiles + 1) & ~1) : p.m_tiles) * p.n_tiles;
    int grid = num_sms();
    if (a->max_ctas > 0 && a->max_ctas < grid) grid = a->max_ctas;
    if (total < grid) grid = static_cast<int>(total);
    if (k2) grid &= ~1;   // whole pairs (total is even: m_tiles is)

    using Kern = void (*)(const CUtensorMap, const CUtensorMap, const CUtensorMap, const CUtensorMap, const GemmKernelParams);
    static const Kern kernels[12] = {gemm_tc_kernel<0, false, 8, false>, gemm_tc_kernel<0, true, 8, false>,
                                     gemm_tc_kernel<1, false, 8, false>, gemm_tc_kernel<2, false, 8, false>,
                                     gemm_tc_kernel<1, false, 16, false>, gemm_tc_kernel<0, false, 16, false>,
                                     gemm_tc_kernel<0, false, 8, true>, gemm_tc_kernel<0, true, 8, true>,
                                     gemm_tc_kernel<1, false, 8, true>, gemm_tc_kernel<2, false, 8, true>,
                                     gemm_tc_kernel<1, false, 16, true>, gemm_tc_kernel<0, false, 16, true>};
    static bool configured = false;
    if (!configured) {
        for (Kern k : kernels) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
            if (e != cudaSuccess) {
                set_last_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e));
                return MOFA_ERR_CUDA;
            }
        }
        configured = true;
    }
    if (p.gn_stats && a->act != 0) {
        set_last_error("mofa_gemm: gn_stats is implemented for the plain epilogue (act 0), got act %d", a->act);
        return MOFA_ERR_ARG;
    }
    // 16 epilogue warps where the epilogue is the bottleneck (5 k-blocks per tile at K = 320: GEGLU 690 -> 916 TFLOP/s); with
    // longer main loops the 8-warp unrolled body is faster (GEGLU K = 640: 1180 vs 1092, K = 1280: 1410 vs 1333)
    const bool geglu16 = use_ew16;
    const int base = geglu ? 2 : (a->act != 0 ? 3 : (p.gn_stats ? 1 : 0));
    const Kern kern = use_ew16 ? (geglu ? kernels[k2 ? 10 : 4] : kernels[k2 ? 11 : 5]) : kernels[base + (k2 ? 6 : 0)];
    if (k2) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cfg.gridDim = dim3(grid);
        cfg.blockDim = dim3(geglu16 ? 64 + 32 * 16 : kGemmThreads);
        cfg.dynamicSmemBytes = smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tmA, tmA2, tmB, tmOut, p);
        if (e != cudaSuccess) {
            set_last_error("mofa_gemm: cluster launch: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
    } else {
        kern<<<grid, geglu16 ? 64 + 32 * 16 : kGemmThreads, smem_bytes, stream>>>(tmA, tmA2, tmB, tmOut, p);
    }
    return check_launch("mofa_gemm");
}
