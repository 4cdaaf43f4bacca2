// Fused FeedForward (GEGLU -> Linear) for one 128-row tile at a time: the 4C-wide hidden activation never leaves the SM.
//
//   out = alpha * ( GEGLU(x W1^T + b1) W2^T + b2 ) + beta1 * res1 + beta2 * res2          x: [M, C], hidden = 4C
//
// Replaces diffusers 0.24 `FeedForward(activation_fn="geglu")` inside BasicTransformerBlock / TemporalBasicTransformerBlock
// (blocks constructed at /root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py:169-233 and
// models/controlnet_sdv.py:270-309) where C <= 320: the level-0 blocks of SVD (C = 320, hidden 1280, 21 per denoise step).
// Unfused, that block is two GEMMs at the HBM ridge: the first writes a [460800, 1280] fp16 intermediate (1.18 GB) that the
// second immediately re-reads -- 50 GB of DRAM traffic per step that carries no information across kernels.
//
// Per CTA (persistent over M tiles), hidden processed in chunks of 64:
//   S_j   = X W1_j^T            tcgen05.mma  M=128, N=128 (64 value | 64 gate rows of the packed W1), K = C, accumulator in TMEM
//   H_j   = (S_val + b) * GELU(S_gate + b)      16 epilogue warps: S read once into registers (buffer released at once, so
//                                               S_{j+1} is computed during the GELUs), H_j written back to TENSOR MEMORY as
//                                               packed fp16 pairs
//   OUT  += H_j W2_j^T          tcgen05.mma with A = H_j FROM TENSOR MEMORY (as the attention kernel does with P), B = the
//                               [C, 64] slab of W2 in shared memory, accumulator OUT = C <= 320 TMEM columns for the whole tile
// TMEM: OUT 320 | S 128 | H 2 x 32 columns = 512.   Shared memory: X tile C/64 x 16 KB (resident for the tile), W1 ring
// 4 x 16 KB, W2 slab C x 128 B, 16 x 2 KB store staging  (216 KB at C = 320).
//   warp 0 : TMA producer  X tile + W1 k-blocks          warp 2 : TMA producer  W2 slabs (its own thread: it blocks on a
//   warp 1 : MMA issuer    S_{j} then OUT += H_{j-1}              different consumer than the W1 stream)
//   warps 3..18 : epilogue; warp (quarter q, sub s) owns hidden columns [16 s, 16 s + 16) of every chunk and, at the end of a
//                 tile, the 32-column output units s, s + 4, ... (bias, alpha, two residuals, SWIZZLE_64B staging, TMA store)
#include "../../include/mofa_b200.h"
#include <cstdlib>
#include "common.cuh"

namespace mofa {

int make_tmap_f16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
int make_tmap_f16_sw(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                     const uint32_t* box, CUtensorMapSwizzle swizzle);
int num_sms();

namespace ff {

constexpr int kW1Stages = 5;      // one full chunk of W1 k-blocks (C = 320): the next chunk prefetches while this one computes
constexpr int kStgBytes = 1024;   // output staging per epilogue warp: 32 rows x 16 columns
constexpr int kEW = 16;
constexpr int kThreads = 96 + 32 * kEW;
constexpr uint32_t kTile16K = 128 * 64 * 2;

struct Params {
    int C, hidden, m_tiles, n1, n2;
    long long M;
    const __half* b1;   // packed like W1: per 128 rows [64 value | 64 gate]
    const __half* b2;
    const __half* res1;
    long long ldr1;
    const __half* res2;
    long long ldr2;
    float alpha, beta1, beta2;
    int w1_stages;      // W1 ring depth actually used (<= kW1Stages)
    int debug;          // experiment mask (env MOFA_FF_DEBUG): 1 skip S MMAs, 2 skip OUT MMAs, 4 skip the GEGLU math, 8 / 16 skip the W1 / W2 loads
};

__device__ long long ff_dbg[64 * 16];
#define FF_STAMP(slot, gidx)                                                            \
    if ((p.debug & 2048) && blockIdx.x == 0 && (gidx) < 64u) ff_dbg[(gidx) * 16 + (slot)] = clock64();

// experiment helper: suspend-hinted try_wait (default) or a pure test_wait spin
MOFA_DEVICE void mma_wait(bool spin, uint64_t* bar, uint32_t parity) {
    if (spin) {
        while (!mbar_test(bar, parity)) {
        }
    } else {
        mbar_wait(bar, parity);
    }
}

union H8 {
    uint4 u;
    __half2 h2[4];
    __half h[8];
};

__global__ void __launch_bounds__(kThreads, 1)
ff_geglu_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmW1,
                const __grid_constant__ CUtensorMap tmW2, const __grid_constant__ CUtensorMap tmOut, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KB = p.C >> 6;                 // k-blocks of the first GEMM = 64-row pieces of the W2 slab
    const int nch = p.hidden >> 6;           // hidden chunks
    uint8_t* sX = smem;
    uint8_t* sW1 = sX + static_cast<size_t>(KB) * kTile16K;
    uint8_t* sW2 = sW1 + kW1Stages * kTile16K;
    uint8_t* staging = sW2 + static_cast<size_t>(p.C) * 128;
    uint64_t* bars = reinterpret_cast<uint64_t*>(staging + kEW * kStgBytes);
    uint64_t* x_full = bars;
    uint64_t* x_empty = bars + 1;
    uint64_t* w1_full = bars + 2;
    uint64_t* w1_empty = w1_full + kW1Stages;
    uint64_t* w2_full = w1_empty + kW1Stages;
    uint64_t* w2_empty = w2_full + 1;
    uint64_t* s_full = w2_empty + 1;
    uint64_t* s_free = s_full + 1;
    uint64_t* h_full = s_free + 1;    // [2]
    uint64_t* h_free = h_full + 2;    // [2]
    uint64_t* out_full = h_free + 2;
    uint64_t* out_free = out_full + 1;
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(out_free + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmX);
        tma_prefetch_desc(&tmW1);
        tma_prefetch_desc(&tmW2);
        tma_prefetch_desc(&tmOut);
        mbar_init(x_full, 1);
        mbar_init(x_empty, 1);
        for (int i = 0; i < kW1Stages; ++i) {
            mbar_init(&w1_full[i], 1);
            mbar_init(&w1_empty[i], 1);
        }
        mbar_init(w2_full, 1);
        mbar_init(w2_empty, 1);
        mbar_init(s_full, 1);
        mbar_init(s_free, kEW);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&h_full[i], kEW);
            mbar_init(&h_free[i], 1);
        }
        mbar_init(out_full, 1);
        mbar_init(out_free, kEW);
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_OUT = tmem_base;          // C columns
    const uint32_t tmem_S = tmem_base + 320;      // 128 columns: 64 value | 64 gate
    const uint32_t tmem_H = tmem_base + 448;      // 2 x 32 columns of packed fp16 pairs

    if (threadIdx.x == 0) {
        // ===================== TMA producer: X tile, W1 k-blocks =====================
        int st = 0;
        uint32_t ph = 0, t = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++t) {
            mbar_wait(x_empty, (t & 1u) ^ 1u);     // the previous tile's last S MMA has read X
            mbar_arrive_expect_tx(x_full, static_cast<uint32_t>(KB) * kTile16K);
            for (int kb = 0; kb < KB; ++kb) tma_load_2d(&tmX, x_full, sX + kb * kTile16K, kb * 64, tile * 128);
            for (int j = 0; j < nch; ++j) {
                for (int kb = 0; kb < KB; ++kb) {
                    mbar_wait(&w1_empty[st], ph ^ 1u);
                    if (p.debug & 8) {
                        mbar_arrive(&w1_full[st]);
                    } else {
                        mbar_arrive_expect_tx(&w1_full[st], kTile16K);
                        tma_load_2d(&tmW1, &w1_full[st], sW1 + st * kTile16K, kb * 64, j * 128);
                    }
                    if (++st == p.w1_stages) {
                        st = 0;
                        ph ^= 1u;
                    }
                }
            }
        }
    } else if (threadIdx.x == 64) {
        // ===================== TMA producer: W2 slabs =====================
        uint32_t g = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x) {
            for (int j = 0; j < nch; ++j, ++g) {
                mbar_wait(w2_empty, (g & 1u) ^ 1u);   // OUT += H_{g-1} W2_{g-1}^T has completed
                if (p.debug & 16) {
                    mbar_arrive(w2_full);
                    continue;
                }
                mbar_arrive_expect_tx(w2_full, static_cast<uint32_t>(p.C) * 128u);
                for (int pc = 0; pc < KB; ++pc) tma_load_2d(&tmW2, w2_full, sW2 + pc * 8192, j * 64, pc * 64);
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The whole warp walks the (warp-uniform) schedule and one elected lane issues: descriptors, barrier addresses
        // and loop state then live in uniform registers.  With the loop inside a single-thread branch the compiler
        // re-elects a lane and broadcasts every operand before each tcgen05 instruction (~16 instructions per MMA,
        // ~4.2k issue cycles per 64-wide chunk against 1.9k cycles of tensor work: profiles/r2_ff_fused_notes.txt).
        const bool leader = elect_one();
        const bool spin_m = (p.debug & 32) != 0;
        const uint32_t idesc_s = umma_idesc_f16(128, false);
        const uint32_t idesc_o1 = umma_idesc_f16(static_cast<uint32_t>(p.n1), false);
        const uint32_t idesc_o2 = umma_idesc_f16(static_cast<uint32_t>(p.n2 > 0 ? p.n2 : 16), false);
        const uint64_t dX0 = umma_desc_sw128_kmajor(smem_u32(sX));
        const uint64_t dW10 = umma_desc_sw128_kmajor(smem_u32(sW1));
        const uint64_t dw = umma_desc_sw128_kmajor(smem_u32(sW2));
        const uint64_t dw2 = umma_desc_sw128_kmajor(smem_u32(sW2 + p.n1 * 128));
        const bool two = p.n2 > 0;
        const bool skip1 = (p.debug & 1) != 0, skip2 = (p.debug & 2) != 0;
        int st = 0;
        uint32_t ph = 0, t = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++t) {
            mma_wait(spin_m, x_full, t & 1u);
            tc_fence_after();
            for (int j = 0; j <= nch; ++j) {
                if (j < nch) {
                    const uint32_t g = t * static_cast<uint32_t>(nch) + static_cast<uint32_t>(j);
                    if (g > 0) {   // the epilogue warps hold S_{g-1} in registers
                        mma_wait(spin_m, s_free, (g - 1u) & 1u);
                        tc_fence_after();
                    }
                    if (leader) { FF_STAMP(0, g) }
                    for (int kb = 0; kb < KB; ++kb) {
                        mma_wait(spin_m, &w1_full[st], ph);
                        tc_fence_after();
                        if (leader) {
                            const uint64_t da = dX0 + static_cast<uint64_t>(kb) * (kTile16K >> 4);
                            const uint64_t db = dW10 + static_cast<uint64_t>(st) * (kTile16K >> 4);
                            if (!skip1) {
                                umma_f16_ss(tmem_S, da, db, idesc_s, kb != 0 ? 1u : 0u);
                                umma_f16_ss(tmem_S, da + 2, db + 2, idesc_s, 1u);
                                umma_f16_ss(tmem_S, da + 4, db + 4, idesc_s, 1u);
                                umma_f16_ss(tmem_S, da + 6, db + 6, idesc_s, 1u);
                            }
                            if (p.debug & 128) mbar_arrive(&w1_empty[st]); else umma_commit(&w1_empty[st]);
                        }
                        if (++st == p.w1_stages) {
                            st = 0;
                            ph ^= 1u;
                        }
                    }
                    if (leader) {
                        umma_commit(s_full);
                        if (j == nch - 1) umma_commit(x_empty);
                        FF_STAMP(1, g)
                    }
                }
                if (j > 0) {
                    const int jj = j - 1;
                    const uint32_t gg = t * static_cast<uint32_t>(nch) + static_cast<uint32_t>(jj);
                    const uint32_t buf = gg & 1u;
                    mma_wait(spin_m, &h_full[buf], (gg >> 1) & 1u);
                    mma_wait(spin_m, w2_full, gg & 1u);
                    if (jj == 0 && t > 0) mma_wait(spin_m, out_free, (t - 1u) & 1u);   // the previous tile's output was drained
                    tc_fence_after();
                    if (leader) {
                        FF_STAMP(2, gg)
                        const uint32_t ha = tmem_H + buf * 32;
                        if (!skip2) {
                            umma_f16_ts(tmem_OUT, ha, dw, idesc_o1, jj > 0 ? 1u : 0u);
                            if (two) umma_f16_ts(tmem_OUT + p.n1, ha, dw2, idesc_o2, jj > 0 ? 1u : 0u);
#pragma unroll
                            for (int kk = 1; kk < 4; ++kk) {
                                umma_f16_ts(tmem_OUT, ha + kk * 8, dw + 2 * kk, idesc_o1, 1u);
                                if (two) umma_f16_ts(tmem_OUT + p.n1, ha + kk * 8, dw2 + 2 * kk, idesc_o2, 1u);
                            }
                        }
                        umma_commit(w2_empty);
                        umma_commit(&h_free[buf]);
                        FF_STAMP(3, gg)
                        if (jj == nch - 1) umma_commit(out_full);
                    }
                }
            }
        }
        __syncwarp();
    } else if (warp >= 3) {
        // ===================== epilogue =====================
        const int e = warp - 3;
        const int q = warp & 3;        // TMEM lane quarter this warp may touch
        const int sub = e >> 2;        // which 16 hidden columns of a chunk / which output units
        const int r = q * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(q * 32) << 16;
        uint8_t* stg = staging + e * kStgBytes;
        uint32_t t = 0;
        for (int tile = blockIdx.x; tile < p.m_tiles; tile += gridDim.x, ++t) {
            const long long row = static_cast<long long>(tile) * 128 + r;
            const bool valid = row < p.M;
            for (int j = 0; j < nch; ++j) {
                const uint32_t g = t * static_cast<uint32_t>(nch) + static_cast<uint32_t>(j);
                const uint32_t buf = g & 1u;
                // bias of this warp's 16 hidden columns (same for every row: broadcast loads, L1 hits)
                uint4 bvv[2], bgv[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    bvv[i] = make_uint4(0, 0, 0, 0);
                    bgv[i] = make_uint4(0, 0, 0, 0);
                    if (p.b1 && !(p.debug & 1024)) {
                        bvv[i] = __ldg(reinterpret_cast<const uint4*>(p.b1 + j * 128 + sub * 16) + i);
                        bgv[i] = __ldg(reinterpret_cast<const uint4*>(p.b1 + j * 128 + 64 + sub * 16) + i);
                    }
                }
                mma_wait((p.debug & 64) != 0, s_full, g & 1u);
                tc_fence_after();
                if (threadIdx.x == 96) { FF_STAMP(4, g) }
                if (threadIdx.x == 576) { FF_STAMP(9, g) }
                uint32_t va[16], ga[16];
                if (p.debug & 512) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) va[i] = ga[i] = static_cast<uint32_t>(i + lane);
                } else {
                    tmem_ld_32x16(tmem_S + lane_addr + sub * 16, va);
                    tmem_ld_32x16(tmem_S + lane_addr + 64 + sub * 16, ga);
                    tmem_ld_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (threadIdx.x == 96) { FF_STAMP(5, g) }
                if (threadIdx.x == 576) { FF_STAMP(10, g) }
                if (lane == 0) mbar_arrive(s_free);        // S_{g+1} may overwrite the buffer
                if ((p.debug & 4096) && q == 1) __nanosleep(200);   // experiment: leave the MMA warp's scheduler alone
                if (p.debug & 8192) __nanosleep(100);
                uint32_t packed[8];
                if (p.debug & 4) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) packed[i] = va[i] ^ ga[i];
                } else
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    H8 bv, bg;
                    bv.u = bvv[i];
                    bg.u = bgv[i];
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        const int k0 = i * 8 + 2 * c;
                        const float y0 = (__uint_as_float(va[k0]) + __half2float(bv.h[2 * c])) *
                                         gelu_erf_relu_form(__uint_as_float(ga[k0]) + __half2float(bg.h[2 * c]));
                        const float y1 = (__uint_as_float(va[k0 + 1]) + __half2float(bv.h[2 * c + 1])) *
                                         gelu_erf_relu_form(__uint_as_float(ga[k0 + 1]) + __half2float(bg.h[2 * c + 1]));
                        const __half2 hh = __floats2half2_rn(y0, y1);
                        packed[i * 4 + c] = *reinterpret_cast<const uint32_t*>(&hh);
                    }
                }
                if (threadIdx.x == 96) { FF_STAMP(6, g) }
                if (g >= 2) {   // OUT += H_{g-2} W2^T has finished reading this H buffer
                    mbar_wait(&h_free[buf], ((g >> 1) - 1u) & 1u);
                    tc_fence_after();
                }
                if (!(p.debug & 256)) {
                    tmem_st_32x8(tmem_H + lane_addr + buf * 32 + sub * 8, packed);
                    tmem_st_wait();
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&h_full[buf]);
                if (threadIdx.x == 96) { FF_STAMP(7, g) }
                if (threadIdx.x == 576) { FF_STAMP(11, g) }
            }
            // ---- output of the tile: bias, alpha, residuals, fp16, TMA store
            mbar_wait(out_full, t & 1u);
            tc_fence_after();
            const int n_units = p.C >> 4;      // 16-column units: 1 KB of staging per warp (the W1 ring needs the rest)
            const bool has1 = p.res1 != nullptr, has2 = p.res2 != nullptr;
            uint4 r1n[2], r2n[2];              // residuals of the NEXT unit: their latency hides behind this unit's work
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                r1n[i] = r2n[i] = make_uint4(0, 0, 0, 0);
                if (has1 && valid && sub < n_units)
                    r1n[i] = __ldg(reinterpret_cast<const uint4*>(p.res1 + row * p.ldr1 + sub * 16) + i);
                if (has2 && valid && sub < n_units)
                    r2n[i] = __ldg(reinterpret_cast<const uint4*>(p.res2 + row * p.ldr2 + sub * 16) + i);
            }
#pragma unroll 1
            for (int u = sub; u < n_units; u += 4) {
                const int n0 = u * 16;
                uint4 r1v[2], r2v[2], bvv[2];
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    r1v[i] = r1n[i];
                    r2v[i] = r2n[i];
                    bvv[i] = make_uint4(0, 0, 0, 0);
                    if (p.b2) bvv[i] = __ldg(reinterpret_cast<const uint4*>(p.b2 + n0) + i);
                    if (u + 4 < n_units) {
                        if (has1 && valid) r1n[i] = __ldg(reinterpret_cast<const uint4*>(p.res1 + row * p.ldr1 + n0 + 64) + i);
                        if (has2 && valid) r2n[i] = __ldg(reinterpret_cast<const uint4*>(p.res2 + row * p.ldr2 + n0 + 64) + i);
                    }
                }
                uint32_t acc[16];
                tmem_ld_32x16(tmem_OUT + lane_addr + n0, acc);
                tmem_ld_wait();
                if (lane == 0) tma_store_wait_read<0>();   // the previous unit's store has read the staging rows
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    H8 b, a1, a2, o;
                    b.u = bvv[i];
                    a1.u = r1v[i];
                    a2.u = r2v[i];
#pragma unroll
                    for (int c = 0; c < 8; ++c) {
                        float x = (__uint_as_float(acc[i * 8 + c]) + __half2float(b.h[c])) * p.alpha;
                        x = fmaf(p.beta1, __half2float(a1.h[c]), x);
                        x = fmaf(p.beta2, __half2float(a2.h[c]), x);
                        o.h[c] = __float2half_rn(x);
                    }
                    // 16-byte chunk i of the 32-byte staging row (SWIZZLE_32B: address bit 4 ^= bit 7)
                    *reinterpret_cast<uint4*>(stg + lane * 32 + ((i ^ ((lane >> 2) & 1)) << 4)) = o.u;
                }
                fence_proxy_async_smem();
                __syncwarp();
                if (lane == 0) {
                    tma_store_2d(&tmOut, stg, n0, tile * 128 + q * 32);
                    tma_store_commit();
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(out_free);
        }
        if (lane == 0) tma_store_wait_all<0>();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace ff
}  // namespace mofa

using namespace mofa;

extern "C" int mofa_ff_geglu(const void* x, const void* w1_packed, const void* b1_packed, const void* w2, const void* b2,
                             void* out, int64_t M, int32_t C, int32_t hidden, const void* res1, int64_t ldr1,
                             const void* res2, int64_t ldr2, float alpha, float beta1, float beta2,
                             mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !w1_packed || !w2 || !out || M <= 0 || C < 64 || C > 320 || (C % 64) != 0 || hidden < 64 ||
        (hidden % 64) != 0 || (res1 && (ldr1 % 8)) || (res2 && (ldr2 % 8))) {
        set_last_error("mofa_ff_geglu: needs C in {64..320 step 64}, hidden %% 64 == 0, 16-byte rows (C=%d hidden=%d)", C,
                       hidden);
        return MOFA_ERR_ARG;
    }
    ff::Params p;
    memset(&p, 0, sizeof(p));
    p.C = C;
    p.hidden = hidden;
    p.M = M;
    p.m_tiles = static_cast<int>((M + 127) / 128);
    p.n1 = C <= 256 ? C : 192;
    p.n2 = C - p.n1;
    p.b1 = static_cast<const __half*>(b1_packed);
    p.b2 = static_cast<const __half*>(b2);
    p.res1 = static_cast<const __half*>(res1);
    p.ldr1 = ldr1;
    p.res2 = static_cast<const __half*>(res2);
    p.ldr2 = ldr2;
    p.alpha = alpha;
    {
        const char* dbg = getenv("MOFA_FF_DEBUG");
        p.debug = dbg ? atoi(dbg) : 0;
        const char* stg = getenv("MOFA_FF_STAGES");
        p.w1_stages = stg ? atoi(stg) : ff::kW1Stages;
        if (p.w1_stages < 2 || p.w1_stages > ff::kW1Stages) p.w1_stages = ff::kW1Stages;
    }
    p.beta1 = beta1;
    p.beta2 = beta2;

    CUtensorMap tmX, tmW1, tmW2, tmOut;
    int rc;
    {
        uint64_t d[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(M)};
        uint64_t s[1] = {static_cast<uint64_t>(C) * 2};
        uint32_t b[2] = {64, 128};
        if ((rc = make_tmap_f16(&tmX, x, 2, d, s, b)) != MOFA_OK) return rc;
        uint32_t ob[2] = {16, 32};
        if ((rc = make_tmap_f16_sw(&tmOut, out, 2, d, s, ob, CU_TENSOR_MAP_SWIZZLE_32B)) != MOFA_OK) return rc;
    }
    {
        uint64_t d[2] = {static_cast<uint64_t>(C), static_cast<uint64_t>(2) * hidden};
        uint64_t s[1] = {static_cast<uint64_t>(C) * 2};
        uint32_t b[2] = {64, 128};
        if ((rc = make_tmap_f16(&tmW1, w1_packed, 2, d, s, b)) != MOFA_OK) return rc;
    }
    {
        uint64_t d[2] = {static_cast<uint64_t>(hidden), static_cast<uint64_t>(C)};
        uint64_t s[1] = {static_cast<uint64_t>(hidden) * 2};
        uint32_t b[2] = {64, 64};
        if ((rc = make_tmap_f16(&tmW2, w2, 2, d, s, b)) != MOFA_OK) return rc;
    }
    const size_t smem_bytes = static_cast<size_t>(C / 64) * ff::kTile16K + ff::kW1Stages * ff::kTile16K +
                              static_cast<size_t>(C) * 128 + ff::kEW * ff::kStgBytes + 32 * 8 + 16 + 1024;
    static size_t configured = 0;
    if (smem_bytes > configured) {
        cudaError_t e = cudaFuncSetAttribute(ff::ff_geglu_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_bytes));
        if (e != cudaSuccess) {
            set_last_error("mofa_ff_geglu: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        configured = smem_bytes;
    }
    int grid = num_sms();
    if (p.m_tiles < grid) grid = p.m_tiles;
    ff::ff_geglu_kernel<<<grid, ff::kThreads, smem_bytes, stream>>>(tmX, tmW1, tmW2, tmOut, p);
    return check_launch("mofa_ff_geglu");
}

// experiment support (not part of the ABI header): copy the stage timestamps of block 0 (MOFA_FF_DEBUG bit 2048)
extern "C" int mofa_ff_debug_dump(long long* host_out) {
    return cudaMemcpyFromSymbol(host_out, mofa::ff::ff_dbg, sizeof(long long) * 64 * 16) == cudaSuccess ? 0 : 1;
}
