// Spatial self-attention, version 2: two 128-query tiles per CTA ("ping-pong" softmax warpgroups).
//
// Why: with head_dim 64 the kernel is bound by the exponential (MUFU) and by instruction issue of the
// softmax warps, not by the tensor core (per 128x128 tile: 512 MMA cycles vs >= 1024 MUFU cycles per
// SMSP-warp).  Version 1 ran ONE softmax warp per SM sub-partition, so every tcgen05.ld / MUFU latency
// was exposed (measured 315 TFLOP/s at L = 9216).  Here each SM sub-partition hosts two softmax warps
// (one per query tile) that interleave, K/V tiles are loaded once for 256 queries, and the instruction
// mix per score is trimmed: masking only on the ragged last tile, 3-input max, row sums of the unrounded
// probabilities, lazy rescaling (the running maximum is only moved when it grows by more than 2^8, so the
// 64-wide O rescale is skipped on almost every tile; P <= 256 stays well inside fp16).
//
//   warp 0 (1 lane) : TMA producer   Q0,Q1 once; K/V tiles through a 3-stage ring
//   warp 1 (1 lane) : MMA issuer     S_w = Q_w K^T (TMEM, one buffer per query tile), O_w = P_w V
//   warps 2..5      : softmax warpgroup 0 (query tile 0), one row per thread
//   warps 6..9      : softmax warpgroup 1 (query tile 1)
// Issue order per KV tile j:  [P0_j ready] PV0_j, S0_{j+1};  [P1_j ready] PV1_j, S1_{j+1}  -- so while
// warpgroup 0 works on S0_{j+1} the tensor core runs PV1_j / S1_{j+1} and vice versa.
#include "../../include/mofa_b200.h"
#include <stdlib.h>

#include "common.cuh"

namespace mofa {

int make_tmap_f16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
int attn_spatial_v1(const void* qkv, void* out, int32_t frames, int32_t L, int32_t heads, float scale,
                    mofa_stream_t stream_);

namespace v2 {

constexpr int kThreads = 320;  // 2 role warps + 2 x 4 softmax warps: 204 registers per thread available
constexpr int kKVStages = 3;
constexpr uint32_t kTile = 128 * 64 * 2;  // 16 KB
constexpr uint32_t kPBytes = 2 * kTile;   // one 128 x 128 fp16 P tile

struct Params {
    __half* out;
    int L, C, heads, n_kv;
    float scale_log2;
};

MOFA_DEVICE float max3(float a, float b, float c) {
    float d;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
    return d;
}

template <bool kMasked>
MOFA_DEVICE float row_max(uint32_t ts, int kv_valid) {
    float mx = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t s[32];
        tmem_ld_32x32(ts + c * 32, s);
        tmem_ld_wait();
        if (kMasked) {
#pragma unroll
            for (int i = 0; i < 32; ++i)
                if (c * 32 + i >= kv_valid) s[i] = 0xff800000u;  // -inf
        }
#pragma unroll
        for (int i = 0; i < 32; i += 2) mx = max3(mx, __uint_as_float(s[i]), __uint_as_float(s[i + 1]));
    }
    return mx;
}

// p = exp2(s * sl2 - m) -> fp16 -> swizzled K-major P tile; returns the row sum of p
template <bool kMasked>
MOFA_DEVICE float exp_store(uint32_t ts, uint8_t* prow, int r, float sl2, float m, int kv_valid) {
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
        uint32_t s[32];
        tmem_ld_32x32(ts + c * 32, s);
        tmem_ld_wait();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
            uint32_t packed[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                float p0 = fast_exp2(fmaf(__uint_as_float(s[g * 8 + 2 * i]), sl2, -m));
                float p1 = fast_exp2(fmaf(__uint_as_float(s[g * 8 + 2 * i + 1]), sl2, -m));
                if (kMasked) {
                    const int col = c * 32 + g * 8 + 2 * i;
                    if (col >= kv_valid) p0 = 0.f;
                    if (col + 1 >= kv_valid) p1 = 0.f;
                }
                sum0 += p0;
                sum1 += p1;
                const __half2 h = __floats2half2_rn(p0, p1);
                packed[i] = *reinterpret_cast<const uint32_t*>(&h);
            }
            const int col0 = c * 32 + g * 8;
            uint8_t* dst = prow + (col0 >> 6) * kTile + ((((col0 & 63) >> 3) ^ (r & 7)) << 4);
            *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
        }
    }
    return sum0 + sum1;
}

__global__ void __launch_bounds__(kThreads, 1)
attn_spatial2_kernel(const __grid_constant__ CUtensorMap tmQKV, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;                              // 2 x 16 KB
    uint8_t* sKV = sQ + 2 * kTile;                   // kKVStages x (K | V)
    uint8_t* sP = sKV + kKVStages * 2 * kTile;       // 2 x 32 KB (one per query tile)
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + kKVStages;
    uint64_t* s_full = kv_empty + kKVStages;  // [2] per query tile
    uint64_t* p_full = s_full + 2;            // [2]
    uint64_t* o_full = p_full + 2;            // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 256;
    const int head = blockIdx.y;
    const int frame = blockIdx.z;
    const int n_kv = p.n_kv;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmQKV);
        mbar_init(q_full, 1);
        for (int i = 0; i < kKVStages; ++i) {
            mbar_init(&kv_full[i], 1);
            mbar_init(&kv_empty[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&o_full[i], 1);
        }
        fence_barrier_init();
    }
    if (warp == 1) tmem_alloc(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base;        // 2 x 128 columns
    const uint32_t tmem_O = tmem_base + 256;  // 2 x 64 columns

    if (threadIdx.x == 0) {
        // ===================== TMA producer =====================
        mbar_arrive_expect_tx(q_full, 2 * kTile);
        tma_load_3d(&tmQKV, q_full, sQ, head * 64, q0, frame);
        tma_load_3d(&tmQKV, q_full, sQ + kTile, head * 64, q0 + 128, frame);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1);
            uint8_t* sk = sKV + stage * 2 * kTile;
            mbar_arrive_expect_tx(&kv_full[stage], 2 * kTile);
            tma_load_3d(&tmQKV, &kv_full[stage], sk, p.C + head * 64, j * 128, frame);
            tma_load_3d(&tmQKV, &kv_full[stage], sk + kTile, 2 * p.C + head * 64, j * 128, frame);
            if (++stage == kKVStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (threadIdx.x == 32) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_f16(128, false);
        const uint32_t idesc_o = umma_idesc_f16(64, true);
        uint64_t dq[2];
        dq[0] = umma_desc_sw128_kmajor(smem_u32(sQ));
        dq[1] = umma_desc_sw128_kmajor(smem_u32(sQ + kTile));
        mbar_wait(q_full, 0);
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        {
            const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(sKV));
#pragma unroll
            for (int w = 0; w < 2; ++w) {
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S + w * 128, dq[w] + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[w]);
            }
        }
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
            int nstage = stage + 1;
            uint32_t nphase = phase;
            if (nstage == kKVStages) {
                nstage = 0;
                nphase ^= 1;
            }
            const bool have_next = j + 1 < n_kv;
            const uint32_t va = smem_u32(sKV + stage * 2 * kTile + kTile);
            const uint64_t dv = umma_desc_sw128_mnmajor(va, kTile);
            uint64_t dk_next = 0;
#pragma unroll
            for (int w = 0; w < 2; ++w) {
                mbar_wait(&p_full[w], j & 1);
                tc_fence_after();
                const uint32_t pa = smem_u32(sP + w * kPBytes);
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t dp = umma_desc_sw128_kmajor(pa + (kk >> 2) * kTile) + 2 * (kk & 3);
                    umma_f16_ss(tmem_O + w * 64, dp, dv + 128 * kk, idesc_o, kk != 0);
                }
                umma_commit(&o_full[w]);
                if (have_next) {
                    if (w == 0) {
                        mbar_wait(&kv_full[nstage], nphase);
                        tc_fence_after();
                        dk_next = umma_desc_sw128_kmajor(smem_u32(sKV + nstage * 2 * kTile));
                    }
#pragma unroll
                    for (int k = 0; k < 4; ++k)
                        umma_f16_ss(tmem_S + w * 128, dq[w] + 2 * k, dk_next + 2 * k, idesc_s, k != 0);
                    umma_commit(&s_full[w]);
                }
            }
            umma_commit(&kv_empty[stage]);
            stage = nstage;
            phase = nphase;
        }
    } else if (warp >= 2) {
        // ===================== softmax warpgroups =====================
        const int w = (warp - 2) >> 2;  // query tile of this warpgroup
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        const uint32_t ts = tmem_S + lane_addr + w * 128;
        const uint32_t to = tmem_O + lane_addr + w * 64;
        uint8_t* prow = sP + w * kPBytes + r * 128;
        const float sl2 = p.scale_log2;
        float m = -INFINITY, l = 0.f;
        float o[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = 0.f;

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[w], j & 1);
            tc_fence_after();
            const int kv_valid = p.L - j * 128;
            const bool masked = kv_valid < 128;  // warp-uniform
            const float mx = (masked ? row_max<true>(ts, kv_valid) : row_max<false>(ts, kv_valid)) * sl2;
            // lazy rescale: move the reference maximum only when it grows by more than 2^8
            float alpha = 1.0f;
            if (mx > m + 8.0f) {
                alpha = fast_exp2(m - mx);  // first tile: m = -inf -> 0
                m = mx;
            }
            // fold in the previous tile's PV partial (it is relative to the old m), then rescale
            if (j > 0) {
                mbar_wait(&o_full[w], (j - 1) & 1);
                tc_fence_after();
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(to + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(v[i]);
                }
            }
            if (alpha != 1.0f) {
#pragma unroll
                for (int i = 0; i < 64; ++i) o[i] *= alpha;
                l *= alpha;
            }
            l += masked ? exp_store<true>(ts, prow, r, sl2, m, kv_valid) : exp_store<false>(ts, prow, r, sl2, m, kv_valid);
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[w]);
        }
        {
            const int j = n_kv - 1;
            mbar_wait(&o_full[w], j & 1);
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(to + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(v[i]);
            }
        }
        const int qrow = q0 + w * 128 + r;
        if (qrow < p.L) {
            const float inv = 1.0f / l;
            __half* dst = p.out + (static_cast<long long>(frame) * p.L + qrow) * p.C + head * 64;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint32_t wv[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 h = __floats2half2_rn(o[g * 8 + 2 * i] * inv, o[g * 8 + 2 * i + 1] * inv);
                    wv[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace v2
}  // namespace mofa

using namespace mofa;

extern "C" int mofa_attn_spatial(const void* qkv, void* out, int32_t frames, int32_t L, int32_t heads, float scale,
                                 mofa_stream_t stream_) {
    static int use_v1 = -1;
    if (use_v1 < 0) {
        const char* e = getenv("MOFA_ATTN_V1");
        use_v1 = (e && e[0] == '1') ? 1 : 0;
    }
    if (use_v1) return attn_spatial_v1(qkv, out, frames, L, heads, scale, stream_);
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qkv || !out || frames <= 0 || L <= 0 || heads <= 0) {
        set_last_error("mofa_attn_spatial: bad arguments");
        return MOFA_ERR_ARG;
    }
    const int C = heads * 64;
    CUtensorMap tm;
    uint64_t dims[3] = {static_cast<uint64_t>(3 * C), static_cast<uint64_t>(L), static_cast<uint64_t>(frames)};
    uint64_t strides[2] = {static_cast<uint64_t>(3 * C) * 2, static_cast<uint64_t>(L) * 3 * C * 2};
    uint32_t box[3] = {64, 128, 1};
    int rc = make_tmap_f16(&tm, qkv, 3, dims, strides, box);
    if (rc) return rc;
    v2::Params p;
    p.out = static_cast<__half*>(out);
    p.L = L;
    p.C = C;
    p.heads = heads;
    p.n_kv = (L + 127) / 128;
    p.scale_log2 = scale * 1.4426950408889634f;
    const size_t smem_bytes = 2 * v2::kTile + v2::kKVStages * 2 * v2::kTile + 2 * v2::kPBytes + 16 * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(v2::attn_spatial2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_bytes));
        if (e != cudaSuccess) {
            set_last_error("mofa_attn_spatial: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid((L + 255) / 256, heads, frames);
    v2::attn_spatial2_kernel<<<grid, v2::kThreads, smem_bytes, stream>>>(tm, p);
    return check_launch("mofa_attn_spatial");
}
