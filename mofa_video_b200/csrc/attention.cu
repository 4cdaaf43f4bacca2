// Self-attention kernels of the SVD transformer blocks (head_dim = 64).
//
// mofa_attn_spatial : FlashAttention-style kernel on tcgen05.  One CTA = one 128-query tile of one
//   (frame, head).  Roles: warp 0 = TMA producer (Q once, K/V tiles through a 3-stage ring),
//   warp 1 = MMA issuer (S = Q K^T into TMEM, then P V from a shared-memory P tile),
//   warps 2..5 = softmax (one query row per thread: TMEM -> exp2 -> fp16 P tile in the UMMA
//   K-major SWIZZLE_128B layout -> running max / sum, O kept in registers, PV partials read back from
//   TMEM one tile late so the PV MMA of tile j overlaps the softmax of tile j+1).
//   S is double-buffered in TMEM (2 x 128 columns), PV partials too (2 x 64 columns).
//   V is consumed in place from the [tokens, 3C] qkv matrix as an MN-major B operand (no transpose).
// mofa_attn_temporal lives in attn_temporal.cu.
//
// Replaces diffusers Attention(AttnProcessor2_0) -> F.scaled_dot_product_attention for attn1 of
// BasicTransformerBlock / TemporalBasicTransformerBlock (blocks created at
// /root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py:169-233).
#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

int make_tmap_f16(CUtensorMap* m, const void* ptr, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);
int attn_spatial_v1(const void* qkv, void* out, int32_t frames, int32_t L, int32_t heads, float scale,
                    mofa_stream_t stream_);

constexpr int kAttnThreads = 192;
constexpr int kKVStages = 3;
constexpr uint32_t kTileBytes = 128 * 64 * 2;  // 16 KB: one 128 x 64 fp16 tile
constexpr uint32_t kPBytes = 2 * kTileBytes;   // 128 x 128 fp16 P tile = two K-major 64-wide atoms

struct AttnParams {
    __half* out;
    int L, C, heads, n_kv;
    float scale_log2;
};

__global__ void __launch_bounds__(kAttnThreads, 1)
attn_spatial_kernel(const __grid_constant__ CUtensorMap tmQKV, const AttnParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + kTileBytes;                      // kKVStages x (K 16 KB | V 16 KB)
    uint8_t* sP = sKV + kKVStages * 2 * kTileBytes;      // 2 x 32 KB
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + 2 * kPBytes);
    uint64_t* q_full = bars;
    uint64_t* kv_full = bars + 1;
    uint64_t* kv_empty = kv_full + kKVStages;
    uint64_t* s_full = kv_empty + kKVStages;  // [2]
    uint64_t* p_full = s_full + 2;            // [2]
    uint64_t* o_full = p_full + 2;            // [2]
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(o_full + 2);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * 128;
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
    if (warp == 2) tmem_alloc(tmem_ptr_smem, 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_ptr_smem;
    const uint32_t tmem_S = tmem_base;         // 2 x 128 columns
    const uint32_t tmem_O = tmem_base + 256;   // 2 x 64 columns

    if (threadIdx.x == 0) {
        // ===================== TMA producer =====================
        mbar_arrive_expect_tx(q_full, kTileBytes);
        tma_load_3d(&tmQKV, q_full, sQ, head * 64, q0, frame);
        int stage = 0;
        uint32_t phase = 0;
        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&kv_empty[stage], phase ^ 1);
            uint8_t* sk = sKV + stage * 2 * kTileBytes;
            mbar_arrive_expect_tx(&kv_full[stage], 2 * kTileBytes);
            tma_load_3d(&tmQKV, &kv_full[stage], sk, p.C + head * 64, j * 128, frame);
            tma_load_3d(&tmQKV, &kv_full[stage], sk + kTileBytes, 2 * p.C + head * 64, j * 128, frame);
            if (++stage == kKVStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (threadIdx.x == 32) {
        // ===================== MMA issuer =====================
        const uint32_t idesc_s = umma_idesc_f16(128, false);
        const uint32_t idesc_o = umma_idesc_f16(64, true);
        const uint64_t dq = umma_desc_sw128_kmajor(smem_u32(sQ));
        mbar_wait(q_full, 0);
        // S_0
        mbar_wait(&kv_full[0], 0);
        tc_fence_after();
        {
            const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(sKV));
#pragma unroll
            for (int k = 0; k < 4; ++k) umma_f16_ss(tmem_S, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
            umma_commit(&s_full[0]);
        }
        int stage = 0;
        uint32_t phase = 0;  // stage/phase of tile j
        for (int j = 0; j < n_kv; ++j) {
            // issue S_{j+1} early so the tensor core works while the softmax warps chew on S_j
            if (j + 1 < n_kv) {
                int nstage = stage + 1;
                uint32_t nphase = phase;
                if (nstage == kKVStages) {
                    nstage = 0;
                    nphase ^= 1;
                }
                mbar_wait(&kv_full[nstage], nphase);
                tc_fence_after();
                const uint64_t dk = umma_desc_sw128_kmajor(smem_u32(sKV + nstage * 2 * kTileBytes));
                const uint32_t d = tmem_S + ((j + 1) & 1) * 128;
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_f16_ss(d, dq + 2 * k, dk + 2 * k, idesc_s, k != 0);
                umma_commit(&s_full[(j + 1) & 1]);
            }
            // O_j = P_j V_j
            mbar_wait(&p_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            {
                const uint32_t pa = smem_u32(sP + (j & 1) * kPBytes);
                const uint32_t va = smem_u32(sKV + stage * 2 * kTileBytes + kTileBytes);
                const uint64_t dv = umma_desc_sw128_mnmajor(va, kTileBytes);
                const uint32_t d = tmem_O + (j & 1) * 64;
#pragma unroll
                for (int kk = 0; kk < 8; ++kk) {
                    const uint64_t dp = umma_desc_sw128_kmajor(pa + (kk >> 2) * kTileBytes) + 2 * (kk & 3);
                    // V: 16 kv rows x 128 B per K step = 2048 B = 128 sixteen-byte units
                    umma_f16_ss(d, dp, dv + 128 * kk, idesc_o, kk != 0);
                }
                umma_commit(&o_full[j & 1]);
                umma_commit(&kv_empty[stage]);
            }
            if (++stage == kKVStages) {
                stage = 0;
                phase ^= 1;
            }
        }
    } else if (warp >= 2) {
        // ===================== softmax / output =====================
        const int qd = warp & 3;
        const int r = qd * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(qd * 32) << 16;
        float m = -INFINITY, l = 0.f;
        float o[64];
#pragma unroll
        for (int i = 0; i < 64; ++i) o[i] = 0.f;
        const float sl2 = p.scale_log2;

        for (int j = 0; j < n_kv; ++j) {
            mbar_wait(&s_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t ts = tmem_S + lane_addr + (j & 1) * 128;
            const int kv_valid = p.L - j * 128;  // columns >= kv_valid are padding
            // pass 1: row max
            float mx = -INFINITY;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t s[32];
                tmem_ld_32x32(ts + c * 32, s);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    const float v = (c * 32 + i < kv_valid) ? __uint_as_float(s[i]) : -INFINITY;
                    mx = fmaxf(mx, v);
                }
            }
            const float m_new = fmaxf(m, mx * sl2);
            const float alpha = fast_exp2(m - m_new);  // m = -inf on the first tile -> 0
            // pass 2: p = exp2(s*sl2 - m_new) -> fp16 -> swizzled smem; row sum in fp32 of the rounded values
            float lsum = 0.f;
            uint8_t* prow = sP + (j & 1) * kPBytes + r * 128;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                uint32_t s[32];
                tmem_ld_32x32(ts + c * 32, s);
                tmem_ld_wait();
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    uint32_t packed[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int col = c * 32 + g * 8 + 2 * i;
                        float p0 = fast_exp2(__uint_as_float(s[g * 8 + 2 * i]) * sl2 - m_new);
                        float p1 = fast_exp2(__uint_as_float(s[g * 8 + 2 * i + 1]) * sl2 - m_new);
                        if (col >= kv_valid) p0 = 0.f;
                        if (col + 1 >= kv_valid) p1 = 0.f;
                        const __half2 h = __floats2half2_rn(p0, p1);
                        const float2 back = __half22float2(h);
                        lsum += back.x + back.y;
                        packed[i] = *reinterpret_cast<const uint32_t*>(&h);
                    }
                    const int col0 = c * 32 + g * 8;         // first of 8 columns
                    const int atom = col0 >> 6;              // which 64-wide K atom
                    const int chunk = (col0 & 63) >> 3;      // 16-byte chunk inside the 128-byte row
                    uint8_t* dst = prow + atom * kTileBytes + ((chunk ^ (r & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = make_uint4(packed[0], packed[1], packed[2], packed[3]);
                }
            }
            // make the generic-proxy smem writes visible to the tensor core (async proxy), then signal
            fence_proxy_async_smem();
            tc_fence_before();
            mbar_arrive(&p_full[j & 1]);

            // deferred accumulate of PV_{j-1} (relative to the previous max), then rescale to the new max
            if (j > 0) {
                mbar_wait(&o_full[(j - 1) & 1], ((j - 1) >> 1) & 1);
                tc_fence_after();
                const uint32_t to = tmem_O + lane_addr + ((j - 1) & 1) * 64;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(to + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 32; ++i) o[c * 32 + i] = (o[c * 32 + i] + __uint_as_float(v[i])) * alpha;
                }
            }
            l = l * alpha + lsum;
            m = m_new;
        }
        // last partial
        {
            const int j = n_kv - 1;
            mbar_wait(&o_full[j & 1], (j >> 1) & 1);
            tc_fence_after();
            const uint32_t to = tmem_O + lane_addr + (j & 1) * 64;
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                uint32_t v[32];
                tmem_ld_32x32(to + c * 32, v);
                tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 32; ++i) o[c * 32 + i] += __uint_as_float(v[i]);
            }
        }
        if (q0 + r < p.L) {
            const float inv = 1.0f / l;
            __half* dst = p.out + (static_cast<long long>(frame) * p.L + q0 + r) * p.C + head * 64;
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                uint32_t w[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const __half2 h = __floats2half2_rn(o[g * 8 + 2 * i] * inv, o[g * 8 + 2 * i + 1] * inv);
                    w[i] = *reinterpret_cast<const uint32_t*>(&h);
                }
                *reinterpret_cast<uint4*>(dst + g * 8) = make_uint4(w[0], w[1], w[2], w[3]);
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc(tmem_base, 512);
    }
}

}  // namespace mofa

using namespace mofa;

int mofa::attn_spatial_v1(const void* qkv, void* out, int32_t frames, int32_t L, int32_t heads, float scale,
                          mofa_stream_t stream_) {
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
    AttnParams p;
    p.out = static_cast<__half*>(out);
    p.L = L;
    p.C = C;
    p.heads = heads;
    p.n_kv = (L + 127) / 128;
    p.scale_log2 = scale * 1.4426950408889634f;
    const size_t smem_bytes = kTileBytes + kKVStages * 2 * kTileBytes + 2 * kPBytes + 16 * 8 + 16 + 1024;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_spatial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_bytes));
        if (e != cudaSuccess) {
            set_last_error("mofa_attn_spatial: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        configured = true;
    }
    dim3 grid((L + 127) / 128, heads, frames);
    attn_spatial_kernel<<<grid, kAttnThreads, smem_bytes, stream>>>(tm, p);
    return check_launch("mofa_attn_spatial");
}
