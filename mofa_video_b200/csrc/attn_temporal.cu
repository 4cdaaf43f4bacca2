// mofa_attn_temporal: self-attention over the frame axis (T <= 32 tokens, head_dim 64) of
// TemporalBasicTransformerBlock.attn1 (created at
// /root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py:169-233; the reference reshapes
// [B*T, hw, C] -> [B*hw, T, C] and calls F.scaled_dot_product_attention).
//
// The op moves 4 * 128 B per (token, head) and does 2 * 2 * T * 64 FLOPs per token: HBM-bound.  One warp owns one
// (batch item, pixel, head): cp.async copies the T rows of Q, K, V (128 B each, read in place from the fused
// [tokens, 3C] projection with the frame stride hw * 3C -- no [B*hw, T, C] transpose is ever materialised) into a
// swizzled 3 x 4 KB shared-memory slab, S = Q K^T and O = P V run on mma.sync.m16n8k16 (the tile is 32 x 32 x 64: far
// too small for tcgen05's 128-row atoms, and the tensor pipe is idle in this kernel anyway), softmax in registers
// (quad shuffles), O is staged through the Q slab and stored as full 128-B rows.  16 warps per SM keep ~190 KB of
// loads in flight.
#include <cuda_fp16.h>

#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {
int check_launch(const char* what);
void set_last_error(const char* fmt, ...);

namespace {

constexpr int kWarps = 8;
constexpr int kSlab = 32 * 64;  // halves per 32 x 64 tile (4 KB)

__device__ __forceinline__ uint32_t swz(int row, int chunk) {  // byte offset of 16-B chunk `chunk` of row `row`
    return static_cast<uint32_t>(row * 128 + ((chunk ^ (row & 7)) << 4));
}

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t (&r)[4]) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_h2(float a, float b) {
    const __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<const uint32_t*>(&h);
}

__global__ void __launch_bounds__(kWarps * 32, 2)
attn_temporal_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int B, int T, int HW, int heads,
                     float scale_log2) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int w = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const int g = lane >> 2, tg = lane & 3;
    unsigned char* slab = smem_raw + w * (3 * kSlab * 2);
    const uint32_t sQ = smem_u32(slab), sK = sQ + kSlab * 2, sV = sK + kSlab * 2;
    const int C = heads * 64;
    const long long ld = 3LL * C;
    const long long items = static_cast<long long>(B) * HW * heads;
    const long long warps_total = static_cast<long long>(gridDim.x) * kWarps;
    const int mtiles = (T + 15) >> 4;

    // rows >= T of K and V must be zero (a NaN bit pattern there would poison P V through 0 * NaN)
    for (int idx = lane; idx < (32 - T) * 8; idx += 32) {
        const int r = T + (idx >> 3), ch = idx & 7;
        *reinterpret_cast<uint4*>(slab + kSlab * 2 + swz(r, ch)) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(slab + 2 * kSlab * 2 + swz(r, ch)) = make_uint4(0, 0, 0, 0);
        *reinterpret_cast<uint4*>(slab + swz(r, ch)) = make_uint4(0, 0, 0, 0);
    }

    for (long long it = blockIdx.x * static_cast<long long>(kWarps) + w; it < items; it += warps_total) {
        const int h = static_cast<int>(it % heads);
        const long long t0 = it / heads;
        const int px = static_cast<int>(t0 % HW);
        const int b = static_cast<int>(t0 / HW);
        const long long row0 = static_cast<long long>(b) * T * HW + px;  // token t lives in row row0 + t * HW
        __syncwarp();
        for (int idx = lane; idx < T * 8; idx += 32) {
            const int t = idx >> 3, ch = idx & 7;
            const __half* src = qkv + (row0 + static_cast<long long>(t) * HW) * ld + h * 64 + ch * 8;
            const uint32_t off = swz(t, ch);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sQ + off), "l"(src));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sK + off), "l"(src + C));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sV + off), "l"(src + 2 * C));
        }
        asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
        __syncwarp();

        for (int mt = 0; mt < mtiles; ++mt) {
            // ---- S = Q K^T for query rows mt*16 .. mt*16+15, all 32 key columns
            float s[4][4];
#pragma unroll
            for (int n = 0; n < 4; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[n][e] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) {  // d = kk*16 .. +15
                uint32_t a[4];
                {
                    const int r = mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int ch = kk * 2 + (lane >> 4);
                    ldsm_x4(sQ + swz(r, ch), a);
                }
#pragma unroll
                for (int np = 0; np < 2; ++np) {  // key rows np*16 .. +15 -> n-tiles 2np, 2np+1
                    uint32_t bf[4];
                    const int r = np * 16 + (lane & 7) + (lane >> 4) * 8;
                    const int ch = kk * 2 + ((lane >> 3) & 1);
                    ldsm_x4(sK + swz(r, ch), bf);
                    mma16816(s[2 * np], a, bf[0], bf[1]);
                    mma16816(s[2 * np + 1], a, bf[2], bf[3]);
                }
            }
            // ---- softmax over the key axis; this thread holds rows g and g+8, columns n*8 + 2*tg + {0,1}
            float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const int c0 = n * 8 + 2 * tg;
                if (c0 >= T) s[n][0] = s[n][2] = -INFINITY;
                if (c0 + 1 >= T) s[n][1] = s[n][3] = -INFINITY;
                m0 = fmaxf(m0, fmaxf(s[n][0], s[n][1]));
                m1 = fmaxf(m1, fmaxf(s[n][2], s[n][3]));
            }
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
            m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
            m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
            float l0 = 0.f, l1 = 0.f;
            uint32_t p[4][2];
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                const float e0 = fast_exp2((s[n][0] - m0) * scale_log2), e1 = fast_exp2((s[n][1] - m0) * scale_log2);
                const float e2 = fast_exp2((s[n][2] - m1) * scale_log2), e3 = fast_exp2((s[n][3] - m1) * scale_log2);
                l0 += e0 + e1;
                l1 += e2 + e3;
                p[n][0] = pack_h2(e0, e1);
                p[n][1] = pack_h2(e2, e3);
            }
            l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
            l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
            l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
            const float inv0 = 1.0f / l0, inv1 = 1.0f / l1;
            // ---- O = P V : k = key index (2 steps of 16), n = d (8 tiles of 8)
            float o[8][4];
#pragma unroll
            for (int n = 0; n < 8; ++n)
#pragma unroll
                for (int e = 0; e < 4; ++e) o[n][e] = 0.f;
#pragma unroll
            for (int kk = 0; kk < 2; ++kk) {
                const uint32_t a[4] = {p[2 * kk][0], p[2 * kk][1], p[2 * kk + 1][0], p[2 * kk + 1][1]};
#pragma unroll
                for (int np = 0; np < 4; ++np) {  // d = np*16 .. +15
                    uint32_t bf[4];
                    const int r = kk * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
                    const int ch = np * 2 + (lane >> 4);
                    ldsm_x4_t(sV + swz(r, ch), bf);
                    mma16816(o[2 * np], a, bf[0], bf[1]);
                    mma16816(o[2 * np + 1], a, bf[2], bf[3]);
                }
            }
            // ---- stage O (fp16) over the Q rows this m-tile has finished with
            __syncwarp();
#pragma unroll
            for (int n = 0; n < 8; ++n) {
                const int r0 = mt * 16 + g;
                *reinterpret_cast<uint32_t*>(slab + swz(r0, n) + tg * 4) = pack_h2(o[n][0] * inv0, o[n][1] * inv0);
                *reinterpret_cast<uint32_t*>(slab + swz(r0 + 8, n) + tg * 4) = pack_h2(o[n][2] * inv1, o[n][3] * inv1);
            }
        }
        __syncwarp();
        for (int idx = lane; idx < T * 8; idx += 32) {
            const int t = idx >> 3, ch = idx & 7;
            const uint4 v = *reinterpret_cast<const uint4*>(slab + swz(t, ch));
            *reinterpret_cast<uint4*>(out + (row0 + static_cast<long long>(t) * HW) * C + h * 64 + ch * 8) = v;
        }
        // rows >= T of the Q slab were overwritten with O rows of padding queries: harmless (never stored), but they
        // must stay finite for the next item's S; they are (products of finite values), and K/V padding is untouched.
    }
}

}  // namespace
}  // namespace mofa

using namespace mofa;

extern "C" int mofa_attn_temporal(const void* qkv, void* out, int32_t B, int32_t T, int32_t HW, int32_t heads,
                                  float scale, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qkv || !out || B <= 0 || T <= 0 || T > 32 || HW <= 0 || heads <= 0) {
        set_last_error("mofa_attn_temporal: needs 1 <= T <= 32");
        return MOFA_ERR_ARG;
    }
    const size_t smem_bytes = static_cast<size_t>(kWarps) * 3 * kSlab * 2;  // 96 KB, two CTAs per SM
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_temporal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem_bytes));
        if (e != cudaSuccess) {
            set_last_error("mofa_attn_temporal: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        configured = true;
    }
    const long long items = static_cast<long long>(B) * HW * heads;
    long long blocks = (items + kWarps - 1) / kWarps;
    if (blocks > 148LL * 2) blocks = 148LL * 2;
    attn_temporal_kernel<<<static_cast<unsigned>(blocks), kWarps * 32, smem_bytes, stream>>>(
        static_cast<const __half*>(qkv), static_cast<__half*>(out), B, T, HW, heads,
        scale * 1.4426950408889634f);
    return check_launch("mofa_attn_temporal");
}
