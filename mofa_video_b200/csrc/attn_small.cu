// Small-sequence multi-head self-attention for head dims that are not 64: the CLIP ViT-H/14 image encoder of
// /root/reference/MOFA-Video-Traj/pipeline/pipeline.py:114-141 (`self.image_encoder(image).image_embeds`, model built at
// run_gradio.py:98): 16 heads x d = 80, 257 tokens, 32 layers, ONE image per clip -- 10.8 GFLOP of attention per clip in
// total, i.e. nothing to put on tensor cores; what matters is that it is one launch per layer and reads K / V once.
//
//   grid = (ceil(L / 32), heads, sequences), 128 threads: the head's K and V ([L, d] fp16 each) are staged once in shared
//   memory (row stride d + 2 halves: odd word stride, conflict-free for the strided K reads); every warp owns 8 query rows.
//   Per row: lanes split the keys for S = q K^T (fp32), warp-shuffle max / sum, probabilities to shared memory, then lanes
//   split the head dimension for O = P V (each lane accumulates one or two half2 columns), normalised, written as fp16.
#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

constexpr int kSmallAttnWarps = 4;
constexpr int kRowsPerWarp = 8;

__global__ void __launch_bounds__(kSmallAttnWarps * 32)
attn_small_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int heads, int d, float scale_log2) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int C = heads * d;
    const int ldk = d + 2;                       // halves per shared row
    __half* sK = reinterpret_cast<__half*>(smem_raw);
    __half* sV = sK + static_cast<size_t>(L) * ldk;
    float* sP = reinterpret_cast<float*>(sV + static_cast<size_t>(L) * ldk);   // [warps][Lpad]
    const int Lpad = (L + 31) & ~31;
    float* sQ = sP + kSmallAttnWarps * Lpad;                                    // [warps][d]

    const int head = blockIdx.y, seq = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* base = qkv + static_cast<long long>(seq) * L * 3 * C;

    // stage K and V of this head: d / 2 half2 per row, coalesced along d
    const int d2 = d >> 1;
    for (int i = threadIdx.x; i < L * d2; i += blockDim.x) {
        const int r = i / d2, c = i - r * d2;
        const __half2* src = reinterpret_cast<const __half2*>(base + static_cast<long long>(r) * 3 * C + head * d);
        reinterpret_cast<__half2*>(sK + r * ldk)[c] = src[(C >> 1) + c];
        reinterpret_cast<__half2*>(sV + r * ldk)[c] = src[C + c];
    }
    __syncthreads();

    float* myP = sP + warp * Lpad;
    float* myQ = sQ + warp * d;
    const int row0 = blockIdx.x * (kSmallAttnWarps * kRowsPerWarp) + warp * kRowsPerWarp;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int row = row0 + rr;
        if (row >= L) break;                                   // warp-uniform
        const __half* qrow = base + static_cast<long long>(row) * 3 * C + head * d;
        for (int c = lane; c < d; c += 32) myQ[c] = __half2float(qrow[c]) * scale_log2;
        __syncwarp();
        // scores: lane owns keys lane, lane + 32, ...
        float mx = -INFINITY;
        for (int j = lane; j < L; j += 32) {
            const __half2* kr = reinterpret_cast<const __half2*>(sK + j * ldk);
            float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
            for (int c = 0; c < d2; ++c) {
                const float2 kv = __half22float2(kr[c]);
                s0 = fmaf(myQ[2 * c], kv.x, s0);
                s1 = fmaf(myQ[2 * c + 1], kv.y, s1);
            }
            const float s = s0 + s1;
            myP[j] = s;
            mx = fmaxf(mx, s);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float sum = 0.f;
        for (int j = lane; j < L; j += 32) {
            const float p = fast_exp2(myP[j] - mx);
            myP[j] = p;
            sum += p;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        __syncwarp();
        const float inv = 1.0f / sum;
        // O = P V: lane owns half2 columns lane and lane + 32 (d <= 128)
        float2 a0 = make_float2(0.f, 0.f), a1 = make_float2(0.f, 0.f);
        const bool has0 = lane < d2, has1 = lane + 32 < d2;
        for (int j = 0; j < L; ++j) {
            const float p = myP[j];
            const __half2* vr = reinterpret_cast<const __half2*>(sV + j * ldk);
            if (has0) {
                const float2 v = __half22float2(vr[lane]);
                a0.x = fmaf(p, v.x, a0.x);
                a0.y = fmaf(p, v.y, a0.y);
            }
            if (has1) {
                const float2 v = __half22float2(vr[lane + 32]);
                a1.x = fmaf(p, v.x, a1.x);
                a1.y = fmaf(p, v.y, a1.y);
            }
        }
        __half2* orow = reinterpret_cast<__half2*>(out + (static_cast<long long>(seq) * L + row) * C + head * d);
        if (has0) orow[lane] = __floats2half2_rn(a0.x * inv, a0.y * inv);
        if (has1) orow[lane + 32] = __floats2half2_rn(a1.x * inv, a1.y * inv);
        __syncwarp();
    }
}

}  // namespace mofa

using namespace mofa;

extern "C" int mofa_attn_small(const void* qkv, void* out, int32_t n_seq, int32_t L, int32_t heads, int32_t head_dim,
                               float scale, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!qkv || !out || n_seq <= 0 || L <= 0 || heads <= 0 || head_dim <= 0 || (head_dim & 1) || head_dim > 128) {
        set_last_error("mofa_attn_small: needs an even head_dim <= 128 (head_dim=%d)", head_dim);
        return MOFA_ERR_ARG;
    }
    const int Lpad = (L + 31) & ~31;
    const size_t smem = static_cast<size_t>(2) * L * (head_dim + 2) * 2 + static_cast<size_t>(kSmallAttnWarps) * Lpad * 4 +
                        static_cast<size_t>(kSmallAttnWarps) * head_dim * 4;
    if (smem > 220 * 1024) {
        set_last_error("mofa_attn_small: sequence too long for one shared-memory stage (L=%d, head_dim=%d)", L, head_dim);
        return MOFA_ERR_ARG;
    }
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) {
            set_last_error("mofa_attn_small: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        configured = smem;
    }
    dim3 grid((L + kSmallAttnWarps * kRowsPerWarp - 1) / (kSmallAttnWarps * kRowsPerWarp), heads, n_seq);
    attn_small_kernel<<<grid, kSmallAttnWarps * 32, smem, stream>>>(static_cast<const __half*>(qkv),
                                                                   static_cast<__half*>(out), L, heads, head_dim,
                                                                   scale * 1.4426950408889634f);
    return check_launch("mofa_attn_small");
}
