// Small-sequence multi-head self-attention for head dims that are not 64: the CLIP ViT-H/14 image encoder of
// /root/reference/MOFA-Video-Traj/pipeline/pipeline.py:114-141 (`self.image_encoder(image).image_embeds`, model built at
// run_gradio.py:98): 16 heads x d = 80, 257 tokens, 32 layers, ONE image per clip -- 10.8 GFLOP of attention per clip in
// total, i.e. nothing to put on tensor cores; what matters is that it is one launch per layer and reads K / V once.
//
//   grid = (ceil(L / 32), heads, sequences), 128 threads: the head's K and V are staged in shared memory 256 keys at a time
//   (row stride d + 2 halves: odd word stride, conflict-free for the strided K reads; online softmax across the passes, so L is
//   unbounded); every warp owns 8 query rows.  Per row and pass: lanes split the keys for S = q K^T (fp32), warp-shuffle
//   max / sum, probabilities to shared memory, then lanes split the head dimension for O += P V (one or two half2 columns
//   per lane).  The same kernel with a strided token layout serves the UNet's attention when a checkpoint's head_dim is
//   not 64 (the reference class default heads (5,10,10,20), UNET.py:93, gives d = 128 at 1280 channels): correct for every
//   configuration, while the tcgen05 kernels (attn_spatial.cu / attn_temporal.cu) remain the d = 64 production path.
#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

constexpr int kSmallAttnWarps = 4;
constexpr int kRowsPerWarp = 8;
constexpr int kKeyTile = 256;   // keys staged per pass (online softmax across passes)

struct SmallAttnLayout {
    // element offsets: token t of sequence s lives at (s / inner) * outer + (s % inner) * inner_stride + t * tok
    long long tok, outer, inner_stride;
    int inner;
};

__global__ void __launch_bounds__(kSmallAttnWarps * 32)
attn_small_kernel(const __half* __restrict__ qkv, __half* __restrict__ out, int L, int heads, int d, float scale_log2,
                  SmallAttnLayout li, SmallAttnLayout lo) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    const int C = heads * d;
    const int ldk = d + 2;                       // halves per shared row (odd word stride: conflict-free key reads)
    const int TK = L < kKeyTile ? L : kKeyTile;
    __half* sK = reinterpret_cast<__half*>(smem_raw);
    __half* sV = sK + static_cast<size_t>(TK) * ldk;
    float* sP = reinterpret_cast<float*>(sV + static_cast<size_t>(TK) * ldk);   // [warps][TKpad]
    const int TKpad = (TK + 31) & ~31;
    float* sQ = sP + kSmallAttnWarps * TKpad;                                    // [warps][rows][d]

    const int head = blockIdx.y, seq = blockIdx.z;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const __half* base = qkv + (seq / li.inner) * li.outer + (seq % li.inner) * li.inner_stride;
    __half* obase = out + (seq / lo.inner) * lo.outer + (seq % lo.inner) * lo.inner_stride;
    const int d2 = d >> 1;

    float* myP = sP + warp * TKpad;
    float* myQ = sQ + warp * kRowsPerWarp * d;
    const int row0 = blockIdx.x * (kSmallAttnWarps * kRowsPerWarp) + warp * kRowsPerWarp;
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int row = row0 + rr;
        if (row >= L) break;
        const __half* qrow = base + row * li.tok + head * d;
        for (int c = lane; c < d; c += 32) myQ[rr * d + c] = __half2float(qrow[c]) * scale_log2;
    }
    float m_run[kRowsPerWarp], l_run[kRowsPerWarp];
    float2 a0[kRowsPerWarp], a1[kRowsPerWarp];
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        m_run[rr] = -INFINITY;
        l_run[rr] = 0.f;
        a0[rr] = make_float2(0.f, 0.f);
        a1[rr] = make_float2(0.f, 0.f);
    }
    const bool has0 = lane < d2, has1 = lane + 32 < d2;

    for (int k0 = 0; k0 < L; k0 += TK) {
        const int nk = (L - k0) < TK ? (L - k0) : TK;
        __syncthreads();   // previous tile fully consumed
        for (int i = threadIdx.x; i < nk * d2; i += blockDim.x) {
            const int r = i / d2, c = i - r * d2;
            const __half2* src = reinterpret_cast<const __half2*>(base + (k0 + r) * li.tok + head * d);
            reinterpret_cast<__half2*>(sK + r * ldk)[c] = src[(C >> 1) + c];
            reinterpret_cast<__half2*>(sV + r * ldk)[c] = src[C + c];
        }
        __syncthreads();
#pragma unroll
        for (int rr = 0; rr < kRowsPerWarp; ++rr) {
            if (row0 + rr >= L) break;                          // warp-uniform
            const float* q = myQ + rr * d;
            float mx = -INFINITY;
            for (int j = lane; j < nk; j += 32) {               // lane owns keys lane, lane + 32, ...
                const __half2* kr = reinterpret_cast<const __half2*>(sK + j * ldk);
                float s0 = 0.f, s1 = 0.f;
#pragma unroll 4
                for (int c = 0; c < d2; ++c) {
                    const float2 kv = __half22float2(kr[c]);
                    s0 = fmaf(q[2 * c], kv.x, s0);
                    s1 = fmaf(q[2 * c + 1], kv.y, s1);
                }
                const float sc = s0 + s1;
                myP[j] = sc;
                mx = fmaxf(mx, sc);
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
            const float m_new = fmaxf(m_run[rr], mx);
            const float alpha = fast_exp2(m_run[rr] - m_new);   // first tile: exp2(-inf) = 0
            float sum = 0.f;
            for (int j = lane; j < nk; j += 32) {
                const float pj = fast_exp2(myP[j] - m_new);
                myP[j] = pj;
                sum += pj;
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
            __syncwarp();
            l_run[rr] = l_run[rr] * alpha + sum;
            m_run[rr] = m_new;
            float2 b0 = make_float2(a0[rr].x * alpha, a0[rr].y * alpha);
            float2 b1 = make_float2(a1[rr].x * alpha, a1[rr].y * alpha);
            for (int j = 0; j < nk; ++j) {                      // lane owns half2 columns lane and lane + 32
                const float pj = myP[j];
                const __half2* vr = reinterpret_cast<const __half2*>(sV + j * ldk);
                if (has0) {
                    const float2 v = __half22float2(vr[lane]);
                    b0.x = fmaf(pj, v.x, b0.x);
                    b0.y = fmaf(pj, v.y, b0.y);
                }
                if (has1) {
                    const float2 v = __half22float2(vr[lane + 32]);
                    b1.x = fmaf(pj, v.x, b1.x);
                    b1.y = fmaf(pj, v.y, b1.y);
                }
            }
            a0[rr] = b0;
            a1[rr] = b1;
            __syncwarp();
        }
    }
#pragma unroll
    for (int rr = 0; rr < kRowsPerWarp; ++rr) {
        const int row = row0 + rr;
        if (row >= L) break;
        const float inv = 1.0f / l_run[rr];
        __half2* orow = reinterpret_cast<__half2*>(obase + row * lo.tok + head * d);
        if (has0) orow[lane] = __floats2half2_rn(a0[rr].x * inv, a0[rr].y * inv);
        if (has1) orow[lane + 32] = __floats2half2_rn(a1[rr].x * inv, a1[rr].y * inv);
    }
}

}  // namespace mofa

using namespace mofa;

static int launch_attn_small(const void* qkv, void* out, int n_seq, int L, int heads, int head_dim, float scale,
                             const SmallAttnLayout& li, const SmallAttnLayout& lo, cudaStream_t stream, const char* what) {
    if (!qkv || !out || n_seq <= 0 || n_seq > 65535 || L <= 0 || heads <= 0 || heads > 65535 || head_dim <= 0 ||
        (head_dim & 1) || head_dim > 128) {
        set_last_error("%s: needs an even head_dim <= 128 and <= 65535 sequences (head_dim=%d, n_seq=%d)", what, head_dim,
                       n_seq);
        return MOFA_ERR_ARG;
    }
    const int TK = L < kKeyTile ? L : kKeyTile;
    const int TKpad = (TK + 31) & ~31;
    const size_t smem = static_cast<size_t>(2) * TK * (head_dim + 2) * 2 + static_cast<size_t>(kSmallAttnWarps) * TKpad * 4 +
                        static_cast<size_t>(kSmallAttnWarps) * kRowsPerWarp * head_dim * 4;
    static size_t configured = 0;
    if (smem > configured) {
        cudaError_t e = cudaFuncSetAttribute(attn_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             static_cast<int>(smem));
        if (e != cudaSuccess) {
            set_last_error("%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
            return MOFA_ERR_CUDA;
        }
        configured = smem;
    }
    dim3 grid((L + kSmallAttnWarps * kRowsPerWarp - 1) / (kSmallAttnWarps * kRowsPerWarp), heads, n_seq);
    attn_small_kernel<<<grid, kSmallAttnWarps * 32, smem, stream>>>(static_cast<const __half*>(qkv),
                                                                   static_cast<__half*>(out), L, heads, head_dim,
                                                                   scale * 1.4426950408889634f, li, lo);
    return check_launch(what);
}

extern "C" int mofa_attn_small(const void* qkv, void* out, int32_t n_seq, int32_t L, int32_t heads, int32_t head_dim,
                               float scale, mofa_stream_t stream_) {
    const long long C = static_cast<long long>(heads) * head_dim;
    SmallAttnLayout li{3 * C, static_cast<long long>(L) * 3 * C, 0, 1};
    SmallAttnLayout lo{C, static_cast<long long>(L) * C, 0, 1};
    return launch_attn_small(qkv, out, n_seq, L, heads, head_dim, scale, li, lo, static_cast<cudaStream_t>(stream_),
                             "mofa_attn_small");
}

extern "C" int mofa_attn_small_temporal(const void* qkv, void* out, int32_t B, int32_t T, int32_t HW, int32_t heads,
                                        int32_t head_dim, float scale, mofa_stream_t stream_) {
    // qkv [B, T, HW, 3C]: one sequence of T tokens per (b, pixel); token stride HW * 3C
    const long long C = static_cast<long long>(heads) * head_dim;
    if (static_cast<long long>(B) * HW > 65535) {
        // sequences ride on gridDim.z: split the pixels into slabs
        const int slab = 65535 / (B > 0 ? B : 1);
        if (slab <= 0) {
            set_last_error("mofa_attn_small_temporal: batch too large");
            return MOFA_ERR_ARG;
        }
        for (int p0 = 0; p0 < HW; p0 += slab) {
            const int np = (HW - p0) < slab ? (HW - p0) : slab;
            SmallAttnLayout li{static_cast<long long>(HW) * 3 * C, static_cast<long long>(T) * HW * 3 * C, 3 * C, np};
            SmallAttnLayout lo{static_cast<long long>(HW) * C, static_cast<long long>(T) * HW * C, C, np};
            int rc = launch_attn_small(static_cast<const __half*>(qkv) + static_cast<long long>(p0) * 3 * C,
                                       static_cast<__half*>(out) + static_cast<long long>(p0) * C, B * np, T, heads,
                                       head_dim, scale, li, lo, static_cast<cudaStream_t>(stream_),
                                       "mofa_attn_small_temporal");
            if (rc) return rc;
        }
        return MOFA_OK;
    }
    SmallAttnLayout li{static_cast<long long>(HW) * 3 * C, static_cast<long long>(T) * HW * 3 * C, 3 * C, HW};
    SmallAttnLayout lo{static_cast<long long>(HW) * C, static_cast<long long>(T) * HW * C, C, HW};
    return launch_attn_small(qkv, out, B * HW, T, heads, head_dim, scale, li, lo, static_cast<cudaStream_t>(stream_),
                             "mofa_attn_small_temporal");
}
