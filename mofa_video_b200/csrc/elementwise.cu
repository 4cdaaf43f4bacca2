// HBM-bound kernels of the hot path: GroupNorm(+SiLU), LayerNorm, broadcast axpy, im2col, nearest
// upsample, NCHW<->NHWC, small-M linear, sinusoidal timestep embedding, fused CFG + Euler step.
// All activations are fp16 channels-last; statistics and arithmetic are fp32; 16-byte vector accesses.
#include "../../include/mofa_b200.h"
#include <stdlib.h>

#include <cstdlib>
#include "common.cuh"

namespace mofa {

union V8 {
    uint4 u;
    __half h[8];
    __half2 h2[4];
};

// =============================================================================================
// GroupNorm (torch.nn.GroupNorm(32, C, eps) [+ SiLU]) on channels-last data, optional concat input.
// Two streaming passes (statistics, then normalise): block = (VX vectors of 8 channels) x (VY rows), so a block reads
// a contiguous slab with fully coalesced 16-B loads; every thread owns fixed channels (its per-channel scale / shift
// are computed once per block) and keeps kUnroll independent loads in flight.
// =============================================================================================
constexpr int kGnUnroll = 4;

__device__ __forceinline__ const __half* gn_src(const __half* x1, int C1, const __half* x2, int C2, int c,
                                                long long& ld) {
    if (c < C1) {
        ld = C1;
        return x1 + c;
    }
    ld = C2;
    return x2 + (c - C1);
}

__global__ void __launch_bounds__(256, 3)
groupnorm_stats_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2,
                       long long rows_per_stat, int rows_per_block, int slabs_per_stat, long long n_slabs, int groups,
                       float* __restrict__ stats) {
    constexpr int kU = 8;  // independent 16-byte loads in flight per thread
    const int C = C1 + C2;
    const int vecs = C >> 3;
    const int cpg = C / groups;
    const int VX = blockDim.x, VY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    __shared__ float s_sum[64], s_sq[64];
    const int tid = ty * VX + tx;

    // persistent: the grid is sized to the machine (no partial last wave), slabs are dealt round-robin
    for (long long slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        const long long s = slab / slabs_per_stat;
        const long long r_begin = (slab - s * slabs_per_stat) * rows_per_block;
        long long r_end = r_begin + rows_per_block;
        if (r_end > rows_per_stat) r_end = rows_per_stat;
        const long long row0 = s * rows_per_stat;
        if (tid < 64) {
            s_sum[tid] = 0.f;
            s_sq[tid] = 0.f;
        }
        __syncthreads();

        for (int vec = tx; vec < vecs; vec += VX) {
            const int c = vec << 3;
            long long ld;
            const __half* base = gn_src(x1, C1, x2, C2, c, ld) + row0 * ld;
            float sum[8], sq[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) sum[j] = sq[j] = 0.f;
            long long r = r_begin + ty;
            for (; r + static_cast<long long>(kU - 1) * VY < r_end; r += static_cast<long long>(kU) * VY) {
                V8 v[kU];
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    v[u].u = __ldg(reinterpret_cast<const uint4*>(base + (r + static_cast<long long>(u) * VY) * ld));
#pragma unroll
                for (int u = 0; u < kU; ++u)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(v[u].h2[j]);
                        sum[2 * j] += f.x;
                        sq[2 * j] = fmaf(f.x, f.x, sq[2 * j]);
                        sum[2 * j + 1] += f.y;
                        sq[2 * j + 1] = fmaf(f.y, f.y, sq[2 * j + 1]);
                    }
            }
            for (; r < r_end; r += VY) {
                V8 v;
                v.u = __ldg(reinterpret_cast<const uint4*>(base + r * ld));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float f = __half2float(v.h[j]);
                    sum[j] += f;
                    sq[j] = fmaf(f, f, sq[j]);
                }
            }
            // flush: merge the channels of one group before touching shared memory (cpg >= 8: at most two groups)
            int g_prev = c / cpg;
            float a_sum = 0.f, a_sq = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int g = (c + j) / cpg;
                if (g != g_prev) {
                    atomicAdd(&s_sum[g_prev], a_sum);
                    atomicAdd(&s_sq[g_prev], a_sq);
                    a_sum = a_sq = 0.f;
                    g_prev = g;
                }
                a_sum += sum[j];
                a_sq += sq[j];
            }
            atomicAdd(&s_sum[g_prev], a_sum);
            atomicAdd(&s_sq[g_prev], a_sq);
        }
        __syncthreads();
        if (tid < groups) {
            atomicAdd(&stats[(s * groups + tid) * 2 + 0], s_sum[tid]);
            atomicAdd(&stats[(s * groups + tid) * 2 + 1], s_sq[tid]);
        }
        __syncthreads();
    }
}

template <int kU, int kMinBlocks, bool kPrefetch>
__global__ void __launch_bounds__(256, kMinBlocks)
groupnorm_apply_kernel(const __half* __restrict__ x1, int C1, const __half* __restrict__ x2, int C2,
                       const __half* __restrict__ gamma, const __half* __restrict__ beta, __half* __restrict__ out,
                       long long rows_per_stat, int rows_per_block, int slabs_per_stat, long long n_slabs, int groups,
                       float eps, int silu, const float* __restrict__ stats) {
    const int C = C1 + C2;
    const int vecs = C >> 3;
    const int cpg = C / groups;
    const int VX = blockDim.x, VY = blockDim.y;
    const int tx = threadIdx.x, ty = threadIdx.y;
    const float inv_cnt = 1.0f / (static_cast<float>(cpg) * static_cast<float>(rows_per_stat));

    for (long long slab = blockIdx.x; slab < n_slabs; slab += gridDim.x) {
        const long long s = slab / slabs_per_stat;
        const long long r_begin = (slab - s * slabs_per_stat) * rows_per_block;
        long long r_end = r_begin + rows_per_block;
        if (r_end > rows_per_stat) r_end = rows_per_stat;
        const long long row0 = s * rows_per_stat;
        for (int vec = tx; vec < vecs; vec += VX) {
            const int c = vec << 3;
            long long ld;
            const __half* base = gn_src(x1, C1, x2, C2, c, ld) + row0 * ld;
            __half* obase = out + row0 * C + c;
            V8 gm, bt;
            gm.u = __ldg(reinterpret_cast<const uint4*>(gamma + c));
            bt.u = __ldg(reinterpret_cast<const uint4*>(beta + c));
            float sc[8], sh[8];
            int gi_prev = -1;
            float mean = 0.f, rstd = 0.f;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int gi = (c + j) / cpg;
                if (gi != gi_prev) {
                    const float sm = stats[(s * groups + gi) * 2 + 0];
                    const float sq = stats[(s * groups + gi) * 2 + 1];
                    mean = sm * inv_cnt;
                    float var = sq * inv_cnt - mean * mean;
                    var = var < 0.f ? 0.f : var;
                    rstd = rsqrtf(var + eps);
                    gi_prev = gi;
                }
                sc[j] = rstd * __half2float(gm.h[j]);
                sh[j] = __half2float(bt.h[j]) - mean * sc[j];
            }
            // Two batches of kU rows in flight: the next batch is requested before the current one is normalised, so the
            // DRAM round trip overlaps the math and the stores instead of being exposed once per batch.
            long long r = r_begin + ty;
            const long long step = static_cast<long long>(kU) * VY;
            auto full = [&](long long rr) { return rr + static_cast<long long>(kU - 1) * VY < r_end; };
            auto load = [&](V8 (&v)[kU], long long rr) {
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    v[u].u = __ldg(reinterpret_cast<const uint4*>(base + (rr + static_cast<long long>(u) * VY) * ld));
            };
            auto process = [&](const V8 (&v)[kU], long long rr) {
#pragma unroll
                for (int u = 0; u < kU; ++u) {
                    V8 o;
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float2 f = __half22float2(v[u].h2[j]);
                        float y0 = fmaf(f.x, sc[2 * j], sh[2 * j]);
                        float y1 = fmaf(f.y, sc[2 * j + 1], sh[2 * j + 1]);
                        if (silu) {
                            y0 = silu_f(y0);
                            y1 = silu_f(y1);
                        }
                        o.h2[j] = __floats2half2_rn(y0, y1);
                    }
                    *reinterpret_cast<uint4*>(obase + (rr + static_cast<long long>(u) * VY) * C) = o.u;
                }
            };
            if constexpr (!kPrefetch) {   // one batch at a time (round-1 loop; kept for same-box A/Bs: MOFA_GN_APPLY=2)
                for (; full(r); r += step) {
                    V8 v[kU];
                    load(v, r);
                    process(v, r);
                }
            }
            V8 va[kU], vb[kU];
            bool ha = kPrefetch && full(r);
            if (ha) load(va, r);
            while (ha) {
                const long long r2 = r + step;
                const bool hb = full(r2);
                if (hb) load(vb, r2);
                process(va, r);
                r = r2;
                if (!hb) break;
                const long long r3 = r2 + step;
                ha = full(r3);
                if (ha) load(va, r3);
                process(vb, r2);
                r = r3;
            }
            for (; r < r_end; r += VY) {
                V8 v, o;
                v.u = __ldg(reinterpret_cast<const uint4*>(base + r * ld));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float y = fmaf(__half2float(v.h[j]), sc[j], sh[j]);
                    if (silu) y = silu_f(y);
                    o.h[j] = __float2half_rn(y);
                }
                *reinterpret_cast<uint4*>(obase + r * C) = o.u;
            }
        }
    }
}

// =============================================================================================
// LayerNorm over C (<= 2048) per row; LPR lanes share one row (8 / 16 / 32 for C = 320 / 640 / 1280, so every lane
// carries <= VPL 16-byte vectors and a warp keeps 32/LPR rows in flight); optional per-row-group pre-add
// =============================================================================================
template <int LPR, int VPL>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                 __half* __restrict__ out, long long rows, int C, float eps, const __half* __restrict__ add,
                 long long rows_per_group, long long add_period, __half* __restrict__ sum_out) {
    constexpr int RPW = 32 / LPR;
    const int lane = threadIdx.x & 31;
    const int sub = lane % LPR;
    const long long warp = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    const long long row = warp * RPW + lane / LPR;
    const bool live = row < rows;
    const int vecs = C >> 3;
    float v[VPL][8];
    const __half* addrow = (add && live) ? add + ((row / rows_per_group) % add_period) * C : nullptr;
    V8 t[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vec = sub + i * LPR;
        if (live && vec < vecs) t[i].u = __ldg(reinterpret_cast<const uint4*>(x + row * C + vec * 8));
    }
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vec = sub + i * LPR;
        if (live && vec < vecs) {
            if (addrow) {
                V8 a;
                a.u = __ldg(reinterpret_cast<const uint4*>(addrow + vec * 8));
                V8 so;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    // the reference adds in fp16 (hidden_states + emb), so round the sum to fp16
                    so.h[j] = __float2half_rn(__half2float(t[i].h[j]) + __half2float(a.h[j]));
                    v[i][j] = __half2float(so.h[j]);
                }
                if (sum_out) *reinterpret_cast<uint4*>(sum_out + row * C + vec * 8) = so.u;
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[i][j] = __half2float(t[i].h[j]);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) sum += v[i][j];
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float mean = sum / static_cast<float>(C);
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vec = sub + i * LPR;
        if (live && vec < vecs) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float d = v[i][j] - mean;
                sq = fmaf(d, d, sq);
            }
        }
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    const float rstd = rsqrtf(sq / static_cast<float>(C) + eps);
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
        const int vec = sub + i * LPR;
        if (live && vec < vecs) {
            V8 g, b, o;
            g.u = __ldg(reinterpret_cast<const uint4*>(gamma + vec * 8));
            b.u = __ldg(reinterpret_cast<const uint4*>(beta + vec * 8));
#pragma unroll
            for (int j = 0; j < 8; ++j)
                o.h[j] = __float2half_rn((v[i][j] - mean) * rstd * __half2float(g.h[j]) + __half2float(b.h[j]));
            *reinterpret_cast<uint4*>(out + row * C + vec * 8) = o.u;
        }
    }
}

// =============================================================================================
// out = x + scale * y[i % period]
// =============================================================================================
__global__ void __launch_bounds__(256)
axpy_bcast_kernel(const __half* __restrict__ x, const __half* __restrict__ y, __half* __restrict__ out,
                  long long nvec, long long period_vec, float scale) {
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < nvec;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        V8 a, b, o;
        a.u = __ldg(reinterpret_cast<const uint4*>(x) + i);
        b.u = __ldg(reinterpret_cast<const uint4*>(y) + (i % period_vec));
#pragma unroll
        for (int j = 0; j < 8; ++j) o.h[j] = __float2half_rn(__half2float(a.h[j]) + scale * __half2float(b.h[j]));
        reinterpret_cast<uint4*>(out)[i] = o.u;
    }
}

// =============================================================================================
// im2col 3x3, pad 1, stride s (for stride-2 convs and convs whose C is not a multiple of 64)
// =============================================================================================
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C, int stride,
                 int Ho, int Wo, int Kpad) {
    const long long total = static_cast<long long>(n_img) * Ho * Wo * Kpad;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = idx / Kpad;
        const int k = static_cast<int>(idx - row * Kpad);
        __half v = __float2half(0.f);
        if (k < 9 * C) {
            const int tap = k / C;
            const int c = k - tap * C;
            const int ky = tap / 3, kx = tap - ky * 3;
            const int ox = static_cast<int>(row % Wo);
            const long long t = row / Wo;
            const int oy = static_cast<int>(t % Ho);
            const int n = static_cast<int>(t / Ho);
            const int iy = oy * stride + ky - 1;
            const int ix = ox * stride + kx - 1;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((static_cast<long long>(n) * H + iy) * W + ix) * C + c];
        }
        out[idx] = v;
    }
}
// vectorised variant for C % 8 == 0 and Kpad == 9*C
__global__ void __launch_bounds__(256)
im2col3x3_vec_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C,
                     int stride, int Ho, int Wo) {
    const int cv = C >> 3;
    const long long total = static_cast<long long>(n_img) * Ho * Wo * 9 * cv;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(idx % cv);
        long long t = idx / cv;
        const int tap = static_cast<int>(t % 9);
        const long long row = t / 9;
        const int ky = tap / 3, kx = tap - ky * 3;
        const int ox = static_cast<int>(row % Wo);
        t = row / Wo;
        const int oy = static_cast<int>(t % Ho);
        const int n = static_cast<int>(t / Ho);
        const int iy = oy * stride + ky - 1;
        const int ix = ox * stride + kx - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W)
            v = __ldg(reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + iy) * W + ix) * C) + c8);
        reinterpret_cast<uint4*>(out)[idx] = v;
    }
}

__global__ void __launch_bounds__(256)
upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C) {
    const int cv = C >> 3;
    const int Ho = 2 * H, Wo = 2 * W;
    const long long total = static_cast<long long>(n_img) * Ho * Wo * cv;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c8 = static_cast<int>(idx % cv);
        long long t = idx / cv;
        const int ox = static_cast<int>(t % Wo);
        t /= Wo;
        const int oy = static_cast<int>(t % Ho);
        const int n = static_cast<int>(t / Ho);
        reinterpret_cast<uint4*>(out)[idx] = __ldg(
            reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + (oy >> 1)) * W + (ox >> 1)) * C) + c8);
    }
}

__global__ void __launch_bounds__(256)
nchw_to_nhwc_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int C, int HW, int ldo,
                    int c_off) {
    const long long total = static_cast<long long>(n_img) * HW * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        const long long t = idx / C;
        const int p = static_cast<int>(t % HW);
        const long long n = t / HW;
        out[(n * HW + p) * ldo + c_off + c] = x[(n * C + c) * HW + p];
    }
}
__global__ void __launch_bounds__(256)
nhwc_to_nchw_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int C, int HW, int ldi,
                    int c_off) {
    const long long total = static_cast<long long>(n_img) * HW * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int p = static_cast<int>(idx % HW);
        const long long t = idx / HW;
        const int c = static_cast<int>(t % C);
        const long long n = t / C;
        out[idx] = x[(n * HW + p) * ldi + c_off + c];
    }
}

// =============================================================================================
// small-M linear: one warp per output column
// =============================================================================================
__global__ void __launch_bounds__(256)
linear_small_kernel(const __half* __restrict__ a, const __half* __restrict__ w, const __half* __restrict__ bias,
                    __half* __restrict__ out, int M, int N, int K, int act_in, int act_out) {
    const int lane = threadIdx.x & 31;
    const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (n >= N) return;
    float acc[8];
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[m] = 0.f;
    for (int k = lane * 8; k < K; k += 256) {
        V8 wv;
        wv.u = __ldg(reinterpret_cast<const uint4*>(w + static_cast<long long>(n) * K + k));
        float wf[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) wf[j] = __half2float(wv.h[j]);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
            if (m < M) {
                V8 av;
                av.u = __ldg(reinterpret_cast<const uint4*>(a + static_cast<long long>(m) * K + k));
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float f = __half2float(av.h[j]);
                    if (act_in == 1) f = __half2float(__float2half_rn(silu_f(f)));
                    acc[m] += f * wf[j];
                }
            }
        }
    }
#pragma unroll
    for (int m = 0; m < 8; ++m) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc[m] += __shfl_xor_sync(0xffffffffu, acc[m], o);
    }
    if (lane == 0) {
        const float b = bias ? __half2float(bias[n]) : 0.f;
        for (int m = 0; m < M; ++m) {
            float v = acc[m] + b;
            if (act_out == 1) v = silu_f(v);
            out[static_cast<long long>(m) * N + n] = __float2half_rn(v);
        }
    }
}

// diffusers Timesteps(dim, flip_sin_to_cos=True, downscale_freq_shift=0) -> [cos | sin], fp32 math
__global__ void timestep_embedding_kernel(const float* __restrict__ t, __half* __restrict__ out, int M, int dim) {
    const int half_dim = dim >> 1;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= M * half_dim) return;
    const int m = idx / half_dim, j = idx - m * half_dim;
    const float freq = expf(-9.210340371976184f * static_cast<float>(j) / static_cast<float>(half_dim));
    const float arg = t[m] * freq;
    out[m * dim + j] = __float2half_rn(cosf(arg));
    out[m * dim + half_dim + j] = __float2half_rn(sinf(arg));
}

// =============================================================================================
// CFG combine + Euler v-prediction step + next model input
// =============================================================================================
__global__ void __launch_bounds__(256)
cfg_euler_kernel(const __half* __restrict__ noise, __half* __restrict__ latents, const __half* __restrict__ img_lat,
                 __half* __restrict__ next_in, int T, int HW, float g_min, float g_max, float sigma,
                 float sigma_next, const float* __restrict__ sigmas_dev) {
    if (sigmas_dev) {  // CUDA-graph replay: the step's (sigma, sigma_next) live in device memory, not in the launch
        sigma = sigmas_dev[0];
        sigma_next = sigmas_dev[1];
    }
    const long long total = static_cast<long long>(T) * HW;
    const float in_scale = rsqrtf(sigma_next * sigma_next + 1.0f);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / HW);
        const int p = static_cast<int>(idx - static_cast<long long>(t) * HW);
        float x[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) x[c] = __half2float(latents[(static_cast<long long>(t) * 4 + c) * HW + p]);
        if (noise) {
            float g = (T > 1) ? g_min + (g_max - g_min) * static_cast<float>(t) / static_cast<float>(T - 1) : g_min;
            g = __half2float(__float2half_rn(g));  // guidance_scale.to(latents.dtype), pipeline.py:424
            const __half2* nu = reinterpret_cast<const __half2*>(noise + (static_cast<long long>(t) * HW + p) * 4);
            const __half2* nc = reinterpret_cast<const __half2*>(noise + (static_cast<long long>(T + t) * HW + p) * 4);
            const float2 u01 = __half22float2(nu[0]), u23 = __half22float2(nu[1]);
            const float2 c01 = __half22float2(nc[0]), c23 = __half22float2(nc[1]);
            const float u[4] = {u01.x, u01.y, u23.x, u23.y};
            const float cd[4] = {c01.x, c01.y, c23.x, c23.y};
            const float s2 = sigma * sigma + 1.0f;
            const float c_out = -sigma * rsqrtf(s2);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float v = u[c] + g * (cd[c] - u[c]);
                const float x0 = v * c_out + x[c] / s2;
                const float d = (x[c] - x0) / sigma;
                const float xn = x[c] + d * (sigma_next - sigma);
                const __half xh = __float2half_rn(xn);
                latents[(static_cast<long long>(t) * 4 + c) * HW + p] = xh;
                x[c] = __half2float(xh);
            }
        }
#pragma unroll
        for (int b = 0; b < 2; ++b) {
            V8 o;
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                o.h[c] = __float2half_rn(x[c] * in_scale);
                o.h[4 + c] = img_lat[(static_cast<long long>(b) * 4 + c) * HW + p];
            }
            *reinterpret_cast<uint4*>(next_in + ((static_cast<long long>(b) * T + t) * HW + p) * 8) = o.u;
        }
    }
}

static int grid_for(long long work_items, int block = 256) {
    long long g = (work_items + block - 1) / block;
    const long long cap = 148LL * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace mofa

using namespace mofa;

extern "C" int mofa_groupnorm(const void* x1, int32_t C1, const void* x2, int32_t C2, const void* gamma,
                              const void* beta, void* out, int64_t rows, int64_t rows_per_stat, int32_t groups,
                              float eps, int32_t silu, float* stats, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const int C = C1 + (x2 ? C2 : 0);
    if (!x2) C2 = 0;
    if (!x1 || !out || !stats || rows <= 0 || rows_per_stat <= 0 || (rows % rows_per_stat) != 0 || groups <= 0 ||
        groups > 64 || (C % groups) != 0 || (C1 % 8) != 0 || (C2 % 8) != 0) {
        set_last_error("mofa_groupnorm: bad arguments (C1=%d C2=%d rows=%lld rows_per_stat=%lld groups=%d)", C1, C2,
                       (long long)rows, (long long)rows_per_stat, groups);
        return MOFA_ERR_ARG;
    }
    const long long nstat = rows / rows_per_stat;
    const bool stats_ready = (silu & 2) != 0;   // accumulated by the producing GEMM's epilogue (mofa_gemm_args.gn_stats)
    silu &= 1;
    if (!stats_ready) cudaMemsetAsync(stats, 0, sizeof(float) * 2 * groups * nstat, stream);
    const int vecs = C / 8;
    const int nv = (vecs + 255) / 256;
    const int VX = (vecs + nv - 1) / nv;
    int VY = 256 / VX;
    if (VY < 1) VY = 1;
    dim3 block(VX, VY);
    // slabs: a few unrolled iterations per thread; the grid is persistent (blocks-per-SM x 148) so there is no partial wave
    auto plan = [&](int rows_per_thread, int blocks_per_sm, int& rpb, int& slabs_per_stat, long long& n_slabs, int& grid) {
        long long want = static_cast<long long>(VY) * rows_per_thread;
        if (want > rows_per_stat) want = rows_per_stat;
        rpb = static_cast<int>(want < 1 ? 1 : want);
        slabs_per_stat = static_cast<int>((rows_per_stat + rpb - 1) / rpb);
        n_slabs = static_cast<long long>(slabs_per_stat) * nstat;
        const long long cap = 148LL * blocks_per_sm;
        grid = static_cast<int>(n_slabs < cap ? n_slabs : cap);
    };
    int rpb, sps, grid;
    long long n_slabs;
    if (!stats_ready) {
        plan(16, 3, rpb, sps, n_slabs, grid);
        groupnorm_stats_kernel<<<grid, block, 0, stream>>>(static_cast<const __half*>(x1), C1,
                                                           static_cast<const __half*>(x2), C2, rows_per_stat, rpb, sps,
                                                           n_slabs, groups, stats);
        int rc = check_launch("mofa_groupnorm(stats)");
        if (rc) return rc;
    }
    // apply pass: 4 loads of 16 bytes in flight per thread, 3 blocks per SM.  Round-2 microbenchmark of (2 loads, 4 blocks),
    // (2, 6), (1, 8) register / occupancy trade-offs: equal or slower (4.17 TB/s at level 0 for the first two, 3.8 / 3.5 for
    // the spilling ones; profiles/r2_groupnorm_variants.txt)
    // (24 rows per thread where that still leaves two waves of slabs: amortises the per-slab scale / shift set-up and gives
    //  the prefetch something to run ahead on)
    static int variant = -1;
    if (variant < 0) {
        const char* e = getenv("MOFA_GN_APPLY");
        variant = e ? atoi(e) : 0;
    }
    const int bps = variant == 0 ? 2 : 3;
    plan(variant == 2 ? 8 : 24, bps, rpb, sps, n_slabs, grid);
    if (n_slabs < 2 * 148LL * bps) plan(8, bps, rpb, sps, n_slabs, grid);
#define GN_APPLY(KU, MB, PF)                                                                                            \
    groupnorm_apply_kernel<KU, MB, PF><<<grid, block, 0, stream>>>(                                                     \
        static_cast<const __half*>(x1), C1, static_cast<const __half*>(x2), C2, static_cast<const __half*>(gamma),      \
        static_cast<const __half*>(beta), static_cast<__half*>(out), rows_per_stat, rpb, sps, n_slabs, groups, eps, silu, \
        stats)
    if (variant == 0) GN_APPLY(4, 2, true);
    else if (variant == 1) GN_APPLY(3, 3, true);
    else GN_APPLY(4, 3, false);
#undef GN_APPLY
    return check_launch("mofa_groupnorm(apply)");
}

extern "C" int mofa_layernorm(const void* x, const void* gamma, const void* beta, void* out, int64_t rows, int32_t C,
                              float eps, const void* add, int64_t rows_per_group, int64_t add_period, void* sum_out,
                              mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || rows <= 0 || C <= 0 || (C % 8) != 0 || C > 2048) {
        set_last_error("mofa_layernorm: bad arguments (C=%d rows=%lld)", C, (long long)rows);
        return MOFA_ERR_ARG;
    }
    if (rows_per_group <= 0) rows_per_group = 1;
    if (add_period <= 0) add_period = 1;
    const int warps = 8;
    const int vecs = C / 8;
#define LN_LAUNCH(LPR, VPL)                                                                                        \
    do {                                                                                                           \
        const long long nwarps = (rows + (32 / LPR) - 1) / (32 / LPR);                                             \
        const unsigned grid = static_cast<unsigned>((nwarps + warps - 1) / warps);                                 \
        layernorm_kernel<LPR, VPL><<<grid, warps * 32, 0, stream>>>(                                               \
            static_cast<const __half*>(x), static_cast<const __half*>(gamma), static_cast<const __half*>(beta),    \
            static_cast<__half*>(out), rows, C, eps, static_cast<const __half*>(add), rows_per_group, add_period,  \
            static_cast<__half*>(sum_out));                                                                        \
    } while (0)
    if (vecs <= 8 * 5)
        LN_LAUNCH(8, 5);
    else if (vecs <= 16 * 5)
        LN_LAUNCH(16, 5);
    else if (vecs <= 32 * 5)
        LN_LAUNCH(32, 5);
    else
        LN_LAUNCH(32, 8);
#undef LN_LAUNCH
    return check_launch("mofa_layernorm");
}

extern "C" int mofa_axpy_bcast(const void* x, const void* y, void* out, int64_t n, int64_t period, float scale,
                               mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !y || !out || n <= 0 || (n % 8) != 0 || period <= 0 || (period % 8) != 0) {
        set_last_error("mofa_axpy_bcast: n and period must be positive multiples of 8");
        return MOFA_ERR_ARG;
    }
    axpy_bcast_kernel<<<grid_for(n / 8), 256, 0, stream>>>(static_cast<const __half*>(x),
                                                          static_cast<const __half*>(y), static_cast<__half*>(out),
                                                          n / 8, period / 8, scale);
    return check_launch("mofa_axpy_bcast");
}

extern "C" int mofa_im2col3x3(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                              int32_t stride, int32_t Kpad, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || (stride != 1 && stride != 2) || Kpad < 9 * C) {
        set_last_error("mofa_im2col3x3: bad arguments");
        return MOFA_ERR_ARG;
    }
    const int Ho = (H + 2 - 3) / stride + 1, Wo = (W + 2 - 3) / stride + 1;
    if ((C % 8) == 0 && Kpad == 9 * C) {
        const long long total = static_cast<long long>(n_img) * Ho * Wo * 9 * (C / 8);
        im2col3x3_vec_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const __half*>(x),
                                                                 static_cast<__half*>(out), n_img, H, W, C, stride, Ho,
                                                                 Wo);
    } else {
        const long long total = static_cast<long long>(n_img) * Ho * Wo * Kpad;
        im2col3x3_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const __half*>(x),
                                                             static_cast<__half*>(out), n_img, H, W, C, stride, Ho, Wo,
                                                             Kpad);
    }
    return check_launch("mofa_im2col3x3");
}

extern "C" int mofa_upsample2x(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                               mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || (C % 8) != 0) {
        set_last_error("mofa_upsample2x: C must be a multiple of 8");
        return MOFA_ERR_ARG;
    }
    const long long total = static_cast<long long>(n_img) * 4 * H * W * (C / 8);
    upsample2x_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(out),
                                                          n_img, H, W, C);
    return check_launch("mofa_upsample2x");
}

extern "C" int mofa_nchw_to_nhwc(const void* x, void* out, int32_t n_img, int32_t C, int32_t HW, int32_t ldo,
                                 int32_t c_off, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long total = static_cast<long long>(n_img) * C * HW;
    nchw_to_nhwc_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(out),
                                                            n_img, C, HW, ldo, c_off);
    return check_launch("mofa_nchw_to_nhwc");
}
extern "C" int mofa_nhwc_to_nchw(const void* x, void* out, int32_t n_img, int32_t C, int32_t HW, int32_t ldi,
                                 int32_t c_off, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    const long long total = static_cast<long long>(n_img) * C * HW;
    nhwc_to_nchw_kernel<<<grid_for(total), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(out),
                                                            n_img, C, HW, ldi, c_off);
    return check_launch("mofa_nhwc_to_nchw");
}

extern "C" int mofa_linear_small(const void* a, const void* w, const void* bias, void* out, int32_t M, int32_t N,
                                 int32_t K, int32_t act_in, int32_t act_out, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!a || !w || !out || M <= 0 || M > 8 || N <= 0 || K <= 0 || (K % 8) != 0) {
        set_last_error("mofa_linear_small: needs 1<=M<=8 and K %% 8 == 0 (M=%d K=%d)", M, K);
        return MOFA_ERR_ARG;
    }
    const int warps = 8;
    linear_small_kernel<<<(N + warps - 1) / warps, warps * 32, 0, stream>>>(
        static_cast<const __half*>(a), static_cast<const __half*>(w), static_cast<const __half*>(bias),
        static_cast<__half*>(out), M, N, K, act_in, act_out);
    return check_launch("mofa_linear_small");
}

extern "C" int mofa_timestep_embedding(const float* t, void* out, int32_t M, int32_t dim, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!t || !out || M <= 0 || dim <= 0 || (dim & 1)) {
        set_last_error("mofa_timestep_embedding: bad arguments");
        return MOFA_ERR_ARG;
    }
    const int total = M * (dim / 2);
    timestep_embedding_kernel<<<(total + 127) / 128, 128, 0, stream>>>(t, static_cast<__half*>(out), M, dim);
    return check_launch("mofa_timestep_embedding");
}

extern "C" int mofa_cfg_euler_step(const void* noise, void* latents_h, const void* image_latents, void* next_in,
                                   int32_t T, int32_t HW, float g_min, float g_max, float sigma, float sigma_next,
                                   mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!latents_h || !image_latents || !next_in || T <= 0 || HW <= 0) {
        set_last_error("mofa_cfg_euler_step: bad arguments");
        return MOFA_ERR_ARG;
    }
    cfg_euler_kernel<<<grid_for(static_cast<long long>(T) * HW), 256, 0, stream>>>(
        static_cast<const __half*>(noise), static_cast<__half*>(latents_h), static_cast<const __half*>(image_latents),
        static_cast<__half*>(next_in), T, HW, g_min, g_max, sigma, sigma_next, nullptr);
    return check_launch("mofa_cfg_euler_step");
}

extern "C" int mofa_cfg_euler_step_dev(const void* noise, void* latents_h, const void* image_latents, void* next_in,
                                       int32_t T, int32_t HW, float g_min, float g_max, const float* sigmas,
                                       mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!noise || !latents_h || !image_latents || !next_in || !sigmas || T <= 0 || HW <= 0) {
        set_last_error("mofa_cfg_euler_step_dev: bad arguments");
        return MOFA_ERR_ARG;
    }
    cfg_euler_kernel<<<grid_for(static_cast<long long>(T) * HW), 256, 0, stream>>>(
        static_cast<const __half*>(noise), static_cast<__half*>(latents_h), static_cast<const __half*>(image_latents),
        static_cast<__half*>(next_in), T, HW, g_min, g_max, 0.f, 0.f, sigmas);
    return check_launch("mofa_cfg_euler_step_dev");
}

// =============================================================================================
// row softmax in place (VAE mid-block attention computed as GEMM -> softmax -> GEMM, head_dim 512)
// =============================================================================================
namespace mofa {

__global__ void __launch_bounds__(128)
softmax_rows_kernel(__half* __restrict__ x, long long rows, int L, long long ld) {
    const long long row = blockIdx.x;
    if (row >= rows) return;
    __half* xr = x + row * ld;
    const int vecs = L >> 3;
    float v[16][8];
    float mx = -INFINITY;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int vec = threadIdx.x + i * 128;
        if (vec < vecs) {
            V8 t;
            t.u = *reinterpret_cast<const uint4*>(xr + vec * 8);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[i][j] = __half2float(t.h[j]);
                mx = fmaxf(mx, v[i][j]);
            }
        }
    }
    __shared__ float red[4];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mx;
    __syncthreads();
    mx = fmaxf(fmaxf(red[0], red[1]), fmaxf(red[2], red[3]));
    __syncthreads();
    float sum = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int vec = threadIdx.x + i * 128;
        if (vec < vecs) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                v[i][j] = __expf(v[i][j] - mx);
                sum += v[i][j];
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = sum;
    __syncthreads();
    const float inv = 1.0f / (red[0] + red[1] + red[2] + red[3]);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
        const int vec = threadIdx.x + i * 128;
        if (vec < vecs) {
            V8 o;
#pragma unroll
            for (int j = 0; j < 8; ++j) o.h[j] = __float2half_rn(v[i][j] * inv);
            *reinterpret_cast<uint4*>(xr + vec * 8) = o.u;
        }
    }
}

// TemporalDecoder tail: time_conv_out = Conv3d(3, 3, (3,1,1), padding (1,0,0)) over the frames of a chunk, then the
// reference's post-processing (x/2+0.5).clamp(0,1)*255 -> uint8 (optional) -- SURVEY.md §8f row 3 "frame epilogue".
__global__ void __launch_bounds__(256)
vae_time_conv_out_kernel(const __half* __restrict__ y, const float* __restrict__ w, const float* __restrict__ b,
                         float* __restrict__ out_f32, uint8_t* __restrict__ out_u8, int T, long long HW) {
    const long long total = static_cast<long long>(T) * HW;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / HW);
        const long long p = idx - static_cast<long long>(t) * HW;
        float acc[3] = {b[0], b[1], b[2]};
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            const int tt = t + dt - 1;
            if (tt < 0 || tt >= T) continue;
            const __half* src = y + (static_cast<long long>(tt) * HW + p) * 3;
            const float x0 = __half2float(src[0]), x1 = __half2float(src[1]), x2 = __half2float(src[2]);
#pragma unroll
            for (int co = 0; co < 3; ++co)  // w[co][ci][dt]
                acc[co] += w[(co * 3 + 0) * 3 + dt] * x0 + w[(co * 3 + 1) * 3 + dt] * x1 + w[(co * 3 + 2) * 3 + dt] * x2;
        }
        if (out_f32) {
#pragma unroll
            for (int co = 0; co < 3; ++co) out_f32[(static_cast<long long>(t) * 3 + co) * HW + p] = acc[co];
        }
        if (out_u8) {
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                float u = fminf(fmaxf(acc[co] * 0.5f + 0.5f, 0.f), 1.f) * 255.f;
                out_u8[(static_cast<long long>(t) * HW + p) * 3 + co] = static_cast<uint8_t>(__float2int_rn(u));
            }
        }
    }
}

// uint8-only tail, four pixels per thread: 24-byte fp16 reads per temporal tap, ONE 12-byte run of output per thread as three
// aligned 32-bit stores -- the form that also travels well when `out_u8` is another GPU's memory (NVLink peer stores into the
// gather buffer of rank 0, parallel.PeerFrameGather: the frame epilogue fused with the path's only collective).
__global__ void __launch_bounds__(256)
vae_time_conv_out_u8x4_kernel(const __half* __restrict__ y, const float* __restrict__ w, const float* __restrict__ b,
                              uint32_t* __restrict__ out_u8, int T, long long HW) {
    const long long quads = HW >> 2;
    const long long total = static_cast<long long>(T) * quads;
    float wr[27];
#pragma unroll
    for (int i = 0; i < 27; ++i) wr[i] = w[i];
    const float b0 = b[0], b1 = b[1], b2 = b[2];
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int t = static_cast<int>(idx / quads);
        const long long p = (idx - static_cast<long long>(t) * quads) << 2;
        float acc[4][3];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            acc[q][0] = b0;
            acc[q][1] = b1;
            acc[q][2] = b2;
        }
#pragma unroll
        for (int dt = 0; dt < 3; ++dt) {
            const int tt = t + dt - 1;
            if (tt < 0 || tt >= T) continue;
            // 12 fp16 = 24 bytes, 8-byte aligned (p % 4 == 0): three 64-bit loads
            const uint2* src = reinterpret_cast<const uint2*>(y + (static_cast<long long>(tt) * HW + p) * 3);
            uint2 r[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) r[k] = __ldg(src + k);
            const __half* h = reinterpret_cast<const __half*>(r);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float x0 = __half2float(h[q * 3 + 0]), x1 = __half2float(h[q * 3 + 1]),
                            x2 = __half2float(h[q * 3 + 2]);
#pragma unroll
                for (int co = 0; co < 3; ++co)
                    acc[q][co] += wr[(co * 3 + 0) * 3 + dt] * x0 + wr[(co * 3 + 1) * 3 + dt] * x1 +
                                  wr[(co * 3 + 2) * 3 + dt] * x2;
            }
        }
        uint32_t o[3] = {0u, 0u, 0u};
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int co = 0; co < 3; ++co) {
                const float u = fminf(fmaxf(acc[q][co] * 0.5f + 0.5f, 0.f), 1.f) * 255.f;
                const int byte = q * 3 + co;
                o[byte >> 2] |= static_cast<uint32_t>(__float2int_rn(u)) << ((byte & 3) * 8);
            }
        uint32_t* dst = out_u8 + ((static_cast<long long>(t) * HW + p) * 3 >> 2);
        dst[0] = o[0];
        dst[1] = o[1];
        dst[2] = o[2];
    }
}

// ---------------------------------------------------------------------------------------------
// cross-GPU flags for the peer-store gather: publish (after this GPU's stores) / wait (before reading them)
// ---------------------------------------------------------------------------------------------
__global__ void peer_signal_kernel(uint32_t* flag, uint32_t value) {
    __threadfence_system();   // every store this GPU issued before this kernel is visible system-wide first
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
}

__global__ void peer_wait_kernel(const uint32_t* flags, int n, uint32_t value, unsigned long long timeout_ns,
                                 uint32_t* timed_out) {
    const int i = threadIdx.x;
    if (i >= n) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        uint32_t v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        if (static_cast<int32_t>(v - value) >= 0) break;            // epochs only grow
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {                                 // a dead peer must not hang this GPU
            if (timed_out) *timed_out = 1u + static_cast<uint32_t>(i);
            break;
        }
        __nanosleep(200);
    }
}

}  // namespace mofa

extern "C" int mofa_peer_enable(int32_t peer_device) {
    int cur = 0;
    cudaGetDevice(&cur);
    if (cur == peer_device) return MOFA_OK;
    int can = 0;
    cudaDeviceCanAccessPeer(&can, cur, peer_device);
    if (!can) {
        set_last_error("mofa_peer_enable: device %d cannot access device %d (no NVLink / P2P path)", cur, peer_device);
        return MOFA_ERR_CUDA;
    }
    cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
    if (e == cudaErrorPeerAccessAlreadyEnabled) {
        cudaGetLastError();
        return MOFA_OK;
    }
    if (e != cudaSuccess) {
        set_last_error("mofa_peer_enable: %s", cudaGetErrorString(e));
        return MOFA_ERR_CUDA;
    }
    return MOFA_OK;
}

extern "C" int mofa_peer_signal(void* flag, uint32_t value, mofa_stream_t stream_) {
    if (!flag) {
        set_last_error("mofa_peer_signal: null flag");
        return MOFA_ERR_ARG;
    }
    mofa::peer_signal_kernel<<<1, 1, 0, static_cast<cudaStream_t>(stream_)>>>(static_cast<uint32_t*>(flag), value);
    return check_launch("mofa_peer_signal");
}

extern "C" int mofa_peer_wait(const void* flags, int32_t n, uint32_t value, double timeout_s, void* timed_out,
                              mofa_stream_t stream_) {
    if (!flags || n <= 0 || n > 1024) {
        set_last_error("mofa_peer_wait: bad arguments");
        return MOFA_ERR_ARG;
    }
    const unsigned long long ns = static_cast<unsigned long long>((timeout_s > 0 ? timeout_s : 30.0) * 1e9);
    mofa::peer_wait_kernel<<<1, ((n + 31) / 32) * 32, 0, static_cast<cudaStream_t>(stream_)>>>(
        static_cast<const uint32_t*>(flags), n, value, ns, static_cast<uint32_t*>(timed_out));
    return check_launch("mofa_peer_wait");
}

extern "C" int mofa_softmax_rows(void* x, int64_t rows, int32_t L, int64_t ld, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || rows <= 0 || L <= 0 || (L % 8) != 0 || L > 16384 || (ld % 8) != 0) {
        set_last_error("mofa_softmax_rows: needs L %% 8 == 0 and L <= 16384 (L=%d)", L);
        return MOFA_ERR_ARG;
    }
    mofa::softmax_rows_kernel<<<static_cast<unsigned>(rows), 128, 0, stream>>>(static_cast<__half*>(x), rows, L, ld);
    return check_launch("mofa_softmax_rows");
}

extern "C" int mofa_vae_time_conv_out(const void* y, const float* w, const float* b, float* out_f32, void* out_u8,
                                      int32_t T, int64_t HW, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!y || !w || !b || (!out_f32 && !out_u8) || T <= 0 || HW <= 0) {
        set_last_error("mofa_vae_time_conv_out: bad arguments");
        return MOFA_ERR_ARG;
    }
    long long blocks = (static_cast<long long>(T) * HW + 255) / 256;
    if (blocks > 148LL * 16) blocks = 148LL * 16;
    if (out_u8 && !out_f32 && (HW % 4) == 0 && (reinterpret_cast<uintptr_t>(out_u8) % 4) == 0) {
        long long b4 = (static_cast<long long>(T) * (HW / 4) + 255) / 256;
        if (b4 > 148LL * 16) b4 = 148LL * 16;
        mofa::vae_time_conv_out_u8x4_kernel<<<static_cast<unsigned>(b4), 256, 0, stream>>>(
            static_cast<const __half*>(y), w, b, static_cast<uint32_t*>(out_u8), T, HW);
        return check_launch("mofa_vae_time_conv_out");
    }
    mofa::vae_time_conv_out_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const __half*>(y), w, b, out_f32, static_cast<uint8_t*>(out_u8), T, HW);
    return check_launch("mofa_vae_time_conv_out");
}
