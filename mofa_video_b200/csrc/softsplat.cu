// Forward softmax-splatting ('avg' mode) of a first-frame feature map by all flows of a clip.
//
// Restates the arithmetic of the reference's CuPy kernel `softsplat_out`
// (/root/reference/MOFA-Video-Traj/models/softsplat.py:285-335) and of softsplat(..., 'avg')
// (:240-241 ones channel, :253-270 divide by (sum of weights + 1e-7)), with the adapter's flow pyramid
// (nearest 1/s downsample then /s in fp16,
//  /root/reference/MOFA-Video-Traj/models/svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py:302-309)
// fused into the flow fetch.  B200 design: channels-last so one warp owns one source pixel and its lanes
// stream the channel vector with 16-byte vector atomics (red.global.add.v4.f32); all F flows in one
// launch; the ones-channel becomes a separate [F, hs, ws] weight plane; normalise+fp16 cast is a second
// streaming pass.  Hoisted out of the denoise loop by the engine (the inputs are loop-invariant).
#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

__global__ void __launch_bounds__(256)
softsplat_scatter_kernel(const __half* __restrict__ feat, const __half* __restrict__ flow, float* __restrict__ acc,
                         float* __restrict__ wsum, int F, int hs, int ws, int C, int Hf, int Wf, int s) {
    const int lane = threadIdx.x & 31;
    const long long warp_global = (blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x) >> 5;
    const long long warps_total = (static_cast<long long>(gridDim.x) * blockDim.x) >> 5;
    const long long items = static_cast<long long>(F) * hs * ws;
    const float inv_s = 1.0f / static_cast<float>(s);
    for (long long it = warp_global; it < items; it += warps_total) {
        const int x = static_cast<int>(it % ws);
        const long long t = it / ws;
        const int y = static_cast<int>(t % hs);
        const int f = static_cast<int>(t / hs);
        // nearest-downsampled flow, divided by the scale in fp16 like the reference's fp16 tensor op
        const long long fo = (static_cast<long long>(f) * 2) * Hf * Wf + static_cast<long long>(y * s) * Wf + x * s;
        const float fx = __half2float(__float2half_rn(__half2float(flow[fo]) * inv_s));
        const float fy =
            __half2float(__float2half_rn(__half2float(flow[fo + static_cast<long long>(Hf) * Wf]) * inv_s));
        const float ox = static_cast<float>(x) + fx;
        const float oy = static_cast<float>(y) + fy;
        if (!isfinite(ox) || !isfinite(oy)) continue;  // softsplat.py:301-302
        const int nwx = static_cast<int>(floorf(ox));
        const int nwy = static_cast<int>(floorf(oy));
        const int sex = nwx + 1, sey = nwy + 1;
        // softsplat.py:315-318
        const float w_nw = (static_cast<float>(sex) - ox) * (static_cast<float>(sey) - oy);
        const float w_ne = (ox - static_cast<float>(nwx)) * (static_cast<float>(sey) - oy);
        const float w_sw = (static_cast<float>(sex) - ox) * (oy - static_cast<float>(nwy));
        const float w_se = (ox - static_cast<float>(nwx)) * (oy - static_cast<float>(nwy));
        const bool in_w = nwx >= 0 && nwx < ws, in_e = sex >= 0 && sex < ws;
        const bool in_n = nwy >= 0 && nwy < hs, in_s = sey >= 0 && sey < hs;
        const long long plane = static_cast<long long>(f) * hs * ws;
        const long long o_nw = plane + static_cast<long long>(nwy) * ws + nwx;
        const long long o_ne = o_nw + 1, o_sw = o_nw + ws, o_se = o_nw + ws + 1;
        if (lane == 0) {
            if (in_n && in_w) atomicAdd(wsum + o_nw, w_nw);
            if (in_n && in_e) atomicAdd(wsum + o_ne, w_ne);
            if (in_s && in_w) atomicAdd(wsum + o_sw, w_sw);
            if (in_s && in_e) atomicAdd(wsum + o_se, w_se);
        }
        const __half* src = feat + (static_cast<long long>(y) * ws + x) * C;
        for (int c = lane * 4; c < C; c += 128) {
            const uint2 raw = __ldg(reinterpret_cast<const uint2*>(src + c));
            const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&raw.x));
            const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&raw.y));
            if (in_n && in_w)
                atomicAdd(reinterpret_cast<float4*>(acc + o_nw * C + c),
                          make_float4(a.x * w_nw, a.y * w_nw, b.x * w_nw, b.y * w_nw));
            if (in_n && in_e)
                atomicAdd(reinterpret_cast<float4*>(acc + o_ne * C + c),
                          make_float4(a.x * w_ne, a.y * w_ne, b.x * w_ne, b.y * w_ne));
            if (in_s && in_w)
                atomicAdd(reinterpret_cast<float4*>(acc + o_sw * C + c),
                          make_float4(a.x * w_sw, a.y * w_sw, b.x * w_sw, b.y * w_sw));
            if (in_s && in_e)
                atomicAdd(reinterpret_cast<float4*>(acc + o_se * C + c),
                          make_float4(a.x * w_se, a.y * w_se, b.x * w_se, b.y * w_se));
        }
    }
}

__global__ void __launch_bounds__(256)
softsplat_normalize_kernel(const float* __restrict__ acc, const float* __restrict__ wsum, __half* __restrict__ out,
                           long long pixels, int C) {
    const int cv = C >> 2;
    const long long total = pixels * cv;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long px = idx / cv;
        const float inv = 1.0f / (wsum[px] + 0.0000001f);  // softsplat.py:257
        const float4 a = __ldg(reinterpret_cast<const float4*>(acc) + idx);
        __half2 lo = __floats2half2_rn(a.x * inv, a.y * inv);
        __half2 hi = __floats2half2_rn(a.z * inv, a.w * inv);
        uint2 o;
        o.x = *reinterpret_cast<uint32_t*>(&lo);
        o.y = *reinterpret_cast<uint32_t*>(&hi);
        reinterpret_cast<uint2*>(out)[idx] = o;
    }
}

}  // namespace mofa

using namespace mofa;

extern "C" int mofa_softsplat_avg(const void* feat, const void* flow, float* acc, float* wsum, void* out, int32_t F,
                                  int32_t hs, int32_t ws, int32_t C, int32_t Hf, int32_t Wf, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!feat || !flow || !acc || !wsum || !out || F <= 0 || hs <= 0 || ws <= 0 || C <= 0 || (C % 4) != 0 ||
        Hf % hs != 0 || Wf % ws != 0 || Hf / hs != Wf / ws) {
        set_last_error("mofa_softsplat_avg: bad arguments (C=%d hs=%d ws=%d Hf=%d Wf=%d)", C, hs, ws, Hf, Wf);
        return MOFA_ERR_ARG;
    }
    const long long pixels = static_cast<long long>(F) * hs * ws;
    cudaMemsetAsync(acc, 0, sizeof(float) * pixels * C, stream);
    cudaMemsetAsync(wsum, 0, sizeof(float) * pixels, stream);
    long long blocks = (pixels + 7) / 8;  // 8 warps per block, one source pixel per warp per trip
    if (blocks > 148LL * 32) blocks = 148LL * 32;
    softsplat_scatter_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const __half*>(feat), static_cast<const __half*>(flow), acc, wsum, F, hs, ws, C, Hf, Wf, Hf / hs);
    int rc = check_launch("mofa_softsplat_avg(scatter)");
    if (rc) return rc;
    long long nblocks = (pixels * (C / 4) + 255) / 256;
    if (nblocks > 148LL * 32) nblocks = 148LL * 32;
    softsplat_normalize_kernel<<<static_cast<unsigned>(nblocks), 256, 0, stream>>>(acc, wsum,
                                                                                   static_cast<__half*>(out), pixels, C);
    return check_launch("mofa_softsplat_avg(normalize)");
}
