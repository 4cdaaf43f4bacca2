// Streaming kernels used by the CMP sparse-to-dense flow network and by the
// Keypoint adapter's occlusion nets: general im2col, max/avg pooling, align_corners bilinear resize, the
// Fuser softmax-expectation head, and a broadcast copy into a channel slice (concat assembly).
// All tensors fp16 channels-last [n, H, W, C] (rows = pixels).  HBM-bound, once per clip.
#include "../../include/mofa_b200.h"
#include "common.cuh"

namespace mofa {

// out[(n,oy,ox), (ky,kx,c)] = x[n, oy*s + ky*d - pad, ox*s + kx*d - pad, c]; zero outside; K padded to Kpad
__global__ void __launch_bounds__(256)
im2col_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C, int ks, int stride,
              int pad, int dil, int Ho, int Wo, int Kpad) {
    const long long total = static_cast<long long>(n_img) * Ho * Wo * Kpad;
    const int K = ks * ks * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = idx / Kpad;
        const int k = static_cast<int>(idx - row * Kpad);
        __half v = __float2half(0.f);
        if (k < K) {
            const int tap = k / C;
            const int c = k - tap * C;
            const int ky = tap / ks, kx = tap - ky * ks;
            const int ox = static_cast<int>(row % Wo);
            const long long t = row / Wo;
            const int oy = static_cast<int>(t % Ho);
            const int n = static_cast<int>(t / Ho);
            const int iy = oy * stride + ky * dil - pad;
            const int ix = ox * stride + kx * dil - pad;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = x[((static_cast<long long>(n) * H + iy) * W + ix) * C + c];
        }
        out[idx] = v;
    }
}

// mode 0 = max (padding ignored, like nn.MaxPool2d), 1 = average over the full window (nn.AvgPool2d, pad = 0)
__global__ void __launch_bounds__(256)
pool2d_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C, int ks, int stride,
              int pad, int Ho, int Wo, int mode) {
    const long long total = static_cast<long long>(n_img) * Ho * Wo * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        long long t = idx / C;
        const int ox = static_cast<int>(t % Wo);
        t /= Wo;
        const int oy = static_cast<int>(t % Ho);
        const int n = static_cast<int>(t / Ho);
        float acc = mode == 0 ? -INFINITY : 0.f;
        for (int ky = 0; ky < ks; ++ky) {
            const int iy = oy * stride + ky - pad;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < ks; ++kx) {
                const int ix = ox * stride + kx - pad;
                if (ix < 0 || ix >= W) continue;
                const float v = __half2float(x[((static_cast<long long>(n) * H + iy) * W + ix) * C + c]);
                acc = mode == 0 ? fmaxf(acc, v) : acc + v;
            }
        }
        if (mode == 1) acc /= static_cast<float>(ks * ks);
        out[idx] = __float2half_rn(acc);
    }
}

// F.interpolate(mode="bilinear", align_corners=True); writes channels [c_off, c_off+C) of rows of width ldo
__global__ void __launch_bounds__(256)
resize_bilinear_ac_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C, int Ho,
                          int Wo, int ldo, int c_off) {
    const long long total = static_cast<long long>(n_img) * Ho * Wo * C;
    const float sy = Ho > 1 ? static_cast<float>(H - 1) / static_cast<float>(Ho - 1) : 0.f;
    const float sx = Wo > 1 ? static_cast<float>(W - 1) / static_cast<float>(Wo - 1) : 0.f;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx % C);
        long long t = idx / C;
        const int ox = static_cast<int>(t % Wo);
        t /= Wo;
        const int oy = static_cast<int>(t % Ho);
        const int n = static_cast<int>(t / Ho);
        const float fy = oy * sy, fx = ox * sx;
        const int y0 = static_cast<int>(fy), x0 = static_cast<int>(fx);
        const int y1 = y0 + 1 < H ? y0 + 1 : H - 1, x1 = x0 + 1 < W ? x0 + 1 : W - 1;
        const float wy = fy - y0, wx = fx - x0;
        const __half* b = x + static_cast<long long>(n) * H * W * C + c;
        const float v00 = __half2float(b[(static_cast<long long>(y0) * W + x0) * C]);
        const float v01 = __half2float(b[(static_cast<long long>(y0) * W + x1) * C]);
        const float v10 = __half2float(b[(static_cast<long long>(y1) * W + x0) * C]);
        const float v11 = __half2float(b[(static_cast<long long>(y1) * W + x1) * C]);
        const float v = (1.f - wy) * ((1.f - wx) * v00 + wx * v01) + wy * ((1.f - wx) * v10 + wx * v11);
        out[((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c_off + c] = __float2half_rn(v);
    }
}

// Fuser.convert_flow: two softmax distributions over nbins bins -> expected value with bin centres
// (k + 0.5) * 2*fmax/nbins - fmax.  One warp per pixel.  logits [rows, 2*nbins] fp16 -> flow [rows, 2] fp16.
__global__ void __launch_bounds__(256)
cmp_fuser_kernel(const __half* __restrict__ logits, __half* __restrict__ flow, long long rows, int nbins, float fmax) {
    const int lane = threadIdx.x & 31;
    const long long row = blockIdx.x * static_cast<long long>(blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float step = 2.f * fmax / static_cast<float>(nbins);
    for (int comp = 0; comp < 2; ++comp) {
        const __half* l = logits + row * 2 * nbins + comp * nbins;
        float mx = -INFINITY;
        for (int k = lane; k < nbins; k += 32) mx = fmaxf(mx, __half2float(l[k]));
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float se = 0.f, sv = 0.f;
        for (int k = lane; k < nbins; k += 32) {
            const float e = __expf(__half2float(l[k]) - mx);
            se += e;
            sv += e * (static_cast<float>(k) * step - fmax + 0.5f * step);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            se += __shfl_xor_sync(0xffffffffu, se, o);
            sv += __shfl_xor_sync(0xffffffffu, sv, o);
        }
        if (lane == 0) flow[row * 2 + comp] = __float2half_rn(sv / se);
    }
}

// dst[r, c_off + c] = src[r % period_rows, c]  (concat assembly; period_rows < rows broadcasts over frames)
__global__ void __launch_bounds__(256)
copy_cols_kernel(const __half* __restrict__ src, __half* __restrict__ dst, long long rows, int C, long long period_rows,
                 int ldo, int c_off) {
    const long long total = rows * C;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = idx / C;
        const int c = static_cast<int>(idx - r * C);
        dst[r * ldo + c_off + c] = src[(r % period_rows) * C + c];
    }
}

static int grid_for2(long long work_items) {
    long long g = (work_items + 255) / 256;
    const long long cap = 148LL * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace mofa

using namespace mofa;

extern "C" int mofa_im2col(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t ksize,
                           int32_t stride, int32_t pad, int32_t dilation, int32_t Kpad, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || n_img <= 0 || H <= 0 || W <= 0 || C <= 0 || ksize <= 0 || stride <= 0 || dilation <= 0 ||
        Kpad < ksize * ksize * C) {
        set_last_error("mofa_im2col: bad arguments");
        return MOFA_ERR_ARG;
    }
    // pad < 0: asymmetric padding (0 before, -pad after) -- F.pad(x, (0, 1, 0, 1)) + stride-2 conv of the VAE encoder
    const int pad_lo = pad < 0 ? 0 : pad;
    const int pad_tot = pad < 0 ? -pad : 2 * pad;
    const int Ho = (H + pad_tot - dilation * (ksize - 1) - 1) / stride + 1;
    const int Wo = (W + pad_tot - dilation * (ksize - 1) - 1) / stride + 1;
    const long long total = static_cast<long long>(n_img) * Ho * Wo * Kpad;
    im2col_kernel<<<grid_for2(total), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(out), n_img,
                                                       H, W, C, ksize, stride, pad_lo, dilation, Ho, Wo, Kpad);
    return check_launch("mofa_im2col");
}

extern "C" int mofa_pool2d(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C, int32_t ksize,
                           int32_t stride, int32_t pad, int32_t mode, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || ksize <= 0 || stride <= 0 || (mode != 0 && mode != 1)) {
        set_last_error("mofa_pool2d: bad arguments");
        return MOFA_ERR_ARG;
    }
    const int Ho = (H + 2 * pad - ksize) / stride + 1, Wo = (W + 2 * pad - ksize) / stride + 1;
    const long long total = static_cast<long long>(n_img) * Ho * Wo * C;
    pool2d_kernel<<<grid_for2(total), 256, 0, stream>>>(static_cast<const __half*>(x), static_cast<__half*>(out), n_img,
                                                       H, W, C, ksize, stride, pad, Ho, Wo, mode);
    return check_launch("mofa_pool2d");
}

extern "C" int mofa_resize_bilinear_ac(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                       int32_t Ho, int32_t Wo, int32_t ldo, int32_t c_off, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || Ho <= 0 || Wo <= 0 || ldo < c_off + C) {
        set_last_error("mofa_resize_bilinear_ac: bad arguments");
        return MOFA_ERR_ARG;
    }
    const long long total = static_cast<long long>(n_img) * Ho * Wo * C;
    resize_bilinear_ac_kernel<<<grid_for2(total), 256, 0, stream>>>(
        static_cast<const __half*>(x), static_cast<__half*>(out), n_img, H, W, C, Ho, Wo, ldo, c_off);
    return check_launch("mofa_resize_bilinear_ac");
}

extern "C" int mofa_cmp_fuser(const void* logits, void* flow, int64_t rows, int32_t nbins, float fmax,
                              mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!logits || !flow || rows <= 0 || nbins <= 0) {
        set_last_error("mofa_cmp_fuser: bad arguments");
        return MOFA_ERR_ARG;
    }
    cmp_fuser_kernel<<<static_cast<unsigned>((rows + 7) / 8), 256, 0, stream>>>(
        static_cast<const __half*>(logits), static_cast<__half*>(flow), rows, nbins, fmax);
    return check_launch("mofa_cmp_fuser");
}

extern "C" int mofa_copy_cols(const void* src, void* dst, int64_t rows, int32_t C, int64_t period_rows, int32_t ldo,
                              int32_t c_off, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!src || !dst || rows <= 0 || C <= 0 || period_rows <= 0 || ldo < c_off + C) {
        set_last_error("mofa_copy_cols: bad arguments");
        return MOFA_ERR_ARG;
    }
    copy_cols_kernel<<<grid_for2(rows * C), 256, 0, stream>>>(static_cast<const __half*>(src), static_cast<__half*>(dst),
                                                             rows, C, period_rows, ldo, c_off);
    return check_launch("mofa_copy_cols");
}

// ---------------------------------------------------------------------------------------------
// Keypoint / Hybrid adapter helpers
// ---------------------------------------------------------------------------------------------
namespace mofa {

// adapter flow pyramid materialised (K/models/ldmk_ctrlnet.py:409-417): out[(f,y,x), c_off + c] = half(flow[f,c,y*s,x*s] / s)
__global__ void __launch_bounds__(256)
flow_pyramid_kernel(const __half* __restrict__ flow, __half* __restrict__ out, int F, int hs, int ws, int Hf, int Wf,
                    int s, int ldo, int c_off) {
    const long long total = static_cast<long long>(F) * hs * ws * 2;
    const float inv = 1.0f / static_cast<float>(s);
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int c = static_cast<int>(idx & 1);
        long long t = idx >> 1;
        const int x = static_cast<int>(t % ws);
        t /= ws;
        const int y = static_cast<int>(t % hs);
        const int f = static_cast<int>(t / hs);
        const float v = __half2float(flow[((static_cast<long long>(f) * 2 + c) * Hf + y * s) * Wf + x * s]) * inv;
        out[((static_cast<long long>(f) * hs + y) * ws + x) * ldo + c_off + c] = __float2half_rn(v);
    }
}

// out[r, c] = a[r, c] * m + b[r, c] * (1 - m),  m = mask[r % period_rows]   (matting blend / hybrid residual blend)
__global__ void __launch_bounds__(256)
mask_blend_kernel(const __half* __restrict__ a, const __half* __restrict__ b, const __half* __restrict__ mask,
                  __half* __restrict__ out, long long rows, int C, long long period_rows) {
    const int cv = C >> 3;
    const long long total = rows * cv;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long r = idx / cv;
        const float m = __half2float(mask[r % period_rows]);
        union { uint4 u; __half h[8]; } va, vb, vo;
        va.u = __ldg(reinterpret_cast<const uint4*>(a) + idx);
        vb.u = __ldg(reinterpret_cast<const uint4*>(b) + idx);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            vo.h[j] = __float2half_rn(__half2float(va.h[j]) * m + __half2float(vb.h[j]) * (1.0f - m));
        reinterpret_cast<uint4*>(out)[idx] = vo.u;
    }
}

// F.interpolate(scale_factor=1/s) nearest on channels-last data: out[n,y,x,:] = x[n, y*s, x*s, :]
__global__ void __launch_bounds__(256)
downsample_nearest_kernel(const __half* __restrict__ x, __half* __restrict__ out, int n_img, int H, int W, int C, int s) {
    const int Ho = H / s, Wo = W / s, cv = C >> 3;
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
            reinterpret_cast<const uint4*>(x + ((static_cast<long long>(n) * H + oy * s) * W + ox * s) * C) + c8);
    }
}

}  // namespace mofa

extern "C" int mofa_flow_pyramid(const void* flow, void* out, int32_t F, int32_t hs, int32_t ws, int32_t Hf,
                                 int32_t Wf, int32_t ldo, int32_t c_off, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!flow || !out || F <= 0 || hs <= 0 || ws <= 0 || Hf % hs != 0 || Wf % ws != 0 || Hf / hs != Wf / ws ||
        ldo < c_off + 2) {
        set_last_error("mofa_flow_pyramid: bad arguments");
        return MOFA_ERR_ARG;
    }
    flow_pyramid_kernel<<<grid_for2(static_cast<long long>(F) * hs * ws * 2), 256, 0, stream>>>(
        static_cast<const __half*>(flow), static_cast<__half*>(out), F, hs, ws, Hf, Wf, Hf / hs, ldo, c_off);
    return check_launch("mofa_flow_pyramid");
}

extern "C" int mofa_mask_blend(const void* a, const void* b, const void* mask, void* out, int64_t rows, int32_t C,
                               int64_t period_rows, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!a || !b || !mask || !out || rows <= 0 || C <= 0 || (C % 8) != 0 || period_rows <= 0) {
        set_last_error("mofa_mask_blend: bad arguments (C must be a multiple of 8)");
        return MOFA_ERR_ARG;
    }
    mask_blend_kernel<<<grid_for2(rows * (C / 8)), 256, 0, stream>>>(
        static_cast<const __half*>(a), static_cast<const __half*>(b), static_cast<const __half*>(mask),
        static_cast<__half*>(out), rows, C, period_rows);
    return check_launch("mofa_mask_blend");
}

namespace mofa {
// CLIP-side resize of /root/reference/MOFA-Video-Traj/pipeline/pipeline.py:532-640 (_resize_with_antialiasing) fused
// into one pass: separable Gaussian blur (sigma = max((factor-1)/2, 0.001) per axis, odd two-sigma kernel, reflect
// padding) followed by bicubic interpolation (A = -0.75, align_corners = True, clamped taps), fp32 NCHW.
__global__ void __launch_bounds__(256)
resize_antialias_kernel(const float* __restrict__ img, float* __restrict__ out, int planes, int H, int W, int Ho, int Wo,
                        float sig_y, float sig_x, int ks_y, int ks_x) {
    __shared__ float gy[64], gx[64];
    if (threadIdx.x < 64) {
        for (int pass = 0; pass < 2; ++pass) {
            const int ks = pass ? ks_x : ks_y;
            const float sg = pass ? sig_x : sig_y;
            float v = 0.f;
            if (static_cast<int>(threadIdx.x) < ks) {
                const float t = static_cast<float>(static_cast<int>(threadIdx.x) - ks / 2);
                v = expf(-(t * t) / (2.0f * sg * sg));
            }
            (pass ? gx : gy)[threadIdx.x] = v;
        }
    }
    __syncthreads();
    float sy = 0.f, sx = 0.f;
    for (int i = 0; i < ks_y; ++i) sy += gy[i];
    for (int i = 0; i < ks_x; ++i) sx += gx[i];
    const float ny = 1.0f / sy, nx = 1.0f / sx;
    const float ry = Ho > 1 ? static_cast<float>(H - 1) / (Ho - 1) : 0.f;
    const float rx = Wo > 1 ? static_cast<float>(W - 1) / (Wo - 1) : 0.f;
    const int py = (ks_y - 1) / 2, px = (ks_x - 1) / 2;
    const long long total = static_cast<long long>(planes) * Ho * Wo;
    for (long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; idx < total;
         idx += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int ox = static_cast<int>(idx % Wo);
        const int oy = static_cast<int>((idx / Wo) % Ho);
        const int pl = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
        const float* src = img + static_cast<long long>(pl) * H * W;
        const float fy = oy * ry, fx = ox * rx;
        const int y0 = static_cast<int>(floorf(fy)), x0 = static_cast<int>(floorf(fx));
        const float ty = fy - y0, tx = fx - x0;
        float wy[4], wx[4];
        {
            const float A = -0.75f;
            auto c1 = [&](float x) { return ((A + 2.f) * x - (A + 3.f)) * x * x + 1.f; };          // |x| <= 1
            auto c2 = [&](float x) { return ((A * x - 5.f * A) * x + 8.f * A) * x - 4.f * A; };    // 1 < |x| < 2
            wy[0] = c2(ty + 1.f); wy[1] = c1(ty); wy[2] = c1(1.f - ty); wy[3] = c2(2.f - ty);
            wx[0] = c2(tx + 1.f); wx[1] = c1(tx); wx[2] = c1(1.f - tx); wx[3] = c2(2.f - tx);
        }
        float acc = 0.f;
        for (int a = 0; a < 4; ++a) {
            int yy = y0 - 1 + a;
            yy = yy < 0 ? 0 : (yy > H - 1 ? H - 1 : yy);     // bicubic taps are clamped to the (blurred) image
            float row = 0.f;
            for (int b = 0; b < 4; ++b) {
                int xx = x0 - 1 + b;
                xx = xx < 0 ? 0 : (xx > W - 1 ? W - 1 : xx);
                float blur = 0.f;                             // blurred pixel (yy, xx): reflect padding
                for (int i = 0; i < ks_y; ++i) {
                    int sy_ = yy + i - py;
                    sy_ = sy_ < 0 ? -sy_ : (sy_ > H - 1 ? 2 * (H - 1) - sy_ : sy_);
                    float line = 0.f;
                    for (int j = 0; j < ks_x; ++j) {
                        int sx_ = xx + j - px;
                        sx_ = sx_ < 0 ? -sx_ : (sx_ > W - 1 ? 2 * (W - 1) - sx_ : sx_);
                        line = fmaf(gx[j] * nx, src[static_cast<long long>(sy_) * W + sx_], line);
                    }
                    blur = fmaf(gy[i] * ny, line, blur);
                }
                row = fmaf(wx[b], blur, row);
            }
            acc = fmaf(wy[a], row, acc);
        }
        out[idx] = acc;
    }
}
}  // namespace mofa

extern "C" int mofa_resize_antialias(const void* img, void* out, int32_t planes, int32_t H, int32_t W, int32_t Ho,
                                     int32_t Wo, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!img || !out || planes <= 0 || H <= 1 || W <= 1 || Ho <= 0 || Wo <= 0) {
        mofa::set_last_error("mofa_resize_antialias: bad arguments");
        return MOFA_ERR_ARG;
    }
    auto sigma = [](int in, int o) {
        const float s = (static_cast<float>(in) / o - 1.0f) / 2.0f;
        return s > 0.001f ? s : 0.001f;
    };
    auto ksize = [](float s) {
        const float k = 4.0f * s > 3.0f ? 4.0f * s : 3.0f;
        int ks = static_cast<int>(k);
        return ks % 2 == 0 ? ks + 1 : ks;
    };
    const float sy = sigma(H, Ho), sx = sigma(W, Wo);
    const int ky = ksize(sy), kx = ksize(sx);
    if (ky > 63 || kx > 63 || ky / 2 >= H || kx / 2 >= W) {
        mofa::set_last_error("mofa_resize_antialias: reduction factor too large for the 63-tap blur (%d x %d taps)", ky, kx);
        return MOFA_ERR_ARG;
    }
    const long long total = static_cast<long long>(planes) * Ho * Wo;
    long long g = (total + 255) / 256;
    if (g > 148LL * 8) g = 148LL * 8;
    mofa::resize_antialias_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(
        static_cast<const float*>(img), static_cast<float*>(out), planes, H, W, Ho, Wo, sy, sx, ky, kx);
    return mofa::check_launch("mofa_resize_antialias");
}

namespace mofa {
// Drag-flow post-processing of T/run_gradio.py:251-277, 330-333 in one pass over the output [F, 2, H, W]:
// (flow * brush) at 384^2 -> nearest resize -> x (W/Ws, H/Hs), fp16 roundings where the reference has them, then
// where((flow_in != 0).all(channel), flow_in, flow_out).
__global__ void __launch_bounds__(256)
flow_post_kernel(const __half* __restrict__ fin, const __half* __restrict__ brush, const __half* __restrict__ fout,
                 __half* __restrict__ out, int F, int Hs, int Ws, int H, int W) {
    const long long total = static_cast<long long>(F) * H * W;
    const float sx = static_cast<float>(W) / Ws, sy = static_cast<float>(H) / Hs;
    const bool resize = (H != Hs) || (W != Ws);
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int x = static_cast<int>(i % W);
        const int y = static_cast<int>((i / W) % H);
        const int f = static_cast<int>(i / (static_cast<long long>(W) * H));
        int ys = static_cast<int>(floorf(y * (static_cast<float>(Hs) / H)));
        int xs = static_cast<int>(floorf(x * (static_cast<float>(Ws) / W)));
        ys = ys < Hs - 1 ? ys : Hs - 1;
        xs = xs < Ws - 1 ? xs : Ws - 1;
        const long long s0 = (static_cast<long long>(f) * 2 * Hs + ys) * Ws + xs;
        const long long plane = static_cast<long long>(Hs) * Ws;
        __half u = fin[s0], v = fin[s0 + plane];
        if (brush) {
            const __half b = brush[static_cast<long long>(ys) * Ws + xs];
            u = __hmul(u, b);
            v = __hmul(v, b);
        }
        if (resize) {
            u = __float2half_rn(__half2float(u) * sx);
            v = __float2half_rn(__half2float(v) * sy);
        }
        if (fout) {
            const bool keep = (__half2float(u) != 0.f) && (__half2float(v) != 0.f);
            if (!keep) {
                __half uo = fout[s0], vo = fout[s0 + plane];
                if (resize) {
                    uo = __float2half_rn(__half2float(uo) * sx);
                    vo = __float2half_rn(__half2float(vo) * sy);
                }
                u = uo;
                v = vo;
            }
        }
        const long long d0 = (static_cast<long long>(f) * 2 * H + y) * W + x;
        out[d0] = u;
        out[d0 + static_cast<long long>(H) * W] = v;
    }
}
}  // namespace mofa

extern "C" int mofa_flow_post(const void* flow_in, const void* brush, const void* flow_out, void* out, int32_t F,
                              int32_t Hs, int32_t Ws, int32_t H, int32_t W, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!flow_in || !out || F <= 0 || Hs <= 0 || Ws <= 0 || H <= 0 || W <= 0) {
        mofa::set_last_error("mofa_flow_post: bad arguments");
        return MOFA_ERR_ARG;
    }
    const long long total = static_cast<long long>(F) * H * W;
    long long g = (total + 255) / 256;
    if (g > 148LL * 16) g = 148LL * 16;
    mofa::flow_post_kernel<<<static_cast<unsigned>(g), 256, 0, stream>>>(
        static_cast<const __half*>(flow_in), static_cast<const __half*>(brush), static_cast<const __half*>(flow_out),
        static_cast<__half*>(out), F, Hs, Ws, H, W);
    return mofa::check_launch("mofa_flow_post");
}

extern "C" int mofa_downsample_nearest(const void* x, void* out, int32_t n_img, int32_t H, int32_t W, int32_t C,
                                       int32_t s, mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!x || !out || s <= 0 || H % s != 0 || W % s != 0 || (C % 8) != 0) {
        set_last_error("mofa_downsample_nearest: bad arguments");
        return MOFA_ERR_ARG;
    }
    const long long total = static_cast<long long>(n_img) * (H / s) * (W / s) * (C / 8);
    downsample_nearest_kernel<<<grid_for2(total), 256, 0, stream>>>(static_cast<const __half*>(x),
                                                                   static_cast<__half*>(out), n_img, H, W, C, s);
    return check_launch("mofa_downsample_nearest");
}

// =============================================================================================
// Sparse motion hints -> dense (flow, mask) planes: the control-signal rasterisation in front of CMP.
//   mode 0  /root/reference/MOFA-Video-Traj/run_gradio.py:61-86  get_sparseflow_and_mask_forward:
//           per track k and step i, flow = int64(end - start) * sign written at pixel (int(start.y), int(start.x)) of a
//           per-track plane, planes summed over k  =>  collisions ADD (flow and mask); float64 arithmetic like numpy.
//           One thread per (k, i): integer-valued fp32 atomicAdd (exact, order-independent).
//   mode 1  /root/reference/MOFA-Video-Keypoint/utils/utils.py:81-119  get_sparse_flow + sample_optical_flow:
//           flow = landmarks[t] - landmarks[0] ASSIGNED at (clip(long(y0)), clip(long(x0))); on a collision the highest
//           landmark index wins (what the sequential CPU index_put_ of the reference leaves behind): pass 1 atomicMax of
//           k into an owner plane, pass 2 the owner writes.  Output already in the returned [b, t-1, 2, h, w] layout.
// =============================================================================================
namespace mofa {

__global__ void sparse_hints_add_kernel(const double* __restrict__ pts, int K, int n_steps, int H, int W, int sign,
                                        float* __restrict__ flow, float* __restrict__ mask) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= K * n_steps) return;
    const int k = idx / n_steps, i = idx - k * n_steps;
    const double* p0 = pts + static_cast<long long>(k) * (n_steps + 1) * 2;
    const double sx = p0[0], sy = p0[1];
    const double ex = p0[(i + 1) * 2 + 0], ey = p0[(i + 1) * 2 + 1];
    int px = static_cast<int>(sx), py = static_cast<int>(sy);       // int(): truncation toward zero
    if (px < 0) px += W;                                             // numpy negative-index wrap (host validated range)
    if (py < 0) py += H;
    const long long fx = static_cast<long long>(ex - sx) * sign;     // np.int64(): truncation toward zero
    const long long fy = static_cast<long long>(ey - sy) * sign;
    const long long pix = (static_cast<long long>(i) * H + py) * W + px;
    atomicAdd(&flow[pix * 2 + 0], static_cast<float>(fx));
    atomicAdd(&flow[pix * 2 + 1], static_cast<float>(fy));
    atomicAdd(&mask[pix], 1.0f);
}

template <typename T>
__device__ __forceinline__ void ldmk_pixel(const T* l0, int H, int W, int& py, int& px) {
    long long x = static_cast<long long>(l0[0]), y = static_cast<long long>(l0[1]);   // .long(): toward zero
    y = y < 0 ? 0 : (y > H - 1 ? H - 1 : y);                                          // torch.clip
    x = x < 0 ? 0 : (x > W - 1 ? W - 1 : x);
    py = static_cast<int>(y);
    px = static_cast<int>(x);
}

template <typename T>
__global__ void sparse_ldmk_owner_kernel(const T* __restrict__ lm, int B, int Tn, int K, int H, int W,
                                         int* __restrict__ owner) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * (Tn - 1) * K) return;
    const int k = idx % K, l = (idx / K) % (Tn - 1), b = idx / (K * (Tn - 1));
    int py, px;
    ldmk_pixel(lm + (static_cast<long long>(b) * Tn * K + k) * 2, H, W, py, px);
    atomicMax(&owner[((static_cast<long long>(b) * (Tn - 1) + l) * H + py) * W + px], k);
}

template <typename T>
__global__ void sparse_ldmk_write_kernel(const T* __restrict__ lm, int B, int Tn, int K, int H, int W,
                                         const int* __restrict__ owner, T* __restrict__ flow,
                                         uint8_t* __restrict__ mask) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= B * (Tn - 1) * K) return;
    const int k = idx % K, l = (idx / K) % (Tn - 1), b = idx / (K * (Tn - 1));
    const T* l0 = lm + (static_cast<long long>(b) * Tn * K + k) * 2;
    const T* lt = lm + ((static_cast<long long>(b) * Tn + (l + 1)) * K + k) * 2;
    int py, px;
    ldmk_pixel(l0, H, W, py, px);
    if (owner[((static_cast<long long>(b) * (Tn - 1) + l) * H + py) * W + px] != k) return;
    const long long plane = static_cast<long long>(H) * W;
    const long long o = (static_cast<long long>(b) * (Tn - 1) + l) * 2 * plane + static_cast<long long>(py) * W + px;
    flow[o] = lt[0] - l0[0];
    flow[o + plane] = lt[1] - l0[1];
    mask[o] = 1;
    mask[o + plane] = 1;
}

}  // namespace mofa

extern "C" int mofa_sparse_hints(const void* pts, int32_t mode, int32_t dtype64, int32_t B, int32_t Tn, int32_t K,
                                 int32_t H, int32_t W, int32_t sign, void* flow, void* mask, int32_t* owner_ws,
                                 mofa_stream_t stream_) {
    cudaStream_t stream = static_cast<cudaStream_t>(stream_);
    if (!pts || !flow || !mask || K <= 0 || Tn < 2 || H <= 0 || W <= 0 || (mode == 1 && (!owner_ws || B <= 0))) {
        set_last_error("mofa_sparse_hints: bad arguments");
        return MOFA_ERR_ARG;
    }
    if (mode == 0) {
        const int n_steps = Tn - 1;
        const long long px = static_cast<long long>(n_steps) * H * W;
        cudaMemsetAsync(flow, 0, sizeof(float) * 2 * px, stream);
        cudaMemsetAsync(mask, 0, sizeof(float) * px, stream);
        const int n = K * n_steps;
        mofa::sparse_hints_add_kernel<<<(n + 127) / 128, 128, 0, stream>>>(
            static_cast<const double*>(pts), K, n_steps, H, W, sign < 0 ? -1 : 1, static_cast<float*>(flow),
            static_cast<float*>(mask));
        return check_launch("mofa_sparse_hints(add)");
    }
    if (mode != 1) {
        set_last_error("mofa_sparse_hints: unknown mode %d", mode);
        return MOFA_ERR_ARG;
    }
    const long long px = static_cast<long long>(B) * (Tn - 1) * H * W;
    cudaMemsetAsync(owner_ws, 0xff, sizeof(int32_t) * px, stream);                       // -1
    cudaMemsetAsync(flow, 0, (dtype64 ? sizeof(double) : sizeof(float)) * 2 * px, stream);
    cudaMemsetAsync(mask, 0, 2 * px, stream);
    const int n = B * (Tn - 1) * K;
    const int g = (n + 127) / 128;
    if (dtype64) {
        mofa::sparse_ldmk_owner_kernel<double><<<g, 128, 0, stream>>>(static_cast<const double*>(pts), B, Tn, K, H, W,
                                                                        owner_ws);
        int rc = check_launch("mofa_sparse_hints(owner)");
        if (rc) return rc;
        mofa::sparse_ldmk_write_kernel<double><<<g, 128, 0, stream>>>(static_cast<const double*>(pts), B, Tn, K, H, W,
                                                                        owner_ws, static_cast<double*>(flow),
                                                                        static_cast<uint8_t*>(mask));
    } else {
        mofa::sparse_ldmk_owner_kernel<float><<<g, 128, 0, stream>>>(static_cast<const float*>(pts), B, Tn, K, H, W,
                                                                       owner_ws);
        int rc = check_launch("mofa_sparse_hints(owner)");
        if (rc) return rc;
        mofa::sparse_ldmk_write_kernel<float><<<g, 128, 0, stream>>>(static_cast<const float*>(pts), B, Tn, K, H, W,
                                                                       owner_ws, static_cast<float*>(flow),
                                                                       static_cast<uint8_t*>(mask));
    }
    return check_launch("mofa_sparse_hints(write)");
}
