"""Plain-PyTorch statement of what every C-ABI op in include/mofa_b200.h computes (fp32 math, fp16 storage).

TEST INFRASTRUCTURE ONLY.  Two uses:
  * `-m gpu` kernel tests compare each CUDA op with the function of the same name here;
  * `-m "not gpu"` tests monkeypatch `mofa_video_b200.lib` with these functions to check the engine's
    host logic (weight packing, op sequencing, quirk handling) against the oracle on CPU.
The product never imports this file.
"""
import math

import torch
import torch.nn.functional as F

A_LINEAR, A_CONV3X3, A_TEMPORAL3 = 0, 1, 2
ACT_NONE, ACT_SILU, ACT_GEGLU, ACT_RELU, ACT_SIGMOID = 0, 1, 2, 3, 4
ACT_GELU, ACT_QUICK_GELU = 6, 7


def pick_bn(n, geglu=False):
    from mofa_video_b200 import lib as real
    return real.pick_bn(n, geglu)


def _epilogue(acc, *, N, bn, act, bias, rowbias, rows_per_group, rowbias_mod, res1, res2, alpha, beta1, beta2):
    """acc: fp32 [rows, N] pre-bias accumulator."""
    rows = acc.shape[0]
    v = acc
    if bias is not None:
        v = v + bias.float()[None, :]
    if rowbias is not None:
        ar = torch.arange(rows, device=acc.device)
        grp = ar % rowbias_mod if rowbias_mod > 0 else ar // rows_per_group
        v = v + rowbias.float()[grp]
    if act == ACT_SILU:
        v = F.silu(v)
    elif act == ACT_RELU:
        v = F.relu(v)
    elif act == ACT_SIGMOID:
        v = torch.sigmoid(v)
    elif act == ACT_GELU:
        v = F.gelu(v)
    elif act == ACT_QUICK_GELU:
        v = v * torch.sigmoid(1.702 * v)
    elif act == ACT_GEGLU:
        # weight rows packed per N tile as [bn/2 value | bn/2 gate]
        t = v.view(rows, N // bn, 2, bn // 2)
        v = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(rows, N // 2)
    v = alpha * v
    if res1 is not None:
        v = v + beta1 * res1.float().reshape(rows, -1)
    if res2 is not None:
        v = v + beta2 * res2.float().reshape(rows, -1)
    if act == 5:  # ReLU after the residual add
        v = F.relu(v)
    return v.half()


def gemm(mode, a, w, out, *, N, bn=None, act=ACT_NONE, a2=None, M=0, K=0, K1=0, lda=0, lda2=0, n_img=0, H=0, W=0,
         C=0, B=0, T=0, HW=0, ldc=None, bias=None, rowbias=None, rows_per_group=1, rowbias_mod=0, res1=None, res2=None,
         alpha=1.0, beta1=1.0, beta2=1.0, max_ctas=0, dilation=1, ksize=3, c_off=0, gn_stats=None,
         gn_rows_per_stat=0, gn_groups=32, gn_cpg=0, gn_c_off=0):
    if bn is None:
        bn = pick_bn(N, act == ACT_GEGLU)
    wf = w.float()
    if mode == A_LINEAR:
        af = a.float().reshape(M, -1)[:, : (K1 if a2 is not None else K)]
        if a2 is not None:
            af = torch.cat([af, a2.float().reshape(M, -1)[:, : K - K1]], dim=1)
        acc = af @ wf[:, :K].t()
    elif mode == A_CONV3X3:
        x = a.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2)
        wk = wf.reshape(N, ksize, ksize, C).permute(0, 3, 1, 2)
        acc = F.conv2d(x, wk, padding=dilation * (ksize // 2), dilation=dilation).permute(0, 2, 3, 1)
        acc = acc.reshape(n_img * H * W, N)
    else:
        x = a.float().reshape(B, T, HW, C)
        xp = F.pad(x, (0, 0, 0, 0, 1, 1))
        cols = torch.cat([xp[:, 0:T], xp[:, 1:T + 1], xp[:, 2:T + 2]], dim=-1)  # (kt, c)
        acc = cols.reshape(B * T * HW, 3 * C) @ wf.t()
    res = _epilogue(acc, N=N, bn=bn, act=act, bias=bias, rowbias=rowbias, rows_per_group=rows_per_group, rowbias_mod=rowbias_mod,
                    res1=res1, res2=res2, alpha=alpha, beta1=beta1, beta2=beta2)
    n_out = res.shape[1]
    ld = ldc if ldc is not None else n_out
    out.view(-1, ld)[: res.shape[0], c_off:c_off + n_out] = res
    if gn_stats is not None:  # (sum, sum of squares) of the fp16 outputs per (statistic, consumer group)
        cpg = gn_cpg if gn_cpg else n_out // gn_groups
        rf = res.float()
        n_stat = rf.shape[0] // gn_rows_per_stat
        st = gn_stats.view(-1)[: n_stat * gn_groups * 2].view(n_stat, gn_groups, 2)
        st.zero_()
        grp = (gn_c_off + torch.arange(n_out, device=rf.device)) // cpg
        r3 = rf.view(n_stat, gn_rows_per_stat, n_out)
        st[:, :, 0].index_add_(1, grp, r3.sum(1))
        st[:, :, 1].index_add_(1, grp, (r3 * r3).sum(1))
    return out


def linear(a, w, out, **kw):
    M, K = a.shape[0], a.shape[1]
    return gemm(A_LINEAR, a, w, out, N=w.shape[0], M=M, K=K if "K" not in kw else kw.pop("K"), lda=a.stride(0), **kw)


def attn_spatial(qkv, out, frames, L, heads, scale):
    C = heads * 64
    t = qkv.float().reshape(frames, L, 3, heads, 64).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    s = (q @ k.transpose(-1, -2)) * scale
    o = torch.softmax(s, dim=-1) @ v
    out.view(frames, L, heads, 64).copy_(o.permute(0, 2, 1, 3).half())
    return out


FF_FUSED_MAX_C = 0      # product default (lib.FF_FUSED_MAX_C): the fused FeedForward is opt-in
FF_FUSED_LIMIT_C = 320


def ff_geglu(x, w1_packed, b1_packed, w2, b2, out, res1=None, res2=None, alpha=1.0, beta1=1.0, beta2=1.0):
    """mofa_ff_geglu: GEGLU projection (bn = 128 packing: per 128 rows 64 value then 64 gate) -> GELU gate -> Linear."""
    hidden = w2.shape[1]
    h = x.float() @ w1_packed.float().t() + (b1_packed.float() if b1_packed is not None else 0.0)
    t = h.view(x.shape[0], hidden // 64, 2, 64)
    hid = (t[:, :, 0] * F.gelu(t[:, :, 1])).reshape(x.shape[0], hidden).half().float()   # H is fp16 in tensor memory
    v = hid @ w2.float().t() + (b2.float() if b2 is not None else 0.0)
    v = alpha * v
    if res1 is not None:
        v = v + beta1 * res1.float()
    if res2 is not None:
        v = v + beta2 * res2.float()
    out.copy_(v.half())
    return out


def attn_small(qkv, out, n_seq, L, heads, head_dim, scale):
    t = qkv.float().reshape(n_seq, L, 3, heads, head_dim).permute(2, 0, 3, 1, 4)
    q, k, v = t[0], t[1], t[2]
    o = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1) @ v
    out.view(n_seq, L, heads, head_dim).copy_(o.permute(0, 2, 1, 3).half())
    return out


def attn_small_temporal(qkv, out, B, T, HW, heads, head_dim, scale):
    t = qkv.float().reshape(B, T, HW, 3, heads, head_dim).permute(3, 0, 2, 4, 1, 5)  # [3,B,HW,h,T,d]
    q, k, v = t[0], t[1], t[2]
    o = torch.softmax((q @ k.transpose(-1, -2)) * scale, dim=-1) @ v                # [B,HW,h,T,d]
    out.view(B, T, HW, heads, head_dim).copy_(o.permute(0, 3, 1, 2, 4).half())
    return out


def attn_temporal(qkv, out, B, T, HW, heads, scale):
    t = qkv.float().reshape(B, T, HW, 3, heads, 64).permute(3, 0, 2, 4, 1, 5)  # [3,B,HW,h,T,64]
    q, k, v = t[0], t[1], t[2]
    s = (q @ k.transpose(-1, -2)) * scale
    o = torch.softmax(s, dim=-1) @ v  # [B,HW,h,T,64]
    out.view(B, T, HW, heads, 64).copy_(o.permute(0, 3, 1, 2, 4).half())
    return out


def groupnorm(x1, gamma, beta, out, rows_per_stat, eps, silu, stats, x2=None, groups=32, stats_ready=False):
    x = x1.float() if x2 is None else torch.cat([x1.float(), x2.float()], dim=1)
    rows, C = x.shape
    n = rows // rows_per_stat
    xg = x.view(n, rows_per_stat, groups, C // groups)
    if stats_ready:   # statistics accumulated by the producing GEMM's epilogue (gemm(..., gn_stats=))
        st = stats.view(-1)[: n * groups * 2].view(n, groups, 2)
        cnt = float(rows_per_stat * (C // groups))
        mean = (st[:, :, 0] / cnt).view(n, 1, groups, 1)
        var = (st[:, :, 1] / cnt).view(n, 1, groups, 1) - mean * mean
        var = var.clamp_min(0.0)
    else:
        mean = xg.mean(dim=(1, 3), keepdim=True)
        var = xg.var(dim=(1, 3), unbiased=False, keepdim=True)
    y = ((xg - mean) * torch.rsqrt(var + eps)).view(rows, C) * gamma.float() + beta.float()
    if silu:
        y = F.silu(y)
    out.copy_(y.half().view_as(out))
    return out


def layernorm(x, gamma, beta, out, eps=1e-5, add=None, rows_per_group=1, add_period=1, sum_out=None):
    xf = x.float()
    if add is not None:
        idx = (torch.arange(x.shape[0], device=x.device) // rows_per_group) % add_period
        xs = (xf + add.float().view(add_period, -1)[idx]).half()
        if sum_out is not None:
            sum_out.copy_(xs)
        xf = xs.float()
    y = F.layer_norm(xf, (x.shape[1],), gamma.float(), beta.float(), eps)
    out.copy_(y.half())
    return out


def axpy_bcast(x, y, out, scale=1.0):
    n, p = x.numel(), y.numel()
    r = x.float().reshape(n // p, p) + scale * y.float().reshape(1, p)
    out.copy_(r.half().view_as(out))
    return out


def im2col3x3(x, out, n_img, H, W, C, stride, Kpad):
    xi = x.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2)
    cols = F.unfold(xi, 3, padding=1, stride=stride)  # [n, C*9, L] ordered (c, ky, kx)
    Lo = cols.shape[-1]
    cols = cols.view(n_img, C, 9, Lo).permute(0, 3, 2, 1).reshape(n_img * Lo, 9 * C)  # (ky,kx,c)
    o = out.view(n_img * Lo, Kpad)
    o.zero_()
    o[:, : 9 * C] = cols.half()
    return out


def upsample2x(x, out, n_img, H, W, C):
    xi = x.reshape(n_img, H, W, C)
    out.view(n_img, 2 * H, 2 * W, C).copy_(xi.repeat_interleave(2, dim=1).repeat_interleave(2, dim=2))
    return out


def nchw_to_nhwc(x, out, n_img, C, HW, ldo=None, c_off=0):
    ldo = ldo if ldo is not None else C
    out.view(n_img, HW, ldo)[:, :, c_off:c_off + C] = x.reshape(n_img, C, HW).permute(0, 2, 1)
    return out


def nhwc_to_nchw(x, out, n_img, C, HW, ldi=None, c_off=0):
    ldi = ldi if ldi is not None else C
    out.view(n_img, C, HW).copy_(x.view(n_img, HW, ldi)[:, :, c_off:c_off + C].permute(0, 2, 1))
    return out


def linear_small(a, w, bias, out, act_in=0, act_out=0):
    af = a.float()
    if act_in == 1:
        af = F.silu(af).half().float()
    v = af @ w.float().t()
    if bias is not None:
        v = v + bias.float()
    if act_out == 1:
        v = F.silu(v)
    out.copy_(v.half().view_as(out))
    return out


def timestep_embedding(t, out, dim):
    half = dim // 2
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32, device=t.device) / half)
    arg = t.float().reshape(-1, 1) * freq[None]
    out.copy_(torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1).half().view_as(out))
    return out


def softsplat_avg(feat, flow, acc, wsum, out, F_, hs, ws, C, Hf, Wf):
    s = Hf // hs
    fl = (flow.view(F_, 2, Hf, Wf)[:, :, ::s, ::s].float() / s).half().float()
    src = feat.float().view(hs * ws, C)
    ys, xs = torch.meshgrid(torch.arange(hs, device=feat.device), torch.arange(ws, device=feat.device), indexing="ij")
    o = out.view(F_, hs * ws, C)
    for f in range(F_):
        ox = xs.float() + fl[f, 0]
        oy = ys.float() + fl[f, 1]
        fin = torch.isfinite(ox) & torch.isfinite(oy)
        x0 = torch.floor(ox)
        y0 = torch.floor(oy)
        a = torch.zeros(hs * ws, C, device=feat.device)
        wsm = torch.zeros(hs * ws, device=feat.device)
        for dx, dy in ((0, 0), (1, 0), (0, 1), (1, 1)):
            cx, cy = x0 + dx, y0 + dy
            wx = (x0 + 1 - ox) if dx == 0 else (ox - x0)
            wy = (y0 + 1 - oy) if dy == 0 else (oy - y0)
            wgt = (wx * wy)
            ok = fin & (cx >= 0) & (cx < ws) & (cy >= 0) & (cy < hs)
            idx = (cy.clamp(0, hs - 1) * ws + cx.clamp(0, ws - 1)).long().view(-1)
            wv = torch.where(ok, wgt, torch.zeros_like(wgt)).view(-1)
            wv = torch.nan_to_num(wv, nan=0.0, posinf=0.0, neginf=0.0)
            a.index_add_(0, idx, src * wv[:, None])
            wsm.index_add_(0, idx, wv)
        o[f] = (a / (wsm[:, None] + 1e-7)).half()
    return out


def cfg_euler_step(noise, latents_h, image_latents, next_in, T, HW, g_min, g_max, sigma, sigma_next):
    x = latents_h.float().view(T, 4, HW)
    if noise is not None:
        n = noise.float().view(2, T, HW, 4).permute(0, 1, 3, 2)  # [2,T,4,HW]
        g = torch.linspace(g_min, g_max, T, device=x.device).half().float().view(T, 1, 1) if T > 1 else \
            torch.full((1, 1, 1), g_min, device=x.device)
        v = n[0] + g * (n[1] - n[0])
        s2 = sigma * sigma + 1.0
        x0 = v * (-sigma / math.sqrt(s2)) + x / s2
        d = (x - x0) / sigma
        xn = (x + d * (sigma_next - sigma)).half()
        latents_h.copy_(xn.view_as(latents_h))
        x = xn.float()
    scaled = (x / math.sqrt(sigma_next * sigma_next + 1.0)).half()  # [T,4,HW]
    o = next_in.view(2, T, HW, 8)
    o[:, :, :, :4] = scaled.permute(0, 2, 1)[None]
    o[:, :, :, 4:] = image_latents.view(2, 1, 4, HW).permute(0, 1, 3, 2)
    return next_in


def cfg_euler_step_dev(noise, latents_h, image_latents, next_in, T, HW, g_min, g_max, sigmas):
    return cfg_euler_step(noise, latents_h, image_latents, next_in, T, HW, g_min, g_max, float(sigmas[0]),
                          float(sigmas[1]))


def sparse_hints_add(pts, flow, mask, sign=1):
    """mofa_sparse_hints mode 0 (T/run_gradio.py:61-86): float64 points, collisions add."""
    K, Tn, _ = pts.shape
    _, H, W, _ = flow.shape
    flow.zero_()
    mask.zero_()
    p = pts.cpu().numpy()
    for k in range(K):
        px, py = int(p[k, 0, 0]), int(p[k, 0, 1])
        for i in range(Tn - 1):
            d = (p[k, i + 1] - p[k, 0]).astype("int64") * sign
            flow[i, py, px, 0] += float(d[0])
            flow[i, py, px, 1] += float(d[1])
            mask[i, py, px] += 1.0
    return flow, mask


def sparse_hints_assign(landmarks, flow, mask, owner):
    """mofa_sparse_hints mode 1 (K/utils/utils.py:81-119): assignment, the highest landmark index wins."""
    B, Tn, K, _ = landmarks.shape
    H, W = flow.shape[-2:]
    flow.zero_()
    mask.zero_()
    for b in range(B):
        for k in range(K):
            x = min(max(int(landmarks[b, 0, k, 0].long()), 0), W - 1)
            y = min(max(int(landmarks[b, 0, k, 1].long()), 0), H - 1)
            for l in range(Tn - 1):
                flow[b, l, :, y, x] = landmarks[b, l + 1, k] - landmarks[b, 0, k]
                mask[b, l, :, y, x] = 1
    return flow, mask


def profiling():
    return False


def note_graph_replay(n):
    pass


def launch_count():
    return 0


def softmax_rows(x, L=None):
    L = L if L is not None else x.shape[1]
    x[:, :L] = torch.softmax(x[:, :L].float(), dim=-1).half()
    return x


def vae_time_conv_out(y, w, b, out_f32, out_u8, T, HW):
    yy = y.float().view(T, HW, 3).permute(2, 0, 1)[None, :, :, :, None]  # [1, 3, T, HW, 1]
    o = F.conv3d(yy, w.view(3, 3, 3, 1, 1), b, padding=(1, 0, 0))[0, :, :, :, 0]  # [3, T, HW]
    if out_f32 is not None:
        out_f32.view(T, 3, HW).copy_(o.permute(1, 0, 2))
    if out_u8 is not None:
        u = ((o / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8)
        out_u8.view(T, HW, 3).copy_(u.permute(1, 2, 0))


def conv_out_size(n, k, stride, pad, dil=1):
    """pad < 0 means asymmetric padding: 0 before, -pad after (mofa_im2col)."""
    return (n + (-pad if pad < 0 else 2 * pad) - dil * (k - 1) - 1) // stride + 1


def im2col(x, out, n_img, H, W, C, ksize, stride, pad, dilation, Kpad):
    xi = x.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2)
    if pad < 0:  # asymmetric: 0 before, -pad after
        xi = F.pad(xi, (0, -pad, 0, -pad))
        pad = 0
    cols = F.unfold(xi, ksize, dilation=dilation, padding=pad, stride=stride)  # [n, C*k*k, L] ordered (c, ky, kx)
    Lo = cols.shape[-1]
    kk = ksize * ksize
    cols = cols.view(n_img, C, kk, Lo).permute(0, 3, 2, 1).reshape(n_img * Lo, kk * C)
    o = out.view(n_img * Lo, Kpad)
    o.zero_()
    o[:, : kk * C] = cols.half()
    return out


def pool2d(x, out, n_img, H, W, C, ksize, stride, pad, mode):
    xi = x.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2)
    o = F.max_pool2d(xi, ksize, stride, pad) if mode == 0 else F.avg_pool2d(xi, ksize, stride, pad)
    out.view(n_img, o.shape[2], o.shape[3], C).copy_(o.permute(0, 2, 3, 1).half())
    return out


def resize_bilinear_ac(x, out, n_img, H, W, C, Ho, Wo, ldo=None, c_off=0):
    ldo = ldo if ldo is not None else C
    xi = x.float().reshape(n_img, H, W, C).permute(0, 3, 1, 2)
    o = F.interpolate(xi, size=(Ho, Wo), mode="bilinear", align_corners=True)
    out.view(n_img, Ho, Wo, ldo)[..., c_off:c_off + C] = o.permute(0, 2, 3, 1).half()
    return out


def cmp_fuser(logits, flow, nbins=99, fmax=50.0):
    l = logits.float()
    step = 2 * fmax / nbins
    mesh = torch.arange(nbins, device=l.device).float() * step - fmax + step / 2
    fx = (torch.softmax(l[:, :nbins], -1) * mesh).sum(-1)
    fy = (torch.softmax(l[:, nbins:], -1) * mesh).sum(-1)
    flow.copy_(torch.stack([fx, fy], -1).half())
    return flow


def copy_cols(src, dst, rows, C, period_rows, ldo, c_off):
    idx = torch.arange(rows, device=src.device) % period_rows
    dst.view(rows, ldo)[:, c_off:c_off + C] = src.view(-1, C)[idx]
    return dst


def flow_pyramid(flow, out, F_, hs, ws, Hf, Wf, ldo, c_off):
    s = Hf // hs
    fl = (flow.view(F_, 2, Hf, Wf)[:, :, ::s, ::s].float() / s).half()  # [F,2,hs,ws]
    out.view(F_, hs, ws, ldo)[..., c_off:c_off + 2] = fl.permute(0, 2, 3, 1)
    return out


def mask_blend(a, b, mask, out, period_rows=None):
    rows = a.shape[0]
    period_rows = period_rows if period_rows is not None else rows
    m = mask.float().view(-1)[torch.arange(rows, device=a.device) % period_rows][:, None]
    out.copy_((a.float() * m + b.float() * (1 - m)).half())
    return out


def downsample_nearest(x, out, n_img, H, W, C, s):
    out.view(n_img, H // s, W // s, C).copy_(x.view(n_img, H, W, C)[:, ::s, ::s])
    return out


def flow_post(flow_in, out, F, Hs, Ws, H, W, brush=None, flow_out=None):
    """T/run_gradio.py:251-255, 268-275, 330-333 in the reference's own op order and fp16 roundings."""
    import torch.nn.functional as Fn

    def resize(fl):
        if (H, W) == (Hs, Ws):
            return fl
        r = Fn.interpolate(fl.float(), (H, W), mode="nearest").half()
        r[:, 0] = (r[:, 0].float() * (W / Ws)).half()
        r[:, 1] = (r[:, 1].float() * (H / Hs)).half()
        return r
    a = flow_in.view(F, 2, Hs, Ws)
    if brush is not None:
        a = a * brush.view(1, 1, Hs, Ws)
    a = resize(a)
    if flow_out is not None:
        b = resize(flow_out.view(F, 2, Hs, Ws))
        keep = (a != 0).all(dim=1, keepdim=True).expand_as(a)
        a = torch.where(keep, a, b)
    out.view(F, 2, H, W).copy_(a)
    return out


def resize_antialias(img, out):
    """pipeline.py:532-640 as restated (two separable passes + F.interpolate) in oracle/pipeline.py."""
    from oracle.pipeline import resize_with_antialiasing
    out.copy_(resize_with_antialiasing(img.float().cpu(), tuple(out.shape[-2:])).to(out.device))
    return out
