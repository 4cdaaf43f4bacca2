"""-m gpu: every CUDA op (through the C ABI) against its plain-PyTorch statement in tests/ref_ops.py."""
import math

import pytest
import torch

import ref_ops as R

pytestmark = pytest.mark.gpu

DEV = "cuda"


def L():
    from mofa_video_b200 import lib
    return lib


def rnd(*shape, scale=1.0, seed=None):
    g = torch.Generator(device="cpu")
    g.manual_seed(seed if seed is not None else (hash(shape) % 100000))
    return (torch.randn(*shape, generator=g) * scale).half().to(DEV)


def close(a, b, atol, rtol, what=""):
    a, b = a.float(), b.float()
    err = (a - b).abs()
    tol = atol + rtol * b.abs()
    bad = err > tol
    if bad.any():
        idx = bad.nonzero()[:8].tolist()
        raise AssertionError(
            f"{what}: {int(bad.sum())}/{bad.numel()} mismatches, max err {err.max().item():.4g}, "
            f"ref absmax {b.abs().max().item():.4g}; first idx {idx}; got {a[bad][:6].tolist()} want {b[bad][:6].tolist()}")


# ------------------------------------------------------------------------------------------ GEMM
@pytest.mark.parametrize("M,K,N,bn", [
    (128, 64, 64, 64),          # one tile, one k-block
    (300, 128, 64, 64),         # ragged M
    (1000, 320, 320, 160),      # bn = 160 (level-0 channel count)
    (256, 72, 320, 160),        # K not a multiple of 64 (conv_in im2col): TMA zero-fills the K tail
    (128 * 300, 256, 640, 160),  # 1200 tiles > 148 CTAs: persistent loop, phase wrap, TMEM double buffering
    (512, 2880, 4, 16),         # tiny N (conv_out): scalar tail stores
    (700, 1280, 960, 240),      # bn = 240: not a multiple of 64 -> direct-store epilogue
    (1000, 320, 320, 192),      # TMA-store epilogue, narrower last N tile (192 + 128)
    (700, 1280, 960, 256),      # 256,256,256,192
    (300, 128, 640, 256),       # ragged M with TMA-store clipping, last tile 128
    (128 * 150, 320, 960, None),  # auto tile, many tiles per CTA
])
def test_gemm_linear_plain(M, K, N, bn):
    lib = L()
    a, w = rnd(M, K, scale=0.5), rnd(N, K, scale=0.05)
    out = torch.zeros(M, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.linear(a, w, out, bn=bn)
    R.linear(a, w, ref, bn=bn)
    torch.cuda.synchronize()
    close(out, ref, 2e-2, 1e-2, f"linear {M}x{K}x{N}")


@pytest.mark.parametrize("bn", [160, 192, 256])
def test_gemm_linear_epilogue(bn):
    lib = L()
    M, K, N = 900, 320, 320
    a, w = rnd(M, K, scale=0.5), rnd(N, K, scale=0.05)
    bias, rowbias = rnd(N), rnd(3, N)
    res1, res2 = rnd(M, N), rnd(M, N)
    kw = dict(bias=bias, rowbias=rowbias, rows_per_group=300, res1=res1, res2=res2, alpha=0.7, beta1=1.0, beta2=-0.5,
              act=R.ACT_SILU, bn=bn)
    out = torch.zeros(M, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.linear(a, w, out, **kw)
    R.linear(a, w, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 2e-2, 1e-2, "linear epilogue")


def test_gemm_linear_split_k():
    lib = L()
    M, K1, K2, N = 500, 128, 64, 128
    a, a2, w = rnd(M, K1, scale=0.5), rnd(M, K2, scale=0.5), rnd(N, K1 + K2, scale=0.05)
    out = torch.zeros(M, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    kw = dict(N=N, M=M, K=K1 + K2, K1=K1, lda=K1, lda2=K2, a2=a2, bn=128)
    lib.gemm(lib.A_LINEAR, a, w, out, **kw)
    R.gemm(R.A_LINEAR, a, w, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 2e-2, 1e-2, "split-K linear")


@pytest.mark.parametrize("bn", [256, 128])
def test_gemm_geglu(bn):
    lib = L()
    M, K, N = 600, 320, 2560
    a, w, bias = rnd(M, K, scale=0.5), rnd(N, K, scale=0.05), rnd(N, scale=0.1)
    out = torch.zeros(M, N // 2, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    kw = dict(bias=bias, act=R.ACT_GEGLU, bn=bn)
    lib.linear(a, w, out, **kw)                    # K = 320: the 16-epilogue-warp instantiation
    R.linear(a, w, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 2e-2, 1e-2, "geglu")
    K2 = 640                                       # K = 640: the 8-warp instantiation
    a2, w2 = rnd(M, K2, scale=0.4), rnd(N, K2, scale=0.04)
    lib.linear(a2, w2, out, **kw)
    R.linear(a2, w2, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 2e-2, 1e-2, "geglu K=640")
    with pytest.raises(RuntimeError):              # GEGLU carries a bias only
        lib.linear(a, w, out, res1=rnd(M, N // 2), **kw)


@pytest.mark.parametrize("n_img,H,W,C,N", [
    (2, 8, 32, 64, 64),     # exact tiles (BW=32, BH=4)
    (3, 18, 32, 128, 160),  # ragged rows
    (2, 9, 16, 64, 128),    # BW=16, BH=8, ragged
    (1, 72, 128, 320, 320),  # level-0 shape of config 2, one frame
    (2, 4, 4, 64, 32),      # tiny map (config-1 bottom level)
])
def test_gemm_conv3x3(n_img, H, W, C, N):
    lib = L()
    x, w = rnd(n_img * H * W, C, scale=0.5), rnd(N, 9 * C, scale=0.03)
    bias, rowbias = rnd(N), rnd(n_img, N)
    out = torch.zeros(n_img * H * W, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    kw = dict(N=N, n_img=n_img, H=H, W=W, C=C, bias=bias, rowbias=rowbias, rows_per_group=H * W)
    lib.gemm(lib.A_CONV3X3, x, w, out, **kw)
    R.gemm(R.A_CONV3X3, x, w, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 3e-2, 1e-2, f"conv3x3 {n_img}x{H}x{W}x{C}->{N}")


@pytest.mark.parametrize("B,T,HW,C,N", [(2, 5, 128, 64, 64), (2, 7, 144, 128, 128), (1, 25, 576, 320, 320),
                                        (2, 3, 16, 64, 64)])
def test_gemm_temporal3(B, T, HW, C, N):
    lib = L()
    x, w = rnd(B * T * HW, C, scale=0.5), rnd(N, 3 * C, scale=0.05)
    res = rnd(B * T * HW, N)
    out = torch.zeros(B * T * HW, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    kw = dict(N=N, B=B, T=T, HW=HW, C=C, res1=res, alpha=0.4)
    lib.gemm(lib.A_TEMPORAL3, x, w, out, **kw)
    R.gemm(R.A_TEMPORAL3, x, w, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 3e-2, 1e-2, f"temporal3 {B}x{T}x{HW}x{C}->{N}")


@pytest.mark.parametrize("M,C", [(1000, 320), (460, 64), (700, 256), (128, 192)])
def test_ff_geglu_fused(M, C):
    """mofa_ff_geglu (GEGLU projection -> GELU gate -> Linear, hidden activation kept in tensor memory) vs its statement,
    and vs the two-GEMM path on the same packed weights."""
    lib = L()
    hidden = 4 * C
    x = rnd(M, C, scale=0.7)
    w1, b1 = rnd(2 * hidden, C, scale=0.08), rnd(2 * hidden, scale=0.2)
    w2, b2 = rnd(C, hidden, scale=0.05), rnd(C, scale=0.2)
    r1, r2 = rnd(M, C), rnd(M, C)
    for kw in (dict(), dict(res1=r1), dict(res1=r1, res2=r2, alpha=0.6, beta1=0.6, beta2=0.4)):
        out = torch.zeros(M, C, dtype=torch.half, device=DEV)
        ref = torch.zeros_like(out)
        lib.ff_geglu(x, w1, b1, w2, b2, out, **kw)
        R.ff_geglu(x, w1, b1, w2, b2, ref, **kw)
        torch.cuda.synchronize()
        close(out, ref, 3e-2, 1e-2, f"fused FF M={M} C={C} {sorted(kw)}")
    # the unfused path on the same weights (GEGLU GEMM with 128-row groups, then the output GEMM)
    f = torch.empty(M, hidden, dtype=torch.half, device=DEV)
    two = torch.empty(M, C, dtype=torch.half, device=DEV)
    lib.linear(x, w1, f, bias=b1, act=lib.ACT_GEGLU, bn=128)
    lib.linear(f, w2, two, bias=b2, res1=r1)
    one = torch.empty_like(two)
    lib.ff_geglu(x, w1, b1, w2, b2, one, res1=r1)
    torch.cuda.synchronize()
    close(one, two, 2e-2, 1e-2, "fused vs two GEMMs")


@pytest.mark.parametrize("M,K,N", [(700, 320, 960), (1000, 320, 320), (300, 72, 320), (513, 384, 96), (260, 128, 1280)])
def test_gemm_linear_short_k_16_epilogue_warps(M, K, N):
    """LINEAR, K <= 384, plain epilogue: the 16-epilogue-warp instantiation (32-column units, SWIZZLE_64B staging)."""
    lib = L()
    a, w = rnd(M, K, scale=0.5), rnd(N, K, scale=0.05)
    bias, res1, res2 = rnd(N, scale=0.2), rnd(M, N), rnd(M, N)
    out = torch.zeros(M, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    for kw in (dict(bias=bias, res1=res1), dict(bias=bias, res1=res1, res2=res2, alpha=0.7, beta1=0.5, beta2=-1.25),
               dict(bias=bias, res1=res1, rowbias=rnd(4, N), rows_per_group=(M + 3) // 4),
               dict(res1=res1, rowbias=rnd(3, N), rowbias_mod=3, alpha=1.5), dict()):
        out.zero_()
        lib.linear(a, w, out, **kw)
        R.linear(a, w, ref, **kw)
        torch.cuda.synchronize()
        close(out, ref, 2e-2, 1e-2, f"short-K linear {M}x{K}x{N} {sorted(kw)}")


@pytest.mark.parametrize("n_img,H,W,C,N", [
    (50, 9, 16, 64, 128),    # level 3 of 576x1024: tiles of {16 px, 1 row, 8 images}; 50 images = 6 full groups + 2
    (6, 8, 8, 64, 64),       # level 3 of 512x512: two whole images per tile
    (5, 18, 32, 64, 192),    # level 2: {32 px, 2 rows, 2 images}, odd image count
    (3, 5, 3, 64, 64),       # W not a power of two, everything ragged
])
def test_gemm_conv3x3_small_maps_multi_image_tiles(n_img, H, W, C, N):
    lib = L()
    x, w = rnd(n_img * H * W, C, scale=0.5), rnd(N, 9 * C, scale=0.03)
    bias, rowbias, res = rnd(N), rnd(n_img, N), rnd(n_img * H * W, N)
    out = torch.zeros(n_img * H * W, N, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    kw = dict(N=N, n_img=n_img, H=H, W=W, C=C, bias=bias, rowbias=rowbias, rows_per_group=H * W, res1=res)
    lib.gemm(lib.A_CONV3X3, x, w, out, **kw)
    R.gemm(R.A_CONV3X3, x, w, ref, **kw)
    torch.cuda.synchronize()
    close(out, ref, 3e-2, 1e-2, f"conv3x3 small map {n_img}x{H}x{W}x{C}->{N}")


@pytest.mark.parametrize("mode,shape", [
    ("conv", (4, 16, 32, 64, 320, 16 * 32)),       # spatial statistics (one per image), N = 320: groups of 10
    ("conv", (50, 9, 16, 64, 128, 9 * 16)),        # several images inside one warp's 32 rows
    ("conv", (4, 12, 20, 64, 128, 2 * 12 * 20)),   # temporal statistics (2 frames per batch item), ragged tiles
    ("temporal", (2, 5, 144, 64, 64, 5 * 144)),    # temporal conv -> temporal GroupNorm
    ("temporal", (2, 3, 200, 128, 640, 200)),      # temporal conv -> per-frame GroupNorm (resblock output), cpg 20
    ("linear", (700, 128, 192, 350)),              # linear, ragged M, statistic boundary inside a tile
])
def test_gemm_groupnorm_statistics_in_epilogue(mode, shape):
    """gn_stats: (sum, sum of squares) of the fp16 outputs per (statistic, group), accumulated by the epilogue; then
    mofa_groupnorm(stats_ready) == the two-pass GroupNorm."""
    lib = L()
    if mode == "conv":
        n_img, H, W, C, N, rps = shape
        rows = n_img * H * W
        x, w = rnd(rows, C, scale=0.5), rnd(N, 9 * C, scale=0.03)
        kw = dict(N=N, n_img=n_img, H=H, W=W, C=C, bias=rnd(N), res1=rnd(rows, N))
        m = lib.A_CONV3X3
    elif mode == "temporal":
        B, T, HW, C, N, rps = shape
        rows = B * T * HW
        x, w = rnd(rows, C, scale=0.5), rnd(N, 3 * C, scale=0.05)
        kw = dict(N=N, B=B, T=T, HW=HW, C=C, bias=rnd(N), res1=rnd(rows, N), alpha=0.4)
        m = lib.A_TEMPORAL3
    else:
        rows, K, N, rps = shape
        x, w = rnd(rows, K, scale=0.5), rnd(N, K, scale=0.05)
        kw = dict(N=N, M=rows, K=K, lda=K, bias=rnd(N))
        m = lib.A_LINEAR
    n_stat = rows // rps
    out = torch.zeros(rows, N, dtype=torch.half, device=DEV)
    st = torch.full((n_stat * 64,), 123.0, device=DEV)          # must be zeroed by the call
    lib.gemm(m, x, w, out, gn_stats=st, gn_rows_per_stat=rps, **kw)
    torch.cuda.synchronize()
    of = out.float().view(n_stat, rps, 32, N // 32)
    want = torch.stack([of.sum(dim=(1, 3)), (of * of).sum(dim=(1, 3))], dim=-1)      # [n_stat, 32, 2]
    got = st.view(n_stat, 32, 2)
    close(got, want, 0.05, 2e-3, f"gn statistics {mode} {shape}")
    gamma, beta = rnd(N), rnd(N)
    y1, y2 = torch.empty_like(out), torch.empty_like(out)
    lib.groupnorm(out, gamma, beta, y1, rps, 1e-5, True, st, stats_ready=True)
    st2 = torch.empty(n_stat * 64, device=DEV)
    lib.groupnorm(out, gamma, beta, y2, rps, 1e-5, True, st2)
    close(y1, y2, 4e-3, 4e-3, f"groupnorm from epilogue statistics {mode} {shape}")


# ------------------------------------------------------------------------------------------ attention
@pytest.mark.parametrize("frames,Lq,heads", [(2, 128, 1), (3, 144, 2), (2, 576, 5), (1, 2304, 2), (2, 64, 1),
                                             (1, 9216, 1)])
def test_attn_spatial(frames, Lq, heads):
    lib = L()
    C = heads * 64
    qkv = rnd(frames * Lq, 3 * C, scale=1.0)
    out = torch.zeros(frames * Lq, C, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.attn_spatial(qkv, out, frames, Lq, heads, 0.125)
    R.attn_spatial(qkv, ref, frames, Lq, heads, 0.125)
    torch.cuda.synchronize()
    close(out, ref, 3e-3, 2e-2, f"attn_spatial L={Lq}")


@pytest.mark.parametrize("B,T,HW,heads", [(2, 25, 144, 5), (1, 14, 16, 10), (2, 8, 100, 1), (2, 32, 50, 2),
                                            (2, 3, 256, 1), (1, 17, 2500, 3)])
def test_attn_temporal(B, T, HW, heads):
    lib = L()
    C = heads * 64
    qkv = rnd(B * T * HW, 3 * C)
    out = torch.zeros(B * T * HW, C, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.attn_temporal(qkv, out, B, T, HW, heads, 0.125)
    R.attn_temporal(qkv, ref, B, T, HW, heads, 0.125)
    torch.cuda.synchronize()
    close(out, ref, 3e-3, 2e-2, "attn_temporal")


# ------------------------------------------------------------------------------------------ norms & elementwise
@pytest.mark.parametrize("rows,rps,C1,C2,silu", [(4 * 144, 144, 320, 0, True), (2 * 576, 576, 640, 320, True),
                                                 (2 * 5 * 64, 5 * 64, 1280, 0, False), (512, 256, 128, 0, True),
                                                 (3 * 300, 300, 1280, 640, True), (2 * 5000, 5000, 320, 0, True),
                                                 (2 * 3 * 16, 3 * 16, 64, 0, False), (2 * 700, 700, 1280, 1280, True)])
def test_groupnorm(rows, rps, C1, C2, silu):
    lib = L()
    x1 = rnd(rows, C1, scale=2.0) + 0.5
    x2 = rnd(rows, C2) if C2 else None
    C = C1 + C2
    gamma, beta = rnd(C) * 0.1 + 1.0, rnd(C) * 0.1
    out = torch.zeros(rows, C, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    stats = torch.zeros(rows // rps * 64, dtype=torch.float32, device=DEV)
    lib.groupnorm(x1, gamma, beta, out, rps, 1e-5, silu, stats, x2=x2)
    R.groupnorm(x1, gamma, beta, ref, rps, 1e-5, silu, stats, x2=x2)
    torch.cuda.synchronize()
    close(out, ref, 4e-3, 4e-3, "groupnorm")


@pytest.mark.parametrize("rows,C,with_add", [(1000, 320, False), (777, 640, True), (300, 1280, True), (99, 64, True),
                                             (65, 256, False), (40, 2048, True)])
def test_layernorm(rows, C, with_add):
    lib = L()
    x = rnd(rows, C, scale=2.0)
    gamma, beta = rnd(C) * 0.1 + 1.0, rnd(C) * 0.1
    out = torch.zeros(rows, C, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    kw = {}
    if with_add:
        kw = dict(add=rnd(7, C), rows_per_group=10, add_period=7, sum_out=torch.zeros_like(out))
    lib.layernorm(x, gamma, beta, out, 1e-5, **kw)
    kwr = dict(kw)
    if with_add:
        kwr["sum_out"] = torch.zeros_like(out)
    R.layernorm(x, gamma, beta, ref, 1e-5, **kwr)
    torch.cuda.synchronize()
    close(out, ref, 4e-3, 4e-3, "layernorm")
    if with_add:
        close(kw["sum_out"], kwr["sum_out"], 1e-3, 1e-3, "layernorm sum_out")


def test_axpy_im2col_upsample_layout():
    lib = L()
    x, y = rnd(50 * 64, 320), rnd(25 * 64, 320)
    out, ref = torch.zeros_like(x), torch.zeros_like(x)
    lib.axpy_bcast(x, y, out, 3.0)
    R.axpy_bcast(x, y, ref, 3.0)
    close(out, ref, 2e-3, 2e-3, "axpy_bcast")
    for (n, H, W, C, s, Kpad) in [(2, 8, 12, 64, 2, 576), (2, 9, 7, 3, 1, 32), (1, 16, 16, 8, 1, 72), (2, 6, 6, 16, 2, 144)]:
        xi = rnd(n * H * W, C)
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        o = torch.zeros(n * Ho * Wo, Kpad, dtype=torch.half, device=DEV)
        r = torch.zeros_like(o)
        lib.im2col3x3(xi, o, n, H, W, C, s, Kpad)
        R.im2col3x3(xi, r, n, H, W, C, s, Kpad)
        assert torch.equal(o, r), f"im2col {n,H,W,C,s}"
    xi = rnd(2 * 5 * 6, 64)
    o = torch.zeros(2 * 10 * 12, 64, dtype=torch.half, device=DEV)
    r = torch.zeros_like(o)
    lib.upsample2x(xi, o, 2, 5, 6, 64)
    R.upsample2x(xi, r, 2, 5, 6, 64)
    assert torch.equal(o, r)
    xc = rnd(3, 4, 35)
    o = torch.zeros(3, 35, 8, dtype=torch.half, device=DEV)
    r = torch.zeros_like(o)
    lib.nchw_to_nhwc(xc, o, 3, 4, 35, 8, 4)
    R.nchw_to_nhwc(xc, r, 3, 4, 35, 8, 4)
    assert torch.equal(o, r)
    back = torch.zeros(3, 4, 35, dtype=torch.half, device=DEV)
    lib.nhwc_to_nchw(o, back, 3, 4, 35, 8, 4)
    assert torch.equal(back, xc)


def test_linear_small_and_timestep():
    lib = L()
    a, w, b = rnd(2, 1280), rnd(640, 1280, scale=0.05), rnd(640)
    out = torch.zeros(2, 640, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.linear_small(a, w, b, out, 1, 1)
    R.linear_small(a, w, b, ref, 1, 1)
    close(out, ref, 3e-3, 1e-2, "linear_small")
    t = torch.tensor([1.6377, 6.0, 128.0, 0.02], dtype=torch.float32, device=DEV)
    out = torch.zeros(4, 320, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.timestep_embedding(t, out, 320)
    R.timestep_embedding(t, ref, 320)
    close(out, ref, 2e-3, 0, "timestep_embedding")


@pytest.mark.parametrize("hs,ws,C,s", [(9, 16, 64, 8), (18, 32, 320, 4), (6, 6, 1280, 2)])
def test_softsplat(hs, ws, C, s):
    lib = L()
    Fn = 5
    feat = rnd(hs * ws, C)
    flow = (torch.randn(Fn, 2, hs * s, ws * s) * 3.0 * s).half().to(DEV)
    flow[0, 0, 0, 0] = float("inf")
    acc = torch.zeros(Fn * hs * ws * C, dtype=torch.float32, device=DEV)
    wsum = torch.zeros(Fn * hs * ws, dtype=torch.float32, device=DEV)
    out = torch.zeros(Fn * hs * ws, C, dtype=torch.half, device=DEV)
    ref = torch.zeros_like(out)
    lib.softsplat_avg(feat, flow, acc, wsum, out, Fn, hs, ws, C, hs * s, ws * s)
    R.softsplat_avg(feat, flow, None, None, ref, Fn, hs, ws, C, hs * s, ws * s)
    torch.cuda.synchronize()
    close(out, ref, 2e-3, 2e-3, "softsplat")
    # identity: zero flow => out == in / (1 + 1e-7)
    flow0 = torch.zeros_like(flow)
    lib.softsplat_avg(feat, flow0, acc, wsum, out, Fn, hs, ws, C, hs * s, ws * s)
    close(out.view(Fn, hs * ws, C)[2], feat, 1e-3, 1e-3, "softsplat zero-flow identity")


def test_cfg_euler():
    lib = L()
    T, HW = 5, 96
    noise = rnd(2 * T * HW, 4)
    lat = rnd(T, 4, HW, scale=5.0)
    img = rnd(2, 4, HW)
    nxt = torch.zeros(2 * T * HW, 8, dtype=torch.half, device=DEV)
    lat_r, nxt_r = lat.clone(), torch.zeros_like(nxt)
    lib.cfg_euler_step(noise, lat, img, nxt, T, HW, 1.0, 3.0, 7.5, 4.2)
    R.cfg_euler_step(noise, lat_r, img, nxt_r, T, HW, 1.0, 3.0, 7.5, 4.2)
    torch.cuda.synchronize()
    close(lat, lat_r, 1e-2, 2e-3, "euler latents")
    close(nxt, nxt_r, 5e-3, 2e-3, "euler next_in")
    lib.cfg_euler_step(None, lat, img, nxt, T, HW, 1.0, 3.0, 0.0, 7.5)
    R.cfg_euler_step(None, lat_r, img, nxt_r, T, HW, 1.0, 3.0, 0.0, 7.5)
    close(nxt, nxt_r, 5e-3, 2e-3, "euler prologue")
