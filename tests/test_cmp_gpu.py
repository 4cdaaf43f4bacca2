"""-m gpu: CMP helper kernels against tests/ref_ops.py, and the native CMP network against the reference-pinned oracle."""
import pytest
import torch

import ref_ops as R
from oracle import cmp as ocmp

pytestmark = pytest.mark.gpu
DEV = "cuda"


def h(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).half().to(DEV)


def test_im2col_pool_resize_fuser_copy():
    from mofa_video_b200 import lib
    for (n, H, W, C, k, s, p, d) in [(2, 12, 10, 3, 7, 2, 3, 1), (1, 9, 9, 4, 5, 2, 2, 1), (2, 8, 8, 16, 3, 2, 1, 1),
                                     (2, 8, 6, 32, 1, 2, 0, 1), (1, 10, 10, 8, 3, 1, 2, 2)]:
        x = h(n * H * W, C, seed=k)
        Ho, Wo = lib.conv_out_size(H, k, s, p, d), lib.conv_out_size(W, k, s, p, d)
        kpad = (k * k * C + 7) // 8 * 8
        o = torch.zeros(n * Ho * Wo, kpad, dtype=torch.half, device=DEV)
        r = torch.zeros_like(o)
        lib.im2col(x, o, n, H, W, C, k, s, p, d, kpad)
        R.im2col(x, r, n, H, W, C, k, s, p, d, kpad)
        assert torch.equal(o, r), (k, s, p, d)
    for (k, s, p, mode) in [(3, 2, 1, 0), (2, 2, 0, 0), (8, 8, 0, 0), (2, 2, 0, 1)]:
        n, H, W, C = 2, 16, 24, 40
        x = h(n * H * W, C, seed=3)
        Ho, Wo = (H + 2 * p - k) // s + 1, (W + 2 * p - k) // s + 1
        o = torch.zeros(n * Ho * Wo, C, dtype=torch.half, device=DEV)
        r = torch.zeros_like(o)
        lib.pool2d(x, o, n, H, W, C, k, s, p, mode)
        R.pool2d(x, r, n, H, W, C, k, s, p, mode)
        assert (o.float() - r.float()).abs().max().item() < 2e-3
    n, H, W, C = 2, 6, 7, 24
    x = h(n * H * W, C, seed=4)
    o = torch.zeros(n * 12 * 15, 64, dtype=torch.half, device=DEV)
    r = torch.zeros_like(o)
    lib.resize_bilinear_ac(x, o, n, H, W, C, 12, 15, 64, 16)
    R.resize_bilinear_ac(x, r, n, H, W, C, 12, 15, 64, 16)
    assert (o.float() - r.float()).abs().max().item() < 4e-3
    lg = h(500, 198, seed=5, scale=3.0)
    o, r = torch.zeros(500, 2, dtype=torch.half, device=DEV), torch.zeros(500, 2, dtype=torch.half, device=DEV)
    lib.cmp_fuser(lg, o)
    R.cmp_fuser(lg, r)
    assert (o.float() - r.float()).abs().max().item() < 5e-2
    src = h(30, 16, seed=6)
    o, r = torch.zeros(90, 40, dtype=torch.half, device=DEV), torch.zeros(90, 40, dtype=torch.half, device=DEV)
    lib.copy_cols(src, o, 90, 16, 30, 40, 8)
    R.copy_cols(src, r, 90, 16, 30, 40, 8)
    assert torch.equal(o, r)


def test_gemm_relu_sigmoid_dilation_and_relu_after_residual():
    from mofa_video_b200 import lib
    M, K, N = 500, 128, 128
    a, w, b, res = h(M, K, seed=1, scale=0.5), h(N, K, seed=2, scale=0.05), h(N, seed=3), h(M, N, seed=4)
    for act in (R.ACT_RELU, R.ACT_SIGMOID, 5):
        o, r = torch.zeros(M, N, dtype=torch.half, device=DEV), torch.zeros(M, N, dtype=torch.half, device=DEV)
        lib.linear(a, w, o, bias=b, act=act, res1=res)
        R.linear(a, w, r, bias=b, act=act, res1=res)
        assert (o.float() - r.float()).abs().max().item() < 2e-2, act
    for d in (2, 4):
        n, H, W, C = 2, 20, 24, 64
        x, wk = h(n * H * W, C, seed=5, scale=0.5), h(64, 9 * C, seed=6, scale=0.03)
        o, r = torch.zeros(n * H * W, 64, dtype=torch.half, device=DEV), torch.zeros(n * H * W, 64, dtype=torch.half, device=DEV)
        kw = dict(N=64, n_img=n, H=H, W=W, C=C, dilation=d, act=R.ACT_RELU)
        lib.gemm(lib.A_CONV3X3, x, wk, o, **kw)
        R.gemm(R.A_CONV3X3, x, wk, r, **kw)
        assert (o.float() - r.float()).abs().max().item() < 2e-2, d


def test_cmp_network_matches_pinned_oracle():
    from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import CMP_demo
    m = ocmp.CMP().eval()
    sd = ocmp.seeded_state_dict(m, seed=3)
    sd["flow_decoder.head.weight"] = sd["flow_decoder.head.weight"] * 25.0
    m.load_state_dict(sd)
    g = torch.Generator().manual_seed(5)
    B, H, W = 4, 128, 192
    image = torch.rand(1, 3, H, W, generator=g).repeat(B, 1, 1, 1)
    sparse, mask = torch.zeros(B, 2, H, W), torch.zeros(B, 2, H, W)
    for b in range(B):
        for _ in range(12):
            y, x = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, W, (1,), generator=g))
            sparse[b, :, y, x] = torch.randn(2, generator=g) * 10
            mask[b, :, y, x] = 1
    with torch.no_grad():
        ref = ocmp.cmp_demo_run(m, image, sparse, mask)
    demo = CMP_demo(state_dict={"module." + k: v for k, v in sd.items()})
    out = demo.run(image.cuda(), sparse.cuda(), mask.cuda())
    assert out.shape == ref.shape and out.dtype == image.dtype
    err = (out.float().cpu() - ref).abs()
    assert ref.abs().max() > 1.0
    assert err.mean().item() < 0.03 and err.max().item() < 0.3, (err.mean().item(), err.max().item())
