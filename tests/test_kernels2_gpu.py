"""-m gpu: VAE-decoder helper kernels against their plain-PyTorch statements (tests/ref_ops.py)."""
import pytest
import torch

import ref_ops as R

pytestmark = pytest.mark.gpu
DEV = "cuda"


def test_softmax_rows():
    from mofa_video_b200 import lib
    g = torch.Generator().manual_seed(0)
    for rows, L in [(64, 9216), (33, 144), (128, 576), (5, 16384)]:
        x = (torch.randn(rows, L, generator=g) * 3).half().to(DEV)
        ref = x.clone()
        lib.softmax_rows(x)
        R.softmax_rows(ref)
        torch.cuda.synchronize()
        assert (x.float() - ref.float()).abs().max().item() < 2e-3 * max(ref.float().max().item(), 1e-3) + 1e-6
        assert abs(x.float().sum(-1).mean().item() - 1.0) < 2e-2


def test_vae_time_conv_out():
    from mofa_video_b200 import lib
    g = torch.Generator().manual_seed(1)
    T, HW = 5, 1000
    y = torch.randn(T * HW, 3, generator=g).half().to(DEV)
    w = (torch.randn(3, 3, 3, generator=g) * 0.3).to(DEV)
    b = (torch.randn(3, generator=g) * 0.1).to(DEV)
    o32, ou8 = torch.zeros(T, 3, HW, device=DEV), torch.zeros(T, HW, 3, dtype=torch.uint8, device=DEV)
    r32, ru8 = torch.zeros_like(o32), torch.zeros_like(ou8)
    lib.vae_time_conv_out(y, w, b, o32, ou8, T, HW)
    R.vae_time_conv_out(y, w, b, r32, ru8, T, HW)
    torch.cuda.synchronize()
    assert torch.allclose(o32, r32, atol=1e-4, rtol=1e-4)
    assert (ou8.int() - ru8.int()).abs().max().item() <= 1


def test_resize_antialias_matches_reference_resize():
    """CLIP-side resize (pipeline.py:532-640): fused blur + bicubic kernel vs the two-pass PyTorch restatement."""
    from mofa_video_b200 import lib
    g = torch.Generator().manual_seed(0)
    for (H, W, Ho, Wo) in [(576, 1024, 224, 224), (128, 192, 224, 224), (384, 384, 224, 224), (300, 700, 64, 96)]:
        img = torch.rand(1, 3, H, W, generator=g).cuda()
        out = torch.empty(1, 3, Ho, Wo, dtype=torch.float32, device="cuda")
        ref = torch.empty_like(out)
        lib.resize_antialias(img, out)
        R.resize_antialias(img, ref)
        assert (out - ref).abs().max().item() < 2e-5, (H, W)
