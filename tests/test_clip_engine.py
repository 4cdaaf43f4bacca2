"""CLIP image encoder on the engine vs transformers' CLIPVisionModelWithProjection in fp32.
CPU: host logic / packing with the PyTorch statements of the kernels, both MLP activations.  GPU: the real ViT-H/14
(32 layers, 16 x 80 heads, 257 tokens) through mofa_attn_small + the tcgen05 GEMMs; tolerance 1e-2 * max|ref| on the
projected embedding (fp16 storage through 32 layers; measured in the test output)."""
import pytest
import torch

import ref_ops


def _model(hidden, inter, layers, heads, image, proj, act, gain=3.0, seed=0):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    torch.manual_seed(seed)
    cfg = CLIPVisionConfig(hidden_size=hidden, intermediate_size=inter, num_hidden_layers=layers,
                           num_attention_heads=heads, image_size=image, patch_size=14, projection_dim=proj,
                           hidden_act=act)
    m = CLIPVisionModelWithProjection(cfg).eval()
    with torch.no_grad():
        for p in m.parameters():
            p.copy_((p * gain).half().float())
    return m


@pytest.mark.parametrize("act", ["gelu", "quick_gelu"])
def test_clip_engine_host_logic_cpu(act):
    from mofa_video_b200.clip_engine import NativeClipVision, is_clip_vision_with_projection
    m = _model(128, 256, 3, 4, 56, 64, act)
    assert is_clip_vision_with_projection(m) and not is_clip_vision_with_projection(torch.nn.Linear(3, 3))
    x = torch.rand(2, 3, 56, 56, generator=torch.Generator().manual_seed(1)).half().float()
    with torch.no_grad():
        ref = m(x).image_embeds
    nat = NativeClipVision(m, ops=ref_ops, device="cpu")
    got = nat(x).image_embeds.float()
    assert got.shape == ref.shape
    assert ((got - ref).abs().max() / ref.abs().max()).item() < 8e-3
    assert next(nat.parameters()).dtype == torch.float16
    with pytest.raises(ValueError):
        nat(torch.rand(1, 3, 28, 28))


@pytest.mark.gpu
def test_attn_small_kernel():
    from mofa_video_b200 import lib
    g = torch.Generator().manual_seed(2)
    for (n, L, heads, d) in ((1, 257, 16, 80), (2, 50, 4, 32), (1, 17, 3, 128)):
        C = heads * d
        qkv = (torch.randn(n * L, 3 * C, generator=g)).half().cuda()
        out, ref = torch.empty(n * L, C, dtype=torch.half, device="cuda"), torch.empty(n * L, C, dtype=torch.half)
        lib.attn_small(qkv, out, n, L, heads, d, d ** -0.5)
        ref_ops.attn_small(qkv.cpu(), ref, n, L, heads, d, d ** -0.5)
        assert (out.float().cpu() - ref.float()).abs().max().item() < 3e-3, (n, L, heads, d)


@pytest.mark.gpu
def test_gemm_gelu_epilogues():
    from mofa_video_b200 import lib
    g = torch.Generator().manual_seed(3)
    a = torch.randn(257, 128, generator=g).half().cuda()
    w = (torch.randn(320, 128, generator=g) * 0.2).half().cuda()
    b = torch.randn(320, generator=g).half().cuda()
    for act in (lib.ACT_GELU, lib.ACT_QUICK_GELU):
        out, ref = torch.empty(257, 320, dtype=torch.half, device="cuda"), torch.empty(257, 320, dtype=torch.half)
        lib.linear(a, w, out, bias=b, act=act)
        ref_ops.linear(a.cpu(), w.cpu(), ref, bias=b.cpu(), act=act)
        assert (out.float().cpu() - ref.float()).abs().max().item() < 6e-3


@pytest.mark.gpu
def test_clip_vit_h_native_vs_fp32_module():
    from mofa_video_b200.clip_engine import NativeClipVision
    m = _model(1280, 5120, 32, 16, 224, 1024, "gelu", gain=1.0)
    x = torch.rand(1, 3, 224, 224, generator=torch.Generator().manual_seed(4)).half().float()
    with torch.no_grad():
        ref = m(x).image_embeds
    nat = NativeClipVision(m)
    got = nat(x.cuda()).image_embeds.float().cpu()
    e = ((got - ref).abs().max() / ref.abs().max()).item()
    print(f"\n[fullsize] CLIP ViT-H/14 native vs fp32 transformers: max rel err {e:.2e}")
    assert got.shape == (1, 1024) and torch.isfinite(got).all() and e < 1e-2
