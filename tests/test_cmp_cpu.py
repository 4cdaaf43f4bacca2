"""-m "not gpu": host logic of the native CMP network (BatchNorm folding, conv dispatch, concat assembly, dedupe of
identical frames) with kernels replaced by tests/ref_ops.py, against the reference-pinned oracle (oracle/cmp.py).

Tolerance: the reference runs CMP in fp32 (FCN.py:54 `.float()`); the engine stores activations in fp16.  The flow is an
expectation over 99 bins spanning +-50 px: mean |err| <= 0.03 px and max |err| <= 0.3 px are required here (measured 0.004 / 0.034)."""
import torch

import ref_ops
from mofa_video_b200.cmp_engine import CmpNet
from oracle import cmp as ocmp


def _inputs(B, H, W, seed=5, identical=False):
    g = torch.Generator().manual_seed(seed)
    image = torch.rand(1 if identical else B, 3, H, W, generator=g)
    if identical:
        image = image.repeat(B, 1, 1, 1)
    sparse, mask = torch.zeros(B, 2, H, W), torch.zeros(B, 2, H, W)
    for b in range(B):
        for _ in range(10):
            y, x = int(torch.randint(0, H, (1,), generator=g)), int(torch.randint(0, W, (1,), generator=g))
            sparse[b, :, y, x] = torch.randn(2, generator=g) * 10
            mask[b, :, y, x] = 1
    return image, sparse, mask


def _run(B, H, W, identical):
    m = ocmp.CMP().eval()
    sd = ocmp.seeded_state_dict(m, seed=3)
    # sharpen the logits so the expected flow has pixel-scale magnitude (random weights give near-uniform bins)
    sd["flow_decoder.head.weight"] = sd["flow_decoder.head.weight"] * 25.0
    m.load_state_dict(sd)
    image, sparse, mask = _inputs(B, H, W, identical=identical)
    with torch.no_grad():
        ref = ocmp.cmp_demo_run(m, image, sparse, mask)
    net = CmpNet({"module." + k: v for k, v in sd.items()}, ref_ops, "cpu")
    out = net.forward(image, sparse, mask).float()
    err = (out - ref).abs()
    return err.mean().item(), err.max().item(), ref.abs().max().item()


def test_cmp_engine_distinct_frames():
    mean, mx, scale = _run(2, 128, 128, identical=False)
    assert scale > 1.0
    assert mean < 0.03 and mx < 0.3, (mean, mx)


def test_cmp_engine_identical_frames_dedupe():
    mean, mx, _ = _run(3, 64, 128, identical=True)
    assert mean < 0.03 and mx < 0.3, (mean, mx)


def test_cmp_demo_entry_point_reads_reference_layout(tmp_path):
    """CMP_demo(configfn, load_iter) as T/run_gradio.py builds it (FCN.py:26-49): the experiment's config.yaml
    (model.module.{image_encoder, flow_decoder, nbins, fmax}) and `checkpoints/ckpt_iter_<n>.pth.tar` with the
    'state_dict' key and 'module.'-prefixed names; .run() -> flow in the caller's dtype."""
    import warnings

    import yaml

    from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import CMP_demo
    exp = tmp_path / "experiments" / "semiauto_annot" / "resnet50_vip+mpii_liteflow"
    (exp / "checkpoints").mkdir(parents=True)
    cfg = {"model": {"arch": "CMP", "module": {"arch": "CMP", "image_encoder": "resnet50", "sparse_encoder": "shallownet8x",
                                               "flow_decoder": "MotionDecoderSkipLayer", "skip_layer": True,
                                               "img_enc_dim": 256, "sparse_enc_dim": 16, "output_dim": 198,
                                               "decoder_combo": [1, 2, 4], "nbins": 99, "fmax": 50}}}
    with open(exp / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    m = ocmp.CMP().eval()
    sd = ocmp.seeded_state_dict(m, seed=3)
    sd["flow_decoder.head.weight"] = sd["flow_decoder.head.weight"] * 25.0
    m.load_state_dict(sd)
    torch.save({"step": 42000, "state_dict": {"module." + k: v for k, v in sd.items()}},
               exp / "checkpoints" / "ckpt_iter_42000.pth.tar")
    demo = CMP_demo(str(exp / "config.yaml"), 42000, device="cpu", ops=ref_ops).to("cpu")
    demo.requires_grad_(False)
    image, sparse, mask = _inputs(2, 128, 128)
    out = demo.run(image, sparse, mask)
    with torch.no_grad():
        ref = ocmp.cmp_demo_run(m, image, sparse, mask)
    assert out.dtype == image.dtype and out.shape == ref.shape
    assert (out - ref).abs().max().item() < 0.3
    with warnings.catch_warnings(record=True) as w:       # missing checkpoint: warning + random init, like the reference
        warnings.simplefilter("always")
        CMP_demo(str(exp / "config.yaml"), 7, device="cpu", ops=ref_ops)
        assert any("no checkpoint found" in str(x.message) for x in w)
    cfg["model"]["module"]["image_encoder"] = "alexnet_fcn_32x"
    with open(exp / "config.yaml", "w") as f:
        yaml.safe_dump(cfg, f)
    import pytest
    with pytest.raises(NotImplementedError):
        CMP_demo(str(exp / "config.yaml"), 42000, device="cpu", ops=ref_ops)
