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
