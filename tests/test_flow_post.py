"""drag-flow post-processing -- host helper + op statement vs the oracle (CPU), kernel vs the op
statement (GPU)."""
import pytest
import torch

import ref_ops as R
from oracle import pipeline as opipe


def _inputs(seed=0, F_=5, hs=24, ws=24):
    g = torch.Generator().manual_seed(seed)
    fin = (torch.randn(1, F_, 2, hs, ws, generator=g) * 6).half()
    fin[:, :, :, 3:9, 2:12] = 0                      # a hole the out-mask flow must fill
    fin[:, 1, 0, 15, 15] = 0                         # only one channel zero: still replaced (all(dim=channel))
    fout = (torch.randn(1, F_, 2, hs, ws, generator=g) * 3).half()
    brush = (torch.rand(hs, ws, generator=g) > 0.3).half()
    return fin, fout, brush


@pytest.mark.parametrize("H,W,use_brush,use_out", [(36, 60, True, True), (24, 24, False, True), (48, 24, True, False)])
def test_flow_post_statement_matches_oracle(H, W, use_brush, use_out):
    from mofa_video_b200.utils.flow_post import drag_flow_postprocess
    fin, fout, brush = _inputs()
    ref = opipe.drag_flow_post(fin, H, W, brush if use_brush else None, fout if use_out else None)

    class Ops:  # the helper drives R.flow_post on CPU tensors
        flow_post = staticmethod(R.flow_post)
    out = drag_flow_postprocess(fin, H, W, brush if use_brush else None, fout if use_out else None, ops=Ops)
    assert out.shape == ref.shape and torch.equal(out, ref.half())


@pytest.mark.gpu
@pytest.mark.parametrize("H,W,use_brush,use_out", [(576, 1024, True, True), (384, 384, False, True),
                                                   (320, 512, True, False)])
def test_flow_post_kernel(H, W, use_brush, use_out):
    from mofa_video_b200 import lib
    fin, fout, brush = _inputs(seed=1, F_=24, hs=384, ws=384)
    fin, fout, brush = fin.cuda(), fout.cuda(), brush.cuda()
    out = torch.zeros(24, 2, H, W, dtype=torch.half, device="cuda")
    ref = torch.zeros_like(out)
    lib.flow_post(fin, out, 24, 384, 384, H, W, brush=brush if use_brush else None, flow_out=fout if use_out else None)
    R.flow_post(fin, ref, 24, 384, 384, H, W, brush=brush if use_brush else None, flow_out=fout if use_out else None)
    assert torch.equal(out, ref)
