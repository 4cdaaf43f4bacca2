"""softsplat() with the reference's signature (/root/reference/MOFA-Video-Traj/models/softsplat.py:232-274),
backed by mofa_softsplat_avg.  Only strMode='avg' with tenMetric=None -- the one mode the adapter uses
(FCN.py:231) -- is implemented; other modes raise (the reference asserts on bad modes, :233-238)."""
import torch

from mofa_video_b200 import lib


def softsplat(tenIn: torch.Tensor, tenFlow: torch.Tensor, tenMetric, strMode: str):
    assert strMode.split('-')[0] in ['sum', 'avg', 'linear', 'soft']
    if strMode != 'avg' or tenMetric is not None:
        raise NotImplementedError("mofa_video_b200.softsplat implements strMode='avg', tenMetric=None only")
    if not tenIn.is_cuda:
        raise RuntimeError("softsplat needs CUDA tensors (the reference's CPU branch is assert(False), softsplat.py:347)")
    N, C, H, W = tenIn.shape
    assert tenFlow.shape == (N, 2, H, W)
    if C % 4:
        raise ValueError("channel count must be a multiple of 4")
    outs = []
    for n in range(N):
        feat = torch.empty(H * W, C, dtype=torch.float16, device=tenIn.device)
        lib.nchw_to_nhwc(tenIn[n:n + 1].half().contiguous(), feat, 1, C, H * W)
        acc = torch.empty(H * W * C, dtype=torch.float32, device=tenIn.device)
        wsum = torch.empty(H * W, dtype=torch.float32, device=tenIn.device)
        out = torch.empty(H * W, C, dtype=torch.float16, device=tenIn.device)
        lib.softsplat_avg(feat, tenFlow[n:n + 1].half().contiguous(), acc, wsum, out, 1, H, W, C, H, W)
        o = torch.empty(1, C, H, W, dtype=torch.float16, device=tenIn.device)
        lib.nhwc_to_nchw(out, o, 1, C, H * W)
        outs.append(o)
    return torch.cat(outs, 0).to(tenIn.dtype)
