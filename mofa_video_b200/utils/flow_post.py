"""Drag-flow post-processing between CMP and the pipeline: the arithmetic of
/root/reference/MOFA-Video-Traj/run_gradio.py:251-255 (motion-brush mask), :268-275 (nearest resize 384x384 -> HxW with
per-axis rescale) and :330-333 (in-mask / out-mask merge), as ONE kernel over the output flow."""
import torch

from mofa_video_b200 import lib as _lib


def drag_flow_postprocess(flow_inmask, height, width, brush_mask=None, flow_outmask=None, ops=None):
    """flow_inmask / flow_outmask: [B, T-1, 2, hs, ws] CMP flows (any float dtype, on the GPU); brush_mask: [hs, ws] in
    [0, 1] or None.  Returns the fp16 controlnet_flow [B, T-1, 2, height, width] the pipeline takes."""
    ops = ops if ops is not None else _lib
    b, t, c, hs, ws = flow_inmask.shape
    assert c == 2
    fin = flow_inmask.to(torch.float16).contiguous()
    fout = flow_outmask.to(torch.float16).contiguous() if flow_outmask is not None else None
    brush = brush_mask.to(device=fin.device, dtype=torch.float16).contiguous() if brush_mask is not None else None
    out = torch.empty(b, t, 2, height, width, dtype=torch.float16, device=fin.device)
    ops.flow_post(fin, out, b * t, hs, ws, height, width, brush=brush, flow_out=fout)
    return out
