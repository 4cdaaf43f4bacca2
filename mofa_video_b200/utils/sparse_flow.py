"""Sparse motion hints -> dense (flow, mask) planes on the device, with the reference's names.

  get_sparseflow_and_mask_forward   /root/reference/MOFA-Video-Traj/run_gradio.py:61-86 (also H/run_gradio_*_driven.py:175)
  get_sparse_flow                   /root/reference/MOFA-Video-Keypoint/utils/utils.py:106-119 (+ sample_optical_flow :81-103)

The reference builds these in numpy (O(K * T * H * W) zero-filled per-track planes summed on the CPU) / PyTorch index_put_
and then ships 28 MB per call to the GPU; here the points go to the device (a few hundred bytes) and one scatter kernel
writes the planes where CMP reads them.  Results are device tensors (the reference returns numpy float64 / torch
tensors); values are identical: integer-valued flows, mask counts, float32 landmark differences."""
import numpy as np
import torch

from mofa_video_b200 import lib as _lib


def get_sparseflow_and_mask_forward(resized_all_points, n_steps, H, W, is_backward_flow=False, device="cuda", ops=None):
    """resized_all_points [K, n_steps + 1, 2] (x, y): track starts and interpolated ends.  Returns (s_flow fp32
    [n_steps, H, W, 2], mask fp32 [n_steps, H, W]); tracks that start on the same pixel ADD (np.sum over k, :83-84)."""
    ops = ops if ops is not None else _lib
    pts = np.asarray(resized_all_points, dtype=np.float64)
    K = pts.shape[0]
    if pts.ndim != 3 or pts.shape[1] < n_steps + 1 or pts.shape[2] != 2:
        raise ValueError(f"resized_all_points must be [K, >= n_steps + 1, 2], got {pts.shape}")
    pts = np.ascontiguousarray(pts[:, :n_steps + 1])
    # numpy indexing semantics of `s_flow[k][i][int(start[1]), int(start[0])] = flow` (:79): IndexError outside
    # [-size, size), negative indices wrap
    sx, sy = pts[:, 0, 0].astype(np.int64), pts[:, 0, 1].astype(np.int64)   # astype truncates toward zero like int()
    for v, size, axis in ((sy, H, 0), (sx, W, 1)):
        bad = (v < -size) | (v >= size)
        if bad.any():
            raise IndexError(f"index {int(v[bad][0])} is out of bounds for axis {axis} with size {size}")
    flow = torch.empty(n_steps, H, W, 2, dtype=torch.float32, device=device)
    mask = torch.empty(n_steps, H, W, dtype=torch.float32, device=device)
    if K == 0:
        return flow.zero_(), mask.zero_()
    ops.sparse_hints_add(torch.from_numpy(pts).to(device), flow, mask, -1 if is_backward_flow is True else 1)
    return flow, mask


@torch.no_grad()
def get_sparse_flow(landmarks, h, w, t, ops=None):
    """landmarks [b, t, K, 2] (x, y) on the device.  Returns (sparse_optical_flow [b, t-1, 2, h, w] in landmarks' dtype,
    mask uint8 [b, t-1, 2, h, w]): flow of every landmark relative to frame 0, written at its frame-0 pixel."""
    ops = ops if ops is not None else _lib
    if landmarks.dtype not in (torch.float32, torch.float64):
        landmarks = landmarks.float()
    lm = landmarks[:, :t].contiguous()
    b, _, K, _ = lm.shape
    flow = torch.empty(b, t - 1, 2, h, w, dtype=lm.dtype, device=lm.device)
    mask = torch.empty(b, t - 1, 2, h, w, dtype=torch.uint8, device=lm.device)
    owner = torch.empty(b, t - 1, h, w, dtype=torch.int32, device=lm.device)
    ops.sparse_hints_assign(lm, flow, mask, owner)
    return flow, mask
