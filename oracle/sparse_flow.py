"""ORACLE (test infrastructure): the control-signal rasterisation, restated statement for statement.

  get_sparseflow_and_mask_forward  <- /root/reference/MOFA-Video-Traj/run_gradio.py:61-86 (numpy, float64)
  sample_optical_flow / get_sparse_flow <- /root/reference/MOFA-Video-Keypoint/utils/utils.py:81-119 (torch, CPU)
Pinned: tests/golden/sparse_flow_ref.pt holds outputs of the reference's own two functions, extracted from its files
with ast and executed (oracle/make_goldens.py --sparse-flow); tests/test_oracle.py compares bit-exactly."""
import numpy as np
import torch


def get_sparseflow_and_mask_forward(resized_all_points, n_steps, H, W, is_backward_flow=False):
    K = resized_all_points.shape[0]
    starts = resized_all_points[:, 0]                 # :70
    interpolated_ends = resized_all_points[:, 1:]     # :72
    s_flow = np.zeros((K, n_steps, H, W, 2))
    mask = np.zeros((K, n_steps, H, W))
    for k in range(K):
        for i in range(n_steps):
            start, end = starts[k], interpolated_ends[k][i]
            flow = np.int64(end - start) * (-1 if is_backward_flow is True else 1)   # :78
            s_flow[k][i][int(start[1]), int(start[0])] = flow
            mask[k][i][int(start[1]), int(start[0])] = 1
    return np.sum(s_flow, axis=0), np.sum(mask, axis=0)   # :83-84


def sample_optical_flow(A, B, h, w):
    b, l, k, _ = A.shape
    sparse_optical_flow = torch.zeros((b, l, h, w, 2), dtype=B.dtype, device=B.device)
    mask = torch.zeros((b, l, h, w), dtype=torch.uint8, device=B.device)
    x_coords = torch.clip(A[..., 0].long(), 0, h - 1)       # :87-91
    y_coords = torch.clip(A[..., 1].long(), 0, w - 1)
    b_idx = torch.arange(b)[:, None, None].repeat(1, l, k)
    l_idx = torch.arange(l)[None, :, None].repeat(b, 1, k)
    sparse_optical_flow[b_idx, l_idx, x_coords, y_coords] = B   # :96  (CPU: sequential, the last duplicate wins)
    mask[b_idx, l_idx, x_coords, y_coords] = 1
    return sparse_optical_flow, mask.unsqueeze(-1).repeat(1, 1, 1, 1, 2)


def get_sparse_flow(landmarks, h, w, t):
    landmarks = torch.flip(landmarks, dims=[3])                                   # :109
    pose_flow = (landmarks - landmarks[:, 0:1].repeat(1, t, 1, 1))[:, 1:]         # :111
    according_poses = landmarks[:, 0:1].repeat(1, t - 1, 1, 1)
    pose_flow = torch.flip(pose_flow, dims=[3])                                   # :114
    sparse_optical_flow, mask = sample_optical_flow(according_poses, pose_flow, h, w)
    return sparse_optical_flow.permute(0, 1, 4, 2, 3), mask.permute(0, 1, 4, 2, 3)
