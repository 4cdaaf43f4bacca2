"""Trajectory adapter under the module path the Keypoint / Hybrid sub-projects use
(/root/reference/MOFA-Video-Hybrid/models/traj_ctrlnet.py:185-; same network as the Traj FlowControlNet,
get_warped_frames identical to svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py:223-234)."""
from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import (  # noqa: F401
    FlowControlNet, FlowControlNetOutput)
