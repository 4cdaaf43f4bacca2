"""UNetSpatioTemporalConditionControlNetModel with the reference's entry point
(/root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py:32-504), executed by the
sm_100a engine (mofa_video_b200/engine.py).  Same constructor source (from_pretrained on the SVD `unet`
sub-folder), same forward signature and return convention (:356-365, :501-504), same attributes the pipeline
reads (.config.in_channels / .num_frames / .sample_size / .addition_time_embed_dim,
.add_embedding.linear_1.in_features -- pipeline.py:34-35, 314-317, 377)."""
from dataclasses import dataclass

import torch

from mofa_video_b200.models._base import EngineModel


@dataclass
class UNetSpatioTemporalConditionOutput:
    sample: torch.Tensor = None


class UNetSpatioTemporalConditionControlNetModel(EngineModel):
    kind = "unet"

    def forward(self, sample, timestep, encoder_hidden_states, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True, added_time_ids=None):
        """sample [B, T, C, h, w]; residuals as returned by FlowControlNet.forward ([B*T, C_k, h_k, w_k]).
        Returns (sample,) / UNetSpatioTemporalConditionOutput with sample [B, T, out_channels, h, w] fp16."""
        if down_block_additional_residuals is None or mid_block_additional_residual is None:
            raise ValueError("this UNet variant requires the ControlNet residuals (UNET.py:451-469 adds them "
                             "unconditionally)")
        b, t, c, h, w = sample.shape
        if t != self.config.num_frames:
            raise ValueError(f"num_frames mismatch: sample has {t}, model packed for {self.config.num_frames}")
        self._prepare(encoder_hidden_states, added_time_ids)
        x = self._to_cl(sample)
        res = [self._res_to_cl(r) for r in down_block_additional_residuals]
        mid = self._res_to_cl(mid_block_additional_residual)
        out = self.net.unet_forward(x, self._t_value(timestep), h, w, res, mid)
        out = self._from_cl(out, b * t, h, w).reshape(b, t, -1, h, w)
        if not return_dict:
            return (out,)
        return UNetSpatioTemporalConditionOutput(sample=out)

    __call__ = forward

    def _res_to_cl(self, r):
        if getattr(r, "_mofa_channels_last", False):
            return r
        n, c, h, w = r.shape
        out = torch.empty(n * h * w, c, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(r.to(device=self._device, dtype=torch.float16).contiguous(), out, n, c, h * w)
        return out
