"""Keypoint (facial-landmark) FlowControlNet with the reference's entry point
(/root/reference/MOFA-Video-Keypoint/models/ldmk_ctrlnet.py:187-575), executed by the sm_100a engine
(mofa_video_b200/keypoint_engine.py).  forward keeps the reference signature (:322-339) including `landmarks`
and returns (down_block_res_samples, mid_block_res_sample, controlnet_flow, occlusion_masks) (:569-574)."""
from dataclasses import dataclass
from typing import List, Optional, Tuple

import torch

from mofa_video_b200.keypoint_engine import LdmkAdapterNet
from mofa_video_b200.models._base import EngineModel
from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import \
    FlowControlNet as _TrajFlowControlNet
from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import _identity_key, _same_key


@dataclass
class FlowControlNetOutput:
    down_block_res_samples: Tuple[torch.Tensor] = None
    mid_block_res_sample: torch.Tensor = None
    controlnet_flow: torch.Tensor = None
    occlusion_masks: Optional[List[torch.Tensor]] = None


class FlowControlNet(EngineModel):
    kind = "adapter"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._cond_key = None
        self._masks = None

    def _make_net(self, state_dict, cfg):
        return LdmkAdapterNet(state_dict, cfg, self._ops, self._device)

    @classmethod
    def _net_class(cls):
        return LdmkAdapterNet

    @staticmethod
    def _fresh_state_dict(cfg):
        """from_unet (K/models/controlnet_sdv.py:572-628): landmark embedding / occlusion nets fresh, zero_module convs zero."""
        from mofa_video_b200 import synthetic
        return synthetic.ldmk_adapter_state_dict(cfg, seed=3, zero_std=0.0)

    from_unet = classmethod(_TrajFlowControlNet.from_unet.__func__)
    _UNET_PREFIXES = _TrajFlowControlNet._UNET_PREFIXES

    def prepare_condition(self, controlnet_cond, controlnet_flow, landmarks, force=False):
        """controlnet_cond [B,3,H,W] in [-1,1]; controlnet_flow [B,T-1,2,H,W]; landmarks [B,T,3,H,W] (B = CFG copies).
        Cached on tensor identity only (see svdxt_..._norefine._identity_key); the pipelines pass force=True."""
        key = _identity_key(controlnet_cond, controlnet_flow, landmarks)
        if not force and _same_key(key, self._cond_key):
            return self._masks
        _, _, H, W = controlnet_cond.shape
        if H % 64 or W % 64:
            raise ValueError("height and width must be multiples of 64")
        T = landmarks.shape[1]
        cond = controlnet_cond[:1].to(device=self._device, dtype=torch.float16).contiguous()
        cl = torch.empty(H * W, 3, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(cond, cl, 1, 3, H * W)
        lm = landmarks[0].to(device=self._device, dtype=torch.float16).contiguous()
        lcl = torch.empty(T * H * W, 3, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(lm, lcl, T, 3, H * W)
        flow = controlnet_flow[0].to(device=self._device, dtype=torch.float16).contiguous()
        self._masks = self.net.adapter_cond_branch_ldmk(cl, flow, lcl, H, W)
        self._cond_key = key
        return self._masks

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, controlnet_cond=None,
                controlnet_flow=None, landmarks=None, image_only_indicator=None, return_dict=True, guess_mode=False,
                conditioning_scale=1.0, channels_last_output=False):
        b, t, c, h, w = sample.shape
        if t != self.config.num_frames:
            raise ValueError(f"num_frames mismatch: sample has {t}, model packed for {self.config.num_frames}")
        self._prepare(encoder_hidden_states, added_time_ids)
        masks = self.prepare_condition(controlnet_cond, controlnet_flow, landmarks)
        x = self._to_cl(sample)
        res, mid = self.net.adapter_forward(x, self._t_value(timestep), h, w, conditioning_scale)
        if channels_last_output:
            for r in res + [mid]:
                r._mofa_channels_last = True
            down, midr = res, mid
        else:
            sizes, hh, ww = [(h, w)], h, w
            nlev = len(self.config.block_out_channels)
            lpb = self.config.layers_per_block
            for i in range(nlev):
                sizes += [(hh, ww)] * (lpb if isinstance(lpb, int) else lpb[i])
                if i != nlev - 1:
                    hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
                    sizes.append((hh, ww))
            down = [self._from_cl(r, b * t, *sizes[k]) for k, r in enumerate(res)]
            midr = self._from_cl(mid, b * t, hh, ww)
        # occlusion masks in the reference's shape [B, T-1, 1, hs, ws] per scale (both CFG halves are identical)
        H = controlnet_cond.shape[-2]
        occ = []
        for m, s in zip(masks, (8, 16, 32, 64)):
            hs, ws = H // s, controlnet_cond.shape[-1] // s
            occ.append(m.reshape(1, t - 1, 1, hs, ws).expand(b, -1, -1, -1, -1))
        if not return_dict:
            return (down, midr, controlnet_flow, occ)
        return FlowControlNetOutput(down_block_res_samples=down, mid_block_res_sample=midr,
                                    controlnet_flow=controlnet_flow, occlusion_masks=occ)

    __call__ = forward
