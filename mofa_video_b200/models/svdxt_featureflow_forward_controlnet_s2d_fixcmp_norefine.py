"""FlowControlNet (MOFA-Adapter, trajectory variant) with the reference's entry point
(/root/reference/MOFA-Video-Traj/models/svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py:180-383),
executed by the sm_100a engine.  forward keeps the reference signature (:236-248) and return convention
(:378-383); the loop-invariant conditioning branch (:297-319) is cached per (cond image, flow) pair and
evaluated for one CFG half (both halves are identical, pipeline.py:393-397)."""
from dataclasses import dataclass
from typing import Optional, Tuple

import torch

from mofa_video_b200.models._base import EngineModel


@dataclass
class FlowControlNetOutput:
    down_block_res_samples: Tuple[torch.Tensor] = None
    mid_block_res_sample: torch.Tensor = None
    controlnet_flow: torch.Tensor = None
    cmp_output: Optional[torch.Tensor] = None


def _identity_key(*tensors):
    """(tensor objects, their versions): compared with `is`, so the cache holds the tensors alive and a new tensor
    that happens to land on a recycled address never matches."""
    return tuple(tensors), tuple(t._version for t in tensors)


def _same_key(a, b):
    return (a is not None and b is not None and len(a[0]) == len(b[0])
            and all(x is y for x, y in zip(a[0], b[0])) and a[1] == b[1])


class FlowControlNet(EngineModel):
    kind = "adapter"

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self._cond_key = None

    # -- ControlNetSDVModel.from_unet (models/controlnet_sdv.py:572-628; Training/train_stage2.py:823) ------------
    _UNET_PREFIXES = ("conv_in.", "time_embedding.", "down_blocks.", "mid_block.")   # :617-626 (time_proj has no weights)

    @staticmethod
    def _fresh_state_dict(cfg):
        """Everything from_unet does NOT copy: add_embedding, cond embedding, first-frame encoder (fresh init) and the
        zero_module convs -- exactly zero, weights and biases (controlnet_sdv.py:259-300, FCN.py:86-88,145)."""
        from mofa_video_b200 import synthetic
        return synthetic.adapter_state_dict(cfg, seed=1, zero_std=0.0)

    @classmethod
    def from_unet(cls, unet, controlnet_conditioning_channel_order="rgb",
                  conditioning_embedding_out_channels=(16, 32, 96, 256), load_weights_from_unet=True,
                  conditioning_channels=3, device=None, ops=None):
        """Adapter with the UNet's configuration; conv_in / time_embedding / down_blocks / mid_block take the UNet's
        weights (by reference, no copy) when load_weights_from_unet.  `unet` is any object with .config and
        .state_dict() in the reference layout (this package's UNet, the oracle's, diffusers')."""
        uc = unet.config
        get = (lambda k: uc.get(k)) if isinstance(uc, dict) else (lambda k: getattr(uc, k, None))
        cfg = {k: get(k) for k in ("in_channels", "down_block_types", "block_out_channels", "addition_time_embed_dim",
                                   "transformer_layers_per_block", "cross_attention_dim", "num_attention_heads",
                                   "num_frames", "sample_size", "layers_per_block",
                                   "projection_class_embeddings_input_dim")}
        cfg = {k: v for k, v in cfg.items() if v is not None}
        cfg["conditioning_channels"] = conditioning_channels
        cfg["conditioning_embedding_out_channels"] = tuple(conditioning_embedding_out_channels)
        cfg_full, sd = cls._fresh_state_dict(cfg)
        if load_weights_from_unet:
            usd = unet.state_dict()
            for k in list(sd):
                if k.startswith(cls._UNET_PREFIXES):
                    if tuple(usd[k].shape) != tuple(sd[k].shape):
                        raise ValueError(f"from_unet: {k} has shape {tuple(usd[k].shape)} in the UNet, "
                                         f"{tuple(sd[k].shape)} expected")
                    sd[k] = usd[k]
        if ops is None and device is None and isinstance(unet, EngineModel):
            ops, device = unet._ops, unet._device
        return cls(sd, {k: v for k, v in cfg_full.items() if k not in ("out_channels", "up_block_types")},
                   device=device, ops=ops)

    def prepare_condition(self, controlnet_cond, controlnet_flow, force=False):
        """controlnet_cond [B, 3, H, W] in [-1, 1], controlnet_flow [B, T-1, 2, H, W] (B = CFG copies).
        The branch is skipped only when the SAME tensor objects (held by reference, unchanged `_version`) come back,
        as they do on the per-step forward() path; a recycled allocator address can therefore never alias a previous
        request (`force=True`: the pipelines recompute it once per clip unconditionally)."""
        key = _identity_key(controlnet_cond, controlnet_flow)
        if not force and _same_key(key, self._cond_key):
            return
        if controlnet_cond.shape[0] > 1 and not torch.equal(controlnet_cond[0], controlnet_cond[-1]):
            raise ValueError("the engine hoists the conditioning branch assuming identical CFG halves "
                             "(pipeline.py:393-397); got different condition images per batch item")
        _, _, H, W = controlnet_cond.shape
        if H % 64 or W % 64:
            raise ValueError("height and width must be multiples of 64 (flow pyramid down to 1/64, FCN.py:302-315)")
        cond = controlnet_cond[:1].to(device=self._device, dtype=torch.float16).contiguous()
        cl = torch.empty(H * W, 3, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(cond, cl, 1, 3, H * W)
        flow = controlnet_flow[0].to(device=self._device, dtype=torch.float16).contiguous()
        self.net.adapter_cond_branch(cl, flow, H, W)
        self._cond_key = key

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, controlnet_cond=None,
                controlnet_flow=None, image_only_indicator=None, return_dict=True, guess_mode=False,
                conditioning_scale=1.0, channels_last_output=False):
        b, t, c, h, w = sample.shape
        if t != self.config.num_frames:
            raise ValueError(f"num_frames mismatch: sample has {t}, model packed for {self.config.num_frames}")
        self._prepare(encoder_hidden_states, added_time_ids)
        self.prepare_condition(controlnet_cond, controlnet_flow)
        x = self._to_cl(sample)
        res, mid = self.net.adapter_forward(x, self._t_value(timestep), h, w, conditioning_scale)
        if channels_last_output:
            for r in res + [mid]:
                r._mofa_channels_last = True
            down, midr = res, mid
        else:
            sizes = []
            hh, ww = h, w
            nlev = len(self.config.block_out_channels)
            lpb = self.config.layers_per_block
            sizes.append((hh, ww))
            for i in range(nlev):
                sizes += [(hh, ww)] * (lpb if isinstance(lpb, int) else lpb[i])
                if i != nlev - 1:
                    hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
                    sizes.append((hh, ww))
            down = [self._from_cl(r, b * t, *sizes[k]) for k, r in enumerate(res)]
            midr = self._from_cl(mid, b * t, hh, ww)
        if not return_dict:
            return (down, midr, controlnet_flow, None)
        return FlowControlNetOutput(down_block_res_samples=down, mid_block_res_sample=midr,
                                    controlnet_flow=controlnet_flow, cmp_output=None)

    __call__ = forward


class CMP_demo:
    """Sparse-to-dense flow (CMP) with the reference's entry point (same file, :26-62): CMP_demo(configfn, load_iter),
    .to(device), .requires_grad_(False), .run(image [B,3,H,W] in [0,1], sparse [B,2,H,W], mask [B,2,H,W]) -> flow
    [B,2,H,W] in the caller's dtype.  Reads `<dir(configfn)>/checkpoints/ckpt_iter_<load_iter>.pth.tar` (key
    'state_dict', names prefixed 'module.'); like the reference (models/cmp/utils/common_utils.py:115-116) a missing
    checkpoint is a warning and the network keeps its random initialisation."""

    def __init__(self, configfn=None, load_iter=None, state_dict=None, device=None, ops=None):
        import os
        import warnings

        from mofa_video_b200.cmp_engine import CmpNet
        from mofa_video_b200.models._base import resolve_backend
        nbins, fmax = 99, 50.0
        if configfn is not None:
            import yaml
            with open(configfn) as f:
                cfg = yaml.full_load(f)
            mod = cfg["model"]["module"]
            nbins, fmax = mod.get("nbins", 99), float(mod.get("fmax", 50))
            if mod.get("image_encoder") != "resnet50" or mod.get("flow_decoder") != "MotionDecoderSkipLayer":
                raise NotImplementedError("only the resnet50 + MotionDecoderSkipLayer CMP of MOFA-Video is implemented")
            if state_dict is None:
                fn = os.path.join(os.path.dirname(configfn), "checkpoints", f"ckpt_iter_{load_iter}.pth.tar")
                if os.path.isfile(fn):
                    state_dict = torch.load(fn, map_location="cpu")["state_dict"]
                else:
                    warnings.warn(f"=> no checkpoint found at '{fn}': CMP keeps a random initialisation")
        if state_dict is None:
            from mofa_video_b200 import synthetic
            state_dict = synthetic.cmp_state_dict()
        self._ops, device, _ = resolve_backend(ops, device)
        self.net = CmpNet(state_dict, self._ops, device, nbins=nbins, fmax=fmax)

    def to(self, *a, **k):
        return self

    def requires_grad_(self, flag=False):
        return self

    def eval(self):
        return self

    def run(self, image, sparse, mask):
        dtype = image.dtype
        return self.net.forward(image, sparse, mask).to(dtype)
