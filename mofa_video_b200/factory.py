"""Assemble a FlowControlNetPipeline from random-init weights (no checkpoints / network in this environment):
the same objects T/run_gradio.py:90-159 builds from disk, with synthetic state dicts in the reference layout."""
import torch

from mofa_video_b200 import synthetic
from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import FlowControlNet
from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
    UNetSpatioTemporalConditionControlNetModel
from mofa_video_b200.pipeline.pipeline import FlowControlNetPipeline
from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler


class _TinyImageEncoder(torch.nn.Module):
    """Smoke-test stand-in for CLIPVisionModelWithProjection (same call contract: .image_embeds)."""

    def __init__(self, dim):
        super().__init__()
        self.proj = torch.nn.Linear(3 * 8 * 8, dim)

    def forward(self, x):
        from types import SimpleNamespace
        return SimpleNamespace(image_embeds=self.proj(torch.nn.functional.adaptive_avg_pool2d(x, 8).flatten(1)))


def make_clip_vit_h(projection_dim=1024):
    """Random-init CLIP ViT-H/14 vision tower with projection (transformers)."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = CLIPVisionConfig(hidden_size=1280, intermediate_size=5120, num_hidden_layers=32, num_attention_heads=16,
                           image_size=224, patch_size=14, projection_dim=projection_dim, hidden_act="gelu")
    return CLIPVisionModelWithProjection(cfg)


def build_synthetic_pipeline(config=None, device="cuda", seed=0, tiny_encoders=False, vae_channels=None,
                             native_vae=True, variant="traj"):
    """variant: 'traj' (T/pipeline/pipeline.py), 'keypoint' (K/pipeline/svdxt_pipeline_ctrlnet_loop.py, landmark
    adapter) or 'hybrid' (H/pipeline/pipeline.py, landmark adapter inside the mask + trajectory adapter outside)."""
    cfg_u, sd_u = synthetic.unet_state_dict(config, seed=seed)
    cfg_a, sd_a = synthetic.adapter_state_dict(config, seed=seed + 1)
    unet = UNetSpatioTemporalConditionControlNetModel.from_state_dict(sd_u, cfg_u, device=device)
    del sd_u
    controlnet = FlowControlNet.from_state_dict(sd_a, cfg_a, device=device) if variant != "keypoint" else None
    del sd_a
    face = None
    if variant in ("keypoint", "hybrid"):
        from mofa_video_b200.models.ldmk_ctrlnet import FlowControlNet as LdmkFlowControlNet
        cfg_l, sd_l = synthetic.ldmk_adapter_state_dict(config, seed=seed + 3)
        face = LdmkFlowControlNet.from_state_dict(sd_l, cfg_l, device=device)
        del sd_l
    torch.manual_seed(seed + 2)
    if tiny_encoders:
        vae = AutoencoderKLTemporalDecoder(block_out_channels=vae_channels or (64, 64, 128, 128))
        clip = _TinyImageEncoder(cfg_u["cross_attention_dim"])
    else:
        vae = AutoencoderKLTemporalDecoder()
        clip = make_clip_vit_h(cfg_u["cross_attention_dim"])
    vae = vae.to(device=device, dtype=torch.float16).eval()
    clip = clip.to(device=device, dtype=torch.float16).eval()
    # the pipelines re-host an AutoencoderKLTemporalDecoder-layout VAE on the sm_100a kernels themselves
    # (FlowControlNetPipeline._adopt_vae) -- the same path T/run_gradio.py:init_models takes
    nv = None if native_vae else False
    if variant == "keypoint":
        from mofa_video_b200.pipeline.svdxt_pipeline_ctrlnet_loop import FlowControlNetPipeline as KeypointPipeline
        pipe = KeypointPipeline(vae=vae, image_encoder=clip, unet=unet, controlnet=face,
                                scheduler=EulerDiscreteScheduler(), native_vae=nv)
    elif variant == "hybrid":
        from mofa_video_b200.pipeline.pipeline_hybrid import FlowControlNetPipeline as HybridPipeline
        pipe = HybridPipeline(vae=vae, image_encoder=clip, unet=unet, drag_controlnet=controlnet,
                              face_controlnet=face, scheduler=EulerDiscreteScheduler(), native_vae=nv)
    else:
        pipe = FlowControlNetPipeline(vae=vae, image_encoder=clip, unet=unet, controlnet=controlnet,
                                      scheduler=EulerDiscreteScheduler(), native_vae=nv)
    return pipe.to(device)
