"""ORACLE (test infrastructure): seeded synthetic weights and inputs.

There are no checkpoints and no network in this environment, so every parity test and the benchmark run
on random-init weights in the reference's state-dict layout:
  * torch default nn inits under a fixed seed;
  * every zero_module conv (controlnet_sdv.py:259-300, FCN.py:86-88,145) re-drawn N(0, 0.02^2) so the
    adapter actually contributes (quirk Q19);
  * out-projections (conv2 of every ResNet block, to_out.0, ff.net.2 / ff_in.net.2, proj_out) scaled x0.1
    so 25 fp16 steps stay finite.
"""
import math

import torch

from .models import FlowControlNet, UNetSpatioTemporalConditionControlNetModel

TINY_CONFIG = dict(block_out_channels=(64, 128, 256, 256), num_attention_heads=(1, 2, 4, 4), num_frames=3,
                   cross_attention_dim=64, addition_time_embed_dim=32, projection_class_embeddings_input_dim=96)


def _rescale(model):
    with torch.no_grad():
        for name, p in model.named_parameters():
            if name.endswith(("conv2.weight", "to_out.0.weight", "ff.net.2.weight", "ff_in.net.2.weight",
                              "proj_out.weight")):
                p.mul_(0.1)


def _rerandomise_zero_convs(adapter, gen):
    with torch.no_grad():
        mods = list(adapter.controlnet_down_blocks) + [adapter.controlnet_mid_block,
                                                       adapter.controlnet_cond_embedding.conv_out]
        mods += [z for z in adapter.flow_encoder.zeroconvs]
        for m in mods:
            for p in m.parameters():
                p.copy_(torch.randn(p.shape, generator=gen) * 0.02)


def make_models(config=None, seed=0, adapter_gain=1.0):
    """Returns (unet, adapter) oracle modules in fp32 with fp16-representable weights."""
    cfg = dict(config or {})
    torch.manual_seed(seed)
    unet = UNetSpatioTemporalConditionControlNetModel(**cfg)
    adapter = FlowControlNet(**cfg)
    gen = torch.Generator().manual_seed(seed + 1)
    _rerandomise_zero_convs(adapter, gen)
    _rescale(unet)
    _rescale(adapter)
    if adapter_gain != 1.0:
        with torch.no_grad():
            for m in list(adapter.controlnet_down_blocks) + [adapter.controlnet_mid_block]:
                m.weight.mul_(adapter_gain)
    # round to fp16-representable values so the fp16 engine and the fp32 oracle hold identical weights
    with torch.no_grad():
        for m in (unet, adapter):
            for p in m.parameters():
                p.copy_(p.half().float())
    return unet.eval(), adapter.eval()


def make_flow(T, H, W, seed=1235):
    """[1, T-1, 2, H, W]: smooth rotational field growing with the frame index + unit noise."""
    g = torch.Generator().manual_seed(seed)
    A = 0.06 * min(H, W)
    ys = torch.arange(H, dtype=torch.float32)[:, None] / H
    xs = torch.arange(W, dtype=torch.float32)[None, :] / W
    fx = torch.sin(2 * math.pi * ys) * torch.cos(2 * math.pi * xs)
    fy = torch.cos(2 * math.pi * ys) * torch.sin(2 * math.pi * xs)
    base = torch.stack([fx.expand(H, W), fy.expand(H, W)], 0)
    frames = [(i + 1) / max(T - 1, 1) * A * base + torch.randn(2, H, W, generator=g) for i in range(T - 1)]
    return torch.stack(frames, 0)[None]


def make_image(H, W, seed=1234):
    """[3, H, W] in [0,1]: low-passed noise quantised to uint8 levels."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(1, 3, H + 8, W + 8, generator=g)
    img = torch.nn.functional.avg_pool2d(img, 9, stride=1)
    img = (img - img.min()) / (img.max() - img.min())
    return (img[0] * 255).round() / 255


def make_step_inputs(config, H, W, seed=7, batch=2):
    """Inputs of one adapter+UNet evaluation at latent size H x W (image size 8H x 8W)."""
    g = torch.Generator().manual_seed(seed)
    T = config.get("num_frames", 25)
    ctx_dim = config.get("cross_attention_dim", 1024)
    sample = torch.randn(batch, T, 8, H, W, generator=g).half().float()
    emb = torch.randn(batch, 1, ctx_dim, generator=g).half().float()
    emb[0] = 0  # CFG: the unconditional half sees zero image embeddings (pipeline.py:133-139)
    ids = torch.tensor([[6.0, 128.0, 0.02]] * batch)
    cond = (make_image(8 * H, 8 * W) * 2 - 1).half().float()[None].repeat(batch, 1, 1, 1)
    flow = make_flow(T, 8 * H, 8 * W).half().float().repeat(batch, 1, 1, 1, 1)
    return dict(sample=sample, encoder_hidden_states=emb, added_time_ids=ids, controlnet_cond=cond,
                controlnet_flow=flow)


def make_ldmk_adapter(config, seed=3, gain=20.0):
    """Keypoint (landmark) adapter oracle with every zero-initialised conv re-drawn so the occlusion branch and the
    landmark embedding contribute; fp16-representable weights."""
    from . import keypoint as kp
    torch.manual_seed(seed)
    ad = kp.FlowControlNetLdmk(**config)
    g = torch.Generator().manual_seed(seed + 1)
    with torch.no_grad():
        zero = list(ad.controlnet_down_blocks) + [ad.controlnet_mid_block, ad.controlnet_cond_embedding.conv_out,
                                                  ad.controlnet_ldmk_embedding.conv_out] + list(ad.zero_outs.values())
        for m in zero:
            for p in m.parameters():
                p.copy_(torch.randn(p.shape, generator=g) * 0.02)
        for m in list(ad.controlnet_down_blocks) + [ad.controlnet_mid_block]:
            m.weight.mul_(gain)
        for m in list(ad.zero_outs.values()) + [ad.controlnet_ldmk_embedding.conv_out]:
            m.weight.mul_(30.0)          # make the occlusion branch and the landmark embedding matter
    _rescale(ad)
    with torch.no_grad():
        for p in ad.parameters():
            p.copy_(p.half().float())
    return ad.eval()


def make_vae_and_clip(dim, seed=5):
    """The two third-party constructor arguments of the pipeline, as small seeded PyTorch modules (the same objects
    are handed to the reference pipeline here and to the oracle / engine in the tests)."""
    from types import SimpleNamespace

    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder

    class TinyClip(torch.nn.Module):
        def __init__(self, d):
            super().__init__()
            self.proj = torch.nn.Linear(3 * 8 * 8, d)

        def forward(self, x):
            return SimpleNamespace(image_embeds=self.proj(torch.nn.functional.adaptive_avg_pool2d(x, 8).flatten(1)))

    torch.manual_seed(seed)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=(32, 32, 64, 64)).eval()
    clip = TinyClip(dim).eval()
    return vae, clip
