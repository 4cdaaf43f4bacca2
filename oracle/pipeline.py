"""ORACLE (test infrastructure): FlowControlNetPipeline.__call__ restated step by step from
/root/reference/MOFA-Video-Traj/pipeline/pipeline.py:282-527, runnable on CPU in fp32 with the oracle
networks (oracle/models.py) and scheduler (oracle/scheduler.py).  `vae` and `image_encoder` are the
third-party objects the reference receives in its constructor (:90-108); any module with the same
interface works.  PINNED: tests/golden/pipeline_tiny.pt is produced by executing the reference's own __call__ (oracle/make_goldens.py:
make_pipeline; DiffusionPipeline / VaeImageProcessor base classes stubbed, diffusers blocks from oracle/d24_blocks.py);
run_pipeline reproduces its 2-step latents bit-exactly (tests/test_oracle.py).
"""
import numpy as np
import PIL.Image
import torch
import torch.nn.functional as F


def pil_to_pt(image):
    """VaeImageProcessor.pil_to_numpy + numpy_to_pt: [N, 3, H, W] float32 in [0, 1]."""
    if isinstance(image, PIL.Image.Image):
        image = [image]
    arr = np.stack([np.array(im).astype(np.float32) / 255.0 for im in image], axis=0)
    return torch.from_numpy(arr.transpose(0, 3, 1, 2))


def _gaussian(window_size, sigma):  # pipeline.py:603-617
    x = (torch.arange(window_size, dtype=sigma.dtype) - window_size // 2).expand(sigma.shape[0], -1)
    if window_size % 2 == 0:
        x = x + 0.5
    gauss = torch.exp(-x.pow(2.0) / (2 * sigma.pow(2.0)))
    return gauss / gauss.sum(-1, keepdim=True)


def _filter2d(inp, kernel):  # pipeline.py:580-600
    b, c, h, w = inp.shape
    k = kernel[:, None, ...].to(inp.dtype).expand(-1, c, -1, -1)
    kh, kw = k.shape[-2:]
    pad = [(kw - 1) // 2, (kw - 1) - (kw - 1) // 2, (kh - 1) // 2, (kh - 1) - (kh - 1) // 2]  # :563-577
    inp = F.pad(inp, pad, mode="reflect")
    k = k.reshape(-1, 1, kh, kw)
    inp = inp.view(-1, k.size(0), inp.size(-2), inp.size(-1))
    return F.conv2d(inp, k, groups=k.size(0), padding=0, stride=1).view(b, c, h, w)


def resize_with_antialiasing(inp, size):  # pipeline.py:532-562
    if inp.ndim == 3:
        inp = inp.unsqueeze(0)
    h, w = inp.shape[-2:]
    factors = (h / size[0], w / size[1])
    sigmas = (max((factors[0] - 1.0) / 2.0, 0.001), max((factors[1] - 1.0) / 2.0, 0.001))
    ks = int(max(2.0 * 2 * sigmas[0], 3)), int(max(2.0 * 2 * sigmas[1], 3))
    if (ks[0] % 2) == 0:
        ks = ks[0] + 1, ks[1]
    if (ks[1] % 2) == 0:
        ks = ks[0], ks[1] + 1
    sigma = torch.tensor([sigmas], dtype=inp.dtype)
    kx = _gaussian(ks[1], sigma[:, 1].view(1, 1))
    ky = _gaussian(ks[0], sigma[:, 0].view(1, 1))
    out = _filter2d(_filter2d(inp, kx[..., None, :]), ky[..., None])
    return F.interpolate(out, size=size, mode="bicubic", align_corners=True)


@torch.no_grad()
def prepare_inputs(vae, image_encoder, image, num_frames, generator=None, noise_aug_strength=0.02, dtype=torch.float32):
    """CLIP embedding (Q3, pipeline.py:114-141) and VAE image latents (:339-356), both with the zero CFG half first."""
    if isinstance(image, torch.Tensor) and image.ndim == 3:
        image = image[None]
    img01 = pil_to_pt(image) if not isinstance(image, torch.Tensor) else image
    clip_in = resize_with_antialiasing(img01, (224, 224)).to(dtype)
    emb = image_encoder(clip_in).image_embeds.unsqueeze(1)
    emb = torch.cat([torch.zeros_like(emb), emb])
    img = 2.0 * pil_to_pt(image) - 1.0 if not isinstance(image, torch.Tensor) else 2.0 * image - 1.0
    noise = torch.randn(img.shape, generator=generator, dtype=img.dtype)
    img = img + noise_aug_strength * noise
    image_latents = vae.encode(img.to(dtype)).latent_dist.mode()
    image_latents = torch.cat([torch.zeros_like(image_latents), image_latents]).to(emb.dtype)
    return emb, image_latents.unsqueeze(1).repeat(1, num_frames, 1, 1, 1)


@torch.no_grad()
def run_pipeline(vae, image_encoder, unet, controlnet, scheduler, image, controlnet_condition, controlnet_flow,
                 height=576, width=1024, num_frames=None, num_inference_steps=25, min_guidance_scale=1.0,
                 max_guidance_scale=3.0, noise_aug_strength=0.02, decode_chunk_size=None, generator=None,
                 latents=None, output_type="latent", controlnet_cond_scale=1.0, dtype=torch.float32):
    num_frames = num_frames if num_frames is not None else unet.config.num_frames          # :317
    decode_chunk_size = decode_chunk_size if decode_chunk_size is not None else num_frames  # :318
    do_cfg = max_guidance_scale > 1.0
    assert do_cfg, "Q5: without CFG the reference feeds the latents as condition (pipeline.py:393,396)"
    if isinstance(controlnet_condition, torch.Tensor) and controlnet_condition.ndim == 3:
        controlnet_condition = controlnet_condition[None]
    emb, image_latents = prepare_inputs(vae, image_encoder, image, num_frames, generator, noise_aug_strength, dtype)
    # 5. added time ids, overwritten by constants (Q4)                                       :430-440
    added_time_ids = torch.tensor([[6, 128, 0.02]], dtype=emb.dtype).repeat(2, 1)
    scheduler.set_timesteps(num_inference_steps)
    timesteps = scheduler.timesteps
    shape = (1, num_frames, unet.config.in_channels // 2, height // 8, width // 8)
    if latents is None:
        latents = torch.randn(shape, generator=generator, dtype=emb.dtype)
    latents = latents.to(emb.dtype) * scheduler.init_noise_sigma                           # :267
    cond = 2.0 * pil_to_pt(controlnet_condition) - 1.0 if not isinstance(controlnet_condition, torch.Tensor) \
        else 2.0 * controlnet_condition - 1.0
    cond = torch.cat([cond] * 2).to(latents.dtype)
    flow = torch.cat([controlnet_flow] * 2).to(latents.dtype)
    g = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames).unsqueeze(0).to(latents.dtype)
    g = g[(...,) + (None,) * 3]
    for t in timesteps:                                                                     # :447-511
        lmi = torch.cat([latents] * 2)
        lmi = scheduler.scale_model_input(lmi, t)
        lmi = torch.cat([lmi, image_latents], dim=2)
        dres, mid, flow, _ = controlnet(lmi, t, encoder_hidden_states=emb, controlnet_cond=cond,
                                        controlnet_flow=flow, added_time_ids=added_time_ids,
                                        conditioning_scale=controlnet_cond_scale, return_dict=False)
        noise_pred = unet(lmi, t, encoder_hidden_states=emb, down_block_additional_residuals=dres,
                          mid_block_additional_residual=mid, added_time_ids=added_time_ids, return_dict=False)[0]
        u, c = noise_pred.chunk(2)
        noise_pred = u + g * (c - u)
        latents = scheduler.step(noise_pred, t, latents)
    if output_type == "latent":
        return latents
    lat = latents.flatten(0, 1) / vae.config.scaling_factor                                 # :194-220
    frames = [vae.decode(lat[i:i + decode_chunk_size], num_frames=lat[i:i + decode_chunk_size].shape[0]).sample
              for i in range(0, lat.shape[0], decode_chunk_size)]
    frames = torch.cat(frames, dim=0)
    frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4).float()
    return frames


def drag_flow_post(flow_inmask, height, width, brush_mask=None, flow_outmask=None):
    """T/run_gradio.py:251-255 (brush), :268-275 (nearest resize + rescale), :330-333 (merge), restated op for op on
    [B, T-1, 2, hs, ws] tensors in their own dtype."""
    def resize(fl):
        fb, fl_, _, hs, ws = fl.shape
        if (height, width) == (hs, ws):
            return fl
        r = F.interpolate(fl.flatten(0, 1), (height, width), mode="nearest").reshape(fb, fl_, 2, height, width).clone()
        r[:, :, 0] *= width / ws
        r[:, :, 1] *= height / hs
        return r
    a = flow_inmask
    if brush_mask is not None:
        a = a * brush_mask.to(a.dtype)[None, None, None]
    a = resize(a)
    if flow_outmask is None:
        return a
    b = resize(flow_outmask)
    keep = (a != 0).all(dim=2).unsqueeze(2).expand_as(a)
    return torch.where(keep, a, b)
