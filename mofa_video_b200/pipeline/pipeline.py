"""FlowControlNetPipeline with the reference's entry point
(/root/reference/MOFA-Video-Traj/pipeline/pipeline.py:87-527), driving the sm_100a engine.

Same constructor kwargs (:90-108), same __call__ signature and defaults (:283-311), same output object
(:72-84), same errors (check_inputs :222-234, add-time-id check :34-40).  What changes is the execution:
the 25-step loop (:447-511) keeps the latents, the model input and every activation on the device in
channels-last fp16, runs adapter + UNet through the C-ABI kernels, and replaces the Python-side
torch.cat / scale_model_input / CFG / scheduler.step (:449-454, 495-500) by one fused kernel per step.
Quirks kept on purpose: Q3 un-normalised CLIP input, Q4 hard-coded added_time_ids,
Q5 CFG is mandatory, Q6 per-frame guidance, Q7 RNG placement, Q8 VAE chunking, Q16, Q21, Q22.
"""
import json
import os
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional, Union

import numpy as np
import PIL.Image
import torch
import torch.nn.functional as F

from mofa_video_b200 import lib as _lib


class _NoEvent:
    def record(self):
        pass


@dataclass
class FlowControlNetPipelineOutput:
    frames: Union[List[PIL.Image.Image], np.ndarray, torch.Tensor] = None
    controlnet_flow: Union[List[PIL.Image.Image], np.ndarray, torch.Tensor] = None


def _get_add_time_ids(noise_aug_strength, dtype, batch_size, fps=4, motion_bucket_id=128, unet=None):
    """pipeline.py:24-46."""
    add_time_ids = [fps, motion_bucket_id, noise_aug_strength]
    passed = unet.config.addition_time_embed_dim * len(add_time_ids)
    expected = unet.add_embedding.linear_1.in_features
    if expected != passed:
        raise ValueError(
            f"Model expects an added time embedding vector of length {expected}, but a vector of {passed} was "
            "created. The model has an incorrect config. Please check `unet.config.time_embedding_type` and "
            "`text_encoder_2.config.projection_dim`.")
    return torch.tensor([add_time_ids], dtype=dtype)


def _randn_tensor(shape, generator=None, device=None, dtype=None):
    """diffusers.utils.torch_utils.randn_tensor as the reference calls it (pipeline.py:262, :340): a CPU generator
    draws on the CPU and the result is moved; a CUDA generator cannot fill a CPU tensor (ValueError); a list of
    generators draws one batch item each."""
    device = torch.device(device) if device is not None else torch.device("cpu")
    rand_device = device
    if generator is not None:
        gtype = (generator[0] if isinstance(generator, list) else generator).device.type
        if gtype != device.type and gtype == "cpu":
            rand_device = torch.device("cpu")
        elif gtype != device.type and gtype == "cuda":
            raise ValueError(f"Cannot generate a {device} tensor from a generator of type {gtype}.")
    if isinstance(generator, list) and len(generator) == 1:
        generator = generator[0]
    if isinstance(generator, list):
        one = (1,) + tuple(shape[1:])
        out = torch.cat([torch.randn(one, generator=generator[i], device=rand_device, dtype=dtype)
                         for i in range(shape[0])], dim=0)
    else:
        out = torch.randn(tuple(shape), generator=generator, device=rand_device, dtype=dtype)
    return out.to(device)


# ------------------------------------------------------------------------------------------------
# image pre/post-processing (diffusers VaeImageProcessor behaviour used by the reference, Q21)
# ------------------------------------------------------------------------------------------------
def _to_unit_tensor(image, height=None, width=None):
    """PIL / ndarray / tensor -> float32 [N, 3, H, W] in [0, 1] (PIL is resized like VaeImageProcessor)."""
    if isinstance(image, PIL.Image.Image):
        image = [image]
    if isinstance(image, (list, tuple)) and isinstance(image[0], PIL.Image.Image):
        arrs = []
        for im in image:
            if height is not None and im.size != (width, height):
                im = im.resize((width, height), resample=PIL.Image.LANCZOS)
            arrs.append(np.asarray(im.convert("RGB"), dtype=np.float32) / 255.0)
        return torch.from_numpy(np.stack(arrs, 0)).permute(0, 3, 1, 2).contiguous()
    if isinstance(image, np.ndarray):
        a = image.astype(np.float32) / (255.0 if image.dtype == np.uint8 else 1.0)
        if a.ndim == 3:
            a = a[None]
        return torch.from_numpy(a).permute(0, 3, 1, 2).contiguous()
    if isinstance(image, torch.Tensor):
        t = image.float()
        return t[None] if t.ndim == 3 else t
    raise ValueError(f"unsupported image type {type(image)}")


class FlowControlNetPipeline:
    model_cpu_offload_seq = "image_encoder->unet->vae"
    _callback_tensor_inputs = ["latents"]

    def __init__(self, vae, image_encoder, unet, controlnet, scheduler, feature_extractor=None, ops=None,
                 device=None, native_vae=None, native_clip=None):
        """`ops` / `device` / `native_vae` exist for the CPU host-logic tests only (tests/ref_ops.py states every C-ABI
        op in PyTorch); a product user never passes them: the default binds the CUDA library and fails if it is
        missing, and re-hosts the VAE on the kernels."""
        from mofa_video_b200.models._base import resolve_backend
        self._ops, self._device, _ = resolve_backend(ops, device)  # no fallback: raises if the library is missing
        if native_vae is None:
            native_vae = True
        if native_clip is None:
            native_clip = True
        self.unet, self.controlnet = unet, controlnet
        self.image_encoder = self._adopt_image_encoder(image_encoder) if native_clip else image_encoder
        self.vae = self._adopt_vae(vae) if native_vae else vae
        self.scheduler, self.feature_extractor = scheduler, feature_extractor
        self.vae_scale_factor = 2 ** (len(self.vae.config.block_out_channels) - 1)
        self.timings = {}

    _VAE_KEYS = ("encoder.conv_in.weight", "quant_conv.weight", "decoder.conv_in.weight",
                 "decoder.mid_block.attentions.0.to_q.weight", "decoder.time_conv_out.weight")

    def _adopt_image_encoder(self, enc):
        """transformers' CLIPVisionModelWithProjection (T/run_gradio.py:98) is re-hosted on the kernels
        (clip_engine.NativeClipVision); any other image encoder runs as the module the caller built."""
        from mofa_video_b200.clip_engine import NativeClipVision, is_clip_vision_with_projection
        if isinstance(enc, NativeClipVision) or not is_clip_vision_with_projection(enc):
            return enc
        try:
            return NativeClipVision(enc, ops=self._ops, device=self._device)
        except NotImplementedError as exc:
            import warnings
            warnings.warn(f"image encoder kept as the caller's module: {exc}")
            return enc

    def _adopt_vae(self, vae):
        """What the reference's scripts hand over is diffusers' AutoencoderKLTemporalDecoder (T/run_gradio.py:101-102).
        Any module with that parameter layout is re-hosted on the sm_100a kernels (vae_engine.NativeTemporalDecoderVAE:
        encode, decode and the fused uint8 tail); only a VAE of some other architecture is left to run as the caller
        built it."""
        if hasattr(vae, "decode_uint8") or not hasattr(vae, "state_dict"):
            return vae
        sd = vae.state_dict()
        if not all(k in sd for k in self._VAE_KEYS):
            return vae
        cfg = vae.config
        boc = cfg["block_out_channels"] if isinstance(cfg, dict) else cfg.block_out_channels
        if any(c % 64 for c in boc):
            import warnings
            warnings.warn(f"VAE widths {tuple(boc)} are not multiples of 64 (the SVD VAE is 128/256/512/512): the implicit-"
                          "GEMM kernels cannot host it, it runs as the module the caller built")
            return vae
        from mofa_video_b200.vae_engine import NativeTemporalDecoderVAE
        return NativeTemporalDecoderVAE(vae, ops=self._ops, device=self._device)

    @classmethod
    def from_pretrained(cls, path, unet=None, controlnet=None, image_encoder=None, vae=None, scheduler=None,
                        feature_extractor=None, torch_dtype=None, **_ignored):
        """T/run_gradio.py:147-154: the four models are passed in; the scheduler comes from the SVD folder."""
        if scheduler is None:
            from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler
            fn = os.path.join(path, "scheduler", "scheduler_config.json")
            if os.path.exists(fn):
                with open(fn) as f:
                    scheduler = EulerDiscreteScheduler.from_config(
                        {k: v for k, v in json.load(f).items() if not k.startswith("_")})
            else:
                scheduler = EulerDiscreteScheduler()
        if None in (unet, controlnet, image_encoder, vae):
            raise ValueError("pass unet=, controlnet=, image_encoder= and vae= (as T/run_gradio.py:147-154 does)")
        return cls(vae=vae, image_encoder=image_encoder, unet=unet, controlnet=controlnet, scheduler=scheduler,
                   feature_extractor=feature_extractor)

    def _step_runner(self, unet_net, ad_net, T, h, w, g_min, g_max, cond_scale):
        """graph_step.StepRunner for this (networks, shape, guidance, scale); at most two are kept (each holds the
        activation pool of one captured step)."""
        from mofa_video_b200.graph_step import StepRunner
        cache = self.__dict__.setdefault("_runners", {})
        key = (id(unet_net), id(ad_net), T, h, w, float(g_min), float(g_max), float(cond_scale))
        if key not in cache:
            while len(cache) >= 2:
                cache.pop(next(iter(cache)))
            cache[key] = StepRunner(self._ops, unet_net, ad_net, T, h, w, g_min, g_max, cond_scale, self._device)
            cache[key].use_graph = cache[key].use_graph and getattr(self, "use_cuda_graph", True)
        return cache[key]

    def to(self, device=None, *a, **k):
        if device is not None:
            self._device = torch.device(device)
        return self

    @property
    def guidance_scale(self):
        return self._guidance_scale

    @property
    def num_timesteps(self):
        return self._num_timesteps

    def maybe_free_model_hooks(self):
        pass

    # ------------------------------------------------------------------ pieces of __call__
    def check_inputs(self, image, height, width):
        if not isinstance(image, (torch.Tensor, PIL.Image.Image, list)):
            raise ValueError("`image` has to be of type `torch.FloatTensor` or `PIL.Image.Image` or "
                             f"`List[PIL.Image.Image]` but is {type(image)}")
        if height % 8 != 0 or width % 8 != 0:
            raise ValueError(f"`height` and `width` have to be divisible by 8 but are {height} and {width}.")

    def _encode_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        """pipeline.py:114-141 (Q3: [0,1] image, antialiased bicubic to 224, no CLIP mean/std)."""
        dtype = next(self.image_encoder.parameters()).dtype
        img = _to_unit_tensor(image).to(device=device, dtype=torch.float32).contiguous()
        small = torch.empty(img.shape[0], img.shape[1], 224, 224, dtype=torch.float32, device=device)
        self._ops.resize_antialias(img, small)  # Gaussian pre-blur + bicubic(align_corners) fused (pipeline.py:532-640)
        img = small.to(dtype)
        emb = self.image_encoder(img).image_embeds.unsqueeze(1)
        emb = emb.repeat(1, num_videos_per_prompt, 1)
        if do_classifier_free_guidance:
            emb = torch.cat([torch.zeros_like(emb), emb])
        return emb

    def _encode_vae_image(self, image, device, num_videos_per_prompt, do_classifier_free_guidance):
        """pipeline.py:143-164: .mode() of the posterior, NOT multiplied by scaling_factor (Q8)."""
        lat = self.vae.encode(image.to(device=device)).latent_dist.mode()
        if do_classifier_free_guidance:
            lat = torch.cat([torch.zeros_like(lat), lat])
        return lat.repeat(num_videos_per_prompt, 1, 1, 1)

    def prepare_latents(self, batch_size, num_frames, num_channels_latents, height, width, dtype, device, generator,
                        latents=None):
        shape = (batch_size, num_frames, num_channels_latents // 2, height // self.vae_scale_factor,
                 width // self.vae_scale_factor)
        if isinstance(generator, list) and len(generator) != batch_size:
            raise ValueError(f"You have passed a list of generators of length {len(generator)}, but requested an "
                             f"effective batch size of {batch_size}.")
        if latents is None:
            latents = _randn_tensor(shape, generator=generator, device=device, dtype=dtype)
        else:
            latents = latents.to(device)
        return latents * self.scheduler.init_noise_sigma

    def decode_latents(self, latents, num_frames, decode_chunk_size=14):
        """pipeline.py:194-220 (Q8: chunks see zero temporal padding at their borders)."""
        latents = latents.flatten(0, 1)
        latents = 1 / self.vae.config.scaling_factor * latents
        frames = []
        for i in range(0, latents.shape[0], decode_chunk_size):
            chunk = latents[i:i + decode_chunk_size]
            frames.append(self.vae.decode(chunk, num_frames=chunk.shape[0]).sample)
        frames = torch.cat(frames, dim=0)
        frames = frames.reshape(-1, num_frames, *frames.shape[1:]).permute(0, 2, 1, 3, 4)
        return frames.float()

    def _decode_output(self, latents, num_frames, decode_chunk_size, output_type):
        """Step 9 of the reference __call__ (pipeline.py:513-522): latents [1, T, 4, h, w] -> `frames` of the output."""
        if output_type == "latent":
            return latents
        if output_type in ("pil", "uint8", "uint8_pt") and hasattr(self.vae, "decode_uint8"):
            # native decoder: time_conv_out + (x/2+0.5).clamp*255 -> uint8 fused in the decoder's tail kernel, every chunk
            # written in place into one clip buffer [T, H, W, 3] -- this GPU's, or (frame_sink, multi-GPU serving) slot r
            # of rank 0's gather buffer reached over NVLink, so the epilogue IS the gather
            lat = latents.to(torch.float16).flatten(0, 1) * (1 / self.vae.config.scaling_factor)
            sink = getattr(self, "frame_sink", None)
            n, hh, ww = lat.shape[0], lat.shape[-2] * self.vae_scale_factor, lat.shape[-1] * self.vae_scale_factor
            if sink is not None and output_type == "uint8_pt":
                u8 = sink.begin()                       # waits (on the stream) until the previous clip was consumed
                assert tuple(u8.shape) == (n, hh, ww, 3)
            else:
                u8 = torch.empty(n, hh, ww, 3, dtype=torch.uint8, device=lat.device)
            for i in range(0, n, decode_chunk_size):
                ch = lat[i:i + decode_chunk_size]
                self.vae.decode_uint8(ch, num_frames=ch.shape[0], out=u8[i:i + ch.shape[0]])
            if output_type == "uint8_pt":
                return [u8]
            if output_type == "uint8":
                return [u8.cpu().numpy()]
            return [[PIL.Image.fromarray(f) for f in u8.cpu().numpy()]]
        frames = self.decode_latents(latents.to(self.vae.dtype), num_frames, decode_chunk_size)
        return self._postprocess(frames, "uint8" if output_type == "uint8_pt" else output_type)

    @staticmethod
    def _postprocess(frames, output_type):
        """tensor2vid + VaeImageProcessor.postprocess (pipeline.py:57-69, Q21). frames [B, C, T, H, W]."""
        outs = []
        for b in range(frames.shape[0]):
            vid = (frames[b].permute(1, 0, 2, 3) / 2 + 0.5).clamp(0, 1)  # [T, C, H, W]
            if output_type == "pt":
                outs.append(vid)
                continue
            if output_type == "np":
                outs.append(vid.permute(0, 2, 3, 1).cpu().float().numpy())
                continue
            u8 = (vid * 255).round().to(torch.uint8).permute(0, 2, 3, 1).contiguous().cpu().numpy()
            if output_type == "uint8":
                outs.append(u8)
            elif output_type == "pil":
                outs.append([PIL.Image.fromarray(f) for f in u8])
            else:
                raise ValueError(f"unknown output_type {output_type}")
        return outs

    def _clip_inputs(self, image, controlnet_condition, height, width, num_frames, num_inference_steps, generator,
                     latents, noise_aug_strength, fps, motion_bucket_id, max_guidance_scale, batch_size=1,
                     num_videos_per_prompt=1):
        """Steps 0-6 of the reference __call__ shared by the Traj / Keypoint / Hybrid pipelines (pipeline.py:324-441):
        input checks, CLIP embedding (Q3), VAE latents of the noise-augmented frame (CPU RNG, Q7), added-time-id
        constants (Q4), timesteps, initial latents x init_noise_sigma, the CFG-duplicated condition image in [-1, 1].
        Returns (image_embeddings, image_latents, added_time_ids, latents, cond)."""
        self.check_inputs(image, height, width)
        if batch_size != 1 or num_videos_per_prompt != 1:
            raise NotImplementedError("one clip per call (the reference's batch>1 path is unusable with one image, "
                                      "pipeline.py:378-388,454); shard clips across processes/GPUs instead")
        if max_guidance_scale <= 1.0:
            raise ValueError("max_guidance_scale must be > 1: without CFG the reference substitutes the latents for "
                             "the condition image and flow (pipeline.py:393,396), which is not a usable mode")
        device = self._device
        image_embeddings = self._encode_image(image, device, num_videos_per_prompt, True)
        emb_dtype = image_embeddings.dtype
        img = _to_unit_tensor(image, height, width)
        # the noise is drawn where the preprocessed image lives (pipeline.py:339-341): the CPU for PIL / ndarray inputs
        # (a CUDA generator raises there, as in the reference), the tensor's own device for tensor inputs
        noise = _randn_tensor(img.shape, generator=generator, device=img.device, dtype=img.dtype)
        img = img.to(device) * 2.0 - 1.0
        img = img + noise_aug_strength * noise.to(img.device)
        needs_upcasting = self.vae.dtype == torch.float16 and self.vae.config.force_upcast
        if needs_upcasting:
            self.vae.to(dtype=torch.float32)
        image_latents = self._encode_vae_image(img.to(self.vae.dtype), device, num_videos_per_prompt, True)
        image_latents = image_latents.to(emb_dtype)
        if needs_upcasting:
            self.vae.to(dtype=torch.float16)
        # added time ids: computed from the arguments, then overwritten by constants (Q4, pipeline.py:430-440)
        _get_add_time_ids(noise_aug_strength, emb_dtype, batch_size, fps - 1, motion_bucket_id, unet=self.unet)
        added_time_ids = torch.cat([_get_add_time_ids(0.02, emb_dtype, batch_size, 6, 128, unet=self.unet)] * 2)
        added_time_ids = added_time_ids.to(device)
        self.scheduler.set_timesteps(num_inference_steps, device=device)
        latents = self.prepare_latents(batch_size, num_frames, self.unet.config.in_channels, height, width, emb_dtype,
                                       device, generator, latents)
        cond = _to_unit_tensor(controlnet_condition, height, width) * 2.0 - 1.0
        cond = torch.cat([cond] * 2).to(device, latents.dtype)
        return image_embeddings, image_latents, added_time_ids, latents, cond

    # ------------------------------------------------------------------ the call
    @torch.no_grad()
    def __call__(self, image, controlnet_condition=None, controlnet_flow=None, height: int = 576, width: int = 1024,
                 num_frames: Optional[int] = None, num_inference_steps: int = 25, min_guidance_scale: float = 1.0,
                 max_guidance_scale: float = 3.0, fps: int = 7, motion_bucket_id: int = 127,
                 noise_aug_strength: float = 0.02, decode_chunk_size: Optional[int] = None,
                 num_videos_per_prompt: Optional[int] = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil",
                 callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], return_dict: bool = True,
                 controlnet_cond_scale=1.0, batch_size=1):
        ops = self._ops
        height = height or self.unet.config.sample_size * self.vae_scale_factor
        width = width or self.unet.config.sample_size * self.vae_scale_factor
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        decode_chunk_size = decode_chunk_size if decode_chunk_size is not None else num_frames
        device = self._device
        ev = {k: (torch.cuda.Event(enable_timing=True) if device.type == "cuda" else _NoEvent())
              for k in ("t0", "enc", "loop", "dec")}
        ev["t0"].record()
        image_embeddings, image_latents, added_time_ids, latents, cond = self._clip_inputs(
            image, controlnet_condition, height, width, num_frames, num_inference_steps, generator, latents,
            noise_aug_strength, fps, motion_bucket_id, max_guidance_scale, batch_size, num_videos_per_prompt)
        timesteps = self.scheduler.timesteps
        if controlnet_flow is None or controlnet_flow.shape[1] != num_frames - 1:
            raise ValueError(f"controlnet_flow must be [1, {num_frames - 1}, 2, H, W]")
        controlnet_flow = torch.cat([controlnet_flow] * 2).to(device, latents.dtype)

        g = torch.linspace(min_guidance_scale, max_guidance_scale, num_frames).unsqueeze(0).to(device, latents.dtype)
        self._guidance_scale = g[(...,) + (None,) * 3]
        self._num_timesteps = len(timesteps)

        # ---- engine state for this clip
        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        hw, T = h * w, num_frames
        unet_net, ad_net = self.unet.net, self.controlnet.net
        # the step runner owns the persistent state (latents, model input) and, on CUDA, the captured graph of one
        # denoise step; it must exist before this clip's conditioning is written (persistent addresses, graph_step.py)
        runner = self._step_runner(unet_net, ad_net, T, h, w, min_guidance_scale, max_guidance_scale,
                                   controlnet_cond_scale)
        unet_net.prepare_clip(image_embeddings, added_time_ids)
        ad_net.prepare_clip(image_embeddings, added_time_ids)
        self.controlnet.prepare_condition(cond, controlnet_flow, force=True)
        sig = self.scheduler._sigmas_host
        ts = self.scheduler._timesteps_host
        self.scheduler._step_index = 0
        runner.begin_clip(latents[0].reshape(T, 4, hw), image_latents.reshape(2, 4, hw), ts, sig)
        ev["enc"].record()

        # 8. denoising loop (pipeline.py:447-511): one graph replay (or the same body eagerly) per step
        for i in range(len(ts)):
            runner.step(i)
            self.scheduler._step_index = i + 1
            if callback_on_step_end is not None:
                cur = runner.lat_h.reshape(1, T, 4, h, w)
                kw = {k: cur for k in callback_on_step_end_tensor_inputs if k == "latents"}
                outs = callback_on_step_end(self, i, timesteps[i], kw) or {}
                new = outs.pop("latents", None)
                # whatever the callback left in / returned as `latents` drives the next step, as in the reference
                # (pipeline.py:502-509; an in-place edit of the tensor it was handed counts too): rebuild the fused
                # model input from it
                if new is not None and new is not cur:
                    runner.lat_h.copy_(new.reshape(T, 4, hw))
                runner.rebuild_input(sig[i + 1])
        lat_h = runner.lat_h.clone()    # the runner's storage is reused by the next clip
        latents = lat_h.reshape(1, T, 4, h, w)
        ev["loop"].record()

        frames = self._decode_output(latents, num_frames, decode_chunk_size, output_type)
        ev["dec"].record()
        self._events = ev
        controlnet_flow_out = controlnet_flow
        if not return_dict:
            return frames, controlnet_flow_out
        return FlowControlNetPipelineOutput(frames=frames, controlnet_flow=controlnet_flow_out)

    def last_timings_ms(self):
        """Device times of the last call (after a synchronize): encode (CLIP+VAE enc+cond branch), loop, decode."""
        e = self._events
        torch.cuda.synchronize()
        return {"encode_ms": e["t0"].elapsed_time(e["enc"]), "loop_ms": e["enc"].elapsed_time(e["loop"]),
                "decode_ms": e["loop"].elapsed_time(e["dec"])}
