"""Hybrid pipeline (landmark adapter inside a face mask, trajectory adapter outside) with the reference's entry point
(/root/reference/MOFA-Video-Hybrid/pipeline/pipeline.py:92-116 constructor, :291-322 __call__, :432-507 loop).

Both adapters' conditioning branches are loop-invariant and run once per clip; per denoise step the two trunks run on
the same fused CFG input, their 12 + 1 residuals are blended with the nearest-resized mask (one elementwise kernel per
level, mask row period = h*w so it broadcasts over the 2*T frames; Q13: the mask is not CFG-duplicated), and the UNet +
fused CFG/Euler kernel finish the step."""
from typing import Callable, Dict, List, Optional

import torch
import torch.nn.functional as F

from mofa_video_b200 import lib as _lib
from mofa_video_b200.pipeline.pipeline import FlowControlNetPipeline as _TrajPipeline, FlowControlNetPipelineOutput


def level_masks(mask, h, w, n_levels, T, device):
    """{rows of a residual: fp16 mask [h_l*w_l]} -- F.interpolate(mask, (h_l, w_l), mode='nearest') per level (:479-485)."""
    m4 = mask.to(device=device, dtype=torch.float32).reshape(1, 1, *mask.shape[-2:])
    out, hh, ww = {}, h, w
    for _ in range(n_levels):
        out[2 * T * hh * ww] = F.interpolate(m4, (hh, ww), mode="nearest").reshape(hh * ww).to(torch.float16).contiguous()
        hh, ww = (hh - 1) // 2 + 1, (ww - 1) // 2 + 1
    return out


def denoise_hybrid(ops, unet_net, face_net, drag_net, by_rows, lat, il, sig, tsteps, h, w, g_min, g_max, scale_ldmk,
                   scale_traj, on_step=None):
    """The loop of H/pipeline/pipeline.py:432-507 on channels-last fp16 buffers; lat fp16 [T, 4, hw] is updated in
    place; both adapters' conditioning branches must already be prepared."""
    T, _, hw = lat.shape
    next_in = torch.empty(2 * T * hw, 8, dtype=torch.float16, device=lat.device)
    ops.cfg_euler_step(None, lat, il, next_in, T, hw, g_min, g_max, 0.0, sig[0])
    for i in range(len(tsteps)):
        rf, mf = face_net.adapter_forward(next_in, tsteps[i], h, w, scale_ldmk)
        rd, md = drag_net.adapter_forward(next_in, tsteps[i], h, w, scale_traj)
        for a, b in zip(rf + [mf], rd + [md]):
            m = by_rows[a.shape[0]]
            ops.mask_blend(a, b, m, a, period_rows=m.shape[0])           # face inside the mask, drag outside
        noise = unet_net.unet_forward(next_in, tsteps[i], h, w, rf, mf)
        ops.cfg_euler_step(noise, lat, il, next_in, T, hw, g_min, g_max, sig[i], sig[i + 1])
        if on_step is not None and on_step(i, lat) and i + 1 < len(tsteps):
            ops.cfg_euler_step(None, lat, il, next_in, T, hw, g_min, g_max, 0.0, sig[i + 1])
    return lat


class FlowControlNetPipeline(_TrajPipeline):
    def __init__(self, vae, image_encoder, unet, drag_controlnet, face_controlnet, scheduler, feature_extractor=None,
                 ops=None, device=None, native_vae=None, native_clip=None):
        super().__init__(vae, image_encoder, unet, drag_controlnet, scheduler, feature_extractor, ops=ops, device=device,
                         native_vae=native_vae, native_clip=native_clip)
        self.drag_controlnet, self.face_controlnet = drag_controlnet, face_controlnet

    def _hybrid_runner(self, unet_net, face_net, drag_net, T, h, w, g_min, g_max, scale_ldmk, scale_traj):
        from mofa_video_b200.graph_step import HybridStepRunner
        cache = self.__dict__.setdefault("_runners", {})
        key = ("hybrid", id(unet_net), id(face_net), id(drag_net), T, h, w, float(g_min), float(g_max),
               float(scale_ldmk), float(scale_traj))
        if key not in cache:
            while len(cache) >= 2:
                cache.pop(next(iter(cache)))
            cache[key] = HybridStepRunner(self._ops, unet_net, face_net, drag_net, T, h, w, g_min, g_max, scale_ldmk,
                                          scale_traj, self._device)
            cache[key].use_graph = cache[key].use_graph and getattr(self, "use_cuda_graph", True)
        return cache[key]

    @classmethod
    def from_pretrained(cls, path, unet=None, drag_controlnet=None, face_controlnet=None, image_encoder=None, vae=None,
                        scheduler=None, feature_extractor=None, torch_dtype=None, **_ignored):
        base = _TrajPipeline.from_pretrained(path, unet=unet, controlnet=drag_controlnet, image_encoder=image_encoder,
                                             vae=vae, scheduler=scheduler, feature_extractor=feature_extractor)
        if face_controlnet is None:
            raise ValueError("pass face_controlnet=")
        return cls(vae, image_encoder, unet, drag_controlnet, face_controlnet, base.scheduler, feature_extractor)

    @torch.no_grad()
    def __call__(self, image, controlnet_condition=None, controlnet_flow=None, landmarks=None, drag_flow=None,
                 mask=None, height: int = 576, width: int = 1024, num_frames: Optional[int] = None,
                 num_inference_steps: int = 25, min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0,
                 fps: int = 7, motion_bucket_id: int = 127, noise_aug_strength: float = 0.02,
                 decode_chunk_size: Optional[int] = None, num_videos_per_prompt: Optional[int] = 1, generator=None,
                 latents: Optional[torch.Tensor] = None, output_type: Optional[str] = "pil",
                 callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], return_dict: bool = True,
                 ctrl_scale_traj=1.0, ctrl_scale_ldmk=1.0, batch_size=1):
        ops = self._ops
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        decode_chunk_size = decode_chunk_size if decode_chunk_size is not None else num_frames
        if num_frames != self.unet.config.num_frames:
            raise ValueError(f"num_frames must equal the UNet's num_frames ({self.unet.config.num_frames})")
        device = self._device
        image_embeddings, image_latents, added_time_ids, latents, cond = self._clip_inputs(
            image, controlnet_condition, height, width, num_frames, num_inference_steps, generator, latents,
            noise_aug_strength, fps, motion_bucket_id, max_guidance_scale, batch_size, num_videos_per_prompt)
        flow = torch.cat([controlnet_flow] * 2).to(device, torch.float16)
        dflow = torch.cat([drag_flow] * 2).to(device, torch.float16)
        ldmk = torch.cat([landmarks] * 2).to(device, torch.float16)

        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        hw, T = h * w, num_frames
        unet_net, face, drag = self.unet.net, self.face_controlnet, self.drag_controlnet
        # step runner first: it makes the conditioning tensors persistent (graph_step.py)
        runner = self._hybrid_runner(unet_net, face.net, drag.net, T, h, w, min_guidance_scale, max_guidance_scale,
                                     ctrl_scale_ldmk, ctrl_scale_traj)
        for n in (unet_net, face.net, drag.net):
            n.prepare_clip(image_embeddings, added_time_ids)
        face.prepare_condition(cond, flow, ldmk, force=True)
        drag.prepare_condition(cond, dflow, force=True)
        runner.set_masks(level_masks(mask, h, w, len(self.unet.config.block_out_channels), T, device))
        sig, tsteps = self.scheduler._sigmas_host, self.scheduler._timesteps_host
        self._num_timesteps = len(tsteps)
        self.scheduler._step_index = 0
        runner.begin_clip(latents[0].reshape(T, 4, hw), image_latents.reshape(2, 4, hw), tsteps, sig)
        for i in range(len(tsteps)):
            runner.step(i)
            self.scheduler._step_index = i + 1
            if callback_on_step_end is not None:
                cur = runner.lat_h.reshape(1, T, 4, h, w)
                outs = callback_on_step_end(self, i, self.scheduler.timesteps[i], {"latents": cur}) or {}
                new = outs.pop("latents", None)
                if new is not None and new is not cur:
                    runner.lat_h.copy_(new.reshape(T, 4, hw))
                runner.rebuild_input(sig[i + 1])      # an in-place edit of `cur` counts too
        lat = runner.lat_h.clone()
        latents = lat.reshape(1, T, 4, h, w)
        frames = self._decode_output(latents, num_frames, decode_chunk_size, output_type)
        if not return_dict:
            return frames, controlnet_flow
        return FlowControlNetPipelineOutput(frames=frames, controlnet_flow=controlnet_flow)
