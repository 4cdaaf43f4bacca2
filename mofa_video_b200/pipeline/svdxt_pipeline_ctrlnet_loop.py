"""Keypoint pipeline (landmark-driven portrait animation) with the reference's entry point
(/root/reference/MOFA-Video-Keypoint/pipeline/svdxt_pipeline_ctrlnet_loop.py:287-664): same constructor, same
__call__ signature plus `landmarks`, `window_size`, `stride` (:292-294), periodic / sliding-window sampling for videos
longer than the UNet's 25 frames (:426-429, 445-511).

Engine specifics: every temporal view's conditioning branch (cond pyramid, 96 softsplats, occlusion hourglasses,
landmark embedding) is loop-invariant and computed once per clip; identical views (T = 25 gives [(1,25),(1,25)], quirk
Q12) are evaluated once -- (x + x) / 2 == x exactly; each view runs the fused CFG + Euler kernel on its window of the
latents, and the per-step value/count averaging of the overlapping windows is the only remaining bookkeeping."""
from typing import Callable, Dict, List, Optional

import PIL.Image
import torch

from mofa_video_b200 import lib as _lib
from mofa_video_b200.pipeline.pipeline import FlowControlNetPipeline as _TrajPipeline, FlowControlNetPipelineOutput


def window_views(num_frames, window_size, stride):
    window_num = (num_frames - window_size) // stride + 1
    views = [(1 + i * stride, i * stride + window_size) for i in range(window_num)]
    return views + [(num_frames - window_size + 1, num_frames)]


def unique_views(views):
    """[(view, multiplicity)] in first-seen order: T = window_size yields the same view twice (Q12)."""
    uniq = []
    for v in views:
        for u in uniq:
            if u[0] == v:
                u[1] += 1
                break
        else:
            uniq.append([v, 1])
    return [(v, m) for v, m in uniq]


def views_for_rank(n_views, world, rank):
    """Window k of a long video -> rank k mod world (the windows of one step are independent)."""
    return list(range(rank, n_views, world))


def denoise_windowed(ops, unet_net, ad_net, view_states, lat, il, sig, tsteps, h, w, T, g_min, g_max, scale,
                     on_step=None, shard=None):
    """The loop of svdxt_pipeline_ctrlnet_loop.py:445-511 on channels-last fp16 buffers.
    view_states: [((ts, te), multiplicity, (warped, ldmk))] -- cached conditioning of every distinct view (entries this
    rank does not own may be None);  lat fp16 [F, 4, hw]; il fp16 [2, 4, hw] (image latents are the same for every frame).
    shard = (world, rank): the views of every step are split over the ranks (view k -> rank k mod world) and the
    per-frame (value, count) sums meet in ONE all-reduce per step -- the only exchange of the path that is not a final
    gather ([F, 5, hw] fp32: 1.6 MB at 512x512x37 frames); every rank then holds the same averaged latents.
    Returns the final lat."""
    import torch.distributed as dist
    F_, _, hw = lat.shape
    dev = lat.device
    world, rank = shard if shard is not None else (1, 0)
    next_in = torch.empty(2 * T * hw, 8, dtype=torch.float16, device=dev)
    for i in range(len(tsteps)):
        acc = torch.zeros(F_, 5, hw, dtype=torch.float32, device=dev)      # channels 0-3: value, channel 4: count
        value, count = acc[:, :4], acc[:, 4:5, :1]
        for k, state in enumerate(view_states):
            if k % world != rank:
                continue
            (ts, te), mult, (warped, ldm) = state
            lt = lat[[0] + list(range(ts, te))].contiguous()             # [T, 4, hw] window of the latents
            ad_net.warped, ad_net.ldmk = warped, ldm
            ops.cfg_euler_step(None, lt, il, next_in, T, hw, g_min, g_max, 0.0, sig[i])
            res, mid = ad_net.adapter_forward(next_in, tsteps[i], h, w, scale)
            noise = unet_net.unet_forward(next_in, tsteps[i], h, w, res, mid)
            ops.cfg_euler_step(noise, lt, il, next_in, T, hw, g_min, g_max, sig[i], sig[i + 1])
            if k == 0:                                                   # :501-511
                value[0:te] += mult * lt.float()
                count[0:te] += mult
            else:
                value[ts:te] += mult * lt[1:].float()
                count[ts:te] += mult
        if world > 1:
            dist.all_reduce(acc)                                           # NCCL on device tensors, gloo on CPU tensors
        lat = torch.where(count > 0, value / count.clamp(min=1), value).to(torch.float16)
        if on_step is not None:
            lat = on_step(i, lat)
    return lat


class FlowControlNetPipeline(_TrajPipeline):
    @torch.no_grad()
    def __call__(self, image, controlnet_condition=None, controlnet_flow=None, landmarks=None, height: int = 576,
                 width: int = 1024, num_frames: Optional[int] = None, num_inference_steps: int = 25,
                 min_guidance_scale: float = 1.0, max_guidance_scale: float = 3.0, fps: int = 7,
                 motion_bucket_id: int = 127, noise_aug_strength: float = 0.02, decode_chunk_size: Optional[int] = None,
                 num_videos_per_prompt: Optional[int] = 1, generator=None, latents: Optional[torch.Tensor] = None,
                 output_type: Optional[str] = "pil",
                 callback_on_step_end: Optional[Callable[[int, int, Dict], None]] = None,
                 callback_on_step_end_tensor_inputs: List[str] = ["latents"], return_dict: bool = True,
                 controlnet_cond_scale=1.0, batch_size=1, window_size=25, stride=12, shard_windows=False):
        """shard_windows (engine extension, default off = every process renders its own clip): when torch.distributed is
        initialised, ONE long clip is rendered by all ranks together -- the temporal windows of each denoise step are
        split over the ranks and their latents averaged with one all-reduce per step (all ranks must pass the same
        inputs and seeds; all return the same frames)."""
        ops = self._ops
        num_frames = num_frames if num_frames is not None else self.unet.config.num_frames
        decode_chunk_size = decode_chunk_size if decode_chunk_size is not None else num_frames
        if window_size != self.unet.config.num_frames:
            raise ValueError(f"window_size must equal the UNet's num_frames ({self.unet.config.num_frames})")
        device = self._device
        image_embeddings, image_latents, added_time_ids, latents, cond = self._clip_inputs(
            image, controlnet_condition, height, width, num_frames, num_inference_steps, generator, latents,
            noise_aug_strength, fps, motion_bucket_id, max_guidance_scale, batch_size, num_videos_per_prompt)
        if controlnet_flow.shape[1] != num_frames - 1 or landmarks.shape[1] != num_frames:
            raise ValueError("controlnet_flow must have num_frames-1 frames and landmarks num_frames frames")
        flow = controlnet_flow.to(device, torch.float16)
        ldmk = landmarks.to(device, torch.float16)

        h, w = height // self.vae_scale_factor, width // self.vae_scale_factor
        hw, T = h * w, window_size
        unet_net, ad = self.unet.net, self.controlnet
        unet_net.prepare_clip(image_embeddings, added_time_ids)
        ad.net.prepare_clip(image_embeddings, added_time_ids)
        import torch.distributed as dist
        shard = None
        if shard_windows and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            shard = (dist.get_world_size(), dist.get_rank())
        views = unique_views(window_views(num_frames, window_size, stride))
        sig, tsteps = self.scheduler._sigmas_host, self.scheduler._timesteps_host
        self._num_timesteps = len(tsteps)

        def on_step(i, cur_lat):
            self.scheduler._step_index = i + 1
            if callback_on_step_end is None:
                return cur_lat
            cur = cur_lat.reshape(1, num_frames, 4, h, w)
            outs = callback_on_step_end(self, i, self.scheduler.timesteps[i], {"latents": cur}) or {}
            new = outs.pop("latents", None)
            return cur_lat if new is None else new.reshape(num_frames, 4, hw).to(torch.float16).contiguous()

        if len(views) == 1 and shard is None:
            # num_frames == window_size (configs[2]): the two windows of the reference are the same view (Q12), the
            # value / count average of identical latents is the identity, and the loop is the Traj loop with the landmark
            # adapter -- so it runs through the same captured-step runner (graph_step.StepRunner)
            (ts, te), _ = views[0]
            runner = self._step_runner(unet_net, ad.net, T, h, w, min_guidance_scale, max_guidance_scale,
                                       controlnet_cond_scale)
            unet_net.prepare_clip(image_embeddings, added_time_ids)
            ad.net.prepare_clip(image_embeddings, added_time_ids)
            ad.prepare_condition(cond, torch.cat([flow[:, (ts - 1):(te - 1)]] * 2),
                                 torch.cat([torch.cat([ldmk[:, 0:1], ldmk[:, ts:te]], dim=1)] * 2), force=True)
            runner.begin_clip(latents[0].reshape(num_frames, 4, hw), image_latents.reshape(2, 4, hw), tsteps, sig)
            for i in range(len(tsteps)):
                runner.step(i)
                new = on_step(i, runner.lat_h)
                if callback_on_step_end is not None:
                    if new is not runner.lat_h:
                        runner.lat_h.copy_(new)
                    runner.rebuild_input(sig[i + 1])
            lat = runner.lat_h.clone()
        else:
            states = []
            for k, ((ts, te), mult) in enumerate(views):
                if shard is not None and k % shard[0] != shard[1]:
                    states.append(None)            # another rank's window: no conditioning needed here
                    continue
                # loop-invariant conditioning of every distinct view
                fl = flow[:, (ts - 1):(te - 1)]
                lm = torch.cat([ldmk[:, 0:1], ldmk[:, ts:te]], dim=1)
                ad.prepare_condition(cond, torch.cat([fl] * 2), torch.cat([lm] * 2), force=True)
                states.append(((ts, te), mult, (ad.net.warped, ad.net.ldmk)))
            lat = latents[0].to(torch.float16).reshape(num_frames, 4, hw).contiguous()
            il = image_latents.to(torch.float16).reshape(2, 4, hw).contiguous()
            lat = denoise_windowed(ops, unet_net, ad.net, states, lat, il, sig, tsteps, h, w, T, min_guidance_scale,
                                   max_guidance_scale, controlnet_cond_scale, on_step, shard=shard)
        latents = lat.reshape(1, num_frames, 4, h, w)
        frames = self._decode_output(latents, num_frames, decode_chunk_size, output_type)
        if not return_dict:
            return frames, controlnet_flow
        return FlowControlNetPipelineOutput(frames=frames, controlnet_flow=controlnet_flow)
