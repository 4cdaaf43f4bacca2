"""-m "not gpu": the engine models read the reference's checkpoint layout unchanged
(/root/reference/MOFA-Video-Traj/README.md:20-38: `<dir>/[subfolder/]config.json` +
`diffusion_pytorch_model[.fp16].safetensors`; loaded at T/run_gradio.py:119-128 with subfolder= / variant=)."""
import json
import os

import pytest
import torch
from safetensors.torch import save_file

from mofa_video_b200 import synthetic
from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import FlowControlNet
from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import UNetSpatioTemporalConditionControlNetModel
from oracle import fixtures


def _write(d, cfg, sd, name):
    os.makedirs(d, exist_ok=True)
    with open(os.path.join(d, "config.json"), "w") as f:
        json.dump({"_class_name": "X", "_diffusers_version": "0.24.0", **{k: (list(v) if isinstance(v, tuple) else v)
                                                                         for k, v in cfg.items()}}, f)
    save_file({k: v.contiguous() for k, v in sd.items()}, os.path.join(d, name))


def test_from_pretrained_reads_reference_layout(tmp_path):
    cfg = dict(fixtures.TINY_CONFIG)
    cu, su = synthetic.unet_state_dict(cfg)
    ca, sa = synthetic.adapter_state_dict(cfg)
    svd = tmp_path / "stable-video-diffusion-img2vid-xt-1-1"
    _write(str(svd / "unet"), cu, su, "diffusion_pytorch_model.fp16.safetensors")
    _write(str(tmp_path / "controlnet"), ca, sa, "diffusion_pytorch_model.safetensors")
    unet = UNetSpatioTemporalConditionControlNetModel.from_pretrained(str(svd), subfolder="unet", variant="fp16",
                                                                      low_cpu_mem_usage=True, device="cpu")
    ad = FlowControlNet.from_pretrained(str(tmp_path / "controlnet"), device="cpu")
    assert unet.config.num_frames == cfg["num_frames"] and tuple(unet.config.block_out_channels) == (64, 128, 256, 256)
    assert not hasattr(unet.config, "_class_name")
    ref = UNetSpatioTemporalConditionControlNetModel.from_state_dict(su, cu, device="cpu")
    # identical packed weights whichever way the state dict arrived
    for key in ("conv_in", "conv_out", "norm_out"):
        for ta, tb in zip(unet.net.p[key], ref.net.p[key]):
            assert torch.equal(ta, tb)
    assert ad.dtype == torch.float16 and ad.requires_grad_(False) is ad and ad.to("cpu", dtype=torch.float16) is ad
    with pytest.raises(ValueError):
        ad.to(dtype=torch.float32)                       # the engine is fp16 only
    with pytest.raises(FileNotFoundError):
        FlowControlNet.from_pretrained(str(tmp_path / "nowhere-else" / ".."), subfolder="controlnet_missing")


def test_packed_weight_cache_round_trip(tmp_path):
    """from_pretrained(pack_cache=): the second load skips the repack and runs the same network bit for bit."""
    import ref_ops
    from mofa_video_b200.models import _base
    cfg = dict(fixtures.TINY_CONFIG)
    cu, su = synthetic.unet_state_dict(cfg)
    ca, sa = synthetic.adapter_state_dict(cfg)
    svd = tmp_path / "svd"
    _write(str(svd / "unet"), cu, su, "diffusion_pytorch_model.fp16.safetensors")
    _write(str(tmp_path / "controlnet"), ca, sa, "diffusion_pytorch_model.safetensors")
    cache = str(tmp_path / "pack")
    inp = fixtures.make_step_inputs(cfg, 16, 16)
    outs = []
    with _base.default_backend(ref_ops, "cpu"):
        for attempt in range(2):
            unet = UNetSpatioTemporalConditionControlNetModel.from_pretrained(str(svd), subfolder="unet", variant="fp16",
                                                                              pack_cache=cache)
            ad = FlowControlNet.from_pretrained(str(tmp_path / "controlnet"), pack_cache=cache)
            assert len(os.listdir(cache)) == 2
            down, mid, _, _ = ad.forward(inp["sample"], 1.6377, inp["encoder_hidden_states"], inp["added_time_ids"],
                                         controlnet_cond=inp["controlnet_cond"], controlnet_flow=inp["controlnet_flow"],
                                         return_dict=False)
            outs.append(unet.forward(inp["sample"], 1.6377, inp["encoder_hidden_states"], down, mid,
                                     added_time_ids=inp["added_time_ids"], return_dict=False)[0])
            if attempt == 1:
                with pytest.raises(RuntimeError):
                    unet.state_dict()                   # restored from the cache: no reference-layout tensors in memory
                assert unet.config.num_frames == cfg["num_frames"] and unet.add_embedding.linear_1.in_features == 96
    assert torch.equal(outs[0], outs[1])
    # touching the checkpoint invalidates the entry
    os.utime(str(svd / "unet" / "diffusion_pytorch_model.fp16.safetensors"), ns=(1, 1))
    with _base.default_backend(ref_ops, "cpu"):
        UNetSpatioTemporalConditionControlNetModel.from_pretrained(str(svd), subfolder="unet", variant="fp16",
                                                                   pack_cache=cache)
    assert len(os.listdir(cache)) == 3
