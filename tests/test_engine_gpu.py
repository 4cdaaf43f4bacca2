"""-m gpu: the engine-backed reference entry points (FlowControlNet.forward, UNet.forward,
FlowControlNetPipeline.__call__, softsplat) on the B200 against the fp32 CPU oracle.

Tolerances (fp16 storage / fp32 accumulate against an fp32 oracle, random-init weights):
  single adapter+UNet evaluation : max |err| <= 1e-2 * max |ref|   (measured ~1e-3)
  3-step pipeline latents        : max |err| <= 2e-2 * max |ref|
"""
import pytest
import torch

from oracle import fixtures
from oracle import pipeline as opipe
from oracle import scheduler as osched
from oracle.softsplat import softsplat as oracle_softsplat

pytestmark = pytest.mark.gpu


def rel_err(a, b):
    return ((a.float().cpu() - b.float().cpu()).abs().max() / (b.abs().max() + 1e-12)).item()


def build_pair(cfg, adapter_gain=20.0):
    from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import FlowControlNet
    from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
        UNetSpatioTemporalConditionControlNetModel
    unet, adapter = fixtures.make_models(cfg, seed=0, adapter_gain=adapter_gain)
    e_unet = UNetSpatioTemporalConditionControlNetModel.from_state_dict(unet.state_dict(), unet.config.__dict__)
    e_ad = FlowControlNet.from_state_dict(adapter.state_dict(), adapter.config.__dict__)
    return unet, adapter, e_unet, e_ad


def one_step(cfg, H, W, tol):
    unet, adapter, e_unet, e_ad = build_pair(cfg)
    inp = fixtures.make_step_inputs(cfg, H, W)
    t = torch.tensor(1.6377)
    with torch.no_grad():
        dres, mid, _, _ = adapter(inp["sample"], t, inp["encoder_hidden_states"], inp["added_time_ids"],
                                  controlnet_cond=inp["controlnet_cond"], controlnet_flow=inp["controlnet_flow"],
                                  conditioning_scale=0.8)
        ref = unet(inp["sample"], t, inp["encoder_hidden_states"], dres, mid, added_time_ids=inp["added_time_ids"])[0]
    cu = {k: v.cuda() for k, v in inp.items()}
    down, midr, flow_back, none = e_ad.forward(cu["sample"], 1.6377, cu["encoder_hidden_states"],
                                               cu["added_time_ids"], controlnet_cond=cu["controlnet_cond"],
                                               controlnet_flow=cu["controlnet_flow"], conditioning_scale=0.8,
                                               return_dict=False)
    assert none is None and flow_back is cu["controlnet_flow"] and len(down) == 12
    for k, (a, b) in enumerate(zip(down, dres)):
        assert a.shape == b.shape
        assert rel_err(a, b) < tol, f"adapter residual {k}: {rel_err(a, b)}"
    assert rel_err(midr, mid) < tol
    out = e_unet.forward(cu["sample"], 1.6377, cu["encoder_hidden_states"], down, midr,
                         added_time_ids=cu["added_time_ids"], return_dict=False)[0]
    assert out.shape == ref.shape
    e = rel_err(out, ref)
    assert e < tol, f"unet output rel err {e}"
    return e


def test_step_tiny_config():
    one_step(dict(fixtures.TINY_CONFIG), 16, 16, 1e-2)


def test_step_tiny_rectangular_more_frames():
    cfg = dict(fixtures.TINY_CONFIG)
    cfg["num_frames"] = 5
    one_step(cfg, 16, 24, 1e-2)


def test_step_full_width_channels():
    """Real SVD-XT channel widths / head counts (320,640,1280,1280; 5,10,20,20) at a small latent."""
    cfg = dict(num_frames=2)
    one_step(cfg, 16, 16, 1e-2)


def test_step_full_width_channels_fused_feedforward(monkeypatch):
    """The opt-in fused GEGLU FeedForward (mofa_ff_geglu) on the C = 320 level of the same step."""
    from mofa_video_b200 import lib
    monkeypatch.setattr(lib, "FF_FUSED_MAX_C", lib.FF_FUSED_LIMIT_C)
    n0 = lib.launch_count()
    one_step(dict(num_frames=2), 16, 16, 1e-2)
    assert lib.launch_count() > n0


def test_step_head_dim_128_generic_attention():
    """Reference class-default head geometry (d = 128 on one level): mofa_attn_small / mofa_attn_small_temporal."""
    cfg = dict(fixtures.TINY_CONFIG)
    cfg["num_attention_heads"] = (1, 2, 2, 4)
    one_step(cfg, 16, 24, 1e-2)


def test_attn_small_temporal_and_long_sequence():
    """The generic kernel on the strided temporal layout and across several 256-key passes (online softmax)."""
    import ref_ops
    from mofa_video_b200 import lib
    g = torch.Generator().manual_seed(5)
    B, T, HW, heads, d = 2, 7, 40, 3, 128
    C = heads * d
    qkv = torch.randn(B * T * HW, 3 * C, generator=g).half().cuda()
    out, ref = torch.empty(B * T * HW, C, dtype=torch.half, device="cuda"), torch.empty(B * T * HW, C, dtype=torch.half)
    lib.attn_small_temporal(qkv, out, B, T, HW, heads, d, d ** -0.5)
    ref_ops.attn_small_temporal(qkv.cpu(), ref, B, T, HW, heads, d, d ** -0.5)
    assert (out.float().cpu() - ref.float()).abs().max().item() < 3e-3
    n, L, heads, d = 2, 700, 2, 96
    C = heads * d
    qkv = torch.randn(n * L, 3 * C, generator=g).half().cuda()
    out, ref = torch.empty(n * L, C, dtype=torch.half, device="cuda"), torch.empty(n * L, C, dtype=torch.half)
    lib.attn_small(qkv, out, n, L, heads, d, d ** -0.5)
    ref_ops.attn_small(qkv.cpu(), ref, n, L, heads, d, d ** -0.5)
    assert (out.float().cpu() - ref.float()).abs().max().item() < 3e-3


def test_softsplat_entry_point():
    from mofa_video_b200.models.softsplat import softsplat
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 64, 12, 20, generator=g).half().float()
    fl = (torch.randn(2, 2, 12, 20, generator=g) * 3).half().float()
    ref = oracle_softsplat(x, fl, None, "avg")
    out = softsplat(x.cuda(), fl.cuda(), None, "avg")
    assert out.shape == ref.shape and out.dtype == torch.float32
    assert rel_err(out, ref) < 2e-3
    with pytest.raises(AssertionError):
        softsplat(x.cuda(), fl.cuda(), None, "bogus")


class _TinyClip(torch.nn.Module):
    """Stand-in for CLIPVisionModelWithProjection (a third-party constructor argument of the pipeline)."""

    def __init__(self, dim):
        super().__init__()
        self.proj = torch.nn.Linear(3 * 8 * 8, dim)

    def forward(self, x):
        from types import SimpleNamespace
        f = torch.nn.functional.adaptive_avg_pool2d(x, 8).flatten(1)
        return SimpleNamespace(image_embeds=self.proj(f))


def test_pipeline_three_steps_matches_oracle():
    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from mofa_video_b200.pipeline.pipeline import FlowControlNetPipeline
    from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler
    cfg = dict(fixtures.TINY_CONFIG)
    H, W = 128, 192
    T = cfg["num_frames"]
    unet, adapter, e_unet, e_ad = build_pair(cfg)
    torch.manual_seed(5)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=(32, 32, 64, 64)).eval()
    clip = _TinyClip(cfg["cross_attention_dim"]).eval()
    image = fixtures.make_image(H, W)
    flow = fixtures.make_flow(T, H, W)
    lat0 = torch.randn(1, T, 4, H // 8, W // 8, generator=torch.Generator().manual_seed(9))
    ref = opipe.run_pipeline(vae, clip, unet, adapter, osched.EulerDiscreteScheduler(), image, image, flow,
                             height=H, width=W, num_inference_steps=3, latents=lat0.clone(),
                             generator=torch.Generator().manual_seed(11), output_type="latent")
    pipe = FlowControlNetPipeline(vae=vae.cuda().half(), image_encoder=clip.cuda().half(), unet=e_unet,
                                  controlnet=e_ad, scheduler=EulerDiscreteScheduler())
    out = pipe(image, image, flow, height=H, width=W, num_inference_steps=3, latents=lat0.clone().half(),
               generator=torch.Generator().manual_seed(11), output_type="latent", decode_chunk_size=8)
    lat = out.frames
    assert lat.shape == ref.shape
    assert torch.isfinite(lat).all()
    e = rel_err(lat, ref)
    assert e < 2e-2, f"3-step latents rel err {e}"
    # full call with decode: shapes / dtypes / range of the reference's output contract
    out = pipe(image, image, flow, height=H, width=W, num_inference_steps=2, latents=lat0.clone().half(),
               output_type="pil", decode_chunk_size=2)
    assert len(out.frames) == 1 and len(out.frames[0]) == T and out.frames[0][0].size == (W, H)
    with pytest.raises(ValueError):
        pipe(image, image, flow, height=H + 4, width=W)
    with pytest.raises(ValueError):
        pipe(image, image, flow, height=H, width=W, max_guidance_scale=1.0)


def test_native_vae_decoder_matches_torch_module():
    """Native TemporalDecoder (tcgen05 convs, attention-as-GEMMs, fused tail) vs the fp32 PyTorch module."""
    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from mofa_video_b200.vae_engine import NativeTemporalDecoderVAE
    torch.manual_seed(0)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=(64, 128, 256, 256)).eval()
    with torch.no_grad():
        for n, p in vae.named_parameters():
            if n.endswith("mix_factor"):
                p.fill_(0.3)
            p.copy_(p.half().float())
    z = torch.randn(5, 4, 9, 16, generator=torch.Generator().manual_seed(1)).half().float()
    with torch.no_grad():
        ref = vae.decode(z, num_frames=5).sample
    nat = NativeTemporalDecoderVAE(vae)
    out = nat.decode(z.cuda(), num_frames=5).sample
    assert out.shape == ref.shape
    e = rel_err(out, ref)
    assert e < 2e-2, f"vae decode rel err {e}"
    u8 = nat.decode_uint8(z.cuda(), num_frames=5).cpu()
    want = ((ref / 2 + 0.5).clamp(0, 1) * 255).round().permute(0, 2, 3, 1)
    assert (u8.float() - want).abs().max() <= 3


def test_native_vae_encoder_matches_torch_module():
    """SURVEY.md §8 row a9 on the GPU: conv_in (im2col), TMA implicit-GEMM resnets, asymmetric-pad stride-2 convs,
    d = C attention as GEMMs, quant_conv -- vs the fp32 PyTorch module.  Real channel widths at a small image."""
    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from mofa_video_b200.vae_engine import NativeTemporalDecoderVAE
    torch.manual_seed(4)
    vae = AutoencoderKLTemporalDecoder().eval()
    with torch.no_grad():
        for n, p in vae.named_parameters():
            p.copy_(p.half().float())
    x = (torch.rand(1, 3, 128, 192, generator=torch.Generator().manual_seed(5)) * 2 - 1).half().float()
    with torch.no_grad():
        ref = vae.encode(x).latent_dist.mode()
    nat = NativeTemporalDecoderVAE(vae)
    out = nat.encode(x.cuda()).latent_dist.mode()
    assert out.shape == ref.shape == (1, 4, 16, 24)
    e = rel_err(out, ref)
    assert e < 2e-2, f"vae encode rel err {e}"


def test_step_full_frame_count_tiny_channels():
    """T = 25 (the real clip length) against the oracle on the GPU: temporal attention / conv over 25 frames."""
    cfg = dict(fixtures.TINY_CONFIG)
    cfg["num_frames"] = 25
    one_step(cfg, 16, 16, 1e-2)


def test_entry_points_against_reference_pipeline_latents():
    """The three engine entry points on the B200 (real kernels) DIRECTLY against latents produced by executing the
    reference's own Traj / Keypoint / Hybrid pipelines on CPU in fp32 (tests/golden/README.md)."""
    import os

    import PIL.Image

    from mofa_video_b200.models.ldmk_ctrlnet import FlowControlNet as FaceNet
    from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import FlowControlNet
    from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
        UNetSpatioTemporalConditionControlNetModel as UNet
    from mofa_video_b200.pipeline import pipeline_hybrid as hyb
    from mofa_video_b200.pipeline import svdxt_pipeline_ctrlnet_loop as kpl
    from mofa_video_b200.pipeline.pipeline import FlowControlNetPipeline
    from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler
    gold = os.path.join(os.path.dirname(__file__), "golden")

    def rel(a, b):
        return ((a.float().cpu() - b).abs().max() / b.abs().max()).item()

    g = torch.load(os.path.join(gold, "pipeline_tiny.pt"))
    cfg = g["config"]
    H, W = g["hw"]
    T = cfg["num_frames"]
    unet, drag = fixtures.make_models(cfg, seed=0, adapter_gain=20.0)
    vae, clip = fixtures.make_vae_and_clip(cfg["cross_attention_dim"])
    pil = PIL.Image.fromarray(g["image_u8"].numpy())
    pipe = FlowControlNetPipeline(vae=vae.cuda().half(), image_encoder=clip.cuda().half(),
                                  unet=UNet.from_state_dict(unet.state_dict(), unet.config.__dict__),
                                  controlnet=FlowControlNet.from_state_dict(drag.state_dict(), drag.config.__dict__),
                                  scheduler=EulerDiscreteScheduler())
    lat0 = torch.randn(1, T, 4, H // 8, W // 8, generator=torch.Generator().manual_seed(g["latent_seed"]))
    out = pipe(pil, pil, fixtures.make_flow(T, H, W), height=H, width=W, num_inference_steps=g["steps"],
               latents=lat0.clone().half(), generator=torch.Generator().manual_seed(g["generator_seed"]),
               output_type="latent", controlnet_cond_scale=g["cond_scale"]).frames
    assert rel(out, g["latents"]) < 2e-2

    g = torch.load(os.path.join(gold, "keypoint_pipeline_tiny.pt"))
    cfg = g["config"]
    F_frames = g["frames"]
    unet, drag = fixtures.make_models(cfg, seed=0, adapter_gain=20.0)
    face = fixtures.make_ldmk_adapter(cfg)
    e_unet = UNet.from_state_dict(unet.state_dict(), unet.config.__dict__)
    e_drag = FlowControlNet.from_state_dict(drag.state_dict(), drag.config.__dict__)
    e_face = FaceNet.from_state_dict(face.state_dict(), face.config.__dict__)
    vae, clip = fixtures.make_vae_and_clip(cfg["cross_attention_dim"])
    vae, clip = vae.cuda().half(), clip.cuda().half()
    gen = torch.Generator().manual_seed(3)
    flow = fixtures.make_flow(F_frames, H, W)
    ldmk = torch.rand(1, F_frames, 3, H, W, generator=gen).half().float()
    lat0 = torch.randn(1, F_frames, 4, H // 8, W // 8, generator=gen)
    pipe = kpl.FlowControlNetPipeline(vae=vae, image_encoder=clip, unet=e_unet, controlnet=e_face,
                                      scheduler=EulerDiscreteScheduler())
    out = pipe(pil, pil, flow, ldmk, height=H, width=W, num_frames=F_frames, num_inference_steps=g["steps"],
               latents=lat0.clone().half(), generator=torch.Generator().manual_seed(g["generator_seed"]),
               output_type="latent", window_size=T, stride=g["stride"]).frames
    assert rel(out, g["latents"]) < 3e-2

    g = torch.load(os.path.join(gold, "hybrid_pipeline_tiny.pt"))
    gen = torch.Generator().manual_seed(3)
    flow = fixtures.make_flow(T, H, W)
    ldmk = torch.rand(1, T, 3, H, W, generator=gen).half().float()
    lat0 = torch.randn(1, T, 4, H // 8, W // 8, generator=gen)
    drag_flow = (fixtures.make_flow(T, H, W, seed=99) * 0.5).half().float()
    mask = torch.zeros(1, 1, H, W)
    mask[..., 20:90, 30:100] = 1.0
    pipe = hyb.FlowControlNetPipeline(vae=vae, image_encoder=clip, unet=e_unet, drag_controlnet=e_drag,
                                      face_controlnet=e_face, scheduler=EulerDiscreteScheduler())
    out = pipe(pil, pil, flow, ldmk, drag_flow, mask, height=H, width=W, num_inference_steps=g["steps"],
               latents=lat0.clone().half(), generator=torch.Generator().manual_seed(g["generator_seed"]),
               output_type="latent", ctrl_scale_traj=g["scale_traj"], ctrl_scale_ldmk=g["scale_ldmk"]).frames
    assert rel(out, g["latents"]) < 3e-2
