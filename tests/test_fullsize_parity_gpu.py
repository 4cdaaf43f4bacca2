"""-m gpu: ELEMENT-WISE parity at the BASELINE.json configurations (not at toy shapes, not through identities).

  configs[1]/[3]/[4]  576x1024 : one adapter + UNet evaluation at the real SVD-XT widths / heads on the full 72x128 latent,
                                 2 frames per CFG half (4 of the 50 frames of a step -- spatial attention at L = 9216,
                                 every conv / GroupNorm / FF at its production row count per frame): all 12 adapter
                                 residuals + mid + the UNet output vs oracle.models, max|err| <= 1e-2 * max|ref|
  configs[2]          512x512  : the same with the Keypoint (landmark) adapter: residuals, mid, occlusion masks, UNet
  configs[0]          256x256x14 frames, 2 steps, IN FULL: FlowControlNetPipeline.__call__ (CLIP ViT-H, VAE encode,
                                 2 x (adapter + UNet + CFG + Euler), temporal VAE decode) vs oracle.pipeline.run_pipeline
                                 with the same fp32 modules: latents <= 2e-2 * max|ref|, and the decoded uint8 frames
                                 under the per-pixel tolerance stated in the test
  VAE encode 576x1024          : native fp16-storage encoder vs the fp32 module (the reference upcasts, pipeline.py:343-352)

The fp32 oracle runs on the host cores (16 threads, tests/conftest.py): ~1 min per 576x1024 step, a few minutes in total.
"""
import os

import pytest
import torch

from oracle import fixtures
from oracle import pipeline as opipe
from oracle import scheduler as osched
from test_engine_gpu import one_step, rel_err

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _threads():
    old = torch.get_num_threads()
    torch.set_num_threads(min(32, os.cpu_count() or 1))
    yield
    torch.set_num_threads(old)


def test_traj_step_576x1024_elementwise():
    e = one_step(dict(num_frames=2), 72, 128, 1e-2)
    print(f"\n[fullsize] traj 576x1024 step: unet rel err {e:.2e}")


def test_keypoint_step_512x512_elementwise():
    from mofa_video_b200.models.ldmk_ctrlnet import FlowControlNet as LdmkFlowControlNet
    from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
        UNetSpatioTemporalConditionControlNetModel
    cfg = dict(num_frames=2)
    H = W = 64
    T = 2
    unet, _ = fixtures.make_models(cfg, seed=0)
    ad = fixtures.make_ldmk_adapter(cfg)
    inp = fixtures.make_step_inputs(cfg, H, W)
    landmarks = torch.rand(1, T, 3, 8 * H, 8 * W, generator=torch.Generator().manual_seed(11)).half().float()
    landmarks = landmarks.repeat(2, 1, 1, 1, 1)
    t = torch.tensor(1.6377)
    with torch.no_grad():
        dres, mid, _, occ = ad(inp["sample"], t, inp["encoder_hidden_states"], inp["added_time_ids"],
                               controlnet_cond=inp["controlnet_cond"], controlnet_flow=inp["controlnet_flow"],
                               landmarks=landmarks, conditioning_scale=0.9)
        ref = unet(inp["sample"], t, inp["encoder_hidden_states"], dres, mid, added_time_ids=inp["added_time_ids"])[0]
    e_ad = LdmkFlowControlNet.from_state_dict(ad.state_dict(), ad.config.__dict__)
    e_unet = UNetSpatioTemporalConditionControlNetModel.from_state_dict(unet.state_dict(), unet.config.__dict__)
    cu = {k: v.cuda() for k, v in inp.items()}
    down, midr, _, masks = e_ad(cu["sample"], 1.6377, cu["encoder_hidden_states"], cu["added_time_ids"],
                                controlnet_cond=cu["controlnet_cond"], controlnet_flow=cu["controlnet_flow"],
                                landmarks=landmarks.cuda(), conditioning_scale=0.9, return_dict=False)
    for k, (a, b) in enumerate(zip(down, dres)):
        assert a.shape == b.shape
        assert rel_err(a, b) < 1e-2, f"keypoint residual {k}: {rel_err(a, b)}"
    assert rel_err(midr, mid) < 1e-2
    for m_e, m_o in zip(masks, occ):
        assert m_e.shape == m_o.shape and (m_e.float().cpu() - m_o).abs().max().item() < 1e-2
    out = e_unet.forward(cu["sample"], 1.6377, cu["encoder_hidden_states"], down, midr,
                         added_time_ids=cu["added_time_ids"], return_dict=False)[0]
    e = rel_err(out, ref)
    assert e < 1e-2, f"unet output rel err {e}"
    print(f"\n[fullsize] keypoint 512x512 step: unet rel err {e:.2e}")


def test_config0_full_pipeline_256x256x14f_2steps_latents_and_frames():
    """BASELINE.json configs[0] in full, through the reference-facing call, against the fp32 oracle pipeline.

    Per-pixel tolerance on the decoded uint8 frames (north_star: "outputs match the reference fp32 path ... within a
    stated per-pixel fp tolerance"):  mean |diff| <= 0.25 / 255,  99.9 % of the pixel values within 2 / 255,  max <= 4 / 255
    (measured on B200, round 2: mean 0.065, p99.9 1, max 2; latents 2.1e-3 of max|ref|).  fp16 storage through
    2 x (adapter + UNet) and the 13-block temporal decoder against an all-fp32 oracle."""
    from mofa_video_b200.factory import make_clip_vit_h
    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import FlowControlNet
    from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
        UNetSpatioTemporalConditionControlNetModel
    from mofa_video_b200.pipeline.pipeline import FlowControlNetPipeline
    from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler
    H = W = 256
    T = 14
    cfg = dict(num_frames=T)
    unet, adapter = fixtures.make_models(cfg, seed=0, adapter_gain=20.0)
    torch.manual_seed(5)
    vae = AutoencoderKLTemporalDecoder().eval()                     # real widths (128, 256, 512, 512)
    clip = make_clip_vit_h(1024).eval()                             # CLIP ViT-H/14 vision tower, random init
    with torch.no_grad():
        for m in (vae, clip):
            for p in m.parameters():
                p.copy_(p.half().float())
    image = fixtures.make_image(H, W)
    flow = fixtures.make_flow(T, H, W)
    lat0 = torch.randn(1, T, 4, H // 8, W // 8, generator=torch.Generator().manual_seed(9))
    kw = dict(height=H, width=W, num_inference_steps=2, decode_chunk_size=8)
    ref_lat = opipe.run_pipeline(vae, clip, unet, adapter, osched.EulerDiscreteScheduler(), image, image, flow,
                                 latents=lat0.clone(), generator=torch.Generator().manual_seed(11),
                                 output_type="latent", **kw)
    with torch.no_grad():                                            # tensor2vid (pipeline.py:57-69) on the oracle frames
        lat = ref_lat.flatten(0, 1) / vae.config.scaling_factor
        ref_frames = torch.cat([vae.decode(lat[i:i + 8], num_frames=lat[i:i + 8].shape[0]).sample
                                for i in range(0, T, 8)], 0)
        ref_u8 = ((ref_frames / 2 + 0.5).clamp(0, 1) * 255).round().to(torch.uint8).permute(0, 2, 3, 1)  # [T,H,W,3]
    e_unet = UNetSpatioTemporalConditionControlNetModel.from_state_dict(unet.state_dict(), unet.config.__dict__)
    e_ad = FlowControlNet.from_state_dict(adapter.state_dict(), adapter.config.__dict__)
    import copy
    pipe = FlowControlNetPipeline(vae=copy.deepcopy(vae).cuda().half(), image_encoder=copy.deepcopy(clip).cuda().half(),
                                  unet=e_unet, controlnet=e_ad, scheduler=EulerDiscreteScheduler())
    assert hasattr(pipe.vae, "decode_uint8")                         # the VAE was re-hosted on the kernels
    out = pipe(image, image, flow, latents=lat0.clone().half(), generator=torch.Generator().manual_seed(11),
               output_type="latent", **kw)
    e = rel_err(out.frames, ref_lat)
    assert torch.isfinite(out.frames).all() and e < 2e-2, f"config0 latents rel err {e}"
    out = pipe(image, image, flow, latents=lat0.clone().half(), generator=torch.Generator().manual_seed(11),
               output_type="uint8", **kw)
    got = torch.from_numpy(out.frames[0])
    assert got.shape == ref_u8.shape == (T, H, W, 3)
    d = (got.int() - ref_u8.int()).abs().float()
    mean, q999, mx = d.mean().item(), torch.quantile(d.flatten()[:: 7], 0.999).item(), d.max().item()
    print(f"\n[fullsize] config0 256x256x14f 2 steps: latents rel err {e:.2e}; frames |diff| mean {mean:.3f} "
          f"p99.9 {q999:.1f} max {mx:.0f} (of 255); ref frame std {ref_u8.float().std().item():.1f}")
    assert ref_u8.float().std().item() > 5.0, "degenerate reference frames (all-constant) would make this test vacuous"
    assert mean <= 0.25 and q999 <= 2.0 and mx <= 4.0


def test_vae_encode_576x1024_fp16_storage_vs_fp32_module():
    """Row a9: the reference upcasts the VAE to fp32 for the encode (pipeline.py:343-352); the engine keeps fp16 storage
    with fp32 accumulation / statistics.  Full-size image, real widths, against the fp32 module (evaluated in fp32 on the
    same GPU by PyTorch, test infrastructure).  Measured on B200 (round 2): max 2.6e-3 * max|ref|, rms 1.8e-3 -- the
    rounding of ~25 fp16 activation stores; the reference itself rounds this latent to fp16 right after the encode
    (`image_latents.to(image_embeddings.dtype)`, pipeline.py:349: 5e-4), and the end-to-end frames of configs[0] above
    agree to 2/255 with this encoder in the path.  Bound: 4e-3, finite (DESIGN.md section 5 states the deviation)."""
    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from mofa_video_b200.vae_engine import NativeTemporalDecoderVAE
    torch.manual_seed(5)
    vae = AutoencoderKLTemporalDecoder().eval()
    with torch.no_grad():
        for p in vae.parameters():
            p.copy_(p.half().float())
    x = (fixtures.make_image(576, 1024) * 2 - 1)[None]
    x = (x + 0.02 * torch.randn(x.shape, generator=torch.Generator().manual_seed(3))).half().float()
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = False
    try:
        with torch.no_grad():
            ref = vae.cuda().float().encode(x.cuda()).latent_dist.mode().cpu()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    nat = NativeTemporalDecoderVAE(vae.half())
    got = nat.encode(x.cuda().half()).latent_dist.mode().float().cpu()
    assert got.shape == ref.shape == (1, 4, 72, 128) and torch.isfinite(got).all()
    e = rel_err(got, ref)
    rms = ((got - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt()).item()
    print(f"\n[fullsize] VAE encode 576x1024 fp16-storage vs fp32: max rel err {e:.2e}, rms rel err {rms:.2e}")
    assert e < 4e-3 and rms < 3e-3, (e, rms)
