"""-m gpu: Keypoint / Hybrid rows on the B200 through the C-ABI: helper kernels vs
tests/ref_ops.py, 7x7 / column-offset implicit-GEMM convs, the landmark adapter and the windowed / hybrid loops vs the
fp32 oracle (oracle/keypoint.py, hourglass pinned by tests/golden/hourglass_small.pt)."""
import pytest
import torch
import torch.nn.functional as F

import ref_ops as R
from oracle import fixtures
from oracle import keypoint as kp
from test_keypoint_cpu import _loop_setup, from_cl, make_ldmk_adapter, rel, to_cl

pytestmark = pytest.mark.gpu
DEV = "cuda"


def h(*s, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*s, generator=g) * scale).half().to(DEV)


def test_flow_pyramid_mask_blend_downsample():
    from mofa_video_b200 import lib
    Fn, Hf, Wf = 3, 64, 128
    flow = h(Fn, 2, Hf, Wf, seed=1, scale=5.0)
    for s in (8, 16, 32, 64):
        hs, ws = Hf // s, Wf // s
        o = torch.zeros(Fn * hs * ws, 72, dtype=torch.half, device=DEV)
        r = torch.zeros_like(o)
        lib.flow_pyramid(flow, o, Fn, hs, ws, Hf, Wf, 72, 35)
        R.flow_pyramid(flow, r, Fn, hs, ws, Hf, Wf, 72, 35)
        assert (o.float() - r.float()).abs().max().item() < 2e-3, s
    rows, C, period = 6 * 40, 96, 40
    a, b = h(rows, C, seed=2), h(rows, C, seed=3)
    for m in (torch.rand(period, generator=torch.Generator().manual_seed(4)).half().to(DEV),
              torch.rand(rows, generator=torch.Generator().manual_seed(5)).half().to(DEV)):
        o, r = torch.empty_like(a), torch.empty_like(a)
        lib.mask_blend(a, b, m, o, period_rows=m.shape[0])
        R.mask_blend(a, b, m, r, period_rows=m.shape[0])
        assert (o.float() - r.float()).abs().max().item() < 2e-3
    a2 = a.clone()
    lib.mask_blend(a2, b, m, a2)                       # in place (the hybrid loop blends into the face residual)
    assert (a2.float() - r.float()).abs().max().item() < 2e-3
    n, H, W, C = 3, 8, 12, 64
    x = h(n * H * W, C, seed=6)
    for s in (2, 4):
        o = torch.zeros(n * (H // s) * (W // s), C, dtype=torch.half, device=DEV)
        r = torch.zeros_like(o)
        lib.downsample_nearest(x, o, n, H, W, C, s)
        R.downsample_nearest(x, r, n, H, W, C, s)
        assert torch.equal(o, r)


def test_conv7x7_and_column_offset():
    """ksize=7 implicit GEMM (matting heads) and a conv writing into a column slice of a wider buffer (in-place concat)."""
    from mofa_video_b200 import lib
    n, H, W, C, N = 3, 9, 16, 64, 32
    x = h(n * H * W, C, seed=1, scale=0.5)
    wt = (torch.randn(N, C, 7, 7, generator=torch.Generator().manual_seed(2)) * 0.02).half()
    b = h(N, seed=3)
    ref = F.conv2d(x.float().cpu().reshape(n, H, W, C).permute(0, 3, 1, 2), wt.float(), b.float().cpu(), padding=3)
    wk = wt.permute(0, 2, 3, 1).reshape(N, 49 * C).contiguous().to(DEV)
    out = torch.zeros(n * H * W, 96, dtype=torch.half, device=DEV)
    lib.gemm(lib.A_CONV3X3, x, wk, out, N=N, n_img=n, H=H, W=W, C=C, ksize=7, bias=b, ldc=96, c_off=64)
    got = out[:, 64:].float().cpu().reshape(n, H, W, N).permute(0, 3, 1, 2)
    assert rel(got, ref) < 3e-3
    assert out[:, :64].abs().max().item() == 0
    # N = 1 (matting mask head, sigmoid)
    w1 = (torch.randn(1, C, 7, 7, generator=torch.Generator().manual_seed(4)) * 0.02).half()
    ref1 = torch.sigmoid(F.conv2d(x.float().cpu().reshape(n, H, W, C).permute(0, 3, 1, 2), w1.float(), None, padding=3))
    o1 = torch.zeros(n * H * W, 1, dtype=torch.half, device=DEV)
    lib.gemm(lib.A_CONV3X3, x, w1.permute(0, 2, 3, 1).reshape(1, 49 * C).contiguous().to(DEV), o1, N=1, n_img=n, H=H,
             W=W, C=C, ksize=7, act=4)
    assert (o1.float().cpu().reshape(n, 1, H, W) - ref1).abs().max().item() < 3e-3


def _gpu_nets(s):
    from mofa_video_b200 import engine, lib
    from mofa_video_b200.keypoint_engine import LdmkAdapterNet
    u = engine.Net("unet", s["unet"].state_dict(), s["unet"].config.__dict__, lib, DEV)
    d = engine.Net("adapter", s["drag"].state_dict(), s["drag"].config.__dict__, lib, DEV)
    f = LdmkAdapterNet(s["face"].state_dict(), s["face"].config.__dict__, lib, DEV)
    for n in (u, d, f):
        n.prepare_clip(s["emb"].to(DEV), s["ids"].to(DEV))
    return u, d, f


def test_keypoint_adapter_matches_oracle_gpu():
    from mofa_video_b200.models.ldmk_ctrlnet import FlowControlNet
    cfg = dict(fixtures.TINY_CONFIG)
    H = W = 16
    T = cfg["num_frames"]
    ad = make_ldmk_adapter(cfg)
    inp = fixtures.make_step_inputs(cfg, H, W)
    landmarks = torch.rand(1, T, 3, 8 * H, 8 * W, generator=torch.Generator().manual_seed(11)).half().float()
    landmarks = landmarks.repeat(2, 1, 1, 1, 1)
    t = torch.tensor(1.6377)
    with torch.no_grad():
        dres, mid, _, occ = ad(inp["sample"], t, inp["encoder_hidden_states"], inp["added_time_ids"],
                               controlnet_cond=inp["controlnet_cond"], controlnet_flow=inp["controlnet_flow"],
                               landmarks=landmarks, conditioning_scale=0.9)
    model = FlowControlNet.from_state_dict(ad.state_dict(), ad.config.__dict__)
    cu = {k: v.to(DEV) for k, v in inp.items()}
    out = model(cu["sample"], t, cu["encoder_hidden_states"], cu["added_time_ids"], controlnet_cond=cu["controlnet_cond"],
                controlnet_flow=cu["controlnet_flow"], landmarks=landmarks.to(DEV), conditioning_scale=0.9,
                return_dict=False)
    for k, (a, b) in enumerate(zip(out[0], dres)):
        e = rel(a.float().cpu().reshape(b.shape), b)
        assert e < 8e-3, f"residual {k}: {e}"
    assert rel(out[1].float().cpu().reshape(mid.shape), mid) < 8e-3
    for m_e, m_o in zip(out[3], occ):
        assert m_e.shape == m_o.shape
        assert (m_e.float().cpu() - m_o).abs().max().item() < 8e-3


def test_windowed_and_hybrid_loops_match_oracle_gpu():
    from mofa_video_b200 import lib
    from mofa_video_b200.pipeline import pipeline_hybrid as hyb
    from mofa_video_b200.pipeline import svdxt_pipeline_ctrlnet_loop as kpl
    cfg = dict(fixtures.TINY_CONFIG)
    H = W = 16
    T, F_frames, stride = cfg["num_frames"], 6, 2
    s = _loop_setup(cfg, H, W, F_frames)
    u, d, f = _gpu_nets(s)
    ref = kp.keypoint_denoise(s["unet"], s["face"], s["osch"], s["lat0"], s["il"], s["emb"], s["cond"],
                              s["flow"].repeat(2, 1, 1, 1, 1), s["ldmk"].repeat(2, 1, 1, 1, 1), 2, T, stride, scale=0.9)
    states = []
    cond_cl = to_cl(s["cond"][:1]).to(DEV)
    for (ts, te), mult in kpl.unique_views(kpl.window_views(F_frames, T, stride)):
        fl = s["flow"][0, (ts - 1):(te - 1)].half().contiguous().to(DEV)
        lm = torch.cat([s["ldmk"][0, 0:1], s["ldmk"][0, ts:te]])
        f.adapter_cond_branch_ldmk(cond_cl, fl, to_cl(lm).to(DEV), 8 * H, 8 * W)
        states.append(((ts, te), mult, (f.warped, f.ldmk)))
    lat = s["lat0"][0].half().reshape(F_frames, 4, H * W).contiguous().to(DEV)
    il = s["il"][:, 0].half().reshape(2, 4, H * W).contiguous().to(DEV)
    out = kpl.denoise_windowed(lib, u, f, states, lat, il, s["sch"]._sigmas_host, s["sch"]._timesteps_host, H, W, T,
                               1.0, 3.0, 0.9)
    e = rel(out.float().cpu().reshape(1, F_frames, 4, H, W), ref)
    assert e < 2e-2, e

    s = _loop_setup(cfg, H, W, T)
    u, d, f = _gpu_nets(s)
    drag_flow = (fixtures.make_flow(T, 8 * H, 8 * W, seed=99) * 0.5).half().float()
    mask = torch.zeros(1, 1, 8 * H, 8 * W)
    mask[..., 20:90, 30:100] = 1.0
    ref = kp.hybrid_denoise(s["unet"], s["face"], s["drag"], s["osch"], s["lat0"], s["il"], s["emb"], s["cond"],
                            s["flow"].repeat(2, 1, 1, 1, 1), drag_flow.repeat(2, 1, 1, 1, 1),
                            s["ldmk"].repeat(2, 1, 1, 1, 1), mask, 2, scale_ldmk=0.9, scale_traj=1.1)
    f.adapter_cond_branch_ldmk(cond_cl, s["flow"][0].half().contiguous().to(DEV), to_cl(s["ldmk"][0]).to(DEV), 8 * H,
                               8 * W)
    d.adapter_cond_branch(cond_cl, drag_flow[0].half().contiguous().to(DEV), 8 * H, 8 * W)
    by_rows = hyb.level_masks(mask, H, W, 4, T, DEV)
    lat = s["lat0"][0].half().reshape(T, 4, H * W).contiguous().to(DEV)
    il = s["il"][:, 0].half().reshape(2, 4, H * W).contiguous().to(DEV)
    out = hyb.denoise_hybrid(lib, u, f, d, by_rows, lat, il, s["sch"]._sigmas_host, s["sch"]._timesteps_host, H, W, 1.0,
                             3.0, 0.9, 1.1)
    e = rel(out.float().cpu().reshape(1, T, 4, H, W), ref)
    assert e < 2e-2, e


def test_keypoint_and_hybrid_entry_points():
    """FlowControlNetPipeline.__call__ of the Keypoint and Hybrid variants (reference signatures) against the oracle
    prelude (oracle/pipeline.py:prepare_inputs) + oracle loops."""
    from mofa_video_b200.models.autoencoder_kl_temporal_decoder import AutoencoderKLTemporalDecoder
    from mofa_video_b200.models.ldmk_ctrlnet import FlowControlNet as FaceNet
    from mofa_video_b200.models.traj_ctrlnet import FlowControlNet as DragNet
    from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
        UNetSpatioTemporalConditionControlNetModel as UNet
    from mofa_video_b200.pipeline import pipeline_hybrid as hyb
    from mofa_video_b200.pipeline import svdxt_pipeline_ctrlnet_loop as kpl
    from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler
    from oracle import pipeline as opipe
    from oracle.scheduler import EulerDiscreteScheduler as OSched
    from test_engine_gpu import _TinyClip
    cfg = dict(fixtures.TINY_CONFIG)
    Himg, Wimg = 128, 192
    T, F_frames, stride = cfg["num_frames"], 5, 1
    unet, drag = fixtures.make_models(cfg, seed=0, adapter_gain=20.0)
    face = make_ldmk_adapter(cfg)
    torch.manual_seed(5)
    vae = AutoencoderKLTemporalDecoder(block_out_channels=(32, 32, 64, 64)).eval()
    clip = _TinyClip(cfg["cross_attention_dim"]).eval()
    image = fixtures.make_image(Himg, Wimg)
    g = torch.Generator().manual_seed(3)
    flow = fixtures.make_flow(F_frames, Himg, Wimg)                                  # [1, F-1, 2, H, W]
    ldmk = torch.rand(1, F_frames, 3, Himg, Wimg, generator=g).half().float()
    lat0 = torch.randn(1, F_frames, 4, Himg // 8, Wimg // 8, generator=g)
    e_unet = UNet.from_state_dict(unet.state_dict(), unet.config.__dict__)
    e_face = FaceNet.from_state_dict(face.state_dict(), face.config.__dict__)
    e_drag = DragNet.from_state_dict(drag.state_dict(), drag.config.__dict__)
    vae_cu, clip_cu = vae.cuda().half(), clip.cuda().half()

    # ---- Keypoint: F = 5 frames through T = 3 windows
    emb, il = opipe.prepare_inputs(vae.float().cpu(), clip.float().cpu(), image, F_frames,
                                   torch.Generator().manual_seed(11))
    osch = OSched()
    osch.set_timesteps(2)
    cond = (2.0 * image - 1.0)[None].repeat(2, 1, 1, 1)
    ref = kp.keypoint_denoise(unet, face, osch, lat0 * osch.init_noise_sigma, il, emb, cond,
                              flow.repeat(2, 1, 1, 1, 1), ldmk.repeat(2, 1, 1, 1, 1), 2, T, stride)
    pipe = kpl.FlowControlNetPipeline(vae=vae_cu.cuda().half(), image_encoder=clip_cu.cuda().half(), unet=e_unet,
                                      controlnet=e_face, scheduler=EulerDiscreteScheduler())
    out = pipe(image, image, flow, ldmk, height=Himg, width=Wimg, num_frames=F_frames, num_inference_steps=2,
               latents=lat0.clone().half(), generator=torch.Generator().manual_seed(11), output_type="latent",
               window_size=T, stride=stride)
    assert out.frames.shape == ref.shape
    e = rel(out.frames.float().cpu(), ref)
    assert e < 3e-2, f"keypoint windowed latents rel err {e}"
    out = pipe(image, image, flow, ldmk, height=Himg, width=Wimg, num_frames=F_frames, num_inference_steps=1,
               latents=lat0.clone().half(), output_type="pt", window_size=T, stride=stride, decode_chunk_size=2)
    assert len(out.frames) == 1 and out.frames[0].shape == (F_frames, 3, Himg, Wimg)   # tensor2vid: list per clip
    assert torch.isfinite(out.frames[0]).all()

    # ---- Hybrid: T frames, two adapters + mask
    emb, il = opipe.prepare_inputs(vae.float().cpu(), clip.float().cpu(), image, T, torch.Generator().manual_seed(11))
    osch = OSched()
    osch.set_timesteps(2)
    drag_flow = (fixtures.make_flow(T, Himg, Wimg, seed=99) * 0.5).half().float()
    mask = torch.zeros(1, 1, Himg, Wimg)
    mask[..., 20:90, 30:100] = 1.0
    ref = kp.hybrid_denoise(unet, face, drag, osch, lat0[:, :T] * osch.init_noise_sigma, il, emb, cond,
                            flow[:, :T - 1].repeat(2, 1, 1, 1, 1), drag_flow.repeat(2, 1, 1, 1, 1),
                            ldmk[:, :T].repeat(2, 1, 1, 1, 1), mask, 2, scale_ldmk=0.9, scale_traj=1.1)
    pipe = hyb.FlowControlNetPipeline(vae=vae_cu.cuda().half(), image_encoder=clip_cu.cuda().half(), unet=e_unet,
                                      drag_controlnet=e_drag, face_controlnet=e_face,
                                      scheduler=EulerDiscreteScheduler())
    out = pipe(image, image, flow[:, :T - 1], ldmk[:, :T], drag_flow, mask, height=Himg, width=Wimg,
               num_inference_steps=2, latents=lat0[:, :T].clone().half(), generator=torch.Generator().manual_seed(11),
               output_type="latent", ctrl_scale_traj=1.1, ctrl_scale_ldmk=0.9)
    e = rel(out.frames.float().cpu(), ref)
    assert e < 3e-2, f"hybrid latents rel err {e}"
