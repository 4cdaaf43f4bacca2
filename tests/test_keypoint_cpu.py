"""-m "not gpu": host logic of the Keypoint adapter (landmark embedding, occlusion hourglass batching, in-place decoder
concats, Q14 add placement) and of the windowed / hybrid loops, with kernels replaced by tests/ref_ops.py, against the
fp32 oracle (oracle/keypoint.py)."""
import torch

import ref_ops
from mofa_video_b200 import engine
from mofa_video_b200.keypoint_engine import LdmkAdapterNet
from oracle import fixtures
from oracle import keypoint as kp


make_ldmk_adapter = fixtures.make_ldmk_adapter


def to_cl(x):
    n, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(n * h * w, c).half().contiguous()


def from_cl(x, n, h, w):
    return x.float().reshape(n, h, w, -1).permute(0, 3, 1, 2)


def rel(a, b):
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def test_keypoint_adapter_matches_oracle():
    cfg = dict(fixtures.TINY_CONFIG)
    H, W = 16, 16
    T = cfg["num_frames"]
    ad = make_ldmk_adapter(cfg)
    inp = fixtures.make_step_inputs(cfg, H, W)
    g = torch.Generator().manual_seed(11)
    landmarks = torch.rand(1, T, 3, 8 * H, 8 * W, generator=g).half().float().repeat(2, 1, 1, 1, 1)
    t = torch.tensor(1.6377)
    with torch.no_grad():
        dres, mid, _, occ = ad(inp["sample"], t, inp["encoder_hidden_states"], inp["added_time_ids"],
                               controlnet_cond=inp["controlnet_cond"], controlnet_flow=inp["controlnet_flow"],
                               landmarks=landmarks, conditioning_scale=0.9)
    net = LdmkAdapterNet(ad.state_dict(), ad.config.__dict__, ref_ops, "cpu")
    net.prepare_clip(inp["encoder_hidden_states"], inp["added_time_ids"])
    masks = net.adapter_cond_branch_ldmk(to_cl(inp["controlnet_cond"][:1]), inp["controlnet_flow"][0].half().contiguous(),
                                         to_cl(landmarks[0]), 8 * H, 8 * W)
    res, midr = net.adapter_forward(to_cl(inp["sample"].flatten(0, 1)), float(t), H, W, conditioning_scale=0.9)
    hs = [(16, 16)] * 3 + [(8, 8)] * 3 + [(4, 4)] * 3 + [(2, 2)] * 3
    for k, (a, b) in enumerate(zip(res, dres)):
        e = rel(from_cl(a, 2 * T, *hs[k]), b)
        assert e < 5e-3, f"residual {k}: {e}"
    assert rel(from_cl(midr, 2 * T, 2, 2), mid) < 5e-3
    # occlusion masks [B, T-1, 1, hs, ws] of the oracle vs [T-1, hs*ws] of the engine (one CFG half)
    for m_e, m_o in zip(masks, occ):
        assert (m_e.float() - m_o[0, :, 0].reshape(T - 1, -1)).abs().max().item() < 5e-3
    # the landmark embedding and the occlusion branch must actually contribute in this fixture
    with torch.no_grad():
        d2, _, _, _ = ad(inp["sample"], t, inp["encoder_hidden_states"], inp["added_time_ids"],
                         controlnet_cond=inp["controlnet_cond"], controlnet_flow=inp["controlnet_flow"],
                         landmarks=torch.zeros_like(landmarks), conditioning_scale=0.9)
    assert (d2[0] - dres[0]).abs().max() > 1e-3


# ------------------------------------------------------------------------------------------------ loops
def _loop_setup(cfg, H, W, F_frames, seed=21):
    """Oracle models + engine nets (ref_ops, CPU) + prepared tensors for the windowed / hybrid loops."""
    from oracle.scheduler import EulerDiscreteScheduler as OracleScheduler
    from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler
    unet, drag = fixtures.make_models(cfg, seed=0, adapter_gain=20.0)
    face = make_ldmk_adapter(cfg)
    g = torch.Generator().manual_seed(seed)
    ctx = cfg["cross_attention_dim"]
    emb = torch.randn(2, 1, ctx, generator=g).half().float()
    emb[0] = 0
    ids = torch.tensor([[6.0, 128.0, 0.02]] * 2)
    cond = (fixtures.make_image(8 * H, 8 * W) * 2 - 1).half().float()[None].repeat(2, 1, 1, 1)
    flow = fixtures.make_flow(F_frames, 8 * H, 8 * W).half().float()                 # [1, F-1, 2, H, W]
    ldmk = torch.rand(1, F_frames, 3, 8 * H, 8 * W, generator=g).half().float()
    il1 = torch.randn(1, 1, 4, H, W, generator=g).half().float()
    il = torch.cat([torch.zeros_like(il1), il1]).repeat(1, F_frames, 1, 1, 1)        # CFG: uncond half is zero
    osch = OracleScheduler()
    sch = EulerDiscreteScheduler()
    sch.set_timesteps(2)
    osch.set_timesteps(2)
    lat0 = (torch.randn(1, F_frames, 4, H, W, generator=g) * float(osch.init_noise_sigma)).half().float()
    u_net = engine.Net("unet", unet.state_dict(), unet.config.__dict__, ref_ops, "cpu")
    d_net = engine.Net("adapter", drag.state_dict(), drag.config.__dict__, ref_ops, "cpu")
    f_net = LdmkAdapterNet(face.state_dict(), face.config.__dict__, ref_ops, "cpu")
    for n in (u_net, d_net, f_net):
        n.prepare_clip(emb, ids)
    return dict(unet=unet, drag=drag, face=face, emb=emb, ids=ids, cond=cond, flow=flow, ldmk=ldmk, il=il, lat0=lat0,
                osch=osch, sch=sch, u_net=u_net, d_net=d_net, f_net=f_net)


def test_windowed_loop_matches_oracle():
    from mofa_video_b200.pipeline import svdxt_pipeline_ctrlnet_loop as kpl
    cfg = dict(fixtures.TINY_CONFIG)
    H = W = 16
    T, F_frames, stride = cfg["num_frames"], 6, 2
    s = _loop_setup(cfg, H, W, F_frames)
    ref = kp.keypoint_denoise(s["unet"], s["face"], s["osch"], s["lat0"], s["il"], s["emb"], s["cond"],
                              s["flow"].repeat(2, 1, 1, 1, 1), s["ldmk"].repeat(2, 1, 1, 1, 1), 2, T, stride,
                              scale=0.9)
    views = kpl.unique_views(kpl.window_views(F_frames, T, stride))
    assert len(views) >= 3 and kpl.window_views(F_frames, T, stride) == kp.window_views(F_frames, T, stride)
    states = []
    for (ts, te), mult in views:
        fl = s["flow"][0, (ts - 1):(te - 1)].half().contiguous()
        lm = torch.cat([s["ldmk"][0, 0:1], s["ldmk"][0, ts:te]])
        s["f_net"].adapter_cond_branch_ldmk(to_cl(s["cond"][:1]), fl, to_cl(lm), 8 * H, 8 * W)
        states.append(((ts, te), mult, (s["f_net"].warped, s["f_net"].ldmk)))
    lat = s["lat0"][0].half().reshape(F_frames, 4, H * W).contiguous()
    il = s["il"][:, 0].half().reshape(2, 4, H * W).contiguous()
    out = kpl.denoise_windowed(ref_ops, s["u_net"], s["f_net"], states, lat, il, s["sch"]._sigmas_host,
                               s["sch"]._timesteps_host, H, W, T, 1.0, 3.0, 0.9)
    e = rel(out.float().reshape(1, F_frames, 4, H, W), ref)
    assert e < 1e-2, e
    # T == num_frames: two identical views collapse into one evaluation with multiplicity 2
    assert kpl.unique_views(kpl.window_views(25, 25, 12)) == [((1, 25), 2)]


def test_hybrid_loop_matches_oracle():
    from mofa_video_b200.pipeline import pipeline_hybrid as hyb
    cfg = dict(fixtures.TINY_CONFIG)
    H = W = 16
    T = cfg["num_frames"]
    s = _loop_setup(cfg, H, W, T)
    g = torch.Generator().manual_seed(5)
    drag_flow = (fixtures.make_flow(T, 8 * H, 8 * W, seed=99) * 0.5).half().float()
    mask = torch.zeros(1, 1, 8 * H, 8 * W)
    mask[..., 20:90, 30:100] = 1.0
    ref = kp.hybrid_denoise(s["unet"], s["face"], s["drag"], s["osch"], s["lat0"], s["il"], s["emb"], s["cond"],
                            s["flow"].repeat(2, 1, 1, 1, 1), drag_flow.repeat(2, 1, 1, 1, 1),
                            s["ldmk"].repeat(2, 1, 1, 1, 1), mask, 2, scale_ldmk=0.9, scale_traj=1.1)
    s["f_net"].adapter_cond_branch_ldmk(to_cl(s["cond"][:1]), s["flow"][0].half().contiguous(), to_cl(s["ldmk"][0]),
                                        8 * H, 8 * W)
    s["d_net"].adapter_cond_branch(to_cl(s["cond"][:1]), drag_flow[0].half().contiguous(), 8 * H, 8 * W)
    by_rows = hyb.level_masks(mask, H, W, 4, T, "cpu")
    lat = s["lat0"][0].half().reshape(T, 4, H * W).contiguous()
    il = s["il"][:, 0].half().reshape(2, 4, H * W).contiguous()
    out = hyb.denoise_hybrid(ref_ops, s["u_net"], s["f_net"], s["d_net"], by_rows, lat, il, s["sch"]._sigmas_host,
                             s["sch"]._timesteps_host, H, W, 1.0, 3.0, 0.9, 1.1)
    e = rel(out.float().reshape(1, T, 4, H, W), ref)
    assert e < 1e-2, e
    # the mask must select: an all-ones mask gives a different result
    ref1 = kp.hybrid_denoise(s["unet"], s["face"], s["drag"], s["osch"], s["lat0"], s["il"], s["emb"], s["cond"],
                             s["flow"].repeat(2, 1, 1, 1, 1), drag_flow.repeat(2, 1, 1, 1, 1),
                             s["ldmk"].repeat(2, 1, 1, 1, 1), torch.ones_like(mask), 2, scale_ldmk=0.9, scale_traj=1.1)
    assert (ref1 - ref).abs().max() > 1e-4
