"""Run N denoise steps (adapter + UNet + fused CFG/Euler) of the config-2 workload without CLIP/VAE.
Used under ncu (launch list / --set full captures) and for quick step timing."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import lib, synthetic  # noqa: E402
from mofa_video_b200.models.svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine import FlowControlNet  # noqa: E402
from mofa_video_b200.models.unet_spatio_temporal_condition_controlnet import \
    UNetSpatioTemporalConditionControlNetModel  # noqa: E402
from mofa_video_b200.utils.scheduling_euler_discrete_karras_fix import EulerDiscreteScheduler  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    ap.add_argument("--height", type=int, default=576)
    ap.add_argument("--width", type=int, default=1024)
    ap.add_argument("--frames", type=int, default=25)
    ap.add_argument("--profile", action="store_true")
    ap.add_argument("--detail", action="store_true", help="time every op, keyed by shape (implies --profile)")
    a = ap.parse_args()
    dev = "cuda"
    H, W, T = a.height, a.width, a.frames
    h, w = H // 8, W // 8
    cfg = {"num_frames": T}
    cu, sdu = synthetic.unet_state_dict(cfg)
    unet = UNetSpatioTemporalConditionControlNetModel.from_state_dict(sdu, cu)
    del sdu
    ca, sda = synthetic.adapter_state_dict(cfg)
    ad = FlowControlNet.from_state_dict(sda, ca)
    del sda
    g = torch.Generator().manual_seed(0)
    emb = torch.randn(2, 1, 1024, generator=g).half().to(dev)
    emb[0] = 0
    ids = torch.tensor([[6.0, 128.0, 0.02]] * 2, device=dev, dtype=torch.float16)
    cond = (torch.rand(2, 3, H, W, generator=g) * 2 - 1).half().to(dev)
    cond[1] = cond[0]
    flow = (torch.randn(1, T - 1, 2, H, W, generator=g) * 10).half().to(dev).repeat(2, 1, 1, 1, 1)
    unet.net.prepare_clip(emb, ids)
    ad.net.prepare_clip(emb, ids)
    ad.prepare_condition(cond, flow)
    sch = EulerDiscreteScheduler()
    sch.set_timesteps(25)
    sig, ts = sch._sigmas_host, sch._timesteps_host
    lat = (torch.randn(T, 4, h * w, generator=g) * sig[0]).half().to(dev)
    img_lat = torch.randn(2, 4, h * w, generator=g).half().to(dev)
    img_lat[0] = 0
    nxt = torch.empty(2 * T * h * w, 8, dtype=torch.half, device=dev)
    lib.cfg_euler_step(None, lat, img_lat, nxt, T, h * w, 1.0, 3.0, 0.0, sig[0])

    def step(i):
        res, mid = ad.net.adapter_forward(nxt, ts[i], h, w, 1.0)
        n = unet.net.unet_forward(nxt, ts[i], h, w, res, mid)
        lib.cfg_euler_step(n, lat, img_lat, nxt, T, h * w, 1.0, 3.0, sig[i], sig[i + 1])

    for i in range(a.warmup):
        step(i)
    if a.detail:
        a.profile = True
        skip = {"gemm", "linear", "attn_spatial", "load", "pick_bn", "profile_start", "profile_stop", "launch_count",
                "launch_count_reset", "conv_out_size"}

        def wrap(name, fn):
            def f(*args, **kw):
                t = next((x for x in args if isinstance(x, torch.Tensor)), None)
                nbytes = sum(x.numel() * x.element_size() for x in args if isinstance(x, torch.Tensor))
                with lib._Timed(name, float(nbytes), f"shape={tuple(t.shape) if t is not None else ()}"):
                    return fn(*args, **kw)
            return f
        for name in dir(lib):
            fn = getattr(lib, name)
            if callable(fn) and not name.startswith("_") and name not in skip and getattr(fn, "__module__", "") == lib.__name__ \
                    and not isinstance(fn, type):
                setattr(lib, name, wrap(name, fn))
    torch.cuda.synchronize()
    if a.profile:
        lib.profile_start()
    lib.launch_count_reset()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.steps):
        step(a.warmup + i)
    e1.record()
    enqueue = time.perf_counter() - t0
    torch.cuda.synchronize()
    host = time.perf_counter() - t0
    print(f"steps={a.steps} device_ms_per_step={e0.elapsed_time(e1) / a.steps:.2f} host_s={host:.2f} enqueue_s={enqueue:.2f} "
          f"launches_per_step={lib.launch_count() / a.steps:.0f} finite={bool(torch.isfinite(lat.float()).all())}")
    if a.profile:
        rec = lib.profile_stop(detail=a.detail)
        items = sorted(rec.items(), key=lambda kv: -kv[1]["ms"]) if a.detail else sorted(rec.items())
        tot = sum(v["ms"] for v in rec.values()) / a.steps
        print(f"  sum of timed ops: {tot:.2f} ms/step")
        for k, v in items[:60]:
            print(f"  {k}: {v['ms'] / a.steps:.2f} ms/step, {v['work'] / max(v['ms'], 1e-9) / 1e9:.1f} T(FLOP|B)/s, "
                  f"{v['launches'] // a.steps} launches/step")


if __name__ == "__main__":
    main()
