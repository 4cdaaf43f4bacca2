"""BASELINE.json configs[2] (Keypoint, 512x512x25f) and configs[3] (Hybrid, 576x1024x25f) at full size on one B200,
random-init weights, a few denoise steps: checks that the full-size shapes run through every kernel (TMA boxes,
occlusion hourglasses at 2C+2 = 642..2562 channels, mask blend) and prints per-phase device times."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import factory  # noqa: E402


def run(variant, H, W, steps, frames):
    torch.cuda.empty_cache()
    pipe = factory.build_synthetic_pipeline({"num_frames": 25}, variant=variant, tiny_encoders=True)
    g = torch.Generator().manual_seed(0)
    image = torch.rand(3, H, W, generator=g)
    flow = torch.randn(1, frames - 1, 2, H, W, generator=g) * 8
    ldmk = torch.rand(1, frames, 3, H, W, generator=g)
    kw = dict(height=H, width=W, num_frames=frames, num_inference_steps=steps, output_type="latent",
              generator=torch.Generator().manual_seed(1))
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if variant == "keypoint":
            out = pipe(image, image, flow, ldmk, **kw)
        else:
            mask = torch.zeros(1, 1, H, W)
            mask[..., H // 4: H // 2, W // 3: 2 * W // 3] = 1.0
            out = pipe(image, image, flow, ldmk, flow * 0.5, mask, **kw)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
    lat = out.frames
    print(f"{variant}: {H}x{W}x{frames}f, {steps} steps: {dt * 1e3:.0f} ms (second call, incl. encode + conditioning), "
          f"latents {tuple(lat.shape)} finite={bool(torch.isfinite(lat.float()).all())} "
          f"peak_mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB")
    del pipe


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    run("keypoint", 512, 512, a.steps, 25)
    run("keypoint", 512, 512, a.steps, 37)   # 37 frames: three distinct windows of 25
    run("hybrid", 576, 1024, a.steps, 25)
