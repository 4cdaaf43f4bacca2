"""Spatial attention at the level-0 shape in isolation (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import lib  # noqa: E402

frames, L, heads = 10, 9216, 5
C = heads * 64
qkv = torch.randn(frames * L, 3 * C, device="cuda").half()
out = torch.empty(frames * L, C, dtype=torch.half, device="cuda")
for _ in range(3):
    lib.attn_spatial(qkv, out, frames, L, heads, 0.125)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
lib.attn_spatial(qkv, out, frames, L, heads, 0.125)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1)
print(f"attn L={L}: {ms:.3f} ms {4.0 * frames * heads * L * L * 64 / ms / 1e9:.1f} TFLOP/s")
