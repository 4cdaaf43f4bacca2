"""One GEMM case in isolation (for ncu --set full captures and quick A/B timing)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import lib  # noqa: E402

case = sys.argv[1] if len(sys.argv) > 1 else "qkv320"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = "cuda"
M = 460800


def h(*s, scale=0.1):
    return (torch.randn(*s, device=dev) * scale).half()


if case == "qkv320":
    a, w = h(M, 320), h(960, 320)
    out = torch.empty(M, 960, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out)
    flops = 2.0 * M * 960 * 320
elif case == "geglu320":
    a, w, b = h(M, 320), h(2560, 320), h(2560)
    out = torch.empty(M, 1280, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, act=2, bn=256)
    flops = 2.0 * M * 2560 * 320
elif case == "proj320res":
    a, w, b, r = h(M, 320), h(320, 320), h(320), h(M, 320)
    out = torch.empty(M, 320, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, res1=r)
    flops = 2.0 * M * 320 * 320
elif case == "proj320rb":  # attention out-projection + collapsed cross-attention vector + residual (rb=1, res=1)
    a, w, b, r = h(M, 320), h(320, 320), h(320), h(M, 320)
    rb = h(2, 320)
    out = torch.empty(M, 320, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, res1=r, rowbias=rb, rows_per_group=M // 2)
    flops = 2.0 * M * 320 * 320
elif case == "ff2":
    a, w, b, r = h(M, 1280), h(320, 1280), h(320), h(M, 320)
    out = torch.empty(M, 320, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, res1=r)
    flops = 2.0 * M * 320 * 1280
for _ in range(2):
    fn()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    fn()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print(f"{case}: {ms:.4f} ms  {flops / ms / 1e9:.1f} TFLOP/s")
