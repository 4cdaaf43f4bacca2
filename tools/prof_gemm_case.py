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
elif case == "geglu1280":   # level 2: 225 M tiles (odd)
    M = 28800
    a, w, b = h(M, 1280), h(10240, 1280, scale=0.03), h(10240)
    out = torch.empty(M, 5120, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, act=2, bn=256)
    flops = 2.0 * M * 10240 * 1280
elif case == "lin5120":     # level 2 FeedForward output projection
    M = 28800
    a, w, b, r = h(M, 5120), h(1280, 5120, scale=0.02), h(1280), h(M, 1280)
    out = torch.empty(M, 1280, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, res1=r)
    flops = 2.0 * M * 1280 * 5120
elif case == "ff2":
    a, w, b, r = h(M, 1280), h(320, 1280), h(320), h(M, 320)
    out = torch.empty(M, 320, dtype=torch.half, device=dev)
    fn = lambda: lib.linear(a, w, out, bias=b, res1=r)
    flops = 2.0 * M * 320 * 1280
elif case in ("temporal320", "temporal320_stats", "conv320_stats", "conv320"):
    B, T, HW, C = 2, 25, 9216, 320
    rows = B * T * HW
    a, r = h(rows, C), h(rows, C)
    st = torch.zeros(B * T * 64, device=dev)
    out = torch.empty(rows, C, dtype=torch.half, device=dev)
    if case.startswith("conv320"):
        w, b = h(C, 9 * C, scale=0.03), h(C)
        kw = dict(gn_stats=st, gn_rows_per_stat=HW) if case.endswith("_stats") else {}
        fn = lambda: lib.gemm(lib.A_CONV3X3, a, w, out, N=C, n_img=B * T, H=72, W=128, C=C, bias=b, res1=r, **kw)
        flops = 2.0 * rows * C * 9 * C
    else:
        w, b = h(C, 3 * C, scale=0.05), h(C)
        kw = dict(gn_stats=st, gn_rows_per_stat=HW) if case.endswith("_stats") else {}
        fn = lambda: lib.gemm(lib.A_TEMPORAL3, a, w, out, N=C, B=B, T=T, HW=HW, C=C, bias=b, res1=r, alpha=0.5, **kw)
        flops = 2.0 * rows * C * 3 * C
elif case == "ff_fused":
    x, w1, b1, w2, b2, r = h(M, 320), h(2560, 320), h(2560), h(320, 1280, scale=0.05), h(320), h(M, 320)
    out = torch.empty(M, 320, dtype=torch.half, device=dev)
    fn = lambda: lib.ff_geglu(x, w1, b1, w2, b2, out, res1=r)
    flops = 2.0 * M * 320 * 1280 * 3
elif case == "ff_unfused":
    x, w1, b1, w2, b2, r = h(M, 320), h(2560, 320), h(2560), h(320, 1280, scale=0.05), h(320), h(M, 320)
    f = torch.empty(M, 1280, dtype=torch.half, device=dev)
    out = torch.empty(M, 320, dtype=torch.half, device=dev)

    def fn():
        lib.linear(x, w1, f, bias=b1, act=2, bn=256)
        lib.linear(f, w2, out, bias=b2, res1=r)
    flops = 2.0 * M * 320 * 1280 * 3
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
