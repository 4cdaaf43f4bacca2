"""Device-timed micro-benchmarks of the hot kernels at config-2 shapes (run under gpurun).
Prints one JSON line per case: achieved TFLOP/s or GB/s, CUDA-event timed, L2 flushed between reps."""
import json
import sys
import os

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import lib  # noqa: E402

DEV = "cuda"
flush = None


def timeit(fn, reps=5, warm=2):
    global flush
    if flush is None:
        flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=DEV)
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(reps):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def h(*shape, scale=0.1):
    return (torch.randn(*shape, device=DEV) * scale).half()


def report(name, ms, flops=None, bytes_=None, **kw):
    d = {"case": name, "ms": round(ms, 4)}
    if flops:
        d["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_:
        d["gbs"] = round(bytes_ / ms / 1e6, 1)
    d.update(kw)
    print(json.dumps(d), flush=True)


def main():
    frames = 50
    # linear GEMMs (M = frames*hw)
    for (hw, C) in [(9216, 320), (2304, 640), (576, 1280)]:
        M = frames * hw
        x = h(M, C)
        for (N, K, act, nm) in [(3 * C, C, 0, "qkv"), (C, C, 0, "proj"), (8 * C, C, 2, "ff1_geglu"), (C, 4 * C, 0, "ff2")]:
            a = x if K == C else h(M, K)
            w = h(N, K)
            out = torch.empty(M, N // 2 if act == 2 else N, dtype=torch.half, device=DEV)
            ms = timeit(lambda: lib.linear(a, w, out, act=act))
            report(f"linear_{nm}_C{C}", ms, flops=2.0 * M * N * K, M=M, N=N, K=K, bn=lib.pick_bn(N, act == 2))
    # conv3x3 implicit GEMM
    for (H, W, C, N) in [(72, 128, 320, 320), (36, 64, 640, 640), (18, 32, 1280, 1280), (9, 16, 1280, 1280),
                         (72, 128, 960, 320)]:
        x = h(frames * H * W, C)
        w = h(N, 9 * C)
        out = torch.empty(frames * H * W, N, dtype=torch.half, device=DEV)
        ms = timeit(lambda: lib.gemm(lib.A_CONV3X3, x, w, out, N=N, n_img=frames, H=H, W=W, C=C))
        report(f"conv3x3_{H}x{W}_C{C}_N{N}", ms, flops=2.0 * frames * H * W * N * 9 * C)
    # temporal conv
    for (hw, C) in [(9216, 320), (2304, 640), (576, 1280)]:
        x = h(frames * hw, C)
        w = h(C, 3 * C)
        out = torch.empty(frames * hw, C, dtype=torch.half, device=DEV)
        ms = timeit(lambda: lib.gemm(lib.A_TEMPORAL3, x, w, out, N=C, B=2, T=25, HW=hw, C=C))
        report(f"temporal3_hw{hw}_C{C}", ms, flops=2.0 * frames * hw * C * 3 * C)
    # spatial attention
    for (L, heads) in [(9216, 5), (2304, 10), (576, 20), (144, 20)]:
        C = heads * 64
        qkv = h(frames * L, 3 * C, scale=1.0)
        out = torch.empty(frames * L, C, dtype=torch.half, device=DEV)
        ms = timeit(lambda: lib.attn_spatial(qkv, out, frames, L, heads, 0.125), reps=3, warm=1)
        report(f"attn_spatial_L{L}_h{heads}", ms, flops=4.0 * frames * heads * L * L * 64)
    # temporal attention
    for (hw, heads) in [(9216, 5), (2304, 10)]:
        C = heads * 64
        qkv = h(frames * hw, 3 * C, scale=1.0)
        out = torch.empty(frames * hw, C, dtype=torch.half, device=DEV)
        ms = timeit(lambda: lib.attn_temporal(qkv, out, 2, 25, hw, heads, 0.125))
        report(f"attn_temporal_hw{hw}_h{heads}", ms, bytes_=frames * hw * C * 2 * 4)
    # groupnorm / layernorm
    for (hw, C) in [(9216, 320), (2304, 640)]:
        x = h(frames * hw, C, scale=1.0)
        g, b = h(C), h(C)
        out = torch.empty_like(x)
        stats = torch.zeros(frames * 64, dtype=torch.float32, device=DEV)
        ms = timeit(lambda: lib.groupnorm(x, g, b, out, hw, 1e-5, True, stats))
        report(f"groupnorm_silu_hw{hw}_C{C}", ms, bytes_=x.numel() * 2 * 3)
        ms = timeit(lambda: lib.layernorm(x, g, b, out))
        report(f"layernorm_hw{hw}_C{C}", ms, bytes_=x.numel() * 2 * 2)
    # softsplat at the four pyramid levels (24 flows)
    flow = (torch.randn(24, 2, 576, 1024, device=DEV) * 20).half()
    for (s, C) in [(8, 320), (16, 320), (32, 640), (64, 1280)]:
        hs, ws = 576 // s, 1024 // s
        feat = h(hs * ws, C, scale=1.0)
        acc = torch.zeros(24 * hs * ws * C, dtype=torch.float32, device=DEV)
        wsum = torch.zeros(24 * hs * ws, dtype=torch.float32, device=DEV)
        out = torch.empty(24 * hs * ws, C, dtype=torch.half, device=DEV)
        ms = timeit(lambda: lib.softsplat_avg(feat, flow, acc, wsum, out, 24, hs, ws, C, 576, 1024))
        # algorithmic bytes: fp16 source read + flow read + fp16 write per flow frame
        alg = 24 * (hs * ws * C * 2 + hs * ws * 2 * 2 + hs * ws * C * 2)
        report(f"softsplat_s{s}_C{C}", ms, bytes_=alg)


if __name__ == "__main__":
    main()
