"""GroupNorm apply-pass variants (MOFA_GN_VARIANT, read once per process) at the denoise step's shapes: GB/s of read+write."""
import os
import subprocess
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def one():
    from mofa_video_b200 import lib
    lib.load()
    for rows, C, hw in ((460800, 320, 9216), (115200, 640, 2304), (28800, 1280, 576), (7200, 1280, 144)):
        x = torch.randn(rows, C, device="cuda").half()
        g, b = torch.ones(C, device="cuda").half(), torch.zeros(C, device="cuda").half()
        out = torch.empty_like(x)
        st = torch.empty((rows // hw) * 64, device="cuda")
        lib.groupnorm(x, g, b, out, hw, 1e-5, True, st)            # fills st
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        ts = []
        for _ in range(6):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            lib.groupnorm(x, g, b, out, hw, 1e-5, True, st, stats_ready=True)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[1]
        print(f"variant {os.environ.get('MOFA_GN_VARIANT', '0')} rows={rows} C={C}: {ms * 1e3:.1f} us, "
              f"{2 * x.numel() * 2 / ms / 1e6:.0f} GB/s", flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "one":
        one()
    else:
        for v in range(4):
            subprocess.run([sys.executable, __file__, "one"], env=dict(os.environ, MOFA_GN_VARIANT=str(v)))
