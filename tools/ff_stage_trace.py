"""Stage timestamps of the fused FeedForward kernel (block 0, first 64 chunks): MOFA_FF_DEBUG must include 2048.

Columns per chunk g (SM clocks relative to chunk 2's first stamp):
  m_sfree  MMA warp passed the s_free wait          m_s_iss  S MMAs + commit issued
  m_hfull  MMA warp passed h_full / w2_full         m_o_iss  OUT MMAs + commits issued
  e_sfull  epilogue warp 3 passed the s_full wait   e_ld     its tcgen05.ld completed
  e_math   GEGLU done                               e_hfull  H stored, h_full arrived
  m_top    MMA warp before the s_free wait          l_*      the same three epilogue stamps of the LAST epilogue warp
  m_kb4    MMA warp passed w1_full of k-block 4     p_kb4    W1 producer issued the load of k-block 4
"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mofa_video_b200 import lib  # noqa: E402

M = 460800
dev = "cuda"
h = lambda *s, scale=0.1: (torch.randn(*s, device=dev) * scale).half()
x, w1, b1, w2, b2, r = h(M, 320), h(2560, 320), h(2560), h(320, 1280, scale=0.05), h(320), h(M, 320)
out = torch.empty(M, 320, dtype=torch.half, device=dev)
for _ in range(2):
    lib.ff_geglu(x, w1, b1, w2, b2, out, res1=r)
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 1024)()
assert lib.load().mofa_ff_debug_dump(buf) == 0
v = [[buf[g * 16 + s] for s in range(14)] for g in range(64)]
t0 = v[2][0]
names = ["m_sfree", "m_s_iss", "m_hfull", "m_o_iss", "e_sfull", "e_ld", "e_math", "e_hfull", "m_top", "l_sfull", "l_ld",
         "l_hfull", "m_kb4", "p_kb4"]
print("chunk " + " ".join(f"{n:>8}" for n in names))
for g in range(2, 46):
    print(f"{g:5d} " + " ".join(f"{v[g][s] - t0:8d}" for s in range(14)))
