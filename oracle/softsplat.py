"""ORACLE (test infrastructure): softmax-splatting forward, device-agnostic restatement of
/root/reference/MOFA-Video-Traj/models/softsplat.py:232-274 (softsplat) and the CUDA-C string
`softsplat_out` at :285-335 (one thread per (n,c,y,x); four bounds-checked atomicAdds).

Pinned by algebraic identities only (the kernel needs CuPy+GPU; CPU branch is assert(False), :347-348):
zero flow => x/(1+1e-7); integer flow => exact shift; weights of the four corners sum to 1.
"""
import torch


def softsplat_out(tenIn, tenFlow):
    """tenIn [N,C,H,W] fp32, tenFlow [N,2,H,W] fp32 -> summation splat [N,C,H,W] (softsplat.py:285-335)."""
    N, C, H, W = tenIn.shape
    out = tenIn.new_zeros(N, C, H * W)
    gy, gx = torch.meshgrid(torch.arange(H, device=tenIn.device), torch.arange(W, device=tenIn.device), indexing="ij")
    fltX = gx[None].to(tenIn.dtype) + tenFlow[:, 0]
    fltY = gy[None].to(tenIn.dtype) + tenFlow[:, 1]
    finite = torch.isfinite(fltX) & torch.isfinite(fltY)        # :301-302
    fx = torch.where(finite, fltX, torch.zeros_like(fltX))
    fy = torch.where(finite, fltY, torch.zeros_like(fltY))
    nwx = torch.floor(fx)
    nwy = torch.floor(fy)
    src = tenIn.reshape(N, C, H * W)
    corners = (
        (nwx, nwy, (nwx + 1 - fx) * (nwy + 1 - fy)),        # northwest :315
        (nwx + 1, nwy, (fx - nwx) * (nwy + 1 - fy)),        # northeast :316
        (nwx, nwy + 1, (nwx + 1 - fx) * (fy - nwy)),        # southwest :317
        (nwx + 1, nwy + 1, (fx - nwx) * (fy - nwy)),        # southeast :318
    )
    for cx, cy, w in corners:
        ok = finite & (cx >= 0) & (cx < W) & (cy >= 0) & (cy < H)   # :320-334
        idx = (cy.clamp(0, H - 1) * W + cx.clamp(0, W - 1)).long().reshape(N, 1, H * W).expand(N, C, H * W)
        wv = torch.where(ok, w, torch.zeros_like(w)).reshape(N, 1, H * W)
        out.scatter_add_(2, idx, src * wv)
    return out.reshape(N, C, H, W)


def softsplat(tenIn, tenFlow, tenMetric, strMode):
    """softsplat.py:232-274; only the modes the adapter uses ('avg') plus 'sum'."""
    assert strMode in ("sum", "avg")
    assert tenMetric is None
    if strMode == "avg":
        tenIn = torch.cat([tenIn, tenIn.new_ones([tenIn.shape[0], 1, tenIn.shape[2], tenIn.shape[3]])], 1)
    tenOut = softsplat_out(tenIn.float(), tenFlow.float())   # custom_fwd(cast_inputs=float32), :279
    if strMode == "avg":
        tenNormalize = tenOut[:, -1:, :, :] + 0.0000001      # :256-257
        tenOut = tenOut[:, :-1, :, :] / tenNormalize         # :270
    return tenOut
