"""Keypoint (facial-landmark) MOFA-Adapter on the sm_100a kernels -- SURVEY.md §8 rows a13.

Reference: /root/reference/MOFA-Video-Keypoint/models/ldmk_ctrlnet.py (FlowControlNet :187-575) and
/root/reference/MOFA-Video-Keypoint/models/occlusion/hourglass.py (ForegroundMatting :227-280).
On top of the trajectory adapter (engine.Net kind 'adapter') it adds, all loop-invariant and therefore computed once
per clip / per temporal window for one CFG half:
  * a landmark-image embedding (second conv stack) added where the trunk has block_out_channels[0] channels (Q14),
  * per (scale, flow frame) an occlusion hourglass on cat(feature, scaled flow, warped feature) -- six 3x3 conv+ReLU
    without resampling, two 7x7 heads (sigmoid matting mask, matting image), blend, 1x1 `zero_outs`.
The 24 flow frames of a scale are batched into one GEMM M dimension; the 2C+2-channel concat is zero-padded to a
multiple of 64 so every conv (3x3 and 7x7) is an implicit-GEMM tcgen05 launch; decoder concats are written in place.
"""
import torch

from mofa_video_b200 import engine

ACT_RELU, ACT_SIGMOID = 3, 4


class LdmkAdapterNet(engine.Net):
    def __init__(self, state_dict, config, ops, device):
        super().__init__("adapter", state_dict, config, ops, device)
        pk = engine._Packer(state_dict, device)
        ce = "controlnet_ldmk_embedding"
        convs = [(ce + ".conv_in", 1)]
        for k in range(3):                      # block_out_channels (16, 32, 64, 128): ldmk_ctrlnet.py:232-236
            convs += [(f"{ce}.blocks.{2 * k}", 1), (f"{ce}.blocks.{2 * k + 1}", 2)]
        convs.append((ce + ".conv_out", 1))
        self.p["ldmk_convs"] = [self._pack_im2col_conv(pk, n, s) for n, s in convs]
        self.occ = {}
        boc = self.boc
        for s, C in ((8, boc[0]), (16, boc[0]), (32, boc[1]), (64, boc[2])):
            pre = f"occlusions.{s}"
            cp = (2 * C + 2 + 63) // 64 * 64
            o = {"C": C, "cp": cp, "zero": pk.conv1(f"zero_outs.{s}")}
            enc = [self._conv_k(state_dict, f"{pre}.hourglass.encoder.down_blocks.{i}.conv", 3,
                                cin_pad=cp if i == 0 else None) for i in range(3)]
            dec = [self._conv_k(state_dict, f"{pre}.hourglass.decoder.up_blocks.{i}.conv", 3) for i in range(3)]
            o["enc"], o["dec"] = enc, dec
            o["mask"] = self._conv_k(state_dict, f"{pre}.matting_mask", 7)
            o["matting"] = self._conv_k(state_dict, f"{pre}.matting", 7)
            self.occ[s] = o
        self.ldmk = None

    def _conv_k(self, sd, name, k, cin_pad=None):
        w, b = sd[name + ".weight"].float(), sd[name + ".bias"].float()
        cout, cin = w.shape[:2]
        if cin_pad is not None and cin_pad > cin:
            w = torch.cat([w, torch.zeros(cout, cin_pad - cin, k, k)], dim=1)
            cin = cin_pad
        assert cin % 64 == 0, (name, cin)
        wk = w.permute(0, 2, 3, 1).reshape(cout, k * k * cin)
        return {"w": wk.to(self.device, torch.float16).contiguous(), "b": b.to(self.device, torch.float16).contiguous(),
                "k": k, "cin": cin, "cout": cout}

    def _conv(self, c, x, n, H, W, act, out=None, ldc=None, c_off=0):
        ops = self.ops
        if out is None:
            out = self.new(n * H * W, c["cout"])
        ops.gemm(ops.A_CONV3X3, x, c["w"], out, N=c["cout"], n_img=n, H=H, W=W, C=c["cin"], ksize=c["k"],
                 bias=c["b"], act=act, ldc=ldc, c_off=c_off)
        return out

    # ------------------------------------------------------------------ per-clip (per-window) conditioning
    def adapter_cond_branch_ldmk(self, cond_nhwc, flow_h, landmarks_cl, Himg, Wimg):
        """cond_nhwc [Himg*Wimg, 3]; flow_h fp16 [T-1, 2, Himg, Wimg]; landmarks_cl [T*Himg*Wimg, 3] (one CFG half).
        Fills self.warped (slot 0 = feature, slots 1.. = occlusion-refined warps) and self.ldmk; returns the
        occlusion masks, one fp16 tensor [T-1, hs*ws] per scale (ldmk_ctrlnet.py:291-320, 394-470)."""
        ops, T = self.ops, self.T
        Fn = T - 1
        self.adapter_cond_branch(cond_nhwc, flow_h, Himg, Wimg)      # plain features + softsplat warps
        masks = []
        for (ft, hs, ws), wp in zip(self.cond_feats, self.warped):
            s = Himg // hs
            o = self.occ[s]
            C, cp = o["C"], o["cp"]
            hw = hs * ws
            rows = Fn * hw
            warped = wp[hw:]
            X = torch.zeros(rows, cp, dtype=torch.float16, device=self.device)
            ops.copy_cols(ft, X, rows, C, hw, cp, 0)                 # reference image feature, same for every frame
            ops.flow_pyramid(flow_h, X, Fn, hs, ws, Himg, Wimg, cp, C)
            ops.copy_cols(warped, X, rows, C, rows, cp, C + 2)
            e1 = self._conv(o["enc"][0], X, Fn, hs, ws, ACT_RELU)
            e2 = self._conv(o["enc"][1], e1, Fn, hs, ws, ACT_RELU)
            e3 = self._conv(o["enc"][2], e2, Fn, hs, ws, ACT_RELU)
            c2, c1 = e2.shape[1], e1.shape[1]
            cat1 = self.new(rows, 2 * c2)                            # [e2 | up0(e3)]  (hourglass.py:91-98)
            ops.copy_cols(e2, cat1, rows, c2, rows, 2 * c2, 0)
            self._conv(o["dec"][0], e3, Fn, hs, ws, ACT_RELU, out=cat1, ldc=2 * c2, c_off=c2)
            cat2 = self.new(rows, 2 * c1)                            # [e1 | up1(cat1)]
            ops.copy_cols(e1, cat2, rows, c1, rows, 2 * c1, 0)
            self._conv(o["dec"][1], cat1, Fn, hs, ws, ACT_RELU, out=cat2, ldc=2 * c1, c_off=c1)
            hg = self._conv(o["dec"][2], cat2, Fn, hs, ws, ACT_RELU)
            m = self._conv(o["mask"], hg, Fn, hs, ws, ACT_SIGMOID)   # [rows, 1]
            mat = self._conv(o["matting"], hg, Fn, hs, ws, 0)
            blended = self.new(rows, C)
            ops.mask_blend(warped, mat, m, blended)                  # warped*m + matting*(1-m)   (:278)
            zw, zb = o["zero"]
            ops.linear(blended, zw, warped, bias=zb)                 # zero_outs, in place of the plain warp (:314)
            masks.append(m.reshape(Fn, hw))
        # landmark embedding pyramid
        x, H, W = landmarks_cl, Himg, Wimg
        convs = self.p["ldmk_convs"]
        for k, c in enumerate(convs):
            last = k == len(convs) - 1
            x, H, W = self.conv_im2col(c, x, T, H, W, act=0 if last else ops.ACT_SILU)
        C0 = x.shape[1]
        if self.persistent:   # a captured step graph reads these: same storage for every clip (engine.Net.pbuf)
            xb = self.pbuf(("ldmk", 1), *x.shape)
            xb.copy_(x)
            x = xb
        self.ldmk = {H: x}
        for s in (2, 4):
            d = self.pbuf(("ldmk", s), T * (H // s) * (W // s), C0)
            ops.downsample_nearest(x, d, T, H, W, C0, s)
            self.ldmk[H // s] = d
        return masks
