"""CMP sparse-to-dense flow network on the sm_100a kernels.

Reference forward being executed (inference only, eval-mode BatchNorm):
  CMP_demo.run            /root/reference/MOFA-Video-Traj/models/svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py:51-62
  CMP.forward             /root/reference/MOFA-Video-Traj/models/cmp/models/modules/cmp.py:27-35
  ResNet-50 (dilated)     .../models/backbone/resnet.py:94-168
  ShallowNet              .../models/modules/shallownet.py:4-37
  MotionDecoderSkipLayer  .../models/modules/decoder.py:96-215
  Fuser.convert_flow      .../utils/visualize_utils.py:6-19

B200 design: BatchNorm folded into the conv weights at pack time; every conv is a tcgen05 GEMM (3x3 stride-1 convs,
incl. the dilation-2/4 stages, as implicit GEMM through shifted TMA boxes; the 7x7/5x5/strided ones through im2col);
ReLU / residual+ReLU in the GEMM epilogue; channel concats are written in place (ldc / channel-slice writes);
pools, align_corners resizes and the softmax-expectation head are streaming kernels.  Exact work removal: the caller
feeds 24 identical copies of frame 0 (T/run_gradio.py:293-294), so the ResNet-50 image encoder (78 % of CMP's FLOPs)
and both skip convs run once and are broadcast over the frames.
"""
import math

import torch

ACT_NONE, ACT_RELU, ACT_RELU_AFTER_RES = 0, 3, 5


def _fold(sd, conv, bn, eps=1e-5):
    w = sd[conv + ".weight"].float()
    b = sd[conv + ".bias"].float() if (conv + ".bias") in sd else torch.zeros(w.shape[0])
    if bn is not None:
        g, beta = sd[bn + ".weight"].float(), sd[bn + ".bias"].float()
        mean, var = sd[bn + ".running_mean"].float(), sd[bn + ".running_var"].float()
        s = g / torch.sqrt(var + eps)
        w = w * s.view(-1, 1, 1, 1)
        b = (b - mean) * s + beta
    return w, b


class CmpNet:
    def __init__(self, state_dict, ops, device, nbins=99, fmax=50.0):
        self.ops, self.device, self.nbins, self.fmax = ops, torch.device(device), nbins, fmax
        sd = {(k[7:] if k.startswith("module.") else k): v for k, v in state_dict.items()}
        self.sd = sd
        p = {}
        ie, fe, fd = "image_encoder.", "flow_encoder.features.", "flow_decoder."
        p["stem"] = self._conv(ie + "conv1", ie + "bn1", 7, 2, 3)
        layers = []
        inpl = 64
        for li, (planes, blocks, stride, dil) in enumerate([(64, 3, 1, 1), (128, 4, 2, 1), (256, 6, 1, 2),
                                                            (512, 3, 1, 4)]):
            blks = []
            for bi in range(blocks):
                pre = f"{ie}layer{li + 1}.{bi}."
                s = stride if bi == 0 else 1
                blk = {"c1": self._conv(pre + "conv1", pre + "bn1", 1, 1, 0),
                       "c2": self._conv(pre + "conv2", pre + "bn2", 3, s, dil, dil),
                       "c3": self._conv(pre + "conv3", pre + "bn3", 1, 1, 0),
                       "down": self._conv(pre + "downsample.0", pre + "downsample.1", 1, s, 0)
                       if (pre + "downsample.0.weight") in sd else None}
                blks.append(blk)
            layers.append(blks)
            inpl = planes * 4
        p["layers"] = layers
        p["conv5"] = self._conv(ie + "conv5", None, 1, 1, 0)
        p["s1"] = self._conv(fe + "0", fe + "1", 5, 2, 2)
        p["s2"] = self._conv(fe + "4", fe + "5", 3, 1, 1)
        self.enc_dim = sd[ie + "conv5.weight"].shape[0] + sd[fe + "4.weight"].shape[0]  # 272
        self.x_pad = (self.enc_dim + 63) // 64 * 64                                       # 320
        for k in (1, 2, 4, 8):
            off = 0 if k == 1 else 1
            pre = f"{fd}decoder{k}."
            p[f"dec{k}"] = [self._conv(pre + f"{off}", pre + f"{off + 1}", 3, 1, 1, cin_pad=self.x_pad),
                            self._conv(pre + f"{off + 3}", pre + f"{off + 4}", 3, 1, 1),
                            self._conv(pre + f"{off + 6}", pre + f"{off + 7}", 3, 1, 1)]
        p["fusion8"] = self._conv(fd + "fusion8.0", fd + "fusion8.1", 3, 1, 1)
        p["skip4"] = self._conv(fd + "skipconv4.0", fd + "skipconv4.1", 3, 1, 1)
        p["fusion4"] = self._conv(fd + "fusion4.0", fd + "fusion4.1", 3, 1, 1)
        p["skip2"] = self._conv(fd + "skipconv2.0", fd + "skipconv2.1", 3, 1, 1)
        p["fusion2"] = self._conv(fd + "fusion2.0", fd + "fusion2.1", 3, 1, 1, cin_pad=192)
        p["head"] = self._conv(fd + "head", None, 1, 1, 0)
        self.p = p
        del self.sd

    # ------------------------------------------------------------------ packing
    def _conv(self, conv, bn, k, stride, pad, dil=1, cin_pad=None):
        w, b = _fold(self.sd, conv, bn)
        cout, cin = w.shape[:2]
        if cin_pad is not None and cin_pad > cin:
            w = torch.cat([w, torch.zeros(cout, cin_pad - cin, k, k)], dim=1)
            cin = cin_pad
        kk = k * k * cin
        tma = (k == 3 and stride == 1 and pad == dil and cin % 64 == 0)
        kpad = kk if (tma or k == 1) else (kk + 7) // 8 * 8
        wp = torch.zeros(cout, kpad)
        wp[:, :kk] = w.permute(0, 2, 3, 1).reshape(cout, kk)
        return {"w": wp.to(self.device, torch.float16).contiguous(), "b": b.to(self.device, torch.float16).contiguous(),
                "k": k, "s": stride, "p": pad, "d": dil, "cin": cin, "cout": cout, "kpad": kpad, "tma": tma}

    def new(self, *shape, dtype=torch.float16):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    # ------------------------------------------------------------------ ops
    def conv(self, c, x, n, H, W, act=ACT_NONE, res=None, out=None, ldc=None):
        """x [n*H*W, cin] -> ([n*Ho*Wo, cout] (or a slice of `out` rows of width ldc), Ho, Wo)."""
        ops = self.ops
        Ho = ops.conv_out_size(H, c["k"], c["s"], c["p"], c["d"])
        Wo = ops.conv_out_size(W, c["k"], c["s"], c["p"], c["d"])
        rows = n * Ho * Wo
        if out is None:
            out = self.new(rows, c["cout"])
        kw = dict(bias=c["b"], act=act, res1=res, ldc=ldc)
        if c["tma"]:
            ops.gemm(ops.A_CONV3X3, x, c["w"], out, N=c["cout"], n_img=n, H=H, W=W, C=c["cin"], dilation=c["d"], **kw)
        elif c["k"] == 1 and c["s"] == 1:
            ops.gemm(ops.A_LINEAR, x, c["w"], out, N=c["cout"], M=rows, K=c["cin"], lda=c["cin"], **kw)
        else:
            cols = self.new(rows, c["kpad"])
            ops.im2col(x, cols, n, H, W, c["cin"], c["k"], c["s"], c["p"], c["d"], c["kpad"])
            ops.gemm(ops.A_LINEAR, cols, c["w"], out, N=c["cout"], M=rows, K=c["kpad"], lda=c["kpad"], **kw)
        return out, Ho, Wo

    def pool(self, x, n, H, W, C, k, stride, pad, mode):
        Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
        out = self.new(n * Ho * Wo, C)
        self.ops.pool2d(x, out, n, H, W, C, k, stride, pad, mode)
        return out, Ho, Wo

    def bottleneck(self, blk, x, n, H, W):
        h, _, _ = self.conv(blk["c1"], x, n, H, W, ACT_RELU)
        h, Ho, Wo = self.conv(blk["c2"], h, n, H, W, ACT_RELU)
        res = x
        if blk["down"] is not None:
            res, _, _ = self.conv(blk["down"], x, n, H, W)
        out, _, _ = self.conv(blk["c3"], h, n, Ho, Wo, ACT_RELU_AFTER_RES, res=res)
        return out, Ho, Wo

    def image_encoder(self, img_cl, n, H, W):
        """img_cl [n*H*W, 3] in [-1,1] -> (enc [n*(H/8)*(W/8), 256], conv1 feat @1/2, layer1 feat @1/4)."""
        p = self.p
        conv1, H2, W2 = self.conv(p["stem"], img_cl, n, H, W, ACT_RELU)
        x, H4, W4 = self.pool(conv1, n, H2, W2, 64, 3, 2, 1, 0)
        feats = {}
        h, w = H4, W4
        for li, blks in enumerate(p["layers"]):
            for blk in blks:
                x, h, w = self.bottleneck(blk, x, n, h, w)
            if li == 0:
                feats["layer1"] = (x, h, w)
        return x, h, w, (conv1, H2, W2), feats["layer1"]

    def forward(self, image, sparse, mask):
        """image [B,3,H,W] in [0,1]; sparse, mask [B,2,H,W] -> flow [B,2,H,W] fp16 (CMP_demo.run semantics)."""
        ops, p = self.ops, self.p
        B, _, H, W = image.shape
        dev = self.device
        same = B == 1 or bool((image[1:] == image[:1]).all())
        n_img = 1 if same else B
        img = (image[:n_img].to(dev, torch.float32) * 2 - 1).to(torch.float16).contiguous()      # FCN.py:53
        img_cl = self.new(n_img * H * W, 3)
        ops.nchw_to_nhwc(img, img_cl, n_img, 3, H * W)
        sp = torch.cat([sparse, mask], dim=1).to(dev, torch.float16).contiguous()                # FCN.py:54
        sp_cl = self.new(B * H * W, 4)
        ops.nchw_to_nhwc(sp, sp_cl, B, 4, H * W)

        # image encoder (once when all frames are identical), conv5 written as [.., 256]
        x, h8, w8, (conv1, H2, W2), (layer1, H4, W4) = self.image_encoder(img_cl, n_img, H, W)
        enc, _, _ = self.conv(p["conv5"], x, n_img, h8, w8)
        # sparse encoder on every frame
        s, hs, ws = self.conv(p["s1"], sp_cl, B, H, W, ACT_RELU)
        s, hs, ws = self.pool(s, B, hs, ws, 16, 2, 2, 0, 0)
        s, hs, ws = self.conv(p["s2"], s, B, hs, ws, ACT_RELU)
        s, hs, ws = self.pool(s, B, hs, ws, 16, 2, 2, 0, 1)
        assert (hs, ws) == (h8, w8)
        hw8 = h8 * w8
        X = torch.zeros(B * hw8, self.x_pad, dtype=torch.float16, device=dev)                   # cat(img_enc, sparse_enc)
        ops.copy_cols(enc, X, B * hw8, enc.shape[1], n_img * hw8, self.x_pad, 0)
        ops.copy_cols(s, X, B * hw8, 16, B * hw8, self.x_pad, enc.shape[1])

        # decoder: four pooled branches -> 512-channel concat at 1/8
        cat8 = self.new(B * hw8, 512)
        for bi, k in enumerate((1, 2, 4, 8)):
            if k == 1:
                t, th, tw = X, h8, w8
            else:
                t, th, tw = self.pool(X, B, h8, w8, self.x_pad, k, k, 0, 0)
            d = p[f"dec{k}"]
            t, _, _ = self.conv(d[0], t, B, th, tw, ACT_RELU)
            t, _, _ = self.conv(d[1], t, B, th, tw, ACT_RELU)
            if k == 1:
                self.conv(d[2], t, B, th, tw, ACT_RELU, out=cat8, ldc=512)
            else:
                t, _, _ = self.conv(d[2], t, B, th, tw, ACT_RELU)
                ops.resize_bilinear_ac(t, cat8, B, th, tw, 128, h8, w8, 512, 128 * bi)
        f8, _, _ = self.conv(p["fusion8"], cat8, B, h8, w8, ACT_RELU)
        # 1/4: up(f8) | skipconv4(layer1)
        cat4 = self.new(B * H4 * W4, 384)
        ops.resize_bilinear_ac(f8, cat4, B, h8, w8, 256, H4, W4, 384, 0)
        sk4, _, _ = self.conv(p["skip4"], layer1, n_img, H4, W4, ACT_RELU)
        ops.copy_cols(sk4, cat4, B * H4 * W4, 128, n_img * H4 * W4, 384, 256)
        f4, _, _ = self.conv(p["fusion4"], cat4, B, H4, W4, ACT_RELU)
        # 1/2: up(f4) | skipconv2(conv1) | zero pad to 192 channels
        cat2 = torch.zeros(B * H2 * W2, 192, dtype=torch.float16, device=dev)
        ops.resize_bilinear_ac(f4, cat2, B, H4, W4, 128, H2, W2, 192, 0)
        sk2, _, _ = self.conv(p["skip2"], conv1, n_img, H2, W2, ACT_RELU)
        ops.copy_cols(sk2, cat2, B * H2 * W2, 32, n_img * H2 * W2, 192, 128)
        f2, _, _ = self.conv(p["fusion2"], cat2, B, H2, W2, ACT_RELU)
        logits, _, _ = self.conv(p["head"], f2, B, H2, W2)
        flow = self.new(B * H2 * W2, 2)
        ops.cmp_fuser(logits, flow, self.nbins, self.fmax)
        if (H2, W2) != (H, W):                                                                   # FCN.py:56-60
            up = self.new(B * H * W, 2)
            ops.resize_bilinear_ac(flow, up, B, H2, W2, 2, H, W, 2, 0)
            flow = up
        out = self.new(B, 2, H, W)
        ops.nhwc_to_nchw(flow, out, B, 2, H * W)
        return out
