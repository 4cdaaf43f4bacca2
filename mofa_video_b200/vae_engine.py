"""Native (sm_100a kernel) execution of the VAE temporal decoder -- SURVEY.md §8 row a10 / §8f row 1.

Replaces diffusers 0.24 `TemporalDecoder.forward` as called by
/root/reference/MOFA-Video-Traj/pipeline/pipeline.py:194-220 (decode_latents: chunks of `decode_chunk_size`
frames, each decoded with num_frames = chunk length, quirk Q8).  Topology:
conv_in 4->512; mid: SpatioTemporalResBlock, single-head attention (d = 512, GroupNorm pre-norm, bias, residual),
SpatioTemporalResBlock; 4 up blocks of 3 SpatioTemporalResBlock(temb=None, eps 1e-6 / temporal 1e-5,
merge 'learned', switch_spatial_to_temporal_mix) + nearest-2x conv; GroupNorm-SiLU-conv 128->3; time_conv_out.

Everything reuses the denoise-loop kernels: 3x3 convs = implicit-GEMM tcgen05 (C = 512/256/128), temporal convs,
GroupNorm+SiLU, nearest upsample; the d=512 attention runs as  S = Q K^T (GEMM, alpha = 1/sqrt(512)) ->
mofa_softmax_rows -> O = P V^T (GEMM) with V^T produced directly by a GEMM whose "weight" operand is the token
matrix, and the value bias folded after PV (rows of P sum to 1).  The tail kernel fuses time_conv_out with the
reference's uint8 post-processing.
"""
import math
from types import SimpleNamespace

import torch

from mofa_video_b200 import engine


class VaeDecoderNet(engine.Net):
    def __init__(self, state_dict, ops, device, block_out_channels=(128, 256, 512, 512), latent_channels=4,
                 scaling_factor=0.18215):
        # deliberately not calling Net.__init__: this network has its own topology, only the block runners are shared
        self.kind, self.ops, self.device = "vae_decoder", ops, torch.device(device)
        self.pk_bn = ops.pick_bn
        self.temb = engine._TembBank()
        self.xattn = []
        self.B, self.T = 1, 1
        self.scaling_factor = scaling_factor
        pk = engine._Packer(state_dict, device)
        boc = tuple(block_out_channels)
        p = {}
        p["conv_in"] = self._pack_im2col_conv(pk, "decoder.conv_in", 1)
        C = boc[-1]
        p["mid_res"] = [self._res(pk, f"decoder.mid_block.resnets.{i}", C, C) for i in range(2)]
        a = "decoder.mid_block.attentions.0"
        p["attn"] = {"norm": pk.norm(a + ".group_norm"), "q": pk.lin(a + ".to_q"), "k": pk.lin(a + ".to_k"),
                     "v": pk.lin(a + ".to_v"), "o": pk.lin(a + ".to_out.0"), "C": C}
        rb = list(reversed(boc))
        ups = []
        cin = rb[0]
        for i, cout in enumerate(rb):
            blk = {"res": [self._res(pk, f"decoder.up_blocks.{i}.resnets.{j}", cin if j == 0 else cout, cout)
                           for j in range(3)],
                   "up": pk.conv3(f"decoder.up_blocks.{i}.upsamplers.0.conv") if i != len(rb) - 1 else None}
            ups.append(blk)
            cin = cout
        p["up"] = ups
        p["norm_out"] = pk.norm("decoder.conv_norm_out")
        p["conv_out"] = pk.conv3("decoder.conv_out")
        w = state_dict["decoder.time_conv_out.weight"]  # [3, 3, 3, 1, 1] (co, ci, kt)
        p["tconv_w"] = w[:, :, :, 0, 0].detach().to(device=device, dtype=torch.float32).contiguous()
        p["tconv_b"] = state_dict["decoder.time_conv_out.bias"].detach().to(device=device, dtype=torch.float32).contiguous()
        self.p = p

    def _res(self, pk, pre, cin, cout):
        return self._pack_res(pk, pre, cin, cout, 1e-6, temporal_eps=1e-5, switch=True)

    # ------------------------------------------------------------------ blocks
    def attention(self, x, hw):
        """Single-head attention over the hw tokens of each frame (head_dim = C = 512)."""
        ops, a = self.ops, self.p["attn"]
        rows, C = x.shape
        frames = rows // hw
        stats = self.new(frames * 64, dtype=torch.float32)
        xn = self.new(rows, C)
        ops.groupnorm(x, a["norm"][0], a["norm"][1], xn, hw, 1e-6, False, stats)
        q, k = self.new(rows, C), self.new(rows, C)
        ops.linear(xn, a["q"][0], q, bias=a["q"][1])
        ops.linear(xn, a["k"][0], k, bias=a["k"][1])
        o = self.new(rows, C)
        s = self.new(hw, hw)
        vt = self.new(C, hw)
        scale = 1.0 / math.sqrt(C)
        for f in range(frames):
            sl = slice(f * hw, (f + 1) * hw)
            ops.linear(a["v"][0], xn[sl], vt)                      # V^T [C, hw] = Wv . X^T   (bias folded below)
            ops.linear(q[sl], k[sl], s, alpha=scale)               # S = scale * Q K^T
            ops.softmax_rows(s)
            ops.linear(s, vt, o[sl], bias=a["v"][1])               # O = P V  (+ bv: rows of P sum to 1)
        out = self.new(rows, C)
        ops.linear(o, a["o"][0], out, bias=a["o"][1], res1=x)
        return out

    def decode_chunk(self, z_cl, n, h, w, out_f32=None, out_u8=None):
        """z_cl: fp16 channels-last latents [n*h*w, 4] (already divided by scaling_factor) of ONE chunk of n frames.
        Writes the decoded frames: out_f32 [n, 3, 8h, 8w] float32 and/or out_u8 [n, 8h, 8w, 3] uint8."""
        ops, p = self.ops, self.p
        self.B, self.T = 1, n
        H, W = h, w
        hw = H * W
        x, _, _ = self.conv_im2col(p["conv_in"], z_cl, n, H, W)
        x = self.resblock(p["mid_res"][0], x, None, None, hw, H, W)
        x = self.attention(x, hw)
        x = self.resblock(p["mid_res"][1], x, None, None, hw, H, W)
        for blk in p["up"]:
            for r in blk["res"]:
                x = self.resblock(r, x, None, None, hw, H, W)
            if blk["up"] is not None:
                C = x.shape[1]
                up = self.new(n * 4 * hw, C)
                ops.upsample2x(x, up, n, H, W, C)
                del x
                H, W = 2 * H, 2 * W
                hw = H * W
                wgt, b = blk["up"]
                x = self.new(n * hw, wgt.shape[0])
                ops.gemm(ops.A_CONV3X3, up, wgt, x, N=wgt.shape[0], n_img=n, H=H, W=W, C=C, bias=b)
                del up
        stats = self.new(n * 64, dtype=torch.float32)
        hn = self.new(x.shape[0], x.shape[1])
        ops.groupnorm(x, p["norm_out"][0], p["norm_out"][1], hn, hw, 1e-6, True, stats)
        wgt, b = p["conv_out"]
        y = self.new(n * hw, 3)
        ops.gemm(ops.A_CONV3X3, hn, wgt, y, N=3, n_img=n, H=H, W=W, C=x.shape[1], bias=b, bn=16)
        ops.vae_time_conv_out(y, p["tconv_w"], p["tconv_b"], out_f32, out_u8, n, hw)


class VaeEncoderNet(VaeDecoderNet):
    """SD-style VAE encoder (diffusers 0.24 `Encoder` + `quant_conv`, called by
    /root/reference/MOFA-Video-Traj/pipeline/pipeline.py:143-164 on ONE 576x1024 image per clip):
    conv_in 3->128, four DownEncoderBlock2D (two plain ResnetBlock2D each, eps 1e-6; stride-2 conv after
    F.pad(0,1,0,1) on the first three), mid (resnet, single-head attention d = 512, resnet), GroupNorm-SiLU-conv 512->8,
    1x1 quant_conv; the latent is the mean half.  The reference upcasts this module to fp32 (:343-352); here it runs on
    the fp16 tensor-core kernels with fp32 accumulation (difference bounded by tests/test_vae_cpu.py /
    test_engine_gpu.py)."""

    def __init__(self, state_dict, ops, device, block_out_channels=(128, 256, 512, 512), latent_channels=4):
        self.kind, self.ops, self.device = "vae_encoder", ops, torch.device(device)
        self.pk_bn = ops.pick_bn
        self.temb = engine._TembBank()
        self.xattn = []
        self.B, self.T = 1, 1
        pk = engine._Packer(state_dict, device)
        boc = tuple(block_out_channels)
        p = {"conv_in": self._pack_im2col_conv(pk, "encoder.conv_in", 1)}
        blocks, cin = [], boc[0]
        for i, cout in enumerate(boc):
            pre = f"encoder.down_blocks.{i}"
            blk = {"res": [self._res2d(pk, f"{pre}.resnets.{j}", cin if j == 0 else cout, cout) for j in range(2)],
                   "down": None}
            if i != len(boc) - 1:
                w, b = pk.conv3(f"{pre}.downsamplers.0.conv")
                blk["down"] = {"w": w, "b": b, "C": cout}
            blocks.append(blk)
            cin = cout
        p["blocks"] = blocks
        C = boc[-1]
        p["mid_res"] = [self._res2d(pk, f"encoder.mid_block.resnets.{i}", C, C) for i in range(2)]
        a = "encoder.mid_block.attentions.0"
        p["attn"] = {"norm": pk.norm(a + ".group_norm"), "q": pk.lin(a + ".to_q"), "k": pk.lin(a + ".to_k"),
                     "v": pk.lin(a + ".to_v"), "o": pk.lin(a + ".to_out.0"), "C": C}
        p["norm_out"] = pk.norm("encoder.conv_norm_out")
        p["conv_out"] = pk.conv3("encoder.conv_out")
        # quant_conv (1x1, 8 -> 8): only the mean rows are needed; padded to 8 output columns for 16-byte rows
        wq = state_dict["quant_conv.weight"].float().reshape(2 * latent_channels, 2 * latent_channels)
        bq = state_dict["quant_conv.bias"].float()
        wq8, bq8 = torch.zeros(8, 2 * latent_channels), torch.zeros(8)
        wq8[:latent_channels], bq8[:latent_channels] = wq[:latent_channels], bq[:latent_channels]
        p["quant"] = (wq8.to(device, torch.float16).contiguous(), bq8.to(device, torch.float16).contiguous())
        self.latent_channels = latent_channels
        self.p = p

    def _res2d(self, pk, pre, cin, cout):
        r = {"n1": pk.norm(pre + ".norm1"), "c1": pk.conv3(pre + ".conv1"), "n2": pk.norm(pre + ".norm2"),
             "c2": pk.conv3(pre + ".conv2"), "cin": cin, "cout": cout}
        r["sc"] = pk.conv1(pre + ".conv_shortcut") if (pre + ".conv_shortcut.weight") in pk.sd else None
        return r

    def res2d(self, r, x, n, H, W):
        """ResnetBlock2D without time embedding: x + conv2(silu(gn(conv1(silu(gn(x)))))) (shortcut conv if cin != cout)."""
        ops = self.ops
        hw = H * W
        stats = self.new(n * 64, dtype=torch.float32)
        h = self.new(n * hw, r["cin"])
        ops.groupnorm(x, r["n1"][0], r["n1"][1], h, hw, 1e-6, True, stats)
        h1 = self.new(n * hw, r["cout"])
        ops.gemm(ops.A_CONV3X3, h, r["c1"][0], h1, N=r["cout"], n_img=n, H=H, W=W, C=r["cin"], bias=r["c1"][1])
        h2 = self.new(n * hw, r["cout"])
        ops.groupnorm(h1, r["n2"][0], r["n2"][1], h2, hw, 1e-6, True, stats)
        if r["sc"] is not None:
            sc = self.new(n * hw, r["cout"])
            ops.linear(x, r["sc"][0], sc, bias=r["sc"][1])
        else:
            sc = x
        out = self.new(n * hw, r["cout"])
        ops.gemm(ops.A_CONV3X3, h2, r["c2"][0], out, N=r["cout"], n_img=n, H=H, W=W, C=r["cout"], bias=r["c2"][1],
                 res1=sc)
        return out

    def encode(self, x_cl, n, H, W):
        """x_cl: fp16 channels-last image(s) [n*H*W, 3] in [-1, 1].  Returns the latent mean, fp16 [n*(H/8)*(W/8), 8]
        (columns >= latent_channels are zero padding)."""
        ops, p = self.ops, self.p
        x, _, _ = self.conv_im2col(p["conv_in"], x_cl, n, H, W)
        for blk in p["blocks"]:
            for r in blk["res"]:
                x = self.res2d(r, x, n, H, W)
            if blk["down"] is not None:
                d = blk["down"]
                C = d["C"]
                Ho, Wo = ops.conv_out_size(H, 3, 2, -1), ops.conv_out_size(W, 3, 2, -1)
                cols = self.new(n * Ho * Wo, 9 * C)
                ops.im2col(x, cols, n, H, W, C, 3, 2, -1, 1, 9 * C)   # F.pad(x, (0,1,0,1)) + 3x3 stride 2
                del x
                x = self.new(n * Ho * Wo, C)
                ops.linear(cols, d["w"], x, bias=d["b"])
                del cols
                H, W = Ho, Wo
        hw = H * W
        x = self.res2d(p["mid_res"][0], x, n, H, W)
        x = self.attention(x, hw)
        x = self.res2d(p["mid_res"][1], x, n, H, W)
        stats = self.new(n * 64, dtype=torch.float32)
        hn = self.new(x.shape[0], x.shape[1])
        ops.groupnorm(x, p["norm_out"][0], p["norm_out"][1], hn, hw, 1e-6, True, stats)
        wgt, b = p["conv_out"]
        mom = self.new(n * hw, wgt.shape[0])
        ops.gemm(ops.A_CONV3X3, hn, wgt, mom, N=wgt.shape[0], n_img=n, H=H, W=W, C=x.shape[1], bias=b, bn=16)
        lat = self.new(n * hw, 8)
        ops.linear(mom, p["quant"][0], lat, bias=p["quant"][1], bn=16)
        return lat, H, W


class NativeTemporalDecoderVAE:
    """`vae` object for FlowControlNetPipeline: encode() and decode() both run on the sm_100a kernels (the wrapped
    PyTorch module only supplies the weights and the config).  Same interface as AutoencoderKLTemporalDecoder."""

    def __init__(self, torch_vae, ops=None, device=None):
        from mofa_video_b200.models._base import resolve_backend
        ops, device, _ = resolve_backend(ops, device)
        self.torch_vae = torch_vae
        cfg = torch_vae.config          # our stand-in keeps a namespace, diffusers a FrozenDict
        self.config = SimpleNamespace(**(dict(cfg) if isinstance(cfg, dict) else vars(cfg)))
        self.config.force_upcast = False  # nothing to upcast: encode() runs on the fp16 tensor-core kernels too
        self._ops, self._device = ops, device
        sd = torch_vae.state_dict()
        self.net = VaeDecoderNet(sd, self._ops, self._device, block_out_channels=self.config.block_out_channels,
                                 latent_channels=self.config.latent_channels,
                                 scaling_factor=self.config.scaling_factor)
        self.enc = VaeEncoderNet(sd, self._ops, self._device, block_out_channels=self.config.block_out_channels,
                                 latent_channels=self.config.latent_channels)

    @property
    def dtype(self):
        return torch.float16

    def to(self, *a, **k):
        return self  # weights are packed once for the kernels; the wrapped module is not used at run time

    def encode(self, x):
        """x [n, 3, H, W] in [-1, 1] -> .latent_dist.mode() [n, 4, H/8, W/8] in x's dtype (pipeline.py:143-164)."""
        n, c, H, W = x.shape
        if H % 8 or W % 8 or c != 3:
            raise ValueError("encode expects [n, 3, H, W] with H, W multiples of 8")
        xc = torch.empty(n * H * W, 3, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(x.to(device=self._device, dtype=torch.float16).contiguous(), xc, n, 3, H * W)
        lat, h, w = self.enc.encode(xc, n, H, W)
        L = self.config.latent_channels
        full = torch.empty(n, 8, h, w, dtype=torch.float16, device=self._device)
        self._ops.nhwc_to_nchw(lat, full, n, 8, h * w)
        mean = full[:, :L].to(x.dtype)
        return SimpleNamespace(latent_dist=SimpleNamespace(mode=lambda: mean, mean=mean))

    def _cl(self, z):
        n, c, h, w = z.shape
        zc = torch.empty(n * h * w, c, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(z.to(device=self._device, dtype=torch.float16).contiguous(), zc, n, c, h * w)
        return zc

    def decode(self, z, num_frames=1):
        """z [n, 4, h, w] (latents / scaling_factor, as pipeline.py:198 passes them) -> .sample [n, 3, 8h, 8w]."""
        n, c, h, w = z.shape
        if n != num_frames:
            raise ValueError("the native decoder decodes one chunk of num_frames frames per call (batch of 1 clip)")
        out = torch.empty(n, 3, 8 * h, 8 * w, dtype=torch.float32, device=self._device)
        self.net.decode_chunk(self._cl(z), n, h, w, out_f32=out)
        return SimpleNamespace(sample=out)

    def decode_uint8(self, z, num_frames=1, out=None):
        """Same, but the fused tail writes post-processed uint8 frames [n, 8h, 8w, 3] directly -- into `out` when given
        (any uint8 memory this GPU can address: a slice of a clip buffer, or another GPU's memory mapped over NVLink,
        parallel.PeerFrameGather)."""
        n, c, h, w = z.shape
        if out is None:
            out = torch.empty(n, 8 * h, 8 * w, 3, dtype=torch.uint8, device=self._device)
        else:
            assert out.dtype == torch.uint8 and out.is_contiguous() and tuple(out.shape) == (n, 8 * h, 8 * w, 3)
        self.net.decode_chunk(self._cl(z), n, h, w, out_u8=out)
        return out
