"""AutoencoderKLTemporalDecoder stand-in (the reference takes this object from diffusers 0.24,
/root/reference/MOFA-Video-Traj/run_gradio.py:101-102, pipeline.py:143-164, 194-220).

This PyTorch module is the weight container / parameter-name contract of the VAE (a real `vae` checkpoint loads into
it) and the fp32 statement the native path is tested against; at run time `vae_engine.NativeTemporalDecoderVAE` wraps it
and executes both encode and decode on the sm_100a kernels.  A caller may still pass this
module (or diffusers' own) to the pipeline directly -- the pipeline only uses `.encode(x).latent_dist.mode()`,
`.decode(z, num_frames=n).sample`, `.config`, `.dtype`.
Topology restates diffusers 0.24 `Encoder` / `TemporalDecoder`
(SURVEY.md App. A.2): encoder (128,256,512,512) x2 ResnetBlock2D + mid res-attn(d=512)-res -> 8ch ->
quant_conv, mode() = mean; decoder conv_in 4->512, mid res-attn-res, 4 up blocks of 3
SpatioTemporalResBlock(merge 'learned', switch_spatial_to_temporal_mix) + nearest-2x conv, GN-SiLU-conv 128->3,
time_conv_out Conv3d(3,3,(3,1,1)).  Parameter names follow diffusers so a real `vae` checkpoint loads.
"""
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F


class _Res2D(nn.Module):
    def __init__(self, cin, cout, eps=1e-6):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=eps)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=eps)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.conv_shortcut = nn.Conv2d(cin, cout, 1) if cin != cout else None

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return (x if self.conv_shortcut is None else self.conv_shortcut(x)) + h


class _TRes(nn.Module):
    def __init__(self, c, eps=1e-5):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, c, eps=eps)
        self.conv1 = nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0))
        self.norm2 = nn.GroupNorm(32, c, eps=eps)
        self.conv2 = nn.Conv3d(c, c, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        return x + h


class _Mixer(nn.Module):
    def __init__(self):
        super().__init__()
        self.mix_factor = nn.Parameter(torch.zeros(1))


class _STRes(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.spatial_res_block = _Res2D(cin, cout, 1e-6)
        self.temporal_res_block = _TRes(cout, 1e-5)
        self.time_mixer = _Mixer()

    def forward(self, x, num_frames):
        x = self.spatial_res_block(x)
        bf, c, h, w = x.shape
        b = bf // num_frames
        xs = x.reshape(b, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        xt = self.temporal_res_block(xs)
        a = 1.0 - torch.sigmoid(self.time_mixer.mix_factor).to(x.dtype)  # switch_spatial_to_temporal_mix
        out = a * xs + (1.0 - a) * xt
        return out.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class _Attn(nn.Module):  # single head, head_dim = channels
    def __init__(self, c):
        super().__init__()
        self.group_norm = nn.GroupNorm(32, c, eps=1e-6)
        self.to_q, self.to_k, self.to_v = nn.Linear(c, c), nn.Linear(c, c), nn.Linear(c, c)
        self.to_out = nn.ModuleList([nn.Linear(c, c), nn.Dropout(0.0)])

    def forward(self, x):
        n, c, h, w = x.shape
        t = self.group_norm(x).view(n, c, h * w).transpose(1, 2)
        q, k, v = self.to_q(t)[:, None], self.to_k(t)[:, None], self.to_v(t)[:, None]
        o = F.scaled_dot_product_attention(q, k, v)[:, 0]
        o = self.to_out[0](o).transpose(1, 2).reshape(n, c, h, w)
        return o + x


class _Down(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, stride=2, padding=0)

    def forward(self, x):
        return self.conv(F.pad(x, (0, 1, 0, 1)))


class _Up(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _EncBlock(nn.Module):
    def __init__(self, cin, cout, down):
        super().__init__()
        self.resnets = nn.ModuleList([_Res2D(cin, cout), _Res2D(cout, cout)])
        self.downsamplers = nn.ModuleList([_Down(cout)]) if down else None

    def forward(self, x):
        for r in self.resnets:
            x = r(x)
        return self.downsamplers[0](x) if self.downsamplers is not None else x


class _Mid2D(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([_Res2D(c, c), _Res2D(c, c)])
        self.attentions = nn.ModuleList([_Attn(c)])

    def forward(self, x):
        return self.resnets[1](self.attentions[0](self.resnets[0](x)))


class Encoder(nn.Module):
    def __init__(self, boc=(128, 256, 512, 512), latent=4):
        super().__init__()
        self.conv_in = nn.Conv2d(3, boc[0], 3, padding=1)
        self.down_blocks = nn.ModuleList()
        c = boc[0]
        for i, o in enumerate(boc):
            self.down_blocks.append(_EncBlock(c, o, i != len(boc) - 1))
            c = o
        self.mid_block = _Mid2D(c)
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 2 * latent, 3, padding=1)

    def forward(self, x):
        x = self.conv_in(x)
        for b in self.down_blocks:
            x = b(x)
        x = self.mid_block(x)
        return self.conv_out(F.silu(self.conv_norm_out(x)))


class _MidT(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.resnets = nn.ModuleList([_STRes(c, c), _STRes(c, c)])
        self.attentions = nn.ModuleList([_Attn(c)])

    def forward(self, x, nf):
        return self.resnets[1](self.attentions[0](self.resnets[0](x, nf)), nf)


class _UpT(nn.Module):
    def __init__(self, cin, cout, up):
        super().__init__()
        self.resnets = nn.ModuleList([_STRes(cin if i == 0 else cout, cout) for i in range(3)])
        self.upsamplers = nn.ModuleList([_Up(cout)]) if up else None

    def forward(self, x, nf):
        for r in self.resnets:
            x = r(x, nf)
        return self.upsamplers[0](x) if self.upsamplers is not None else x


class TemporalDecoder(nn.Module):
    def __init__(self, boc=(128, 256, 512, 512), latent=4):
        super().__init__()
        self.conv_in = nn.Conv2d(latent, boc[-1], 3, padding=1)
        self.mid_block = _MidT(boc[-1])
        rb = list(reversed(boc))
        self.up_blocks = nn.ModuleList()
        c = rb[0]
        for i, o in enumerate(rb):
            self.up_blocks.append(_UpT(c, o, i != len(rb) - 1))
            c = o
        self.conv_norm_out = nn.GroupNorm(32, c, eps=1e-6)
        self.conv_out = nn.Conv2d(c, 3, 3, padding=1)
        self.time_conv_out = nn.Conv3d(3, 3, (3, 1, 1), padding=(1, 0, 0))

    def forward(self, z, num_frames):
        x = self.conv_in(z)
        x = self.mid_block(x, num_frames)
        for b in self.up_blocks:
            x = b(x, num_frames)
        x = self.conv_out(F.silu(self.conv_norm_out(x)))
        bf, c, h, w = x.shape
        x = x.reshape(bf // num_frames, num_frames, c, h, w).permute(0, 2, 1, 3, 4)
        x = self.time_conv_out(x)
        return x.permute(0, 2, 1, 3, 4).reshape(bf, c, h, w)


class _Dist:
    def __init__(self, moments):
        self.mean = moments.chunk(2, dim=1)[0]

    def mode(self):
        return self.mean


class AutoencoderKLTemporalDecoder(nn.Module):
    def __init__(self, block_out_channels=(128, 256, 512, 512), latent_channels=4, scaling_factor=0.18215,
                 force_upcast=True):
        super().__init__()
        self.encoder = Encoder(block_out_channels, latent_channels)
        self.decoder = TemporalDecoder(block_out_channels, latent_channels)
        self.quant_conv = nn.Conv2d(2 * latent_channels, 2 * latent_channels, 1)
        self.config = SimpleNamespace(block_out_channels=tuple(block_out_channels), latent_channels=latent_channels,
                                      scaling_factor=scaling_factor, force_upcast=force_upcast)

    @property
    def dtype(self):
        return next(self.parameters()).dtype

    def encode(self, x):
        return SimpleNamespace(latent_dist=_Dist(self.quant_conv(self.encoder(x))))

    def decode(self, z, num_frames=1):
        return SimpleNamespace(sample=self.decoder(z, num_frames))

    def forward(self, x, num_frames=1):
        return self.decode(self.encode(x).latent_dist.mode(), num_frames)
