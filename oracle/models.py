"""ORACLE (test infrastructure): fp32 restatement of the reference's two networks.

  UNetSpatioTemporalConditionControlNetModel  /root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py
      __init__ :69-245, forward :356-504 (ControlNet residual adds INSIDE the down loop, :451-459 = quirk Q1)
  ControlNetSDVModel trunk                    /root/reference/MOFA-Video-Traj/models/controlnet_sdv.py:156-309
  FlowControlNet (MOFA-Adapter)               /root/reference/MOFA-Video-Traj/models/svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py
      FlowControlNetConditioningEmbeddingSVD :66-101, FlowControlNetFirstFrameEncoder :130-155,
      get_warped_frames :223-234, forward :236-383 (warp adds = quirk Q2)

PINNED (forward graphs, constructor bookkeeping, state-dict layout): tests/golden/networks_tiny.pt is produced by
executing the reference's own three files above on CPU with this oracle's state dict loaded strictly
(oracle/make_goldens.py:make_networks); tests/test_oracle.py requires this file's forward to reproduce it to 1e-6
(measured: bit-exact).  The two adapter-only encoders are additionally pinned by tests/golden/adapter_encoders.pt.
What remains unpinned is the arithmetic INSIDE the diffusers blocks, which come from oracle/d24_blocks.py on both
sides of that comparison (see there).
"""
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import d24_blocks as D
from .softsplat import softsplat

DEFAULT_CONFIG = dict(
    in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                      "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                    "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1,
    num_attention_heads=(5, 10, 20, 20),   # SVD-XT unet/config.json (class default is (5,10,10,20), UNET.py:93)
    num_frames=25, conditioning_channels=3, conditioning_embedding_out_channels=(16, 32, 96, 256),
)


def make_config(**kw):
    c = dict(DEFAULT_CONFIG)
    c.update(kw)
    return c


class _Cfg:
    def __init__(self, d):
        self.__dict__.update(d)


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


class _TimeEmbedMixin:
    def _time_embed(self, sample, timestep, added_time_ids):
        """UNET.py:386-417 == FCN.py:251-282."""
        timesteps = timestep
        if not torch.is_tensor(timesteps):
            timesteps = torch.tensor([timesteps], dtype=torch.float64, device=sample.device)
        elif timesteps.ndim == 0:
            timesteps = timesteps[None].to(sample.device)
        batch_size = sample.shape[0]
        timesteps = timesteps.expand(batch_size)
        t_emb = self.time_proj(timesteps).to(sample.dtype)
        emb = self.time_embedding(t_emb)
        time_embeds = self.add_time_proj(added_time_ids.flatten()).reshape(batch_size, -1).to(emb.dtype)
        return emb + self.add_embedding(time_embeds)


class UNetSpatioTemporalConditionControlNetModel(nn.Module, _TimeEmbedMixin):
    def __init__(self, **config):
        super().__init__()
        c = make_config(**config)
        self.config = _Cfg(c)
        boc = c["block_out_channels"]
        n = len(boc)
        heads, cad = _tup(c["num_attention_heads"], n), _tup(c["cross_attention_dim"], n)
        lpb, tlpb = _tup(c["layers_per_block"], n), _tup(c["transformer_layers_per_block"], n)
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        ted = boc[0] * 4
        self.time_proj = D.Timesteps(boc[0], True, 0)
        self.time_embedding = D.TimestepEmbedding(boc[0], ted)
        self.add_time_proj = D.Timesteps(c["addition_time_embed_dim"], True, 0)
        self.add_embedding = D.TimestepEmbedding(c["projection_class_embeddings_input_dim"], ted)
        self.down_blocks = nn.ModuleList()
        out_ch = boc[0]
        for i, t in enumerate(c["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            self.down_blocks.append(D.get_down_block(
                t, num_layers=lpb[i], transformer_layers_per_block=tlpb[i], in_channels=in_ch, out_channels=out_ch,
                temb_channels=ted, add_downsample=i != n - 1, cross_attention_dim=cad[i],
                num_attention_heads=heads[i]))
        self.mid_block = D.UNetMidBlockSpatioTemporal(boc[-1], ted, transformer_layers_per_block=tlpb[-1],
                                                      cross_attention_dim=cad[-1], num_attention_heads=heads[-1])
        self.up_blocks = nn.ModuleList()
        rboc, rheads, rlpb = list(reversed(boc)), list(reversed(heads)), list(reversed(lpb))
        rcad, rtlpb = list(reversed(cad)), list(reversed(tlpb))
        out_ch = rboc[0]
        for i, t in enumerate(c["up_block_types"]):
            prev, out_ch = out_ch, rboc[i]
            in_ch = rboc[min(i + 1, n - 1)]
            self.up_blocks.append(D.get_up_block(
                t, num_layers=rlpb[i] + 1, transformer_layers_per_block=rtlpb[i], in_channels=in_ch,
                out_channels=out_ch, prev_output_channel=prev, temb_channels=ted, add_upsample=i != n - 1,
                resnet_eps=1e-5, cross_attention_dim=rcad[i], num_attention_heads=rheads[i]))
        self.conv_norm_out = nn.GroupNorm(32, boc[0], eps=1e-5)
        self.conv_out = nn.Conv2d(boc[0], c["out_channels"], 3, padding=1)

    def forward(self, sample, timestep, encoder_hidden_states, down_block_additional_residuals=None,
                mid_block_additional_residual=None, return_dict=True, added_time_ids=None):
        batch_size, num_frames = sample.shape[:2]
        emb = self._time_embed(sample, timestep, added_time_ids)
        sample = sample.flatten(0, 1)
        emb = emb.repeat_interleave(num_frames, dim=0)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(num_frames, dim=0)
        sample = self.conv_in(sample)
        ioi = torch.zeros(batch_size, num_frames, dtype=sample.dtype, device=sample.device)
        down_res = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, res = blk(sample, emb, encoder_hidden_states, ioi)
            else:
                sample, res = blk(sample, emb, ioi)
            down_res += res
            # Q1: zip over ALL skips collected so far, on every iteration (UNET.py:451-459)
            down_res = tuple(r + a for r, a in zip(down_res, down_block_additional_residuals))
        sample = self.mid_block(sample, emb, encoder_hidden_states, ioi)
        sample = sample + mid_block_additional_residual
        for up in self.up_blocks:
            res = down_res[-len(up.resnets):]
            down_res = down_res[:-len(up.resnets)]
            if up.has_cross_attention:
                sample = up(sample, res, emb, encoder_hidden_states, ioi)
            else:
                sample = up(sample, res, emb, ioi)
        sample = self.conv_out(F.silu(self.conv_norm_out(sample)))
        sample = sample.reshape(batch_size, num_frames, *sample.shape[1:])
        return (sample,)


def zero_module(m):  # controlnet_sdv.py:779-782
    for p in m.parameters():
        nn.init.zeros_(p)
    return m


class FlowControlNetConditioningEmbeddingSVD(nn.Module):  # FCN.py:66-101
    def __init__(self, conditioning_embedding_channels, conditioning_channels=3, block_out_channels=(16, 32, 96, 256)):
        super().__init__()
        self.conv_in = nn.Conv2d(conditioning_channels, block_out_channels[0], 3, padding=1)
        self.blocks = nn.ModuleList()
        for i in range(len(block_out_channels) - 1):
            cin, cout = block_out_channels[i], block_out_channels[i + 1]
            self.blocks.append(nn.Conv2d(cin, cin, 3, padding=1))
            self.blocks.append(nn.Conv2d(cin, cout, 3, padding=1, stride=2))
        self.conv_out = zero_module(nn.Conv2d(block_out_channels[-1], conditioning_embedding_channels, 3, padding=1))

    def forward(self, conditioning):
        e = F.silu(self.conv_in(conditioning))
        for b in self.blocks:
            e = F.silu(b(e))
        return self.conv_out(e)


class FlowControlNetFirstFrameEncoderLayer(nn.Module):  # FCN.py:106-126
    def __init__(self, c_in, c_out, is_downsample=False):
        super().__init__()
        self.conv_in = nn.Conv2d(c_in, c_out, 3, padding=1, stride=2 if is_downsample else 1)

    def forward(self, feature):
        return F.silu(self.conv_in(feature))


class FlowControlNetFirstFrameEncoder(nn.Module):  # FCN.py:130-155
    def __init__(self, c_in=320, channels=(320, 640, 1280), downsamples=(True, True, True), use_zeroconv=True):
        super().__init__()
        self.encoders = nn.ModuleList()
        self.zeroconvs = nn.ModuleList()
        for ch, ds in zip(channels, downsamples):
            self.encoders.append(FlowControlNetFirstFrameEncoderLayer(c_in, ch, is_downsample=ds))
            self.zeroconvs.append(zero_module(nn.Conv2d(ch, ch, 1)) if use_zeroconv else nn.Identity())
            c_in = ch

    def forward(self, first_frame):
        feature, deep = first_frame, []
        for enc, zc in zip(self.encoders, self.zeroconvs):
            feature = enc(feature)
            deep.append(zc(feature))
        return deep


class FlowControlNet(nn.Module, _TimeEmbedMixin):
    """ControlNetSDVModel.__init__ (controlnet_sdv.py:156-309) + FlowControlNet overrides (FCN.py:213-221)."""

    def __init__(self, **config):
        super().__init__()
        c = make_config(**config)
        self.config = _Cfg(c)
        boc = c["block_out_channels"]
        n = len(boc)
        heads, cad = _tup(c["num_attention_heads"], n), _tup(c["cross_attention_dim"], n)
        lpb, tlpb = _tup(c["layers_per_block"], n), _tup(c["transformer_layers_per_block"], n)
        self.conv_in = nn.Conv2d(c["in_channels"], boc[0], 3, padding=1)
        ted = boc[0] * 4
        self.time_proj = D.Timesteps(boc[0], True, 0)
        self.time_embedding = D.TimestepEmbedding(boc[0], ted)
        self.add_time_proj = D.Timesteps(c["addition_time_embed_dim"], True, 0)
        self.add_embedding = D.TimestepEmbedding(c["projection_class_embeddings_input_dim"], ted)
        self.down_blocks = nn.ModuleList()
        self.controlnet_down_blocks = nn.ModuleList()
        out_ch = boc[0]
        self.controlnet_down_blocks.append(zero_module(nn.Conv2d(out_ch, out_ch, 1)))
        for i, t in enumerate(c["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            final = i == n - 1
            self.down_blocks.append(D.get_down_block(
                t, num_layers=lpb[i], transformer_layers_per_block=tlpb[i], in_channels=in_ch, out_channels=out_ch,
                temb_channels=ted, add_downsample=not final, cross_attention_dim=cad[i],
                num_attention_heads=heads[i]))
            for _ in range(lpb[i]):
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(out_ch, out_ch, 1)))
            if not final:
                self.controlnet_down_blocks.append(zero_module(nn.Conv2d(out_ch, out_ch, 1)))
        self.controlnet_mid_block = zero_module(nn.Conv2d(boc[-1], boc[-1], 1))
        self.mid_block = D.UNetMidBlockSpatioTemporal(boc[-1], ted, transformer_layers_per_block=tlpb[-1],
                                                      cross_attention_dim=cad[-1], num_attention_heads=heads[-1])
        # FCN.py:215 hard-codes FlowControlNetFirstFrameEncoder() = (320 -> 320, 640, 1280), which equals
        # block_out_channels[:3] at the reference config; parametrised so reduced test configs stay consistent
        self.flow_encoder = FlowControlNetFirstFrameEncoder(c_in=boc[0], channels=tuple(boc[:3]))
        self.controlnet_cond_embedding = FlowControlNetConditioningEmbeddingSVD(
            conditioning_embedding_channels=boc[0], block_out_channels=c["conditioning_embedding_out_channels"],
            conditioning_channels=c["conditioning_channels"])

    def get_warped_frames(self, first_frame, flows):  # FCN.py:223-234
        dtype = first_frame.dtype
        warped = [softsplat(first_frame.float(), flows[:, i].float(), None, "avg").to(dtype).unsqueeze(1)
                  for i in range(flows.shape[1])]
        return torch.cat(warped, dim=1)

    def cond_branch(self, controlnet_cond, controlnet_flow):
        """FCN.py:297-319: the loop-invariant part (cond pyramid, flow pyramid, 4x24 splats)."""
        cond = self.controlnet_cond_embedding(controlnet_cond)
        feats = [cond] + self.flow_encoder(cond)
        fb, fl, fc, fh, fw = controlnet_flow.shape
        scale_flows = {}
        for scale in (8, 16, 32, 64):
            sf = F.interpolate(controlnet_flow.reshape(-1, fc, fh, fw), scale_factor=1 / scale)
            scale_flows[scale] = sf.reshape(fb, fl, fc, fh // scale, fw // scale) / scale
        warped_feats = []
        for feat in feats:
            cb, cc, ch, cw = feat.shape
            w = self.get_warped_frames(feat, scale_flows[fh // ch])
            w = torch.cat([feat.unsqueeze(1), w], dim=1)
            wb, wl, wc, wh, ww = w.shape
            warped_feats.append(w.reshape(wb * wl, wc, wh, ww))
        return warped_feats

    def forward(self, sample, timestep, encoder_hidden_states, added_time_ids, controlnet_cond=None,
                controlnet_flow=None, image_only_indicator=None, return_dict=True, guess_mode=False,
                conditioning_scale=1.0):
        batch_size, num_frames = sample.shape[:2]
        emb = self._time_embed(sample, timestep, added_time_ids)
        sample = sample.flatten(0, 1)
        emb = emb.repeat_interleave(num_frames, dim=0)
        encoder_hidden_states = encoder_hidden_states.repeat_interleave(num_frames, dim=0)
        sample = self.conv_in(sample)
        warped = self.cond_branch(controlnet_cond, controlnet_flow)
        ioi = torch.zeros(batch_size, num_frames, dtype=sample.dtype, device=sample.device)
        count, length = 0, len(warped)
        sample = sample + warped[count]                       # FCN.py:328
        count += 1
        down_res = (sample,)
        for blk in self.down_blocks:
            if blk.has_cross_attention:
                sample, res = blk(sample, emb, encoder_hidden_states, ioi)
            else:
                sample, res = blk(sample, emb, ioi)
            sample = sample + warped[min(count, length - 1)]  # FCN.py:348 (Q2)
            count += 1
            down_res += res
        sample = sample + warped[-1]                          # FCN.py:354
        sample = self.mid_block(sample, emb, encoder_hidden_states, ioi)
        down_res = [cb(r) * conditioning_scale for r, cb in zip(down_res, self.controlnet_down_blocks)]
        mid = self.controlnet_mid_block(sample) * conditioning_scale
        return (down_res, mid, controlnet_flow, None)
