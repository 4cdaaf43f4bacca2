"""Random-init weights in the reference's state-dict layout, for benchmarks and smoke
tests: there are no checkpoints and no network here.  This enumerates every parameter name/shape of the SVD
UNet and the trajectory MOFA-Adapter; `tests/test_synthetic.py` checks the enumeration against the oracle
modules' own state_dict() key/shape sets (and thereby against the published 1,524,623,082 UNet parameters).

Init (seeded): N(0, 1/fan_in) weights, zero biases, unit norms, mix_factor 0.5,
out-projections x0.1, zero-convs N(0, 0.02^2) so the adapter contributes.
"""
import math

import torch

SVD_XT_CONFIG = dict(
    sample_size=96, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                      "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                    "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1, num_attention_heads=(5, 10, 20, 20), num_frames=25,
    conditioning_channels=3, conditioning_embedding_out_channels=(16, 32, 96, 256),
)


class _Gen:
    def __init__(self, seed, dtype):
        self.g = torch.Generator().manual_seed(seed)
        self.sd = {}
        self.dtype = dtype

    def w(self, name, shape, fan_in=None, gain=1.0, std=None):
        if std is None:
            fan_in = fan_in if fan_in is not None else int(torch.tensor(shape[1:]).prod())
            std = gain / math.sqrt(fan_in)
        self.sd[name] = (torch.randn(shape, generator=self.g) * std).to(self.dtype)

    def const(self, name, shape, v):
        self.sd[name] = torch.full(shape, v).to(self.dtype)

    def conv(self, name, cout, cin, k=3, gain=1.0, std=None):
        self.w(name + ".weight", (cout, cin, k, k), gain=gain, std=std)
        if std is None:
            self.const(name + ".bias", (cout,), 0.0)
        else:
            self.w(name + ".bias", (cout,), std=std)

    def tconv(self, name, cout, cin, gain=1.0):
        self.w(name + ".weight", (cout, cin, 3, 1, 1), gain=gain)
        self.const(name + ".bias", (cout,), 0.0)

    def lin(self, name, cout, cin, bias=True, gain=1.0):
        self.w(name + ".weight", (cout, cin), gain=gain)
        if bias:
            self.const(name + ".bias", (cout,), 0.0)

    def norm(self, name, c):
        self.const(name + ".weight", (c,), 1.0)
        self.const(name + ".bias", (c,), 0.0)

    # composite blocks --------------------------------------------------------------------------
    def resblock(self, pre, cin, cout, temb):
        sp, tp = pre + ".spatial_res_block", pre + ".temporal_res_block"
        self.norm(sp + ".norm1", cin)
        self.conv(sp + ".conv1", cout, cin)
        self.lin(sp + ".time_emb_proj", cout, temb)
        self.norm(sp + ".norm2", cout)
        self.conv(sp + ".conv2", cout, cout, gain=0.1)
        if cin != cout:
            self.conv(sp + ".conv_shortcut", cout, cin, k=1)
        self.norm(tp + ".norm1", cout)
        self.tconv(tp + ".conv1", cout, cout)
        self.lin(tp + ".time_emb_proj", cout, temb)
        self.norm(tp + ".norm2", cout)
        self.tconv(tp + ".conv2", cout, cout, gain=0.1)
        self.const(pre + ".time_mixer.mix_factor", (1,), 0.5)

    def attn(self, pre, dim, kv_dim):
        self.lin(pre + ".to_q", dim, dim, bias=False)
        self.lin(pre + ".to_k", dim, kv_dim, bias=False)
        self.lin(pre + ".to_v", dim, kv_dim, bias=False)
        self.lin(pre + ".to_out.0", dim, dim, gain=0.1)

    def ff(self, pre, dim):
        self.lin(pre + ".net.0.proj", 8 * dim, dim)
        self.lin(pre + ".net.2", dim, 4 * dim, gain=0.1)

    def transformer(self, pre, C, ctx):
        self.norm(pre + ".norm", C)
        self.lin(pre + ".proj_in", C, C)
        sb, tb = pre + ".transformer_blocks.0", pre + ".temporal_transformer_blocks.0"
        self.norm(sb + ".norm1", C)
        self.attn(sb + ".attn1", C, C)
        self.norm(sb + ".norm2", C)
        self.attn(sb + ".attn2", C, ctx)
        self.norm(sb + ".norm3", C)
        self.ff(sb + ".ff", C)
        self.norm(tb + ".norm_in", C)
        self.ff(tb + ".ff_in", C)
        self.norm(tb + ".norm1", C)
        self.attn(tb + ".attn1", C, C)
        self.norm(tb + ".norm2", C)
        self.attn(tb + ".attn2", C, ctx)
        self.norm(tb + ".norm3", C)
        self.ff(tb + ".ff", C)
        self.lin(pre + ".time_pos_embed.linear_1", 4 * C, C)
        self.lin(pre + ".time_pos_embed.linear_2", C, 4 * C)
        self.const(pre + ".time_mixer.mix_factor", (1,), 0.5)
        self.lin(pre + ".proj_out", C, C, gain=0.1)


def _tup(v, n):
    return tuple(v) if isinstance(v, (tuple, list)) else (v,) * n


def _trunk(g, cfg):
    boc = tuple(cfg["block_out_channels"])
    n = len(boc)
    lpb, ctx = _tup(cfg["layers_per_block"], n), _tup(cfg["cross_attention_dim"], n)
    temb = boc[0] * 4
    g.conv("conv_in", boc[0], cfg["in_channels"])
    g.lin("time_embedding.linear_1", temb, boc[0])
    g.lin("time_embedding.linear_2", temb, temb)
    g.lin("add_embedding.linear_1", temb, cfg["projection_class_embeddings_input_dim"])
    g.lin("add_embedding.linear_2", temb, temb)
    out = boc[0]
    for i, t in enumerate(cfg["down_block_types"]):
        cin, out = out, boc[i]
        for j in range(lpb[i]):
            g.resblock(f"down_blocks.{i}.resnets.{j}", cin if j == 0 else out, out, temb)
            if "CrossAttn" in t:
                g.transformer(f"down_blocks.{i}.attentions.{j}", out, ctx[i])
        if i != n - 1:
            g.conv(f"down_blocks.{i}.downsamplers.0.conv", out, out)
    C = boc[-1]
    g.resblock("mid_block.resnets.0", C, C, temb)
    g.transformer("mid_block.attentions.0", C, ctx[-1])
    g.resblock("mid_block.resnets.1", C, C, temb)
    return boc, n, lpb, ctx, temb


def unet_state_dict(config=None, seed=0, dtype=torch.float16):
    cfg = dict(SVD_XT_CONFIG)
    cfg.update(config or {})
    g = _Gen(seed, dtype)
    boc, n, lpb, ctx, temb = _trunk(g, cfg)
    rboc, rlpb, rctx = list(reversed(boc)), list(reversed(lpb)), list(reversed(ctx))
    out = rboc[0]
    for i, t in enumerate(cfg["up_block_types"]):
        prev, out = out, rboc[i]
        cin = rboc[min(i + 1, n - 1)]
        nl = rlpb[i] + 1
        for j in range(nl):
            skip = cin if j == nl - 1 else out
            rin = prev if j == 0 else out
            g.resblock(f"up_blocks.{i}.resnets.{j}", rin + skip, out, temb)
            if "CrossAttn" in t:
                g.transformer(f"up_blocks.{i}.attentions.{j}", out, rctx[i])
        if i != n - 1:
            g.conv(f"up_blocks.{i}.upsamplers.0.conv", out, out)
    g.norm("conv_norm_out", boc[0])
    g.conv("conv_out", cfg["out_channels"], boc[0])
    return cfg, g.sd


def adapter_state_dict(config=None, seed=1, dtype=torch.float16, zero_std=0.02):
    cfg = dict(SVD_XT_CONFIG)
    cfg.update(config or {})
    g = _Gen(seed, dtype)
    boc, n, lpb, ctx, temb = _trunk(g, cfg)
    k = 0
    g.conv(f"controlnet_down_blocks.{k}", boc[0], boc[0], k=1, std=zero_std)
    for i in range(n):
        for _ in range(lpb[i]):
            k += 1
            g.conv(f"controlnet_down_blocks.{k}", boc[i], boc[i], k=1, std=zero_std)
        if i != n - 1:
            k += 1
            g.conv(f"controlnet_down_blocks.{k}", boc[i], boc[i], k=1, std=zero_std)
    g.conv("controlnet_mid_block", boc[-1], boc[-1], k=1, std=zero_std)
    ce = "controlnet_cond_embedding"
    ceo = tuple(cfg["conditioning_embedding_out_channels"])
    g.conv(ce + ".conv_in", ceo[0], cfg["conditioning_channels"])
    for b in range(len(ceo) - 1):
        g.conv(f"{ce}.blocks.{2 * b}", ceo[b], ceo[b])
        g.conv(f"{ce}.blocks.{2 * b + 1}", ceo[b + 1], ceo[b])
    g.conv(ce + ".conv_out", boc[0], ceo[-1], std=zero_std)
    cin = boc[0]
    for b in range(3):
        g.conv(f"flow_encoder.encoders.{b}.conv_in", boc[b], cin)
        g.conv(f"flow_encoder.zeroconvs.{b}", boc[b], boc[b], k=1, std=zero_std)
        cin = boc[b]
    return cfg, g.sd


def ldmk_adapter_state_dict(config=None, seed=3, dtype=torch.float16, zero_std=0.02):
    """Keypoint (landmark) adapter: the trajectory adapter's trunk plus the landmark embedding, the per-scale occlusion
    hourglasses and `zero_outs` (/root/reference/MOFA-Video-Keypoint/models/ldmk_ctrlnet.py:187-254,
    models/occlusion/hourglass.py:227-246); its flow encoder has no zero-convs (:144-161, use_zeroconv=False)."""
    cfg, sd = adapter_state_dict(config, seed=seed, dtype=dtype, zero_std=zero_std)
    for k in [k for k in sd if k.startswith("flow_encoder.zeroconvs.")]:
        del sd[k]
    g = _Gen(seed + 100, dtype)
    boc = tuple(cfg["block_out_channels"])
    le, lch = "controlnet_ldmk_embedding", (16, 32, 64, 128)
    g.conv(le + ".conv_in", lch[0], cfg["conditioning_channels"])
    for b in range(len(lch) - 1):
        g.conv(f"{le}.blocks.{2 * b}", lch[b], lch[b])
        g.conv(f"{le}.blocks.{2 * b + 1}", lch[b + 1], lch[b])
    g.conv(le + ".conv_out", boc[0], lch[-1], std=zero_std)
    for s_, c in (("8", boc[0]), ("16", boc[0]), ("32", boc[1]), ("64", boc[2])):
        g.conv(f"zero_outs.{s_}", c, c, k=1, std=zero_std)
        pre = f"occlusions.{s_}"
        enc = [2 * c + 2, 128, 256, 512]
        for i in range(3):
            g.conv(f"{pre}.hourglass.encoder.down_blocks.{i}.conv", enc[i + 1], enc[i])
        for i, (cin, cout) in enumerate(((512, 256), (512, 128), (256, 64))):
            g.conv(f"{pre}.hourglass.decoder.up_blocks.{i}.conv", cout, cin)
        g.conv(f"{pre}.matting_mask", 1, 64, k=7)
        g.conv(f"{pre}.matting", c, 64, k=7)
    sd.update(g.sd)
    return cfg, sd


def cmp_state_dict(seed=2, dtype=torch.float32, prefix="module."):
    """Random-init CMP (ResNet-50-dilated + ShallowNet + MotionDecoderSkipLayer) in the reference checkpoint's
    key layout (`module.` prefix of FixModule, models/cmp/models/modules/others.py:3-10).  BN gamma < 1 keeps the
    50-layer residual stack fp16-finite with random weights."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, cout, cin, k, bias):
        sd[name + ".weight"] = (torch.randn(cout, cin, k, k, generator=g) / math.sqrt(cin * k * k)).to(dtype)
        if bias:
            sd[name + ".bias"] = torch.zeros(cout, dtype=dtype)

    def bn(name, c):
        sd[name + ".weight"] = (torch.rand(c, generator=g) * 0.4 + 0.3).to(dtype)
        sd[name + ".bias"] = (torch.randn(c, generator=g) * 0.1).to(dtype)
        sd[name + ".running_mean"] = (torch.randn(c, generator=g) * 0.1).to(dtype)
        sd[name + ".running_var"] = (torch.rand(c, generator=g) + 0.5).to(dtype)
        sd[name + ".num_batches_tracked"] = torch.zeros((), dtype=torch.long)

    ie = "image_encoder."
    conv(ie + "conv1", 64, 3, 7, False)
    bn(ie + "bn1", 64)
    inpl = 64
    for li, (planes, blocks) in enumerate([(64, 3), (128, 4), (256, 6), (512, 3)]):
        for bi in range(blocks):
            pre = f"{ie}layer{li + 1}.{bi}."
            conv(pre + "conv1", planes, inpl, 1, False)
            bn(pre + "bn1", planes)
            conv(pre + "conv2", planes, planes, 3, False)
            bn(pre + "bn2", planes)
            conv(pre + "conv3", planes * 4, planes, 1, False)
            bn(pre + "bn3", planes * 4)
            if bi == 0:
                conv(pre + "downsample.0", planes * 4, inpl, 1, False)
                bn(pre + "downsample.1", planes * 4)
            inpl = planes * 4
    conv(ie + "conv5", 256, 2048, 1, True)
    fe = "flow_encoder.features."
    conv(fe + "0", 16, 4, 5, True)
    bn(fe + "1", 16)
    conv(fe + "4", 16, 16, 3, True)
    bn(fe + "5", 16)
    fd = "flow_decoder."
    for k in (1, 2, 4, 8):
        off = 0 if k == 1 else 1
        for j, cin in enumerate((272, 128, 128)):
            conv(f"{fd}decoder{k}.{off + 3 * j}", 128, cin, 3, True)
            bn(f"{fd}decoder{k}.{off + 3 * j + 1}", 128)
    for name, cout, cin in (("fusion8", 256, 512), ("skipconv4", 128, 256), ("fusion4", 128, 384),
                            ("skipconv2", 32, 64), ("fusion2", 64, 160)):
        conv(f"{fd}{name}.0", cout, cin, 3, True)
        bn(f"{fd}{name}.1", cout)
    conv(fd + "head", 198, 64, 1, True)
    return {prefix + k: v for k, v in sd.items()}
