"""Host-side runner of the SVD UNet and the MOFA-Adapter trunk on the sm_100a kernels.

Weights arrive in the reference's state-dict layout (diffusers key names) and are
repacked once into kernel-native form: conv weights as [Cout, (ky,kx,cin)] / [Cout, (kt,cin)] GEMM
operands, q|k|v fused, GEGLU rows interleaved per N tile, every time_emb_proj of a network concatenated
into one [sum Cout, 1280] matrix, sigmoid(mix_factor) folded into epilogue scalars, the frame position
embedding and the collapsed single-token cross-attention vectors precomputed.

Activations are fp16 channels-last matrices [frames*h*w, C]; every op is one C-ABI call through `ops`
(mofa_video_b200.lib; tests substitute the plain-PyTorch statements of tests/ref_ops.py to check this
host logic on CPU).  There is no PyTorch compute on this path.

Reference graph being executed:
  UNet forward      /root/reference/MOFA-Video-Traj/models/unet_spatio_temporal_condition_controlnet.py:356-504
  adapter forward   /root/reference/MOFA-Video-Traj/models/svdxt_featureflow_forward_controlnet_s2d_fixcmp_norefine.py:236-383
  blocks            diffusers 0.24.0 (restated in oracle/d24_blocks.py)
Exact work removed (SURVEY.md §8d "algorithmic"): single-KV-token cross-attention collapses to a
per-batch vector; the adapter's cond/warp branch is computed once per clip for one CFG half.
"""
import math

import torch


# ------------------------------------------------------------------------------------------------
# packing helpers
# ------------------------------------------------------------------------------------------------
class _Packer:
    def __init__(self, sd, device):
        self.sd = sd
        self.device = device

    def h(self, t):
        return t.detach().to(device=self.device, dtype=torch.float16).contiguous()

    def get(self, key):
        return self.sd[key]

    def vec(self, key):
        return self.h(self.sd[key])

    def conv3(self, key):
        w = self.sd[key + ".weight"]
        o, i = w.shape[:2]
        return self.h(w.permute(0, 2, 3, 1).reshape(o, 9 * i)), self.vec(key + ".bias")

    def tconv(self, key):
        w = self.sd[key + ".weight"]  # [O, I, 3, 1, 1]
        o, i = w.shape[:2]
        return self.h(w[:, :, :, 0, 0].permute(0, 2, 1).reshape(o, 3 * i)), self.vec(key + ".bias")

    def conv1(self, key):
        w = self.sd[key + ".weight"]
        return self.h(w.reshape(w.shape[0], w.shape[1])), self.vec(key + ".bias")

    def lin(self, key, bias=True):
        return self.h(self.sd[key + ".weight"]), (self.vec(key + ".bias") if bias else None)

    def norm(self, key):
        return self.vec(key + ".weight"), self.vec(key + ".bias")

    def geglu(self, key, pick_bn):
        w, b = self.sd[key + ".weight"], self.sd[key + ".bias"]  # [2*inner, dim]: value rows then gate rows
        inner = w.shape[0] // 2
        bn = pick_bn(w.shape[0], True)
        hb = bn // 2
        wv, wg = w[:inner].reshape(inner // hb, hb, -1), w[inner:].reshape(inner // hb, hb, -1)
        bv, bg = b[:inner].reshape(inner // hb, hb), b[inner:].reshape(inner // hb, hb)
        wp = torch.cat([wv, wg], dim=1).reshape(2 * inner, -1)
        bp = torch.cat([bv, bg], dim=1).reshape(2 * inner)
        return self.h(wp), self.h(bp), bn


class _TembBank:
    """All time_emb_proj layers of one network as a single [sum Cout, temb_dim] operand."""

    def __init__(self):
        self.ws, self.bs, self.total = [], [], 0

    def add(self, w, b):
        off = self.total
        self.ws.append(w)
        self.bs.append(b)
        self.total += w.shape[0]
        return off

    def finish(self, pk):
        self.W = pk.h(torch.cat(self.ws, dim=0))
        self.b = pk.h(torch.cat(self.bs, dim=0))
        del self.ws, self.bs


def _frame_pos_embed(sd, prefix, C, T):
    """TransformerSpatioTemporalModel: time_pos_embed(time_proj(arange(T))) -> [T, C] (loop invariant)."""
    half = C // 2
    t = torch.arange(T, dtype=torch.float32)
    freq = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    arg = t[:, None] * freq[None]
    emb = torch.cat([torch.cos(arg), torch.sin(arg)], dim=-1).half().float()  # flip_sin_to_cos; fp16 like the ref
    w1, b1 = sd[prefix + ".linear_1.weight"].float().cpu(), sd[prefix + ".linear_1.bias"].float().cpu()
    w2, b2 = sd[prefix + ".linear_2.weight"].float().cpu(), sd[prefix + ".linear_2.bias"].float().cpu()
    hdn = torch.nn.functional.silu(emb @ w1.t() + b1)
    return (hdn @ w2.t() + b2)


class Net:
    """Packed weights + forward for either network.  kind = 'unet' | 'adapter'."""

    def __init__(self, kind, state_dict, config, ops, device):
        self.kind, self.cfg, self.ops, self.device = kind, config, ops, device
        pk = _Packer(state_dict, device)
        self.pk_bn = ops.pick_bn
        boc = tuple(config["block_out_channels"])
        n = len(boc)
        heads = config["num_attention_heads"]
        heads = tuple(heads) if isinstance(heads, (tuple, list)) else (heads,) * n
        self.boc, self.heads, self.T = boc, heads, config["num_frames"]
        lpb = config["layers_per_block"]
        lpb = tuple(lpb) if isinstance(lpb, (tuple, list)) else (lpb,) * n
        self.temb = _TembBank()
        self.xattn = []  # (Wv, Wo, bo) per collapsed cross-attention, evaluated once per clip

        p = {}
        w = state_dict["conv_in.weight"]
        kin = 9 * w.shape[1]
        self.conv_in_kpad = (kin + 7) // 8 * 8
        wp = torch.zeros(w.shape[0], self.conv_in_kpad)
        wp[:, :kin] = w.permute(0, 2, 3, 1).reshape(w.shape[0], kin)
        p["conv_in"] = (pk.h(wp), pk.vec("conv_in.bias"))
        p["time_embedding"] = (pk.lin("time_embedding.linear_1"), pk.lin("time_embedding.linear_2"))
        p["add_embedding"] = (pk.lin("add_embedding.linear_1"), pk.lin("add_embedding.linear_2"))
        self.time_dim, self.add_dim = boc[0], config["addition_time_embed_dim"]

        down = []
        out_ch = boc[0]
        for i, t in enumerate(config["down_block_types"]):
            in_ch, out_ch = out_ch, boc[i]
            blk = {"res": [], "attn": [], "down": None}
            for j in range(lpb[i]):
                blk["res"].append(self._pack_res(pk, f"down_blocks.{i}.resnets.{j}", in_ch if j == 0 else out_ch,
                                                 out_ch, 1e-6 if "CrossAttn" in t else 1e-5))
                if "CrossAttn" in t:
                    blk["attn"].append(self._pack_tr(pk, f"down_blocks.{i}.attentions.{j}", out_ch, heads[i]))
            if i != n - 1:
                blk["down"] = pk.conv3(f"down_blocks.{i}.downsamplers.0.conv")
            down.append(blk)
        p["down"] = down
        C = boc[-1]
        p["mid"] = {"res": [self._pack_res(pk, "mid_block.resnets.0", C, C, 1e-5),
                            self._pack_res(pk, "mid_block.resnets.1", C, C, 1e-5)],
                    "attn": [self._pack_tr(pk, "mid_block.attentions.0", C, heads[-1])]}
        if kind == "unet":
            up = []
            rboc, rheads, rlpb = list(reversed(boc)), list(reversed(heads)), list(reversed(lpb))
            out_ch = rboc[0]
            for i, t in enumerate(config["up_block_types"]):
                prev, out_ch = out_ch, rboc[i]
                in_ch = rboc[min(i + 1, n - 1)]
                nl = rlpb[i] + 1
                blk = {"res": [], "attn": [], "up": None}
                for j in range(nl):
                    skip = in_ch if j == nl - 1 else out_ch
                    rin = prev if j == 0 else out_ch
                    blk["res"].append(self._pack_res(pk, f"up_blocks.{i}.resnets.{j}", rin + skip, out_ch, 1e-5,
                                                     split=rin))
                    if "CrossAttn" in t:
                        blk["attn"].append(self._pack_tr(pk, f"up_blocks.{i}.attentions.{j}", out_ch, rheads[i]))
                if i != n - 1:
                    blk["up"] = pk.conv3(f"up_blocks.{i}.upsamplers.0.conv")
                up.append(blk)
            p["up"] = up
            p["norm_out"] = pk.norm("conv_norm_out")
            p["conv_out"] = pk.conv3("conv_out")
            self.out_channels = config["out_channels"]
        else:
            p["zero_down"] = [pk.conv1(f"controlnet_down_blocks.{k}") for k in range(sum(lpb) + n)]
            p["zero_mid"] = pk.conv1("controlnet_mid_block")
            ce = "controlnet_cond_embedding"
            convs = [(ce + ".conv_in", 1)]
            nb = len(config["conditioning_embedding_out_channels"]) - 1
            for k in range(nb):
                convs += [(f"{ce}.blocks.{2 * k}", 1), (f"{ce}.blocks.{2 * k + 1}", 2)]
            convs.append((ce + ".conv_out", 1))
            p["cond_convs"] = [self._pack_im2col_conv(pk, name, s) for name, s in convs]
            # the Keypoint variant's first-frame encoder has no zero-convs (K/models/ldmk_ctrlnet.py:144-161)
            p["flow_enc"] = [(self._pack_im2col_conv(pk, f"flow_encoder.encoders.{k}.conv_in", 2),
                              pk.conv1(f"flow_encoder.zeroconvs.{k}")
                              if f"flow_encoder.zeroconvs.{k}.weight" in state_dict else None) for k in range(3)]
        self.temb.finish(pk)
        self.p = p

    # ------------------------------------------------------------------ packed-weight cache
    _TRANSIENT = ("ops", "pk_bn", "device", "_pbufs", "warped", "cond_feats", "xvec", "aug_emb", "ldmk", "B")

    def packed_state(self):
        """Everything __init__ computed from the checkpoint (kernel-native operands, on the CPU) -- what a process
        needs to skip the repack: `Net.from_packed(state, ops, device)`."""
        def cpu(v):
            if isinstance(v, torch.Tensor):
                return v.detach().cpu()
            if isinstance(v, dict):
                return {k: cpu(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(cpu(x) for x in v)
            if isinstance(v, _TembBank):
                b = _TembBank()
                b.__dict__.update({k: cpu(x) for k, x in v.__dict__.items()})
                return b
            return v
        return {"class": type(self).__name__,
                "attrs": {k: cpu(v) for k, v in self.__dict__.items() if k not in self._TRANSIENT}}

    @classmethod
    def from_packed(cls, state, ops, device):
        if state["class"] != cls.__name__:
            raise ValueError(f"packed weights of {state['class']} offered to {cls.__name__}")
        dev = torch.device(device)

        def put(v):
            if isinstance(v, torch.Tensor):
                return v.to(dev)
            if isinstance(v, dict):
                return {k: put(x) for k, x in v.items()}
            if isinstance(v, (list, tuple)):
                return type(v)(put(x) for x in v)
            if isinstance(v, _TembBank):
                v.__dict__.update({k: put(x) for k, x in v.__dict__.items()})
                return v
            return v
        net = cls.__new__(cls)
        net.__dict__.update({k: put(v) for k, v in state["attrs"].items()})
        net.ops, net.device, net.pk_bn = ops, dev, ops.pick_bn
        if hasattr(net, "occ") or "ldmk_convs" in net.p:
            net.ldmk = None
        return net

    # ------------------------------------------------------------------ packing of composite blocks
    def _pack_im2col_conv(self, pk, name, stride):
        w = pk.get(name + ".weight")
        o, i = w.shape[:2]
        k = 9 * i
        kpad = (k + 7) // 8 * 8
        wp = torch.zeros(o, kpad)
        wp[:, :k] = w.permute(0, 2, 3, 1).reshape(o, k)
        return {"w": pk.h(wp), "b": pk.vec(name + ".bias"), "cin": i, "cout": o, "kpad": kpad, "stride": stride}

    def _pack_res(self, pk, pre, cin, cout, eps, split=None, temporal_eps=None, switch=False):
        sp, tp = pre + ".spatial_res_block", pre + ".temporal_res_block"
        r = {"cin": cin, "cout": cout, "eps": eps, "teps": temporal_eps if temporal_eps is not None else eps,
             "split": split}
        r["n1"], r["c1"] = pk.norm(sp + ".norm1"), pk.conv3(sp + ".conv1")
        r["n2"], r["c2"] = pk.norm(sp + ".norm2"), pk.conv3(sp + ".conv2")
        r["sc"] = pk.conv1(sp + ".conv_shortcut") if (sp + ".conv_shortcut.weight") in pk.sd else None
        has_temb = (sp + ".time_emb_proj.weight") in pk.sd  # the VAE's blocks have no time embedding
        r["temb_sp"] = self.temb.add(pk.get(sp + ".time_emb_proj.weight"),
                                     pk.get(sp + ".time_emb_proj.bias")) if has_temb else None
        r["tn1"], r["tc1"] = pk.norm(tp + ".norm1"), pk.tconv(tp + ".conv1")
        r["tn2"], r["tc2"] = pk.norm(tp + ".norm2"), pk.tconv(tp + ".conv2")
        r["temb_tp"] = self.temb.add(pk.get(tp + ".time_emb_proj.weight"),
                                     pk.get(tp + ".time_emb_proj.bias")) if has_temb else None
        a = torch.sigmoid(pk.get(pre + ".time_mixer.mix_factor").float()).item()
        r["alpha"] = (1.0 - a) if switch else a
        return r

    def _pack_tr(self, pk, pre, C, heads):
        sb, tb = pre + ".transformer_blocks.0", pre + ".temporal_transformer_blocks.0"
        t = {"C": C, "heads": heads}
        if C % heads or (C // heads) % 2 or C // heads > 128:
            raise ValueError(f"attention head_dim {C}/{heads}: the kernels take even head dims <= 128")
        t["norm"], t["proj_in"], t["proj_out"] = pk.norm(pre + ".norm"), pk.lin(pre + ".proj_in"), pk.lin(pre + ".proj_out")

        def qkv(a):
            return pk.h(torch.cat([pk.get(a + ".to_q.weight"), pk.get(a + ".to_k.weight"), pk.get(a + ".to_v.weight")], 0))

        def xattn(a):
            self.xattn.append((pk.h(pk.get(a + ".to_v.weight")), pk.h(pk.get(a + ".to_out.0.weight")),
                               pk.vec(a + ".to_out.0.bias")))
            return len(self.xattn) - 1

        t["s_n1"], t["s_qkv"], t["s_o"] = pk.norm(sb + ".norm1"), qkv(sb + ".attn1"), pk.lin(sb + ".attn1.to_out.0")
        t["s_x"] = xattn(sb + ".attn2")
        # FeedForward blocks of width C <= 320 run as ONE fused kernel (mofa_ff_geglu) whose first projection is packed with
        # 128-row GEGLU groups; wider ones as two GEMMs with the widest tile that divides 8C
        fused = C <= getattr(self.ops, "FF_FUSED_MAX_C", 0) and hasattr(self.ops, "ff_geglu")
        t["ff_fused"] = fused
        ff_bn = (lambda n, geglu=True: 128) if fused else self.pk_bn
        t["s_n3"], t["s_ff1"], t["s_ff2"] = pk.norm(sb + ".norm3"), pk.geglu(sb + ".ff.net.0.proj", ff_bn), pk.lin(sb + ".ff.net.2")
        t["t_nin"], t["t_ffi1"], t["t_ffi2"] = pk.norm(tb + ".norm_in"), pk.geglu(tb + ".ff_in.net.0.proj", ff_bn), pk.lin(tb + ".ff_in.net.2")
        t["t_n1"], t["t_qkv"], t["t_o"] = pk.norm(tb + ".norm1"), qkv(tb + ".attn1"), pk.lin(tb + ".attn1.to_out.0")
        t["t_x"] = xattn(tb + ".attn2")
        t["t_n3"], t["t_ff1"], t["t_ff2"] = pk.norm(tb + ".norm3"), pk.geglu(tb + ".ff.net.0.proj", ff_bn), pk.lin(tb + ".ff.net.2")
        t["pos"] = pk.h(_frame_pos_embed(pk.sd, pre + ".time_pos_embed", C, self.T))
        t["alpha"] = torch.sigmoid(pk.get(pre + ".time_mixer.mix_factor").float()).item()
        return t

    # ------------------------------------------------------------------ small utilities
    def new(self, *shape, dtype=torch.float16):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    persistent = False  # set by a pipeline that replays a captured CUDA graph of the step (graph_step.StepRunner)

    def pbuf(self, key, *shape, dtype=torch.float16):
        """Per-clip conditioning tensor.  With `persistent` the same storage is handed out for the same (key, shape)
        on every clip, so a captured step graph -- whose kernel arguments are raw addresses -- stays valid from clip to
        clip; otherwise a fresh tensor (callers such as the Keypoint loop keep several views' tensors alive)."""
        if not self.persistent:
            return self.new(*shape, dtype=dtype)
        cache = self.__dict__.setdefault("_pbufs", {})
        k = (key, tuple(shape), dtype)
        if k not in cache:
            cache[k] = self.new(*shape, dtype=dtype)
        return cache[k]

    # ------------------------------------------------------------------ per-clip / per-step conditioning
    def prepare_clip(self, image_embeddings, added_time_ids):
        """Loop-invariant vectors: collapsed cross-attention outputs [B, C] and the added-time embedding."""
        ops = self.ops
        ctx = image_embeddings.reshape(image_embeddings.shape[0], -1).to(torch.float16).contiguous()  # [B, 1024]
        self.B = ctx.shape[0]
        self.xvec = []
        for k, (wv, wo, bo) in enumerate(self.xattn):
            v = self.new(self.B, wv.shape[0])
            ops.linear_small(ctx, wv, None, v, 0, 0)
            o = self.pbuf(("xvec", k), self.B, wo.shape[0])
            ops.linear_small(v, wo, bo, o, 0, 0)
            self.xvec.append(o)
        ids = added_time_ids.to(device=self.device, dtype=torch.float32).flatten().contiguous()
        e = self.new(ids.numel(), self.add_dim)
        ops.timestep_embedding(ids, e, self.add_dim)
        e = e.reshape(self.B, -1)
        (w1, b1), (w2, b2) = self.p["add_embedding"]
        hdn = self.new(self.B, w1.shape[0])
        ops.linear_small(e, w1, b1, hdn, 0, 1)
        self.aug_emb = self.pbuf("aug_emb", self.B, w2.shape[0])
        ops.linear_small(hdn, w2, b2, self.aug_emb, 0, 0)

    def time_embed(self, t_value):
        """emb = time_embedding(Timesteps(t)) + aug_emb; then every time_emb_proj(SiLU(emb)) in one call."""
        ops = self.ops
        if isinstance(t_value, torch.Tensor):   # [B] fp32 on the device, read at execution time (graph replay)
            tt = t_value
            assert tt.dtype == torch.float32 and tt.numel() == self.B
        else:
            tt = torch.full((self.B,), float(t_value), dtype=torch.float32, device=self.device)
        e = self.new(self.B, self.time_dim)
        ops.timestep_embedding(tt, e, self.time_dim)
        (w1, b1), (w2, b2) = self.p["time_embedding"]
        hdn = self.new(self.B, w1.shape[0])
        ops.linear_small(e, w1, b1, hdn, 0, 1)
        emb = self.new(self.B, w2.shape[0])
        ops.linear_small(hdn, w2, b2, emb, 0, 0)
        emb2 = self.new(self.B, w2.shape[0])
        ops.axpy_bcast(emb, self.aug_emb, emb2, 1.0)
        temb_all = self.new(self.B, self.temb.total)
        ops.linear_small(emb2, self.temb.W, self.temb.b, temb_all, 1, 0)
        return temb_all

    # ------------------------------------------------------------------ blocks
    def resblock2(self, r, x, x2, temb_all, hw, H, W, x_stats=None, want_stats=False):
        """SpatioTemporalResBlock.  GroupNorm statistics travel with the data: every GroupNorm whose input is produced by
        a GEMM of this function (norm2 <- conv1, temporal norm1 <- conv2, temporal norm2 <- temporal conv1) gets its
        (sum, sum of squares) from that GEMM's epilogue and runs its apply pass only.  `x_stats`: statistics of `x` for
        norm1 from the producer of x (previous resblock); `want_stats`: also accumulate the per-frame statistics of this
        block's output for the GroupNorm that consumes it (next resblock / transformer).  Returns (out, stats or None)."""
        ops, B, T = self.ops, self.B, self.T
        rows = x.shape[0]
        n_img = rows // hw
        cin, cout = r["cin"], r["cout"]
        fuse = cout % 64 == 0 and (cout // 32) % 2 == 0     # TMA-store epilogue with an even number of channels per group
        stats = self.new(n_img * 64, dtype=torch.float32)
        h = self.new(rows, cin)
        if x_stats is not None and x2 is None:
            ops.groupnorm(x, r["n1"][0], r["n1"][1], h, hw, r["eps"], True, x_stats, stats_ready=True)
        else:
            ops.groupnorm(x, r["n1"][0], r["n1"][1], h, hw, r["eps"], True, stats, x2=x2)
        h1 = self.new(rows, cout)
        gn = dict(gn_stats=stats, gn_rows_per_stat=hw) if fuse else {}
        ops.gemm(ops.A_CONV3X3, h, r["c1"][0], h1, N=cout, n_img=n_img, H=H, W=W, C=cin, bias=r["c1"][1],
                 rowbias=None if r["temb_sp"] is None else temb_all[:, r["temb_sp"]:r["temb_sp"] + cout],
                 rows_per_group=T * hw, **gn)
        ops.groupnorm(h1, r["n2"][0], r["n2"][1], h1n := self.new(rows, cout), hw, r["eps"], True, stats,
                      stats_ready=fuse)
        if r["sc"] is not None:
            xs = self.new(rows, cout)
            if x2 is not None:
                ops.gemm(ops.A_LINEAR, x, r["sc"][0], xs, N=cout, M=rows, K=cin, K1=x.shape[1], lda=x.shape[1],
                         lda2=x2.shape[1], a2=x2, bias=r["sc"][1])
            else:
                ops.gemm(ops.A_LINEAR, x, r["sc"][0], xs, N=cout, M=rows, K=cin, lda=cin, bias=r["sc"][1])
        else:
            assert x2 is None
            xs = x
        hs = self.new(rows, cout)
        gn = dict(gn_stats=stats, gn_rows_per_stat=T * hw) if fuse else {}
        ops.gemm(ops.A_CONV3X3, h1n, r["c2"][0], hs, N=cout, n_img=n_img, H=H, W=W, C=cout, bias=r["c2"][1], res1=xs,
                 **gn)
        # temporal resnet (GroupNorm statistics span all T frames of a batch item) + learned blend
        g = self.new(rows, cout)
        ops.groupnorm(hs, r["tn1"][0], r["tn1"][1], g, T * hw, r["teps"], True, stats, stats_ready=fuse)
        g1 = self.new(rows, cout)
        ops.gemm(ops.A_TEMPORAL3, g, r["tc1"][0], g1, N=cout, B=B, T=T, HW=hw, C=cout, bias=r["tc1"][1],
                 rowbias=None if r["temb_tp"] is None else temb_all[:, r["temb_tp"]:r["temb_tp"] + cout],
                 rows_per_group=T * hw, **gn)
        ops.groupnorm(g1, r["tn2"][0], r["tn2"][1], g, T * hw, r["teps"], True, stats, stats_ready=fuse)
        out = self.new(rows, cout)
        out_stats = None
        gn = {}
        if want_stats and fuse:
            out_stats = self.new(n_img * 64, dtype=torch.float32)
            gn = dict(gn_stats=out_stats, gn_rows_per_stat=hw)
        # alpha*hs + (1-alpha)*(hs + conv) = hs + (1-alpha)*conv
        ops.gemm(ops.A_TEMPORAL3, g, r["tc2"][0], out, N=cout, B=B, T=T, HW=hw, C=cout, bias=r["tc2"][1],
                 alpha=1.0 - r["alpha"], res1=hs, beta1=1.0, **gn)
        return out, out_stats

    def resblock(self, r, x, x2, temb_all, hw, H, W):
        return self.resblock2(r, x, x2, temb_all, hw, H, W)[0]

    def transformer(self, t, x, hw, x_stats=None):
        ops, B, T = self.ops, self.B, self.T
        rows, C, heads = x.shape[0], t["C"], t["heads"]
        frames = rows // hw
        scale = 1.0 / math.sqrt(C // heads)
        lin = ops.linear
        hn = self.new(rows, C)
        if x_stats is not None:   # accumulated by the epilogue of the GEMM that produced x (resblock(want_stats=True))
            ops.groupnorm(x, t["norm"][0], t["norm"][1], hn, hw, 1e-6, False, x_stats, stats_ready=True)
        else:
            stats = self.new(frames * 64, dtype=torch.float32)
            ops.groupnorm(x, t["norm"][0], t["norm"][1], hn, hw, 1e-6, False, stats)
        h = self.new(rows, C)
        lin(hn, t["proj_in"][0], h, bias=t["proj_in"][1])
        # ---- spatial BasicTransformerBlock
        ops.layernorm(h, t["s_n1"][0], t["s_n1"][1], hn, 1e-5)
        qkv = self.new(rows, 3 * C)
        lin(hn, t["s_qkv"], qkv)
        a = self.new(rows, C)
        d = C // heads
        if d == 64:      # SVD-XT (heads 5,10,20,20): the tcgen05 kernel
            ops.attn_spatial(qkv, a, frames, hw, heads, scale)
        else:            # any other checkpoint geometry (class default (5,10,10,20) -> d = 128): the generic kernel
            ops.attn_small(qkv, a, frames, hw, heads, d, scale)
        h2 = self.new(rows, C)
        lin(a, t["s_o"][0], h2, bias=t["s_o"][1], res1=h, rowbias=self.xvec[t["s_x"]], rows_per_group=T * hw)
        ops.layernorm(h2, t["s_n3"][0], t["s_n3"][1], hn, 1e-5)
        fused = t["ff_fused"]
        w1, b1, bn = t["s_ff1"]
        hsp = self.new(rows, C)
        if fused:   # GEGLU -> Linear in one kernel: the [rows, 4C] activation stays in tensor memory
            f = None
            ops.ff_geglu(hn, w1, b1, t["s_ff2"][0], t["s_ff2"][1], hsp, res1=h2)
        else:
            f = self.new(rows, 4 * C)
            lin(hn, w1, f, bias=b1, act=ops.ACT_GEGLU, bn=bn)
            lin(f, t["s_ff2"][0], hsp, bias=t["s_ff2"][1], res1=h2)
        # ---- temporal block on x_mix = h_spatial + frame position embedding (row order stays (b,t,p))
        hmix = self.new(rows, C)
        ops.layernorm(hsp, t["t_nin"][0], t["t_nin"][1], hn, 1e-5, add=t["pos"], rows_per_group=hw, add_period=T,
                      sum_out=hmix)
        w1, b1, bn = t["t_ffi1"]
        g = self.new(rows, C)
        if fused:
            ops.ff_geglu(hn, w1, b1, t["t_ffi2"][0], t["t_ffi2"][1], g, res1=hmix)
        else:
            lin(hn, w1, f, bias=b1, act=ops.ACT_GEGLU, bn=bn)
            lin(f, t["t_ffi2"][0], g, bias=t["t_ffi2"][1], res1=hmix)
        ops.layernorm(g, t["t_n1"][0], t["t_n1"][1], hn, 1e-5)
        lin(hn, t["t_qkv"], qkv)
        if d == 64 and T <= 32:
            ops.attn_temporal(qkv, a, B, T, hw, heads, scale)
        else:
            ops.attn_small_temporal(qkv, a, B, T, hw, heads, d, scale)
        g2 = self.new(rows, C)
        # diffusers 0.24 quirk: temporal cross-attention row i sees the context of batch item (i % B)
        assert hw % B == 0
        lin(a, t["t_o"][0], g2, bias=t["t_o"][1], res1=g, rowbias=self.xvec[t["t_x"]], rowbias_mod=B)
        ops.layernorm(g2, t["t_n3"][0], t["t_n3"][1], hn, 1e-5)
        w1, b1, bn = t["t_ff1"]
        am = t["alpha"]
        mixed = self.new(rows, C)
        # AlphaBlender: am * h_spatial + (1-am) * (ff + g2)
        if fused:
            ops.ff_geglu(hn, w1, b1, t["t_ff2"][0], t["t_ff2"][1], mixed, res1=g2, res2=hsp, alpha=1.0 - am,
                         beta1=1.0 - am, beta2=am)
        else:
            lin(hn, w1, f, bias=b1, act=ops.ACT_GEGLU, bn=bn)
            lin(f, t["t_ff2"][0], mixed, bias=t["t_ff2"][1], alpha=1.0 - am, res1=g2, beta1=1.0 - am, res2=hsp,
                beta2=am)
        out = self.new(rows, C)
        lin(mixed, t["proj_out"][0], out, bias=t["proj_out"][1], res1=x)
        return out

    def conv_im2col(self, c, x, n_img, H, W, act=0, bn=None):
        ops = self.ops
        s = c["stride"]
        Ho, Wo = (H - 1) // s + 1, (W - 1) // s + 1
        cols = self.new(n_img * Ho * Wo, c["kpad"])
        ops.im2col3x3(x, cols, n_img, H, W, c["cin"], s, c["kpad"])
        out = self.new(n_img * Ho * Wo, c["cout"])
        ops.linear(cols, c["w"], out, bias=c["b"], act=act, bn=bn)
        return out, Ho, Wo

    def down_block(self, blk, x, temb_all, hw, H, W, skips, after_each=None):
        st = None
        for j, r in enumerate(blk["res"]):
            last = j == len(blk["res"]) - 1
            # the output's GroupNorm statistics ride along when the next op is a GroupNorm over the same tensor
            x, st = self.resblock2(r, x, None, temb_all, hw, H, W, x_stats=st,
                                   want_stats=bool(blk["attn"]) or not last)
            if blk["attn"]:
                x = self.transformer(blk["attn"][j], x, hw, x_stats=st)
                st = None
            skips.append((x, hw, H, W))
        if blk["down"] is not None:
            w, b = blk["down"]
            C = x.shape[1]
            n_img = x.shape[0] // hw
            x, H, W = self.conv_im2col({"w": w, "b": b, "cin": C, "cout": w.shape[0], "kpad": 9 * C, "stride": 2}, x,
                                       n_img, H, W)
            hw = H * W
            skips.append((x, hw, H, W))
        return x, hw, H, W

    def mid_block(self, m, x, temb_all, hw, H, W):
        x, st = self.resblock2(m["res"][0], x, None, temb_all, hw, H, W, want_stats=True)
        x = self.transformer(m["attn"][0], x, hw, x_stats=st)
        return self.resblock(m["res"][1], x, None, temb_all, hw, H, W)

    def _conv_in(self, x_in, n_img, H, W):
        ops = self.ops
        w, b = self.p["conv_in"]
        cin = x_in.shape[1]
        cols = self.new(n_img * H * W, self.conv_in_kpad)
        ops.im2col3x3(x_in, cols, n_img, H, W, cin, 1, self.conv_in_kpad)
        out = self.new(n_img * H * W, w.shape[0])
        ops.linear(cols, w, out, bias=b)
        return out

    # ------------------------------------------------------------------ UNet forward
    def unet_forward(self, x_in, t_value, H, W, down_residuals, mid_residual):
        """x_in [B*T*H*W, 8] channels-last model input; residuals from the adapter; returns [B*T*H*W, 4]."""
        assert self.kind == "unet"
        ops, T = self.ops, self.T
        n_img = x_in.shape[0] // (H * W)
        temb_all = self.time_embed(t_value)
        x = self._conv_in(x_in, n_img, H, W)
        hw = H * W
        skips = [(x, hw, H, W)]
        for blk in self.p["down"]:
            x, hw, H, W = self.down_block(blk, x, temb_all, hw, H, W, skips)
        # Q1 (UNET.py:451-459): skip k ends up with multiplicity m_k of its ControlNet residual
        counts = []
        nb = len(self.p["down"])
        for i, blk in enumerate(self.p["down"]):
            counts.append(len(blk["res"]) + (1 if blk["down"] is not None else 0))
        mult, seen = [], 1
        first = [nb]  # conv_in skip is present in all nb iterations
        for i, c in enumerate(counts):
            first += [nb - i] * c
        mult = first
        new_skips = []
        for k, (s, shw, sH, sW) in enumerate(skips):
            o = self.new(*s.shape)
            ops.axpy_bcast(s, down_residuals[k], o, float(mult[k]))
            new_skips.append((o, shw, sH, sW))
        skips = new_skips
        x = self.mid_block(self.p["mid"], x, temb_all, hw, H, W)
        xm = self.new(*x.shape)
        ops.axpy_bcast(x, mid_residual, xm, 1.0)
        x = xm
        for blk in self.p["up"]:
            for j, r in enumerate(blk["res"]):
                s, shw, sH, sW = skips.pop()
                assert shw == hw
                if blk["attn"]:
                    x, st = self.resblock2(r, x, s, temb_all, hw, H, W, want_stats=True)
                    x = self.transformer(blk["attn"][j], x, hw, x_stats=st)
                else:
                    x = self.resblock(r, x, s, temb_all, hw, H, W)
            if blk["up"] is not None:
                C = x.shape[1]
                n_img = x.shape[0] // hw
                up = self.new(n_img * 4 * hw, C)
                ops.upsample2x(x, up, n_img, H, W, C)
                H, W = 2 * H, 2 * W
                hw = H * W
                w, b = blk["up"]
                x = self.new(n_img * hw, w.shape[0])
                ops.gemm(ops.A_CONV3X3, up, w, x, N=w.shape[0], n_img=n_img, H=H, W=W, C=C, bias=b)
        rows = x.shape[0]
        n_img = rows // hw
        stats = self.new(n_img * 64, dtype=torch.float32)
        hn = self.new(rows, x.shape[1])
        ops.groupnorm(x, self.p["norm_out"][0], self.p["norm_out"][1], hn, hw, 1e-5, True, stats)
        w, b = self.p["conv_out"]
        out = self.new(rows, self.out_channels)
        ops.gemm(ops.A_CONV3X3, hn, w, out, N=self.out_channels, n_img=n_img, H=H, W=W, C=x.shape[1], bias=b,
                 bn=16)
        return out

    # ------------------------------------------------------------------ adapter
    def adapter_cond_branch(self, cond_nhwc, flow_h, Himg, Wimg):
        """FCN.py:297-319 once per clip, for ONE CFG half (both halves are identical, pipeline.py:393-397).
        cond_nhwc [Himg*Wimg, 3] fp16 in [-1,1]; flow_h fp16 [T-1, 2, Himg, Wimg].
        Returns 4 tensors [T*hs*ws, C]: slot 0 = the feature itself, slots 1.. = softsplat-warped."""
        assert self.kind == "adapter"
        ops, T = self.ops, self.T
        x, H, W = cond_nhwc, Himg, Wimg
        convs = self.p["cond_convs"]
        for k, c in enumerate(convs):
            last = k == len(convs) - 1
            x, H, W = self.conv_im2col(c, x, 1, H, W, act=0 if last else ops.ACT_SILU)
        feats = [(x, H, W)]
        f = x
        for (enc, zc) in self.p["flow_enc"]:
            f, H, W = self.conv_im2col(enc, f, 1, H, W, act=ops.ACT_SILU)
            if zc is not None:
                z = self.new(H * W, zc[0].shape[0])
                ops.linear(f, zc[0], z, bias=zc[1])
            else:
                z = f
            feats.append((z, H, W))
        self.cond_feats = feats
        warped = []
        Fn = T - 1
        for lvl, (ft, hs, ws) in enumerate(feats):
            C = ft.shape[1]
            out = self.pbuf(("warped", lvl), T * hs * ws, C)
            out[: hs * ws].copy_(ft)
            acc = self.new(Fn * hs * ws * C, dtype=torch.float32)
            wsum = self.new(Fn * hs * ws, dtype=torch.float32)
            ops.softsplat_avg(ft, flow_h, acc, wsum, out[hs * ws:], Fn, hs, ws, C, Himg, Wimg)
            warped.append(out)
        self.warped = warped
        return warped

    def _add_landmarks(self, x, H):
        """x += landmark embedding at this resolution (per frame, broadcast over the CFG batch); no-op for Traj."""
        ld = getattr(self, "ldmk", None)
        if not ld:
            return x
        out = self.new(*x.shape)
        self.ops.axpy_bcast(x, ld[H], out, 1.0)
        return out

    def adapter_forward(self, x_in, t_value, H, W, conditioning_scale=1.0):
        """Trunk (FCN.py:284-376) on the hoisted warped features. Returns (12 residuals, mid residual)."""
        assert self.kind == "adapter"
        ops = self.ops
        n_img = x_in.shape[0] // (H * W)
        temb_all = self.time_embed(t_value)
        x0 = self._conv_in(x_in, n_img, H, W)
        hw = H * W
        wp = self.warped
        x = self.new(*x0.shape)
        ops.axpy_bcast(x0, wp[0], x, 1.0)          # FCN.py:328
        x = self._add_landmarks(x, H)              # Keypoint variant only (ldmk_ctrlnet.py:474)
        skips = [(x, hw, H, W)]
        count, length = 1, len(wp)
        for blk in self.p["down"]:
            x, hw, H, W = self.down_block(blk, x, temb_all, hw, H, W, skips)
            xa = self.new(*x.shape)
            ops.axpy_bcast(x, wp[min(count, length - 1)], xa, 1.0)   # FCN.py:348 (Q2: skips recorded before the add)
            x = xa
            if x.shape[1] == self.boc[0]:
                x = self._add_landmarks(x, H)      # Q14: only where the trunk has boc[0] channels (:501-504)
            count += 1
        xa = self.new(*x.shape)
        ops.axpy_bcast(x, wp[-1], xa, 1.0)         # FCN.py:354
        x = self.mid_block(self.p["mid"], xa, temb_all, hw, H, W)
        res = []
        for (s, shw, sH, sW), (zw, zb) in zip(skips, self.p["zero_down"]):
            o = self.new(s.shape[0], zw.shape[0])
            # conv1x1(x)*scale = scale*(Wx + b)
            ops.linear(s, zw, o, bias=zb, alpha=float(conditioning_scale))
            res.append(o)
        zw, zb = self.p["zero_mid"]
        mid = self.new(x.shape[0], zw.shape[0])
        ops.linear(x, zw, mid, bias=zb, alpha=float(conditioning_scale))
        return res, mid
