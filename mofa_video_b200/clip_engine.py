"""Native (sm_100a kernel) execution of the CLIP ViT image encoder -- SURVEY.md §8 row a2 / §8f-4.

The reference calls `self.image_encoder(image).image_embeds` once per clip
(/root/reference/MOFA-Video-Traj/pipeline/pipeline.py:133) on the module built at run_gradio.py:98
(`CLIPVisionModelWithProjection`, ViT-H/14: 32 layers, width 1280, 16 heads of 80, MLP 5120, 257 tokens, projection
1024).  `NativeClipVision` takes that module's weights and runs the same graph on the engine's kernels:

  patch embedding   14x14 stride-14 conv = im2col (K = 588 -> 592) + tcgen05 GEMM, position embedding added in the epilogue
  pre-LayerNorm     layernorm_kernel
  32 x layer        LN -> fused q|k|v GEMM (+bias) -> mofa_attn_small (d = 80) -> out_proj GEMM (+bias, +residual)
                    LN -> fc1 GEMM (+bias, GELU / quick-GELU in the epilogue) -> fc2 GEMM (+bias, +residual)
  head              post-LayerNorm of the class token, visual_projection (no bias)

0.33 TFLOP and 1.26 GB of fp16 weights per image: weight-streaming bound (~0.2 ms at HBM rate); what this buys is that the
timed clip contains no PyTorch compute at all.  Same call contract as the wrapped module: `enc(pixel_values)` ->
object with `.image_embeds` [n, projection_dim]; `.parameters()` / `.dtype` as the pipeline reads them."""
import math
from types import SimpleNamespace

import torch


class NativeClipVision:
    def __init__(self, clip_module, ops=None, device=None):
        from mofa_video_b200.models._base import resolve_backend
        self.ops, self.device, _ = resolve_backend(ops, device)
        self.module = clip_module
        cfg = clip_module.config
        self.hidden, self.heads = cfg.hidden_size, cfg.num_attention_heads
        self.layers_n, self.patch, self.image_size = cfg.num_hidden_layers, cfg.patch_size, cfg.image_size
        self.eps = cfg.layer_norm_eps
        act = cfg.hidden_act
        if act not in ("gelu", "quick_gelu"):
            raise NotImplementedError(f"CLIP hidden_act {act!r}: the GEMM epilogue implements gelu and quick_gelu")
        self.act = 6 if act == "gelu" else 7
        self.d = self.hidden // self.heads
        if self.d % 2 or self.d > 128 or self.hidden % 64:
            raise NotImplementedError(f"CLIP width {self.hidden} / {self.heads} heads is outside the kernels' range")
        sd = {k: v.detach() for k, v in clip_module.state_dict().items()}
        h = lambda t: t.to(device=self.device, dtype=torch.float16).contiguous()  # noqa: E731
        vm = "vision_model."
        w = sd[vm + "embeddings.patch_embedding.weight"]                      # [hidden, 3, p, p]
        k = 3 * self.patch * self.patch
        self.kpad = (k + 7) // 8 * 8
        wp = torch.zeros(self.hidden, self.kpad)
        wp[:, :k] = w.float().permute(0, 2, 3, 1).reshape(self.hidden, k)
        self.patch_w = h(wp)
        self.grid = self.image_size // self.patch
        self.L = self.grid * self.grid + 1
        pos = sd[vm + "embeddings.position_embedding.weight"].float()        # [L, hidden]
        assert pos.shape[0] == self.L
        self.pos_patches = h(pos[1:])
        self.cls_row = h((sd[vm + "embeddings.class_embedding"].float() + pos[0])[None])   # [1, hidden]
        self.pre_ln = (h(sd[vm + "pre_layrnorm.weight"]), h(sd[vm + "pre_layrnorm.bias"]))
        self.post_ln = (h(sd[vm + "post_layernorm.weight"]), h(sd[vm + "post_layernorm.bias"]))
        self.proj = h(sd["visual_projection.weight"])
        self.layers = []
        for i in range(self.layers_n):
            p = f"{vm}encoder.layers.{i}."
            a = p + "self_attn."
            self.layers.append({
                "ln1": (h(sd[p + "layer_norm1.weight"]), h(sd[p + "layer_norm1.bias"])),
                "qkv_w": h(torch.cat([sd[a + "q_proj.weight"], sd[a + "k_proj.weight"], sd[a + "v_proj.weight"]], 0)),
                "qkv_b": h(torch.cat([sd[a + "q_proj.bias"], sd[a + "k_proj.bias"], sd[a + "v_proj.bias"]], 0)),
                "o": (h(sd[a + "out_proj.weight"]), h(sd[a + "out_proj.bias"])),
                "ln2": (h(sd[p + "layer_norm2.weight"]), h(sd[p + "layer_norm2.bias"])),
                "fc1": (h(sd[p + "mlp.fc1.weight"]), h(sd[p + "mlp.fc1.bias"])),
                "fc2": (h(sd[p + "mlp.fc2.weight"]), h(sd[p + "mlp.fc2.bias"])),
            })

    # -- what the pipeline touches besides the call (pipeline.py:117, run_gradio.py:121-126) ----------------------
    @property
    def dtype(self):
        return torch.float16

    @property
    def config(self):
        return self.module.config

    def parameters(self):
        return iter([torch.empty(0, dtype=torch.float16, device=self.device)])

    def to(self, *a, **k):
        return self

    def requires_grad_(self, flag=False):
        return self

    def eval(self):
        return self

    def new(self, *shape):
        return torch.empty(*shape, dtype=torch.float16, device=self.device)

    def __call__(self, pixel_values):
        """pixel_values [n, 3, S, S] (any float dtype) -> .image_embeds [n, projection_dim] fp16."""
        ops = self.ops
        n, c, S, S2 = pixel_values.shape
        if c != 3 or S != self.image_size or S2 != self.image_size:
            raise ValueError(f"Input image size ({S}*{S2}) doesn't match model ({self.image_size}*{self.image_size}).")
        C, L, g = self.hidden, self.L, self.grid
        x = pixel_values.to(device=self.device, dtype=torch.float16).contiguous()
        xc = self.new(n * S * S, 3)
        ops.nchw_to_nhwc(x, xc, n, 3, S * S)
        cols = self.new(n * g * g, self.kpad)
        ops.im2col(xc, cols, n, S, S, 3, self.patch, self.patch, 0, 1, self.kpad)
        tok = self.new(n * L, C)
        for i in range(n):   # class token row, then the patches (+ position embedding in the GEMM epilogue)
            tok[i * L:i * L + 1].copy_(self.cls_row)
            ops.linear(cols[i * g * g:(i + 1) * g * g], self.patch_w, tok[i * L + 1:(i + 1) * L], res1=self.pos_patches)
        hcur = self.new(n * L, C)
        ops.layernorm(tok, self.pre_ln[0], self.pre_ln[1], hcur, self.eps)
        scale = 1.0 / math.sqrt(self.d)
        # 257 tokens are only 3 M tiles: narrow N tiles (64 weight rows) spread each weight matrix over 60-240 CTAs, so
        # the weight stream (1.26 GB per image) is pulled by the whole machine instead of 15-60 SMs
        nb = 64
        hn, qkv, att = self.new(n * L, C), self.new(n * L, 3 * C), self.new(n * L, C)
        mid = self.new(n * L, self.layers[0]["fc1"][0].shape[0])
        for ly in self.layers:
            ops.layernorm(hcur, ly["ln1"][0], ly["ln1"][1], hn, self.eps)
            ops.linear(hn, ly["qkv_w"], qkv, bias=ly["qkv_b"], bn=nb)
            ops.attn_small(qkv, att, n, L, self.heads, self.d, scale)
            h2 = self.new(n * L, C)
            ops.linear(att, ly["o"][0], h2, bias=ly["o"][1], res1=hcur, bn=nb)
            ops.layernorm(h2, ly["ln2"][0], ly["ln2"][1], hn, self.eps)
            ops.linear(hn, ly["fc1"][0], mid, bias=ly["fc1"][1], act=self.act, bn=nb)
            hcur = self.new(n * L, C)
            ops.linear(mid, ly["fc2"][0], hcur, bias=ly["fc2"][1], res1=h2, bn=nb)
        cls = hcur.view(n, L, C)[:, 0].contiguous()
        pooled = self.new(n, C)
        ops.layernorm(cls, self.post_ln[0], self.post_ln[1], pooled, self.eps)
        emb = self.new(n, self.proj.shape[0])
        ops.linear_small(pooled, self.proj, None, emb, 0, 0)
        return SimpleNamespace(image_embeds=emb, last_hidden_state=hcur.view(n, L, C))


def is_clip_vision_with_projection(m):
    """Duck test for transformers' CLIPVisionModelWithProjection (what run_gradio.py:98 builds)."""
    try:
        return (hasattr(m, "vision_model") and hasattr(m, "visual_projection") and hasattr(m, "config")
                and hasattr(m.vision_model, "encoder") and hasattr(m.config, "num_hidden_layers"))
    except Exception:  # noqa: BLE001
        return False
