"""Shared plumbing of the engine-backed model classes: config bag, checkpoint reading (config.json +
diffusion_pytorch_model[.fp16].safetensors, the layout in /root/reference/MOFA-Video-Traj/README.md:20-38
and /root/reference/MOFA-Video-Hybrid/ckpt_tree.md:60-84), the few nn.Module-like methods the reference's
callers touch (/root/reference/MOFA-Video-Traj/run_gradio.py:119-128), and NCHW<->channels-last
conversion at the API boundary."""
import json
import os
from types import SimpleNamespace

import torch

from mofa_video_b200 import engine
from mofa_video_b200 import lib as _lib

DEFAULT_CONFIG = dict(
    sample_size=None, in_channels=8, out_channels=4,
    down_block_types=("CrossAttnDownBlockSpatioTemporal", "CrossAttnDownBlockSpatioTemporal",
                      "CrossAttnDownBlockSpatioTemporal", "DownBlockSpatioTemporal"),
    up_block_types=("UpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal", "CrossAttnUpBlockSpatioTemporal",
                    "CrossAttnUpBlockSpatioTemporal"),
    block_out_channels=(320, 640, 1280, 1280), addition_time_embed_dim=256,
    projection_class_embeddings_input_dim=768, layers_per_block=2, cross_attention_dim=1024,
    transformer_layers_per_block=1, num_attention_heads=(5, 10, 10, 20), num_frames=25,
    conditioning_channels=3, conditioning_embedding_out_channels=(16, 32, 96, 256),
)


# Where engine objects are built when the caller does not say: the CUDA library on "cuda".  The CPU host-logic tests
# (no GPU in the builder container) switch both through `default_backend(...)` so that UNMODIFIED reference code --
# e.g. T/run_gradio.py:init_models, which never passes ops= / device= -- can be executed against these classes.
_BACKEND = {"ops": None, "device": "cuda"}
PACK_FORMAT = "mofa-b200-pack-v4"   # bump when engine.Net's packed layout changes (invalidates pack caches)


class default_backend:
    """Context manager (tests only): `with default_backend(ref_ops, "cpu"): ...`."""

    def __init__(self, ops, device):
        self.new = {"ops": ops, "device": device}

    def __enter__(self):
        self.old = dict(_BACKEND)
        _BACKEND.update(self.new)
        return self

    def __exit__(self, *exc):
        _BACKEND.update(self.old)
        return False


def resolve_backend(ops=None, device=None):
    """(ops, device, is_product): the caller's choice, else the process default (the CUDA library; load() raises if it
    has not been built -- there is no CPU or PyTorch fallback)."""
    if ops is None:
        ops = _BACKEND["ops"]
    if device is None:
        device = _BACKEND["device"]
    if ops is None:
        _lib.load()
        return _lib, torch.device(device), True
    return ops, torch.device(device), False


def read_checkpoint(path, subfolder=None, variant=None):
    d = os.path.join(path, subfolder) if subfolder else path
    with open(os.path.join(d, "config.json")) as f:
        cfg = {k: v for k, v in json.load(f).items() if not k.startswith("_")}
    names = []
    if variant:
        names.append(f"diffusion_pytorch_model.{variant}.safetensors")
    names += ["diffusion_pytorch_model.safetensors", "diffusion_pytorch_model.fp16.safetensors"]
    for n in names:
        fn = os.path.join(d, n)
        if os.path.exists(fn):
            from safetensors.torch import load_file
            return cfg, load_file(fn)
    raise FileNotFoundError(f"no diffusion_pytorch_model*.safetensors under {d}")


class _Config(SimpleNamespace):
    """Attribute bag that also answers `"key" in config` / config["key"] like diffusers' FrozenDict (the reference's
    from_unet tests membership, controlnet_sdv.py:592-600)."""

    def __contains__(self, key):
        return key in self.__dict__

    def __getitem__(self, key):
        return self.__dict__[key]

    def get(self, key, default=None):
        return self.__dict__.get(key, default)


class EngineModel:
    """Base of UNetSpatioTemporalConditionControlNetModel / FlowControlNet."""

    kind = None

    def __init__(self, state_dict, config=None, device=None, ops=None):
        cfg = dict(DEFAULT_CONFIG)
        cfg.update(config or {})
        self.config = _Config(**cfg)
        self._cfg = cfg
        self._ops, self._device, _ = resolve_backend(ops, device)  # fails loudly when the CUDA library is missing
        self._sd = state_dict  # kept by reference (host memory): state_dict() / from_unet() hand these tensors on
        self.net = self._make_net(state_dict, cfg)
        self.add_embedding = SimpleNamespace(linear_1=SimpleNamespace(
            in_features=int(state_dict["add_embedding.linear_1.weight"].shape[1])))
        self._clip_key = None

    def _make_net(self, state_dict, cfg):
        return engine.Net(self.kind, state_dict, cfg, self._ops, self._device)

    # -- construction --------------------------------------------------------------------------
    @classmethod
    def from_pretrained(cls, path, subfolder=None, variant=None, low_cpu_mem_usage=True, torch_dtype=None,
                        device=None, pack_cache=None, **_ignored):
        """Reads config.json + diffusion_pytorch_model[.variant].safetensors (the reference's call, T/run_gradio.py:
        103-110).  `pack_cache` (or $MOFA_B200_PACK_CACHE): a directory where the kernel-native repack of the checkpoint
        (fused q|k|v, interleaved GEGLU rows, [Cout, ky*kx*cin] conv operands, folded mix factors ...) is kept, keyed by
        the checkpoint file's identity; later processes -- e.g. the 8 ranks of a node -- load that instead of
        re-deriving 4.4 GB of operands."""
        cache_dir = pack_cache or os.environ.get("MOFA_B200_PACK_CACHE")
        if not cache_dir:
            cfg, sd = read_checkpoint(path, subfolder, variant)
            return cls(sd, cfg, device=device)
        import hashlib
        d = os.path.join(path, subfolder) if subfolder else path
        ident = [cls.__module__, cls.__name__, PACK_FORMAT]
        for n in sorted(os.listdir(d)):
            if n == "config.json" or n.endswith(".safetensors"):
                st = os.stat(os.path.join(d, n))
                ident.append((os.path.abspath(os.path.join(d, n)), st.st_size, st.st_mtime_ns))
        ident.append(variant)
        fn = os.path.join(cache_dir, hashlib.sha1(repr(ident).encode()).hexdigest() + ".pt")
        if os.path.exists(fn):
            blob = torch.load(fn, weights_only=False)
            return cls._from_packed(blob, device=device)
        cfg, sd = read_checkpoint(path, subfolder, variant)
        model = cls(sd, cfg, device=device)
        os.makedirs(cache_dir, exist_ok=True)
        tmp = fn + f".tmp{os.getpid()}"
        torch.save({"config": model._cfg, "net": model.net.packed_state(),
                    "add_in_features": model.add_embedding.linear_1.in_features}, tmp)
        os.replace(tmp, fn)        # atomic: concurrent ranks either see the whole file or none
        return model

    @classmethod
    def _from_packed(cls, blob, device=None, ops=None):
        self = cls.__new__(cls)
        self.config = _Config(**blob["config"])
        self._cfg = blob["config"]
        self._ops, self._device, _ = resolve_backend(ops, device)
        self._sd = None
        net_cls = type(self)._net_class()
        self.net = net_cls.from_packed(blob["net"], self._ops, self._device)
        self.add_embedding = SimpleNamespace(linear_1=SimpleNamespace(in_features=int(blob["add_in_features"])))
        self._clip_key = None
        self._cond_key = None
        self._masks = None
        return self

    @classmethod
    def _net_class(cls):
        return engine.Net

    @classmethod
    def from_state_dict(cls, state_dict, config=None, device=None, ops=None):
        return cls(state_dict, config, device=device, ops=ops)

    def state_dict(self):
        """The reference-layout tensors this model was built from (not copies)."""
        if self._sd is None:
            raise RuntimeError("this model was restored from the packed-weight cache: the reference-layout state dict "
                               "is not in memory (load the checkpoint without pack_cache to get it)")
        return self._sd

    # -- nn.Module look-alikes the reference scripts call ---------------------------------------
    @property
    def dtype(self):
        return torch.float16

    @property
    def device(self):
        return self._device

    def to(self, *args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if isinstance(a, torch.dtype) and a != torch.float16:
                raise ValueError("the sm_100a engine computes in fp16 only")
        return self

    def requires_grad_(self, flag=False):
        if flag:
            raise ValueError("inference engine: no gradients")
        return self

    def eval(self):
        return self

    def parameters(self):
        return iter(())

    def enable_xformers_memory_efficient_attention(self, *a, **k):
        return self  # attention is already the fused tcgen05 kernel

    def enable_gradient_checkpointing(self):
        raise ValueError("inference engine: no gradients")

    # -- helpers --------------------------------------------------------------------------------
    def _prepare(self, encoder_hidden_states, added_time_ids):
        # identity, not address: the key holds the tensors, so a new embedding on a recycled address never matches
        key = ((encoder_hidden_states, added_time_ids), (encoder_hidden_states._version, added_time_ids._version))
        old = self._clip_key
        if (old is None or old[0][0] is not key[0][0] or old[0][1] is not key[0][1] or old[1] != key[1]):
            self.net.prepare_clip(encoder_hidden_states.to(self._device), added_time_ids.to(self._device))
            self._clip_key = key

    def _to_cl(self, x5):
        """[B, T, C, h, w] (any float dtype) -> channels-last fp16 [B*T*h*w, C]."""
        b, t, c, h, w = x5.shape
        src = x5.to(device=self._device, dtype=torch.float16).contiguous()
        out = torch.empty(b * t * h * w, c, dtype=torch.float16, device=self._device)
        self._ops.nchw_to_nhwc(src, out, b * t, c, h * w)
        return out

    def _from_cl(self, x, n_img, h, w):
        c = x.shape[1]
        out = torch.empty(n_img, c, h, w, dtype=torch.float16, device=self._device)
        self._ops.nhwc_to_nchw(x, out, n_img, c, h * w)
        return out

    @staticmethod
    def _t_value(timestep):
        return float(timestep)
