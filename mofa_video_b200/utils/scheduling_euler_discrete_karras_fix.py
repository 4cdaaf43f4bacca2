"""EulerDiscreteScheduler with the reference's interface
(/root/reference/MOFA-Video-Traj/utils/scheduling_euler_discrete_karras_fix.py:133-556).

The reference pipeline touches: set_timesteps(n, device), .timesteps, .sigmas, .init_noise_sigma,
.scale_model_input(x, t), .step(pred, t, x).prev_sample, .order and ._step_index (the Keypoint loop
rewinds it).  Those are kept.  The sigma ladder is host arithmetic (numpy, 25 numbers); the per-element
work of scale_model_input / CFG / step is fused on the device in mofa_cfg_euler_step, which the engine
pipeline calls with (sigma, sigma_next) taken from this object, so the Python-side `step` below exists for
API compatibility and for callers that drive the scheduler themselves.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


@dataclass
class EulerDiscreteSchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: Optional[torch.Tensor] = None


class _Config(dict):
    __getattr__ = dict.__getitem__


class EulerDiscreteScheduler:
    order = 1
    # SVD-XT-1.1 scheduler_config.json values are the defaults here (the class defaults of the reference are
    # the generic diffusers ones; from_pretrained overrides them with exactly these numbers)
    _defaults = dict(num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
                     trained_betas=None, prediction_type="v_prediction", interpolation_type="linear",
                     use_karras_sigmas=True, sigma_min=0.002, sigma_max=700.0, timestep_spacing="leading",
                     timestep_type="continuous", steps_offset=1, rescale_betas_zero_snr=False)

    def __init__(self, **kwargs):
        cfg = dict(self._defaults)
        unknown = set(kwargs) - set(cfg)
        if unknown:
            raise TypeError(f"unexpected scheduler config keys: {sorted(unknown)}")
        cfg.update(kwargs)
        self.config = _Config(cfg)
        c = self.config
        n = c.num_train_timesteps
        if c.trained_betas is not None:
            betas = np.asarray(c.trained_betas, dtype=np.float32)
        elif c.beta_schedule == "linear":
            betas = np.linspace(c.beta_start, c.beta_end, n, dtype=np.float32)
        elif c.beta_schedule == "scaled_linear":
            betas = torch.linspace(c.beta_start ** 0.5, c.beta_end ** 0.5, n, dtype=torch.float32).pow(2).numpy()
        else:
            raise NotImplementedError(f"{c.beta_schedule} is not implemented for {self.__class__}")
        self.betas = torch.from_numpy(np.asarray(betas, dtype=np.float32))
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.use_karras_sigmas = bool(c.use_karras_sigmas)
        self.is_scale_input_called = False
        self._step_index = None
        self.num_inference_steps = None
        ladder = self._train_sigmas()[::-1].copy()
        if self.use_karras_sigmas:
            ladder = self._karras(ladder, n)
        self._install(ladder, None)

    @classmethod
    def from_config(cls, config):
        return cls(**{k: v for k, v in dict(config).items() if k in cls._defaults})

    # ------------------------------------------------------------------ ladder construction (host)
    def _train_sigmas(self):
        ac = self.alphas_cumprod
        return (((1 - ac) / ac) ** 0.5).numpy()

    def _karras(self, in_sigmas, count):
        c = self.config
        lo = c.sigma_min if c.get("sigma_min") is not None else float(in_sigmas[-1])
        hi = c.sigma_max if c.get("sigma_max") is not None else float(in_sigmas[0])
        rho = 7.0
        ramp = np.linspace(0, 1, count)
        return (hi ** (1 / rho) + ramp * (lo ** (1 / rho) - hi ** (1 / rho))) ** rho

    def _install(self, sigmas_np, device):
        sig = torch.from_numpy(np.asarray(sigmas_np)).to(dtype=torch.float32, device=device)
        c = self.config
        if c.timestep_type == "continuous" and c.prediction_type == "v_prediction":
            self.timesteps = torch.Tensor([0.25 * s.log() for s in sig]).to(device=device)
        else:
            raise NotImplementedError("only timestep_type='continuous' with v_prediction (the SVD setting)")
        self.sigmas = torch.cat([sig, torch.zeros(1, device=sig.device)])
        self._sigmas_host = [float(v) for v in self.sigmas.cpu()]
        self._timesteps_host = [float(v) for v in self.timesteps.cpu()]

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        c = self.config
        n = c.num_train_timesteps
        if c.timestep_spacing == "leading":
            ratio = n // num_inference_steps
            ts = (np.arange(0, num_inference_steps) * ratio).round()[::-1].copy().astype(np.float32) + c.steps_offset
        elif c.timestep_spacing == "linspace":
            ts = np.linspace(0, n - 1, num_inference_steps, dtype=np.float32)[::-1].copy()
        elif c.timestep_spacing == "trailing":
            ts = np.arange(n, 0, -n / num_inference_steps).round().copy().astype(np.float32) - 1
        else:
            raise ValueError(f"{c.timestep_spacing} is not supported")
        train = self._train_sigmas()
        if c.interpolation_type != "linear":
            raise ValueError("only interpolation_type='linear'")
        sig = np.interp(ts, np.arange(0, len(train)), train)
        if self.use_karras_sigmas:
            sig = self._karras(sig, num_inference_steps)
        self._install(sig, device)
        self._step_index = None

    # ------------------------------------------------------------------ reference-facing accessors
    @property
    def init_noise_sigma(self):
        top = self.sigmas.max()
        if self.config.timestep_spacing in ("linspace", "trailing"):
            return top
        return (top ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _init_step_index(self, timestep):
        t = float(timestep)
        hits = [i for i, v in enumerate(self._timesteps_host) if v == t]
        if not hits:  # tolerate dtype round trips of the timestep value
            hits = [int(np.argmin([abs(v - t) for v in self._timesteps_host]))]
        self._step_index = hits[1] if len(hits) > 1 else hits[0]

    def sigma_pair(self, index=None):
        """(sigma_i, sigma_{i+1}) as host floats -- no device sync (the reference does a .nonzero().item())."""
        i = self._step_index if index is None else index
        return self._sigmas_host[i], self._sigmas_host[i + 1]

    def scale_model_input(self, sample, timestep):
        if self._step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self._step_index]
        self.is_scale_input_called = True
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, timestep, sample, return_dict=True, **_ignored):
        if isinstance(timestep, (int, torch.IntTensor, torch.LongTensor)):
            raise ValueError("pass one of scheduler.timesteps, not an integer index")
        if self._step_index is None:
            self._init_step_index(timestep)
        x = sample.to(torch.float32)
        sigma = self.sigmas[self._step_index].to(x.device)
        nxt = self.sigmas[self._step_index + 1].to(x.device)
        if self.config.prediction_type == "v_prediction":
            x0 = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (x / (sigma ** 2 + 1))
        elif self.config.prediction_type == "epsilon":
            x0 = x - sigma * model_output
        else:
            x0 = model_output
        prev = (x + (x - x0) / sigma * (nxt - sigma)).to(model_output.dtype)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return EulerDiscreteSchedulerOutput(prev_sample=prev, pred_original_sample=x0)
