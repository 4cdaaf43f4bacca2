"""ORACLE (test infrastructure): EulerDiscreteScheduler restated from
/root/reference/MOFA-Video-Traj/utils/scheduling_euler_discrete_karras_fix.py
(__init__ :178-246, init_noise_sigma :248-255, scale_model_input :264-288, set_timesteps :290-350,
_convert_to_karras :376-399, _init_step_index :401-416, step :418-528).

PINNED: tests/golden/scheduler_*.pt are produced by importing that reference file itself (with import
stubs for the absent diffusers base classes) in oracle/make_goldens.py; tests/test_oracle.py compares.
"""
import numpy as np
import torch

SVD_XT_SCHEDULER_CONFIG = dict(   # SVD-XT-1.1 scheduler/scheduler_config.json
    num_train_timesteps=1000, beta_start=0.00085, beta_end=0.012, beta_schedule="scaled_linear",
    prediction_type="v_prediction", interpolation_type="linear", use_karras_sigmas=True, sigma_min=0.002,
    sigma_max=700.0, timestep_spacing="leading", timestep_type="continuous", steps_offset=1,
)


class EulerDiscreteScheduler:
    order = 1

    def __init__(self, **cfg):
        c = dict(SVD_XT_SCHEDULER_CONFIG)
        c.update(cfg)
        self.config = type("Config", (), c)()
        assert c["beta_schedule"] == "scaled_linear"
        self.betas = torch.linspace(c["beta_start"] ** 0.5, c["beta_end"] ** 0.5, c["num_train_timesteps"],
                                    dtype=torch.float32) ** 2                       # :199-201
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.use_karras_sigmas = c["use_karras_sigmas"]
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()[::-1].copy()
        if self.use_karras_sigmas:
            sigmas = self._convert_to_karras(sigmas, c["num_train_timesteps"])
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32)
        self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas])             # continuous + v_prediction :237-238
        self.sigmas = torch.cat([sigmas, torch.zeros(1)])
        self.num_inference_steps = None
        self.is_scale_input_called = False
        self._step_index = None

    @property
    def init_noise_sigma(self):                                                     # :248-255
        max_sigma = self.sigmas.max()
        if self.config.timestep_spacing in ["linspace", "trailing"]:
            return max_sigma
        return (max_sigma ** 2 + 1) ** 0.5

    @property
    def step_index(self):
        return self._step_index

    def _convert_to_karras(self, in_sigmas, num_inference_steps):                   # :376-399
        sigma_min = self.config.sigma_min if self.config.sigma_min is not None else in_sigmas[-1].item()
        sigma_max = self.config.sigma_max if self.config.sigma_max is not None else in_sigmas[0].item()
        rho = 7.0
        ramp = np.linspace(0, 1, num_inference_steps)
        min_inv_rho = sigma_min ** (1 / rho)
        max_inv_rho = sigma_max ** (1 / rho)
        return (max_inv_rho + ramp * (min_inv_rho - max_inv_rho)) ** rho

    def set_timesteps(self, num_inference_steps, device=None):                      # :290-350
        self.num_inference_steps = num_inference_steps
        c = self.config
        assert c.timestep_spacing == "leading" and c.interpolation_type == "linear"
        step_ratio = c.num_train_timesteps // num_inference_steps
        timesteps = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.float32)
        timesteps += c.steps_offset
        sigmas = (((1 - self.alphas_cumprod) / self.alphas_cumprod) ** 0.5).numpy()
        sigmas = np.interp(timesteps, np.arange(0, len(sigmas)), sigmas)
        if self.use_karras_sigmas:
            sigmas = self._convert_to_karras(sigmas, num_inference_steps)
        sigmas = torch.from_numpy(sigmas).to(dtype=torch.float32, device=device)
        self.timesteps = torch.Tensor([0.25 * s.log() for s in sigmas]).to(device=device)
        self.sigmas = torch.cat([sigmas, torch.zeros(1, device=sigmas.device)])
        self._step_index = None

    def _init_step_index(self, timestep):                                           # :401-416
        if isinstance(timestep, torch.Tensor):
            timestep = timestep.to(self.timesteps.device)
        cand = (self.timesteps == timestep).nonzero()
        self._step_index = (cand[1] if len(cand) > 1 else cand[0]).item()

    def scale_model_input(self, sample, timestep):                                  # :264-288
        if self.step_index is None:
            self._init_step_index(timestep)
        sigma = self.sigmas[self.step_index]
        self.is_scale_input_called = True
        return sample / ((sigma ** 2 + 1) ** 0.5)

    def step(self, model_output, timestep, sample):                                 # :418-528 (gamma = 0)
        if self.step_index is None:
            self._init_step_index(timestep)
        sample = sample.to(torch.float32)
        sigma = self.sigmas[self.step_index]
        torch.randn(model_output.shape, dtype=model_output.dtype)                   # drawn and discarded (:487-489, Q18)
        sigma_hat = sigma
        pred_original_sample = model_output * (-sigma / (sigma ** 2 + 1) ** 0.5) + (sample / (sigma ** 2 + 1))
        derivative = (sample - pred_original_sample) / sigma_hat
        dt = self.sigmas[self.step_index + 1] - sigma_hat
        prev_sample = (sample + derivative * dt).to(model_output.dtype)
        self._step_index += 1
        return prev_sample
