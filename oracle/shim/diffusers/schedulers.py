"""diffusers 0.21.4 DDIMScheduler / DDPMScheduler restated for the options the reference uses
(pl_trainer/inference/inference.py:35-50): scaled_linear betas, epsilon prediction, clip_sample=False,
timestep_spacing='leading', eta=0, variance_type='fixed_small'."""
from dataclasses import dataclass

import numpy as np
import torch


@dataclass
class SchedulerOutput:
    prev_sample: torch.Tensor
    pred_original_sample: torch.Tensor


class _Cfg(dict):
    __getattr__ = dict.__getitem__


def _betas(beta_start, beta_end, beta_schedule, n):
    if beta_schedule == "scaled_linear":
        return torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
    if beta_schedule == "linear":
        return torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
    raise NotImplementedError(beta_schedule)


class DDIMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 clip_sample=True, set_alpha_to_one=True, steps_offset=0, prediction_type="epsilon", **kw):
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, steps_offset=steps_offset,
                           clip_sample=clip_sample, prediction_type=prediction_type)
        self.betas = _betas(beta_start, beta_end, beta_schedule, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy().astype(np.int64))

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        ts += self.config.steps_offset
        self.timesteps = torch.from_numpy(ts).to(device)

    def step(self, model_output, timestep, sample, eta=0.0, **kw):
        assert eta == 0.0 and not self.config.clip_sample and self.config.prediction_type == "epsilon"
        prev_timestep = timestep - self.config.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[timestep]
        alpha_prod_t_prev = self.alphas_cumprod[prev_timestep] if prev_timestep >= 0 else self.final_alpha_cumprod
        beta_prod_t = 1 - alpha_prod_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_epsilon = model_output
        pred_sample_direction = (1 - alpha_prod_t_prev) ** 0.5 * pred_epsilon
        prev_sample = alpha_prod_t_prev ** 0.5 * pred_original_sample + pred_sample_direction
        return SchedulerOutput(prev_sample=prev_sample, pred_original_sample=pred_original_sample)


class DDPMScheduler:
    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 variance_type="fixed_small", clip_sample=True, prediction_type="epsilon", **kw):
        self.config = _Cfg(num_train_timesteps=num_train_timesteps, clip_sample=clip_sample,
                           prediction_type=prediction_type, variance_type=variance_type)
        self.betas = _betas(beta_start, beta_end, beta_schedule, num_train_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.one = torch.tensor(1.0)
        self.init_noise_sigma = 1.0
        self.num_inference_steps = None
        self.timesteps = torch.from_numpy(np.arange(0, num_train_timesteps)[::-1].copy())

    def set_timesteps(self, num_inference_steps, device=None):
        self.num_inference_steps = num_inference_steps
        step_ratio = self.config.num_train_timesteps // num_inference_steps
        ts = (np.arange(0, num_inference_steps) * step_ratio).round()[::-1].copy().astype(np.int64)
        self.timesteps = torch.from_numpy(ts).to(device)

    def add_noise(self, original_samples, noise, timesteps):
        ac = self.alphas_cumprod.to(device=original_samples.device, dtype=original_samples.dtype)
        sa = ac[timesteps] ** 0.5
        sb = (1 - ac[timesteps]) ** 0.5
        while sa.dim() < original_samples.dim():
            sa, sb = sa.unsqueeze(-1), sb.unsqueeze(-1)
        return sa * original_samples + sb * noise

    def step(self, model_output, timestep, sample, generator=None, **kw):
        assert not self.config.clip_sample and self.config.prediction_type == "epsilon"
        t = timestep
        prev_t = t - self.config.num_train_timesteps // self.num_inference_steps
        alpha_prod_t = self.alphas_cumprod[t]
        alpha_prod_t_prev = self.alphas_cumprod[prev_t] if prev_t >= 0 else self.one
        beta_prod_t = 1 - alpha_prod_t
        beta_prod_t_prev = 1 - alpha_prod_t_prev
        current_alpha_t = alpha_prod_t / alpha_prod_t_prev
        current_beta_t = 1 - current_alpha_t
        pred_original_sample = (sample - beta_prod_t ** 0.5 * model_output) / alpha_prod_t ** 0.5
        pred_original_sample_coeff = (alpha_prod_t_prev ** 0.5 * current_beta_t) / beta_prod_t
        current_sample_coeff = current_alpha_t ** 0.5 * beta_prod_t_prev / beta_prod_t
        pred_prev_sample = pred_original_sample_coeff * pred_original_sample + current_sample_coeff * sample
        if t > 0:
            noise = torch.randn(model_output.shape, generator=generator, dtype=model_output.dtype).to(
                model_output.device) if model_output.device.type == "cpu" else torch.randn(
                model_output.shape, generator=generator, device=model_output.device, dtype=model_output.dtype)
            variance = torch.clamp((1 - alpha_prod_t_prev) / (1 - alpha_prod_t) * current_beta_t, min=1e-20)
            pred_prev_sample = pred_prev_sample + variance ** 0.5 * noise
        return SchedulerOutput(prev_sample=pred_prev_sample, pred_original_sample=pred_original_sample)


class PNDMScheduler:
    def __init__(self, *a, **k):
        raise NotImplementedError("PNDM is imported but never used by the reference")
