"""DDIM schedule as the reference configures it: linear betas 0.00085 -> 0.012, zero-terminal-SNR
rescale, trailing timestep spacing, v-prediction, eta 0, final_alpha_cumprod = 1
(configs/inference/default.yaml:79-88, scripts/inference.py:186-193, hallo/animate/face_animate.py:285-286, 420).

Only the per-step scalar tables live here (host side, fp64 -> fp32); the update itself is the
`hallo_b200_cfg_ddim_step` kernel.  Mirrors the constructor keywords of diffusers.DDIMScheduler that the
reference passes, so `DDIMScheduler(**sched_kwargs)` keeps working at the call site.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch


class DDIMSchedulerOutput:
    def __init__(self, prev_sample, pred_original_sample=None):
        self.prev_sample = prev_sample
        self.pred_original_sample = pred_original_sample


class DDIMScheduler:
    order = 1
    init_noise_sigma = 1.0

    def __init__(self, num_train_timesteps: int = 1000, beta_start: float = 0.00085, beta_end: float = 0.012,
                 beta_schedule: str = "linear", clip_sample: bool = False, set_alpha_to_one: bool = True,
                 steps_offset: int = 1, prediction_type: str = "v_prediction", rescale_betas_zero_snr: bool = True,
                 timestep_spacing: str = "trailing", **unused):
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, num_train_timesteps, dtype=torch.float32)
        elif beta_schedule == "scaled_linear":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, num_train_timesteps, dtype=torch.float32) ** 2
        else:
            raise NotImplementedError(beta_schedule)
        if prediction_type != "v_prediction" or clip_sample:
            raise NotImplementedError("the Hallo pipeline uses v_prediction without clipping")
        if rescale_betas_zero_snr:
            abar_sqrt = torch.cumprod(1.0 - betas, 0).sqrt()
            s0, sT = abar_sqrt[0].clone(), abar_sqrt[-1].clone()
            abar_sqrt = (abar_sqrt - sT) * s0 / (s0 - sT)
            abar = abar_sqrt ** 2
            alphas = torch.cat([abar[0:1], abar[1:] / abar[:-1]])
            betas = 1 - alphas
        self.alphas_cumprod = torch.cumprod(1.0 - betas, 0)
        self.final_alpha_cumprod = torch.tensor(1.0) if set_alpha_to_one else self.alphas_cumprod[0]
        self.num_train_timesteps = num_train_timesteps
        self.timestep_spacing = timestep_spacing
        self.steps_offset = steps_offset
        self.num_inference_steps: Optional[int] = None
        self.timesteps = torch.arange(num_train_timesteps - 1, -1, -1)

    def set_timesteps(self, num_inference_steps: int, device=None):
        self.num_inference_steps = num_inference_steps
        n, T = num_inference_steps, self.num_train_timesteps
        if self.timestep_spacing == "trailing":
            ts = np.round(np.arange(T, 0, -T / n)) - 1
        elif self.timestep_spacing == "leading":
            ts = (np.arange(0, n) * (T // n)).round()[::-1].copy() + self.steps_offset
        else:
            raise NotImplementedError(self.timestep_spacing)
        self.timesteps = torch.from_numpy(ts.astype(np.int64)).to(device)

    def scale_model_input(self, sample, timestep=None):
        return sample

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, eta: float = 0.0,
             use_clipped_model_output: bool = False, generator=None, variance_noise=None, return_dict: bool = True):
        """diffusers DDIMScheduler.step as the reference calls it (face_animate.py:420: v-prediction, eta 0):
        x0 = sqrt(a_t) x - sqrt(1-a_t) v;  eps = sqrt(a_t) v + sqrt(1-a_t) x;  prev = sqrt(a_prev) x0 + sqrt(1-a_prev) eps.
        Host/PyTorch form of what `hallo_b200_cfg_ddim_step` does on the device -- kept so the scheduler object is a
        complete stand-in at the call site (callbacks, custom loops); the engine's loop does not call it."""
        if self.num_inference_steps is None:
            raise ValueError("Number of inference steps is 'None', you need to run 'set_timesteps' after creating the scheduler")
        if eta != 0.0:
            raise NotImplementedError("eta != 0 (stochastic DDIM) is not part of the reference configuration")
        t = int(timestep)
        prev = t - self.num_train_timesteps // self.num_inference_steps
        a_t = self.alphas_cumprod[t].to(torch.float64)
        a_p = (self.alphas_cumprod[prev] if prev >= 0 else self.final_alpha_cumprod).to(torch.float64)
        sa, sb = float(a_t.sqrt()), float((1 - a_t).sqrt())
        x0 = sa * sample - sb * model_output
        eps = sa * model_output + sb * sample
        prev_sample = float(a_p.sqrt()) * x0 + float((1 - a_p).sqrt()) * eps
        if not return_dict:
            return (prev_sample,)
        return DDIMSchedulerOutput(prev_sample=prev_sample, pred_original_sample=x0)

    def coef_table(self) -> torch.Tensor:
        """[n_steps, 4] = sqrt(a_t), sqrt(1 - a_t), sqrt(a_prev), sqrt(1 - a_prev) per inference step."""
        n = self.num_inference_steps
        rows: List[List[float]] = []
        for t in self.timesteps.tolist():
            prev = t - self.num_train_timesteps // n
            a_t = self.alphas_cumprod[t].double()
            a_p = self.alphas_cumprod[prev].double() if prev >= 0 else self.final_alpha_cumprod.double()
            rows.append([float(a_t.sqrt()), float((1 - a_t).sqrt()), float(a_p.sqrt()), float((1 - a_p).sqrt())])
        return torch.tensor(rows, dtype=torch.float32)


def coef_table_of(sched) -> torch.Tensor:
    """The [n_steps, 4] table for ANY DDIM-style scheduler object that has run set_timesteps: hallo_b200's own class,
    or diffusers.DDIMScheduler as the unmodified scripts/inference.py constructs it (:186-193) -- read through the
    attributes diffusers exposes (alphas_cumprod, final_alpha_cumprod, timesteps, config.num_train_timesteps,
    config.prediction_type)."""
    if hasattr(sched, "coef_table"):
        return sched.coef_table()
    cfg = getattr(sched, "config", None)
    pred = getattr(cfg, "prediction_type", None) if cfg is not None else None
    if pred is None and isinstance(cfg, dict):
        pred = cfg.get("prediction_type")
    if pred not in (None, "v_prediction"):
        raise NotImplementedError(f"the engine's DDIM update is the v-prediction form the reference ships, got {pred!r}")
    T = int(cfg["num_train_timesteps"] if isinstance(cfg, dict) else getattr(cfg, "num_train_timesteps"))
    n = int(sched.num_inference_steps)
    rows = []
    for t in [int(x) for x in sched.timesteps]:
        prev = t - T // n
        a_t = sched.alphas_cumprod[t].double()
        a_p = sched.alphas_cumprod[prev].double() if prev >= 0 else torch.as_tensor(sched.final_alpha_cumprod).double()
        rows.append([float(a_t.sqrt()), float((1 - a_t).sqrt()), float(a_p.sqrt()), float((1 - a_p).sqrt())])
    return torch.tensor(rows, dtype=torch.float32)
