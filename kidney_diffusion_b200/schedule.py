"""Continuous-time Gaussian diffusion schedules (host side, scalar work only).

Mirrors imagen-pytorch 1.18.5 ``GaussianDiffusionContinuousTimes`` as used by ``Imagen.p_sample_loop`` (call chain from
sample_ultra_res.py:183-195).  In sampling every batch element shares the same time, so all per-step quantities are
scalars: they are evaluated here with float32 CPU tensor arithmetic in exactly the reference's expression order and
handed to the CUDA update kernel (kd_ddpm_step) as kernel arguments.
"""
from __future__ import annotations

import math

import torch


def _log(t, eps=1e-12):
    return torch.log(t.clamp(min=eps))


def beta_linear_log_snr(t):
    return -torch.log(torch.expm1(1e-4 + 10 * (t ** 2)))


def alpha_cosine_log_snr(t, s: float = 0.008):
    return -_log((torch.cos((t + s) / (1 + s) * math.pi * 0.5) ** -2) - 1, eps=1e-5)


LOG_SNR = {"linear": beta_linear_log_snr, "cosine": alpha_cosine_log_snr}


def _f32(v):
    return torch.tensor(float(v), dtype=torch.float32) if not torch.is_tensor(v) else v.to(torch.float32)


def log_snr(schedule: str, t) -> torch.Tensor:
    return LOG_SNR[schedule](_f32(t))


def alpha_sigma(schedule: str, t):
    ls = log_snr(schedule, t)
    return torch.sqrt(torch.sigmoid(ls)), torch.sqrt(torch.sigmoid(-ls))


def sampling_times(num_timesteps: int):
    """linspace(1, 0, T + 1) pairs, as float32 tensors (get_sampling_timesteps)."""
    times = torch.linspace(1.0, 0.0, num_timesteps + 1)
    return [(times[k], times[k + 1]) for k in range(num_timesteps)]


def step_scalars(schedule: str, t, t_next) -> dict:
    """Scalars of p_mean_variance / q_posterior / p_sample for one step (all float32-rounded Python floats)."""
    t, t_next = _f32(t), _f32(t_next)
    ls, ls_next = LOG_SNR[schedule](t), LOG_SNR[schedule](t_next)
    alpha, sigma = torch.sqrt(torch.sigmoid(ls)), torch.sqrt(torch.sigmoid(-ls))
    alpha_next, sigma_next = torch.sqrt(torch.sigmoid(ls_next)), torch.sqrt(torch.sigmoid(-ls_next))
    c = -torch.expm1(ls - ls_next)
    var = (sigma_next ** 2) * c
    log_var = _log(var, eps=1e-20)
    nonzero = 1.0 - float(bool(t_next == 0))
    std = torch.tensor(nonzero, dtype=torch.float32) * (0.5 * log_var).exp()
    return dict(
        log_snr=float(ls), alpha=float(alpha), sigma=float(sigma), alpha_next=float(alpha_next), sigma_next=float(sigma_next),
        c=float(c), one_minus_c=float(1 - c), std=float(std),
    )


def renoise_scalars(schedule: str, t_from, t_to):
    """q_sample_from_to(x, from_t, to_t): x * (alpha_to / alpha) + noise * (sigma_to * alpha - sigma * alpha_to) / alpha."""
    alpha, sigma = alpha_sigma(schedule, t_from)
    alpha_to, sigma_to = alpha_sigma(schedule, t_to)
    return float(alpha_to / alpha), float(sigma_to * alpha - sigma * alpha_to), float(alpha)
