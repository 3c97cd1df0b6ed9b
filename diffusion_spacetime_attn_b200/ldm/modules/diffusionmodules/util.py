"""Schedule helpers and small layers with the reference's names
(/root/reference/.../ldm/modules/diffusionmodules/util.py): make_beta_schedule :21-43, make_ddim_timesteps :46-60,
make_ddim_sampling_parameters :63-74, timestep_embedding :151-171, GroupNorm32 :214-216, noise_like :264.
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn as nn


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    if schedule != "linear":
        raise ValueError(f"schedule '{schedule}' is not used by SD-v1 inference and is not built")
    betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    return betas.numpy()


def make_ddim_timesteps(ddim_discr_method, num_ddim_timesteps, num_ddpm_timesteps, verbose=True):
    if ddim_discr_method != "uniform":
        raise NotImplementedError(f'ddim discretization "{ddim_discr_method}" is not built')
    stride = num_ddpm_timesteps // num_ddim_timesteps
    steps_out = np.arange(0, num_ddpm_timesteps, stride) + 1  # +1: final alpha values (reference :56-57)
    if verbose:
        print(f"Selected timesteps for ddim sampler: {steps_out}")
    return steps_out


def make_ddim_sampling_parameters(alphacums, ddim_timesteps, eta, verbose=True):
    alphas = alphacums[ddim_timesteps]
    alphas_prev = np.asarray([alphacums[0]] + alphacums[ddim_timesteps[:-1]].tolist())
    sigmas = eta * np.sqrt((1 - alphas_prev) / (1 - alphas) * (1 - alphas / alphas_prev))
    return sigmas, alphas, alphas_prev


def timestep_embedding(timesteps, dim, max_period=10000, repeat_only=False):
    """Sinusoidal embedding, cosines first (reference :151-171)."""
    if repeat_only:
        return timesteps[:, None].expand(-1, dim)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32, device=timesteps.device) / half)
    args = timesteps[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


class GroupNorm32(nn.GroupNorm):
    def forward(self, x):
        return super().forward(x.float()).type(x.dtype)


def normalization(channels):
    return GroupNorm32(32, channels)


def conv_nd(dims, *args, **kwargs):
    if dims != 2:
        raise ValueError("only 2-D convolutions are used by SD-v1")
    return nn.Conv2d(*args, **kwargs)


def linear(*args, **kwargs):
    return nn.Linear(*args, **kwargs)


def checkpoint(func, inputs, params, flag):
    """Gradient checkpointing with the reference's call shape (util.py:102-116), on torch.utils.checkpoint."""
    if flag and torch.is_grad_enabled():
        from torch.utils.checkpoint import checkpoint as _ckpt

        return _ckpt(func, *inputs, use_reentrant=False)
    return func(*inputs)


def noise_like(shape, device, repeat=False):
    if repeat:
        return torch.randn((1, *shape[1:]), device=device).repeat(shape[0], *((1,) * (len(shape) - 1)))
    return torch.randn(shape, device=device)
