"""`LatentDiffusion` as far as the sampling path uses it (reference ldm/models/diffusion/ddpm.py): noise schedule
buffers (register_schedule :116-160), `apply_model_extra` (:891-905), `DiffusionWrapper.forward` (:1420-1439),
`decode_first_stage` (:705-763), `get_learned_conditioning` (:551-562), `ema_scope` (:170-183).  A plain nn.Module:
the reference's LightningModule training machinery (losses, logging, EMA updates) is never run by txt2img-* and is
out of scope.  Parameter names match the checkpoint (`model.diffusion_model.*`,
`first_stage_model.*`, `cond_stage_model.*`).
"""
from __future__ import annotations

from contextlib import contextmanager
from functools import partial

import numpy as np
import torch
import torch.nn as nn

from ...modules.diffusionmodules.openaimodel import UNetModel
from ...modules.diffusionmodules.util import make_beta_schedule
from ..autoencoder import AutoencoderKL

V1_UNET = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[4, 2, 1],
               num_res_blocks=2, channel_mult=[1, 2, 4, 4], num_heads=8, use_spatial_transformer=True,
               transformer_depth=1, context_dim=768, use_checkpoint=True, legacy=False)
V1_VAE = dict(embed_dim=4, ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                         ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0))


class DiffusionWrapper(nn.Module):
    def __init__(self, diff_model_config, conditioning_key="crossattn"):
        super().__init__()
        params = diff_model_config.get("params", diff_model_config)
        self.diffusion_model = UNetModel(**params)
        if conditioning_key != "crossattn":
            raise NotImplementedError("SD-v1 uses conditioning_key='crossattn'")
        self.conditioning_key = conditioning_key

    def forward(self, x, text_index, t, c_concat=None, c_crossattn=None, coef=None, bboxs_curr=None, step_time=None):
        cc = torch.cat(c_crossattn, 1)
        return self.diffusion_model(x, text_index, t, context=cc, coef=coef, bboxs_curr=bboxs_curr, step_time=step_time)


class LatentDiffusion(nn.Module):
    def __init__(self, unet_config=None, first_stage_config=None, cond_stage_config=None, timesteps=1000,
                 linear_start=0.00085, linear_end=0.0120, scale_factor=0.18215, conditioning_key="crossattn",
                 channels=4, image_size=64, build_first_stage=True, cond_stage_model=None, **ignored):
        super().__init__()
        self.parameterization = "eps"
        self.channels, self.image_size, self.scale_factor = channels, image_size, scale_factor
        self.use_ema = False
        self.model = DiffusionWrapper(unet_config or {"params": V1_UNET}, conditioning_key)
        self.first_stage_model = None
        if build_first_stage:
            fs = (first_stage_config or {"params": V1_VAE})
            self.first_stage_model = AutoencoderKL(**fs.get("params", fs)).eval().requires_grad_(False)
        self.cond_stage_model = cond_stage_model
        self.graph_runner = None  # graphed.GraphedModelRunner when CUDA-graph execution is enabled
        self.register_schedule(timesteps=timesteps, linear_start=linear_start, linear_end=linear_end)

    def register_schedule(self, beta_schedule="linear", timesteps=1000, linear_start=1e-4, linear_end=2e-2):
        betas = make_beta_schedule(beta_schedule, timesteps, linear_start=linear_start, linear_end=linear_end)
        alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
        alphas_cumprod_prev = np.append(1.0, alphas_cumprod[:-1])
        self.num_timesteps = int(timesteps)
        to_torch = partial(torch.tensor, dtype=torch.float32)
        self.register_buffer("betas", to_torch(betas))
        self.register_buffer("alphas_cumprod", to_torch(alphas_cumprod))
        self.register_buffer("alphas_cumprod_prev", to_torch(alphas_cumprod_prev))
        self.register_buffer("sqrt_alphas_cumprod", to_torch(np.sqrt(alphas_cumprod)))
        self.register_buffer("sqrt_one_minus_alphas_cumprod", to_torch(np.sqrt(1.0 - alphas_cumprod)))

    @property
    def device(self):
        return self.betas.device

    @contextmanager
    def ema_scope(self, context=None):
        yield None  # use_ema: False in v1-inference.yaml:18

    def get_learned_conditioning(self, c):
        if self.cond_stage_model is None:
            raise RuntimeError("no cond_stage_model attached (see ldm/modules/encoders/modules.py)")
        return self.cond_stage_model.encode(c)

    def apply_model_extra(self, x_noisy, text_index, t, cond, return_ids=False, coef=None, bboxs_curr=None,
                          step_time=None):
        if self.graph_runner is not None and self.graph_runner.active is not None:
            return self.graph_runner(x_noisy, t, coef)  # context / layout were fixed by begin_prompt()
        if not isinstance(cond, dict):
            cond = {"c_crossattn": cond if isinstance(cond, list) else [cond]}
        return self.model(x_noisy, text_index, t, **cond, coef=coef, bboxs_curr=bboxs_curr, step_time=step_time)

    def decode_first_stage(self, z, predict_cids=False, force_not_quantize=False):
        """Differentiable on purpose (the reference comments @torch.no_grad out, ddpm.py:705)."""
        return self.first_stage_model.decode(1.0 / self.scale_factor * z)
