"""GPU parity against the UNMODIFIED reference executed on a B200 under torch.autocast("cuda") (its only legal mode for
backward): tests/golden/ref_gpu.npz was written by tools/ref_on_gpu.py from baseline/_ref (vendored,
untracked).  r16 = reference fp16-autocast outputs / gradients, r32 = reference fp32 forward, n16 = this library.

Tolerance rule, calibrated by the fixture itself:
    outputs    |n16 - r32| / |r32|  <=  2 * |r16 - r32| / |r32|      (measured gap: blocks 6.6e-4, UNets 1.6-1.8e-3)
    gradients  the fp32 CPU oracle's autograd is the r32 stand-in (the reference cannot back-propagate in fp32):
               err(n16 vs oracle) <= max(2 * err(r16 vs oracle), floor), floor = 5e-3 (d_x, relative L2) / 1e-2 (d_coef)
"""
from __future__ import annotations

import numpy as np
import pytest
import torch

import ref_cases as RC
from diffusion_spacetime_attn_b200 import native
from diffusion_spacetime_attn_b200.ldm.modules.attention import BasicTransformerBlock
from diffusion_spacetime_attn_b200.ldm.modules.diffusionmodules.openaimodel import UNetModel
from oracle import sta_oracle as O

FIX = RC.GOLD / "ref_gpu.npz"


def rel(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).norm() / b.norm())


def coef_err(a, b):
    a, b = torch.as_tensor(a).float().cpu(), torch.as_tensor(b).float().cpu()
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("tag,n,C", RC.BLOCKS, ids=[b[0] for b in RC.BLOCKS])
def test_block_output_and_gradients_vs_reference_on_gpu(tag, n, C):
    fx = np.load(FIX)
    x, G, context, locs, coef = RC.block_case(n, C)
    blk = BasicTransformerBlock(C, 8, C // 8, context_dim=768)
    sd = RC.block_weights({k: tuple(v.shape) for k, v in blk.state_dict().items()})
    blk.load_state_dict(sd)
    blk = blk.cuda().eval().requires_grad_(False)
    sub = RC.token_subsample(n, C)
    # ---- this library: fp16 kernels ----
    xg = x.cuda().half().requires_grad_(True)
    cg = coef.cuda().requires_grad_(True)
    blk.set_local_contexts([c.cuda() for c in locs])
    with torch.autocast("cuda"):
        y = blk(xg, context=context.cuda(), time=981, coef=cg, bboxs_curr=RC.BBOXES)
        (y.float() * G.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert native.device_error() == 0
    # ---- fp32 oracle autograd (r32 stand-in for the gradients) ----
    xo, co = x.half().float().requires_grad_(True), coef.clone().requires_grad_(True)
    yo = O.transformer_block(xo, context, co, [torch.cat([RC.uncond(), c]) for c in locs], O.flat_masks(RC.BBOXES, n), sd, "", 8)
    (yo * G).sum().backward()

    y32, y16 = fx[f"{tag}_y32"], fx[f"{tag}_y16"]
    gap = rel(y16, y32)
    ours = rel(y[:, sub], y32)
    assert ours <= 2 * gap and ours < 5e-3, f"{tag}: |n16-r32|/|r32| = {ours:.3e}, reference fp16 gap {gap:.3e}"
    e_ref, e_ours = rel(fx[f"{tag}_dx16"], xo.grad[:, sub]), rel(xg.grad[:, sub], xo.grad[:, sub])
    assert e_ours <= max(2 * e_ref, 5e-3), f"{tag}: d_x error {e_ours:.3e} vs reference-fp16 error {e_ref:.3e}"
    c_ref, c_ours = coef_err(fx[f"{tag}_dcoef16"], co.grad), coef_err(cg.grad, co.grad)
    assert c_ours <= max(2 * c_ref, 1e-2), f"{tag}: d_coef error {c_ours:.3e} vs reference-fp16 error {c_ref:.3e}"
    # and directly against the reference's own fp16 gradients
    assert coef_err(cg.grad, fx[f"{tag}_dcoef16"]) < 2e-2
    assert rel(xg.grad[:, sub], fx[f"{tag}_dx16"]) < 1e-2


@pytest.mark.gpu
@pytest.mark.parametrize("tag,cfg,latent,seed,t", RC.UNETS, ids=[u[0] for u in RC.UNETS])
def test_unet_output_and_gradients_vs_reference_on_gpu(tag, cfg, latent, seed, t):
    """One UNet evaluation (batch 2 CFG, 2 objects): eps, dL/dx and dL/dalpha.  `unet_full` is the SD-v1 geometry of
    BASELINE.json configs[1] (64x64 latent, 859.5 M parameters)."""
    fx = np.load(FIX)
    x, G, context, locs, coef = RC.unet_case(latent)
    ocfg = O.UNetConfig(**cfg)
    sd = O.seeded_state_dict(O.unet_param_shapes(ocfg), seed)
    kw = {k: v for k, v in cfg.items() if k in ("attention_resolutions", "num_res_blocks", "channel_mult")}
    m = UNetModel(**kw)
    m.load_state_dict(sd, strict=True)
    m = m.cuda().eval().requires_grad_(False)
    m.set_checkpointing(True)
    m.set_local_contexts([c.cuda() for c in locs], first_timestep=981)
    xg, cg = x.cuda().requires_grad_(True), coef.cuda().requires_grad_(True)
    tt = torch.full((2,), t, dtype=torch.long, device="cuda")
    with torch.autocast("cuda"):
        y = m(xg, 0, tt, context=context.cuda(), coef=cg, bboxs_curr=RC.BBOXES, step_time=t)
        (y.float() * G.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert native.device_error() == 0
    del m
    # fp32 oracle autograd on the host (full UNet: a few seconds)
    xo, co = x.clone().requires_grad_(True), coef.clone().requires_grad_(True)
    yo = O.unet_forward(xo, torch.full((2,), t, dtype=torch.long), context, co, RC.BBOXES,
                        [torch.cat([RC.uncond(), c]) for c in locs], sd, ocfg)
    (yo * G).sum().backward()

    y32, y16 = fx[f"{tag}_y32"], fx[f"{tag}_y16"]
    gap, ours = rel(y16, y32), rel(y, y32)
    assert ours <= 2 * gap and ours < 5e-3, f"{tag}: |n16-r32|/|r32| = {ours:.3e}, reference fp16 gap {gap:.3e}"
    assert rel(yo, y32) < 5e-4  # the oracle IS the reference forward (fp32 CPU vs fp32 GPU)
    e_ref, e_ours = rel(fx[f"{tag}_dx16"], xo.grad), rel(xg.grad, xo.grad)
    assert e_ours <= max(2 * e_ref, 5e-3), f"{tag}: d_x error {e_ours:.3e} vs reference-fp16 error {e_ref:.3e}"
    c_ref, c_ours = coef_err(fx[f"{tag}_dcoef16"], co.grad), coef_err(cg.grad, co.grad)
    assert c_ours <= max(2 * c_ref, 1e-2), f"{tag}: d_alpha error {c_ours:.3e} vs reference-fp16 error {c_ref:.3e}"
    assert coef_err(cg.grad, fx[f"{tag}_dcoef16"]) < 2e-2
