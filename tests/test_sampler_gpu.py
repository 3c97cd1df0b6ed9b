"""GPU parity of the fused sampler-step kernel (csrc/sta_sampler.cu, reference ldm/models/diffusion/plms.py:296-358) against
the same arithmetic written with torch ops in fp64, values and gradients (everything is linear: tolerance 1e-5 relative)."""
from __future__ import annotations

import pytest
import torch

from diffusion_spacetime_attn_b200 import native, ops

AB = {0: (1.0, []), 1: (3 / 2, [-1 / 2]), 2: (23 / 12, [-16 / 12, 5 / 12]), 3: (55 / 24, [-59 / 24, 37 / 24, -9 / 24])}


@pytest.mark.gpu
@pytest.mark.parametrize("k", [0, 1, 2, 3])
@pytest.mark.parametrize("B,shape", [(1, (4, 64, 64)), (2, (4, 96, 96)), (3, (4, 6, 6))])
def test_plms_step_matches_torch_expressions(k, B, shape):
    g = torch.Generator().manual_seed(10 * k + B)
    eps = torch.randn(2 * B, *shape, generator=g).cuda().requires_grad_(True)
    x = torch.randn(B, *shape, generator=g).cuda().requires_grad_(True)
    olds = [torch.randn(B, *shape, generator=g).cuda().requires_grad_(True) for _ in range(k)]
    Gx, Ge = torch.randn(B, *shape, generator=g).cuda(), torch.randn(B, *shape, generator=g).cuda()
    s, (w_e, w_old) = 7.5, AB[k]
    a_t, a_prev = 0.61, 0.67
    p_x, p_e = 1 / a_t ** 0.5, -((1 - a_t) ** 0.5) / a_t ** 0.5
    a_x, a_e = a_prev ** 0.5 * p_x, (1 - a_prev) ** 0.5 + a_prev ** 0.5 * p_e
    assert ops.plms_step_usable(eps, x, olds)
    x_prev, e_t, pred = ops.plms_step(eps, x, olds, s, w_e, w_old, a_x, a_e, p_x, p_e)
    ((x_prev * Gx).sum() + (e_t * Ge).sum()).backward()
    got = [x_prev, e_t, pred, eps.grad, x.grad] + [o.grad for o in olds]
    # ---- the reference's expressions (plms.py:308, 346-354, 321-338), fp64 ----
    e64, x64 = eps.detach().double().requires_grad_(True), x.detach().double().requires_grad_(True)
    o64 = [o.detach().double().requires_grad_(True) for o in olds]
    e_u, e_c = e64.chunk(2)
    et = e_u + s * (e_c - e_u)
    ep = w_e * et + sum(w * o for w, o in zip(w_old, o64))
    pred_x0 = (x64 - (1 - a_t) ** 0.5 * ep) / a_t ** 0.5
    xp = a_prev ** 0.5 * pred_x0 + (1 - a_prev) ** 0.5 * ep
    ((xp * Gx.double()).sum() + (et * Ge.double()).sum()).backward()
    want = [xp, et, pred_x0, e64.grad, x64.grad] + [o.grad for o in o64]
    torch.cuda.synchronize()
    assert native.device_error() == 0
    for a, b in zip(got, want):
        assert (a.double() - b.detach()).abs().max().item() <= 1e-5 * b.detach().abs().max().item() + 1e-6


@pytest.mark.gpu
def test_plms_step_gradient_with_only_one_output_used():
    """x_prev unused (last old_eps of a trajectory) or e_t unused: the missing upstream gradient is treated as zero."""
    g = torch.Generator().manual_seed(3)
    eps = torch.randn(2, 4, 8, 8, generator=g).cuda().requires_grad_(True)
    x = torch.randn(1, 4, 8, 8, generator=g).cuda().requires_grad_(True)
    x_prev, e_t, _ = ops.plms_step(eps, x, [], 7.5, 1.0, [], 0.9, -0.2, 1.1, -0.4)
    x_prev.sum().backward()
    e_u, e_c = eps.detach().chunk(2)
    assert torch.allclose(x.grad, torch.full_like(x, 0.9))
    assert torch.allclose(eps.grad, torch.cat([torch.full_like(e_u, -0.2 * (1 - 7.5)), torch.full_like(e_c, -0.2 * 7.5)]), rtol=1e-5)
