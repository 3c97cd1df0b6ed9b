"""GPU parity: fused NHWC GroupNorm(32) [+ SiLU] forward / input-gradient backward (sta_groupnorm_*) against the
reference's GroupNorm32 arithmetic (x.float() -> group_norm -> [SiLU]) evaluated by torch on the CPU in fp32 on the
same fp16-rounded inputs.  Tolerance: |err| <= 2e-3 + 4e-3 |ref| forward (one fp16 rounding), 3e-3 max|ref| + 1e-2 |ref| backward."""
from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

from diffusion_spacetime_attn_b200 import native, ops

# (batch, channels, h, w): every channel count the SD-v1 UNet normalises, at 512^2 and odd sizes
SHAPES = [(2, 320, 64, 64), (2, 640, 32, 32), (2, 1280, 16, 16), (2, 1280, 8, 8), (2, 2560, 8, 8), (2, 1920, 16, 16),
          (2, 960, 32, 32), (2, 640, 64, 64), (1, 320, 12, 12), (3, 64, 5, 7)]


@pytest.mark.gpu
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
def test_groupnorm_silu_fwd_bwd(shape, silu):
    b, c, h, w = shape
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn(b, c, h, w, generator=g) * 1.5 + 0.3).half()
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    dy = (torch.randn(b, c, h, w, generator=g) * 0.1).half()
    xf = x.float().requires_grad_(True)
    ref = F.group_norm(xf, 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    (ref * dy.float()).sum().backward()
    xd = x.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.group_norm_silu(xd, gamma.cuda(), beta.cuda(), 1e-5, silu)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    assert native.device_error() == 0
    assert y.is_contiguous(memory_format=torch.channels_last) or min(h, w, c) == 1
    err = (y.float().cpu() - ref.detach()).abs()
    assert (err <= 2e-3 + 4e-3 * ref.detach().abs()).all(), f"fwd max err {err.max().item():.3e}"
    gref = xf.grad
    gerr = (xd.grad.float().cpu() - gref).abs()
    assert (gerr <= 3e-3 * gref.abs().max() + 1e-2 * gref.abs()).all(), f"bwd max err {gerr.max().item():.3e} (ref max {gref.abs().max().item():.3e})"


@pytest.mark.gpu
def test_groupnorm_accepts_nchw_input_and_rejects_bad_channels():
    x = torch.randn(2, 320, 16, 16).half().cuda()  # NCHW-contiguous: converted once
    gamma, beta = torch.ones(320).cuda(), torch.zeros(320).cuda()
    y = ops.group_norm_silu(x, gamma, beta, 1e-5, False)
    ref = F.group_norm(x.float(), 32, gamma, beta, 1e-5)
    assert (y.float() - ref).abs().max().item() < 5e-3
    with pytest.raises(RuntimeError, match="multiple of 32"):
        ops.group_norm_silu(torch.randn(1, 48, 4, 4).half().cuda(), torch.ones(48).cuda(), torch.zeros(48).cuda())


@pytest.mark.gpu
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("shape", [(2, 320, 64, 64), (2, 1280, 16, 16), (2, 960, 64, 64), (1, 1920, 32, 32), (3, 64, 5, 7)],
                         ids=str)
def test_groupnorm_fork_adds_the_residual_gradient_in_the_kernel(shape, silu):
    """group_norm_silu_fork: y = GN(x) and x itself for the residual branch (openaimodel.py:275, attention.py:345); the
    backward adds the residual gradient inside the GroupNorm kernel (cluster and two-pass paths).  Checked against
    d/dx [ sum(GN(x) dy) + sum(x dr) ] in fp32 on the CPU; also each output used alone."""
    b, c, h, w = shape
    g = torch.Generator().manual_seed(c + h + 1)
    x = (torch.randn(b, c, h, w, generator=g) * 1.5 + 0.3).half()
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    dy = (torch.randn(b, c, h, w, generator=g) * 0.1).half()
    dr = (torch.randn(b, c, h, w, generator=g) * 0.1).half()
    xf = x.float().requires_grad_(True)
    ref = F.group_norm(xf, 32, gamma, beta, 1e-5)
    ref = F.silu(ref) if silu else ref
    ((ref * dy.float()).sum() + (xf * dr.float()).sum()).backward()
    gref = xf.grad
    xd = x.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    y, xr = ops.group_norm_silu_fork(xd, gamma.cuda(), beta.cuda(), 1e-5, silu)
    assert torch.equal(xr, xd) and xr.data_ptr() == xd.data_ptr()
    torch.autograd.backward([y, xr], [dy.cuda(), dr.cuda().to(memory_format=torch.channels_last)])
    torch.cuda.synchronize()
    assert native.device_error() == 0
    gerr = (xd.grad.float().cpu() - gref).abs()
    assert (gerr <= 3e-3 * gref.abs().max() + 1e-2 * gref.abs()).all(), f"bwd max err {gerr.max().item():.3e}"
    # the residual gradient as a channel slice of a wider NHWC tensor (what torch.cat's backward hands over): read in place
    wide = torch.zeros(b, c + 96, h, w, dtype=torch.float16, device="cuda").to(memory_format=torch.channels_last)
    wide[:, 32:32 + c] = dr.cuda()
    copies, real_nhwc = [], ops._nhwc
    ops._nhwc = lambda t: (copies.append(1) if not t.is_contiguous(memory_format=torch.channels_last) else None, real_nhwc(t))[1]
    try:
        _, stats, x_nhwc = ops.groupnorm_fwd(xd.detach(), gamma.cuda(), beta.cuda(), 1e-5, silu)
        g_strided = ops.groupnorm_bwd(x_nhwc, dy.cuda().to(memory_format=torch.channels_last), gamma.cuda(), beta.cuda(),
                                      stats, 1e-5, silu, d_res=wide[:, 32:32 + c])
    finally:
        ops._nhwc = real_nhwc
    assert not copies or min(h, w) == 1, "the channel slice was copied instead of being read in place"
    assert (g_strided.float() - xd.grad.float()).abs().max().item() <= 2e-3 * gref.abs().max().item()
    # only one of the two outputs used
    xe = x.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    _, xr = ops.group_norm_silu_fork(xe, gamma.cuda(), beta.cuda(), 1e-5, silu)
    xr.backward(dr.cuda())
    assert torch.equal(xe.grad, dr.cuda())
    xe.grad = None
    y, _ = ops.group_norm_silu_fork(xe, gamma.cuda(), beta.cuda(), 1e-5, silu)
    y.backward(dy.cuda())
    xo = x.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    ops.group_norm_silu(xo, gamma.cuda(), beta.cuda(), 1e-5, silu).backward(dy.cuda())
    assert (xe.grad.float() - xo.grad.float()).abs().max().item() <= 2e-3 * gref.abs().max().item()  # two-pass path: atomics


# decoder shapes of the SD-v1 UNet at 512^2 (B = 2): (channels of h, channels of the encoder skip, h, w) + an odd one
CAT_SHAPES = [(1280, 1280, 8, 8), (1280, 640, 16, 16), (640, 320, 32, 32), (640, 320, 64, 64), (320, 320, 64, 64), (40, 24, 5, 7)]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", CAT_SHAPES, ids=str)
def test_cat_groupnorm_reads_the_two_parts_in_place_and_splits_the_gradient(shape):
    """cat_group_norm_silu(h, skip) == (GN+SiLU(cat([h, skip], 1)), cat([h, skip], 1)) (openaimodel.py:731 + :258), cluster and
    two-pass kernels.  The concatenation must be exact; the gradient (of sum(GN dy) + sum(cat dr)) comes back as two dense
    tensors and matches fp32 autograd on the CPU."""
    c0, c1, h, w = shape
    g = torch.Generator().manual_seed(c0 + c1 + h)
    xh = (torch.randn(2, c0, h, w, generator=g) * 1.5 + 0.3).half()
    xs = (torch.randn(2, c1, h, w, generator=g) * 0.7 - 0.2).half()
    c = c0 + c1
    gamma, beta = 1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    dy = (torch.randn(2, c, h, w, generator=g) * 0.1).half()
    dr = (torch.randn(2, c, h, w, generator=g) * 0.1).half()
    fh, fs = xh.float().requires_grad_(True), xs.float().requires_grad_(True)
    cat = torch.cat([fh, fs], dim=1)
    ref = F.silu(F.group_norm(cat, 32, gamma, beta, 1e-5))
    ((ref * dy.float()).sum() + (cat * dr.float()).sum()).backward()
    dh = xh.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    ds = xs.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    y, xc = ops.cat_group_norm_silu(dh, ds, gamma.cuda(), beta.cuda(), 1e-5, True)
    assert torch.equal(xc, torch.cat([dh, ds], dim=1)), "the side output is not the exact concatenation"
    err = (y.float().cpu() - ref.detach()).abs()
    assert (err <= 2e-3 + 4e-3 * ref.detach().abs()).all(), f"fwd max err {err.max().item():.3e}"
    torch.autograd.backward([y, xc], [dy.cuda(), dr.cuda().to(memory_format=torch.channels_last)])
    torch.cuda.synchronize()
    assert native.device_error() == 0
    for got, want, name in ((dh.grad, fh.grad, "d_h"), (ds.grad, fs.grad, "d_skip")):
        assert got.is_contiguous(memory_format=torch.channels_last) or min(h, w) == 1, f"{name} is not dense NHWC"
        gerr = (got.float().cpu() - want).abs()
        assert (gerr <= 3e-3 * want.abs().max() + 1e-2 * want.abs()).all(), f"{name} max err {gerr.max().item():.3e}"
    # only the concatenation used
    dh.grad = ds.grad = None
    _, xc = ops.cat_group_norm_silu(dh, ds, gamma.cuda(), beta.cuda(), 1e-5, True)
    xc.backward(dr.cuda())
    assert torch.equal(dh.grad, dr.cuda()[:, :c0]) and torch.equal(ds.grad, dr.cuda()[:, c0:])


@pytest.mark.gpu
def test_cat_groupnorm_rejects_bad_splits():
    gamma, beta = torch.ones(64).cuda(), torch.zeros(64).cuda()
    with pytest.raises(RuntimeError, match="c_split"):
        ops.cat_group_norm_silu(torch.randn(1, 60, 4, 4).half().cuda(), torch.randn(1, 4, 4, 4).half().cuda(), gamma, beta)
    with pytest.raises(RuntimeError, match="concatenate"):
        ops.cat_group_norm_silu(torch.randn(1, 32, 4, 4).half().cuda(), torch.randn(1, 32, 8, 4).half().cuda(), gamma, beta)
