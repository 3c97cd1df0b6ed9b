"""GPU parity of the token-major streaming kernels (csrc/sta_tokens.cu) and of the options added to the GroupNorm and
self-attention-backward entry points, against plain torch fp32 on the same fp16-rounded inputs:

  sta_add_layernorm_fwd/bwd   s = x + bias + residual (one fp16 rounding), y = LayerNorm(s)   (attention.py:274-299)
  sta_geglu_fwd/bwd           value * gelu(gate), exact erf GELU                              (attention.py:42-49)
  sta_groupnorm_* x_bias      GroupNorm(x + xb[:, :, None, None]) [+ SiLU]                    (openaimodel.py:259-268)
  sta_sattn_bwd dqkv stride   dq/dk/dv written as slices of one [b, n, 3C] buffer

Tolerances: forward |err| <= 2e-3 + 4e-3 |ref| (one fp16 rounding of an O(1) value); backward
|err| <= 3e-3 max|ref| + 1e-2 |ref|.
"""
from __future__ import annotations

import pytest
import torch
import torch.nn.functional as F

from diffusion_spacetime_attn_b200 import native, ops

ROWS_CH = [(8192, 320), (2048, 640), (512, 1280), (128, 1280), (7, 8), (33, 2048), (5, 96), (64, 1000)]


def _close(got, ref, atol, rtol, what):
    err = (got.float().cpu() - ref).abs()
    bound = atol + rtol * ref.abs()
    assert (err <= bound).all(), f"{what}: max err {err.max().item():.3e}, worst excess {(err - bound).max().item():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("rows,c", ROWS_CH, ids=[f"{r}x{c}" for r, c in ROWS_CH])
@pytest.mark.parametrize("mode", ["ln", "add_ln", "res_ln", "add"])
def test_add_layernorm_fwd_bwd(rows, c, mode):
    g = torch.Generator().manual_seed(rows + c)
    x = (torch.randn(rows, c, generator=g) * 1.3 + 0.2).half()
    res = (torch.randn(rows, c, generator=g) * 0.8).half()
    bias = 0.3 * torch.randn(c, generator=g)
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    d_y = (torch.randn(rows, c, generator=g) * 0.1).half()
    d_s = (torch.randn(rows, c, generator=g) * 0.1).half()

    xf, rf = x.float().requires_grad_(True), res.float().requires_grad_(True)
    xd, rd = x.cuda().requires_grad_(True), res.cuda().requires_grad_(True)
    gd, bd, biasd = gamma.cuda(), beta.cuda(), bias.cuda()
    if mode == "ln":
        ref_y = F.layer_norm(xf, (c,), gamma, beta, 1e-5)
        (ref_y * d_y.float()).sum().backward()
        y = ops.layer_norm(xd, gd, bd, 1e-5)
        y.backward(d_y.cuda())
        _close(y, ref_y.detach(), 2e-3, 4e-3, "y")
    elif mode in ("add_ln", "res_ln"):
        b_ref = bias if mode == "add_ln" else None
        s_ref = (xf + rf + (b_ref if b_ref is not None else 0)).half().float()  # the kernel rounds s once, then normalises
        s_lin = xf + rf + (b_ref if b_ref is not None else 0)
        s_ste = s_lin + (s_ref - s_lin).detach()
        ref_y = F.layer_norm(s_ste, (c,), gamma, beta, 1e-5)
        ((ref_y * d_y.float()).sum() + (s_ste * d_s.float()).sum()).backward()
        s, y = ops.add_layer_norm(xd, biasd if mode == "add_ln" else None, rd, gd, bd, 1e-5)
        torch.autograd.backward([s, y], [d_s.cuda(), d_y.cuda()])
        _close(s, s_ref.detach(), 1e-3, 1e-3, "s")
        _close(y, ref_y.detach(), 2e-3, 4e-3, "y")
        _close(rd.grad, rf.grad, 3e-3 * rf.grad.abs().max().item(), 1e-2, "d_residual")
    else:
        s_ref = xf + rf + bias
        (s_ref * d_s.float()).sum().backward()
        s = ops.bias_residual_add(xd, biasd, rd)
        s.backward(d_s.cuda())
        _close(s, s_ref.detach(), 2e-3, 2e-3, "s")
        _close(rd.grad, rf.grad, 1e-6, 1e-3, "d_residual")
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(xd.grad, xf.grad, 3e-3 * xf.grad.abs().max().item(), 1e-2, "d_x")


@pytest.mark.gpu
def test_add_layernorm_only_one_output_used():
    """d_s or d_y may be absent (set_materialize_grads(False)): the kernel gets a NULL d_sum / is skipped."""
    g = torch.Generator().manual_seed(0)
    x, r = torch.randn(64, 320, generator=g).half(), torch.randn(64, 320, generator=g).half()
    gamma, beta = torch.ones(320), torch.zeros(320)
    for use in ("y", "s"):
        xd = x.cuda().requires_grad_(True)
        s, y = ops.add_layer_norm(xd, None, r.cuda(), gamma.cuda(), beta.cuda(), 1e-5)
        (y if use == "y" else s).float().sum().backward()
        xf = x.float().requires_grad_(True)
        sf = (xf + r.float())
        (F.layer_norm(sf, (320,), gamma, beta, 1e-5) if use == "y" else sf).sum().backward()
        _close(xd.grad, xf.grad, 2e-3, 1e-2, f"d_x ({use} only)")


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 4096, 320), (2, 64, 1280), (1, 77, 640)], ids=str)
def test_layernorm_fork_adds_the_residual_gradient_in_the_kernel(shape):
    """layer_norm_fork: (LayerNorm(x), x) for `attn1(norm1(x)) + x` (attention.py:274); the residual-stream gradient goes
    through the kernel's d_sum input.  Checked against d/dx [sum(LN(x) dy) + sum(x dr)] in fp32 on the CPU, and each output
    used alone."""
    g = torch.Generator().manual_seed(shape[1])
    c = shape[-1]
    x = (torch.randn(shape, generator=g) * 1.3 + 0.2).half()
    gamma, beta = 1 + 0.1 * torch.randn(c, generator=g), 0.1 * torch.randn(c, generator=g)
    dy, dr = (0.1 * torch.randn(shape, generator=g)).half(), (0.1 * torch.randn(shape, generator=g)).half()
    xf = x.float().requires_grad_(True)
    ((F.layer_norm(xf, (c,), gamma, beta, 1e-5) * dy.float()).sum() + (xf * dr.float()).sum()).backward()
    xd = x.cuda().requires_grad_(True)
    y, xr = ops.layer_norm_fork(xd, gamma.cuda(), beta.cuda(), 1e-5)
    assert xr.data_ptr() == xd.data_ptr()
    torch.autograd.backward([y, xr], [dy.cuda(), dr.cuda()])
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(xd.grad, xf.grad, 3e-3 * xf.grad.abs().max().item(), 1e-2, "d_x (both)")
    xe = x.cuda().requires_grad_(True)
    ops.layer_norm_fork(xe, gamma.cuda(), beta.cuda(), 1e-5)[1].backward(dr.cuda())
    assert torch.equal(xe.grad, dr.cuda())
    xe.grad = None
    ops.layer_norm_fork(xe, gamma.cuda(), beta.cuda(), 1e-5)[0].backward(dy.cuda())
    xo = x.cuda().requires_grad_(True)
    ops.layer_norm(xo, gamma.cuda(), beta.cuda(), 1e-5).backward(dy.cuda())
    assert torch.equal(xe.grad, xo.grad)


@pytest.mark.gpu
def test_add_layernorm_rejects_bad_shapes():
    x = torch.randn(4, 12).half().cuda()
    with pytest.raises(RuntimeError, match="multiple of 8"):
        ops.layer_norm(x, torch.ones(12).cuda(), torch.zeros(12).cuda())
    with pytest.raises(RuntimeError, match="CUDA tensor"):
        ops.layer_norm(torch.randn(4, 16).half(), torch.ones(16), torch.zeros(16))
    with pytest.raises(RuntimeError, match="float32"):
        ops.layer_norm(torch.randn(4, 16).half().cuda(), torch.ones(16).half().cuda(), torch.zeros(16).half().cuda())


GEGLU_SHAPES = [(8192, 1280), (2048, 2560), (512, 5120), (128, 5120), (3, 8), (17, 1000)]


@pytest.mark.gpu
@pytest.mark.parametrize("rows,inner", GEGLU_SHAPES, ids=[f"{r}x{i}" for r, i in GEGLU_SHAPES])
def test_geglu_fwd_bwd(rows, inner):
    g = torch.Generator().manual_seed(rows + inner)
    proj = (torch.randn(rows, 2 * inner, generator=g) * 1.5).half()
    d_out = (torch.randn(rows, inner, generator=g) * 0.1).half()
    pf = proj.float().requires_grad_(True)
    a, gate = pf.chunk(2, dim=-1)
    ref = a * F.gelu(gate)
    (ref * d_out.float()).sum().backward()
    pd = proj.cuda().requires_grad_(True)
    out = ops.geglu(pd)
    out.backward(d_out.cuda())
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(out, ref.detach(), 2e-3, 2e-3, "out")
    _close(pd.grad, pf.grad, 3e-3 * pf.grad.abs().max().item(), 1e-2, "d_proj")


GN_SHAPES = [(2, 320, 64, 64), (2, 640, 32, 32), (2, 1280, 8, 8), (2, 2560, 8, 8), (1, 960, 32, 32), (3, 64, 5, 7)]


@pytest.mark.gpu
@pytest.mark.parametrize("silu", [True, False])
@pytest.mark.parametrize("shape", GN_SHAPES, ids=[str(s) for s in GN_SHAPES])
def test_groupnorm_with_channel_bias(shape, silu):
    b, c, h, w = shape
    g = torch.Generator().manual_seed(c + h)
    x = (torch.randn(b, c, h, w, generator=g) * 1.5 + 0.3).half()
    xb = (torch.randn(b, c, generator=g) * 0.7).half()
    gamma = 1 + 0.1 * torch.randn(c, generator=g)
    beta = 0.1 * torch.randn(c, generator=g)
    dy = (torch.randn(b, c, h, w, generator=g) * 0.1).half()
    xf = x.float().requires_grad_(True)
    ref = F.group_norm(xf + xb.float()[:, :, None, None], 32, gamma, beta, 1e-5)
    if silu:
        ref = F.silu(ref)
    (ref * dy.float()).sum().backward()
    xd = x.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    # the bias is handed over as a column slice of a wider table (row stride c + 64), as UNetModel does
    table = torch.zeros(b, c + 64, dtype=torch.float16, device="cuda")
    table[:, 32:32 + c] = xb.cuda()
    y = ops.group_norm_silu(xd, gamma.cuda(), beta.cuda(), 1e-5, silu, x_bias=table[:, 32:32 + c])
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(y, ref.detach(), 2e-3, 4e-3, "y")
    _close(xd.grad, xf.grad, 3e-3 * xf.grad.abs().max().item(), 1e-2, "d_x")


@pytest.mark.gpu
@pytest.mark.parametrize("b,n,h,d", [(2, 1024, 8, 40), (2, 256, 8, 80), (1, 200, 4, 160)])
def test_self_attention_fused_qkv_gradient_layout(b, n, h, d):
    """d(qkv) written in place as one [b, n, 3C] buffer == the three separate gradients of the chunked call."""
    c = h * d
    g = torch.Generator().manual_seed(n + d)
    qkv = (torch.randn(b, n, 3 * c, generator=g) * 0.7).half().cuda()
    d_out = (torch.randn(b, n, c, generator=g) * 0.1).half().cuda()
    a = qkv.clone().requires_grad_(True)
    ops.self_attention_qkv(a, h).backward(d_out)
    bq = qkv.clone().requires_grad_(True)
    q, k, v = bq.chunk(3, dim=-1)
    ops.self_attention(q, k, v, h).backward(d_out)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    assert a.grad.shape == qkv.shape and a.grad.is_contiguous()
    # dq is accumulated with fp32 reduce-adds whose order differs from launch to launch: compare with a tolerance
    diff = (a.grad.float() - bq.grad.float()).abs().max().item()
    assert diff <= 2e-3 * bq.grad.float().abs().max().item(), diff


@pytest.mark.gpu
@pytest.mark.parametrize("shape", [(2, 1280, 8, 8), (2, 640, 32, 32), (1, 512, 64, 64), (3, 16, 5, 7)], ids=str)
def test_upsample_nearest2x_fwd_bwd(shape):
    """sta_upsample2x_fwd/bwd == F.interpolate(scale_factor=2, nearest) and its autograd (exact forward; the backward
    sums four fp16 values in fp32 and rounds once)."""
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(*shape, generator=g).half()
    dy = torch.randn(shape[0], shape[1], 2 * shape[2], 2 * shape[3], generator=g).half()
    xf = x.float().requires_grad_(True)
    ref = F.interpolate(xf, scale_factor=2, mode="nearest")
    ref.backward(dy.float())
    xd = x.cuda().to(memory_format=torch.channels_last).requires_grad_(True)
    y = ops.upsample_nearest2x(xd)
    y.backward(dy.cuda())
    torch.cuda.synchronize()
    assert native.device_error() == 0
    assert torch.equal(y.float().cpu(), ref.detach())
    _close(xd.grad, xf.grad, 2e-3, 2e-3, "d_x")
