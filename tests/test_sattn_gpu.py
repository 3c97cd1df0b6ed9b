"""GPU parity: sta_sattn_fwd / sta_sattn_bwd (through the C ABI) against the CPU oracle's attention_core.

Tolerances (fp16 storage, fp32 softmax/accumulate — the reference's autocast numerics): outputs are compared
with the oracle evaluated in fp32 ON THE SAME fp16-ROUNDED INPUTS; |err| <= 2e-3 + 1e-2*|ref| element-wise.
"""
from __future__ import annotations

import pytest
import torch

from diffusion_spacetime_attn_b200 import native, ops
from oracle import sta_oracle as O

# (batch, n, heads, head_dim): the four SD-v1 geometries at 512^2 plus ragged / tiny sizes (768^2 gives 144, 576)
SHAPES = [
    (2, 64, 8, 160),
    (2, 256, 8, 160),
    (2, 1024, 8, 80),
    (2, 4096, 8, 40),
    (1, 144, 8, 160),
    (1, 576, 8, 160),
    (1, 2304, 8, 80),
    (1, 100, 2, 40),
    (1, 129, 1, 80),
    (3, 300, 2, 40),
]


def _inputs(b, n, h, d, seed=0, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    q = (torch.randn(b, n, h * d, generator=g) * scale).half()
    k = (torch.randn(b, n, h * d, generator=g) * scale).half()
    v = torch.randn(b, n, h * d, generator=g).half()
    return q, k, v


def _close(got, ref, atol=2e-3, rtol=1e-2):
    err = (got.float().cpu() - ref).abs()
    bound = atol + rtol * ref.abs()
    bad = (err > bound).sum().item()
    assert bad == 0, f"{bad} elements out of tolerance; max abs err {err.max().item():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
def test_sattn_fwd_matches_oracle(shape):
    b, n, h, d = shape
    q, k, v = _inputs(b, n, h, d)
    ref, ref_lse = O.attention_core(q.float(), k.float(), v.float(), h, return_lse=True)
    out, lse = ops.sattn_fwd(q.cuda(), k.cuda(), v.cuda(), h)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(out, ref)
    assert (lse.cpu() - ref_lse).abs().max().item() < 2e-3


@pytest.mark.gpu
def test_sattn_fwd_large_logits_rescale_path():
    """Scores spread over ~+-60 so the running maximum moves by more than 2^8 between tiles (rescale path)."""
    b, n, h, d = 1, 1024, 2, 40
    q, k, v = _inputs(b, n, h, d, seed=3, scale=3.0)
    ref = O.attention_core(q.float(), k.float(), v.float(), h)
    out, _ = ops.sattn_fwd(q.cuda(), k.cuda(), v.cuda(), h)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(out, ref, atol=4e-3, rtol=2e-2)


@pytest.mark.gpu
def test_sattn_fwd_strided_qkv_views():
    """q/k/v given as column slices of one fused [b, n, 3C] projection."""
    b, n, h, d = 2, 256, 8, 40
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(b, n, 3 * h * d, generator=g).half()
    q, k, v = qkv.chunk(3, dim=-1)
    ref = O.attention_core(q.float(), k.float(), v.float(), h)
    dq, dk, dv = qkv.cuda().chunk(3, dim=-1)
    out, _ = ops.sattn_fwd(dq, dk, dv, h)
    torch.cuda.synchronize()
    _close(out, ref)


@pytest.mark.gpu
def test_sattn_rejects_unsupported_head_dim_and_cpu_tensors():
    q = torch.zeros(1, 16, 64, dtype=torch.float16, device="cuda")
    with pytest.raises(RuntimeError, match="head_dim"):
        ops.sattn_fwd(q, q, q, heads=1)  # d = 64 is not built: an error, never a fallback
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.sattn_fwd(q.cpu(), q.cpu(), q.cpu(), heads=1)


BWD_SHAPES = [
    (2, 64, 8, 160),
    (2, 256, 8, 160),
    (2, 1024, 8, 80),
    (1, 4096, 8, 40),
    (1, 576, 4, 160),
    (1, 300, 2, 40),
    (1, 129, 1, 80),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", BWD_SHAPES, ids=[str(s) for s in BWD_SHAPES])
def test_sattn_bwd_matches_oracle_autograd(shape):
    """dQ, dK, dV against torch autograd through the oracle (fp32, same fp16-rounded inputs).
    Tolerance 3e-3 * max|ref| + 2e-2 * |ref| element-wise (P and dS are fp16 MMA operands)."""
    b, n, h, d = shape
    q, k, v = _inputs(b, n, h, d, seed=9)
    g = torch.Generator().manual_seed(10)
    d_out = (torch.randn(b, n, h * d, generator=g) * 0.1).half()
    qf, kf, vf = (t.float().requires_grad_(True) for t in (q, k, v))
    ref = O.attention_core(qf, kf, vf, h)
    (ref * d_out.float()).sum().backward()
    dq, dk, dv = q.cuda(), k.cuda(), v.cuda()
    out, lse = ops.sattn_fwd(dq, dk, dv, h)
    g_q, g_k, g_v = ops.sattn_bwd(dq, dk, dv, out, lse, d_out.cuda(), h)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    for got, want, nm in ((g_q, qf.grad, "dq"), (g_k, kf.grad, "dk"), (g_v, vf.grad, "dv")):
        err = (got.float().cpu() - want).abs()
        bound = 3e-3 * want.abs().max() + 2e-2 * want.abs()
        bad = (err > bound).sum().item()
        assert bad == 0, f"{nm}: {bad} elements out of tolerance, max err {err.max().item():.3e} (ref max {want.abs().max().item():.3e})"


@pytest.mark.gpu
def test_autograd_functions_route_through_kernels():
    b, n, h, d = 2, 256, 8, 40
    q, k, v = (t.cuda().requires_grad_(True) for t in _inputs(b, n, h, d, seed=12))
    before = dict(ops.LAUNCHES)
    out = ops.self_attention(q, k, v, h)
    out.float().square().sum().backward()
    torch.cuda.synchronize()
    assert ops.LAUNCHES["sattn_fwd"] == before["sattn_fwd"] + 1
    assert ops.LAUNCHES["sattn_bwd"] == before["sattn_bwd"] + 3
    assert q.grad is not None and torch.isfinite(q.grad.float()).all()
