"""GPU parity: sta_xattn_fwd / sta_xattn_bwd (through the C ABI) against the CPU oracle's
dual_cross_attention_core (the pre-`to_out` restatement of ldm/modules/attention.py:278-294).

Tolerance: oracle in fp32 on the same fp16-rounded inputs; |err| <= 3e-3 + 1e-2*|ref| element-wise.
"""
from __future__ import annotations

import pytest
import torch

from diffusion_spacetime_attn_b200 import native, ops
from oracle import sta_oracle as O

# (prompts B, n tokens, heads, head_dim, n_obj)
SHAPES = [
    (1, 4096, 8, 40, 2),
    (1, 1024, 8, 80, 2),
    (1, 256, 8, 160, 2),
    (1, 64, 8, 160, 2),
    (1, 4096, 8, 40, 0),
    (1, 1024, 8, 80, 5),
    (2, 576, 8, 160, 6),
    (2, 144, 8, 160, 6),
    (3, 2304, 4, 80, 3),
    (1, 9216, 8, 40, 6),
]


def make_case(B, n, h, d, n_obj, seed=0):
    g = torch.Generator().manual_seed(seed)
    C = h * d
    q = torch.randn(2 * B, n, C, generator=g).half()
    k = torch.randn(B, 2 + n_obj, 77, C, generator=g).half()
    v = torch.randn(B, 2 + n_obj, 77, C, generator=g).half()
    centers = [[(0.3 + 0.4 * ((i + p) % 2)), 0.25 + 0.5 * (((i + p) // 2) % 2)] for p in range(B) for i in range(n_obj)]
    if n_obj:
        masks = torch.stack([O.flat_masks(centers[p * n_obj:(p + 1) * n_obj], n) for p in range(B)])
        coef = (torch.rand(B, n_obj, generator=g) * 3 + 0.5).float()
    else:
        masks = torch.zeros(B, 0, n, dtype=torch.uint8)
        coef = torch.zeros(B, 0)
    return q, k, v, masks, coef


def _close(got, ref, atol=3e-3, rtol=1e-2):
    err = (got.float().cpu() - ref).abs()
    bad = (err > atol + rtol * ref.abs()).sum().item()
    assert bad == 0, f"{bad} elements out of tolerance; max abs err {err.max().item():.3e}"


@pytest.mark.gpu
@pytest.mark.parametrize("shape", SHAPES, ids=[str(s) for s in SHAPES])
def test_xattn_fwd_matches_oracle(shape):
    B, n, h, d, n_obj = shape
    q, k, v, masks, coef = make_case(*shape)
    ref, ref_lse = O.dual_cross_attention_core(q.float(), k.float(), v.float(), masks, coef, h, return_lse=True)
    out, lse = ops.xattn_fwd(q.cuda(), k.cuda(), v.cuda(), masks.cuda() if n_obj else None,
                             coef.cuda() if n_obj else None, h)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(out, ref)
    # LSE is defined for slots 0/1 everywhere and for object slot 2+i on every aligned run of 32 pixels that touches the
    # object's mask (a warp whose pixels are all outside the mask skips that softmax; the backward never reads it there).
    lse = lse.cpu()
    live = torch.ones(B, 1, 2 + n_obj, n, dtype=torch.bool)
    for p in range(B):
        for i in range(n_obj):
            m = masks[p, i].reshape(-1)
            w_live = torch.zeros(n, dtype=torch.bool)
            for t0 in range(0, n, 32):
                if m[t0:t0 + 32].any():
                    w_live[t0:t0 + 32] = True
            live[p, 0, 2 + i] = w_live
    live = live.expand(B, h, 2 + n_obj, n)
    assert (lse - ref_lse)[live].abs().max().item() < 2e-3


@pytest.mark.gpu
def test_xattn_empty_masks_equal_plain_cross_attention():
    """With every mask empty the conditional row is plain attention against the global context."""
    B, n, h, d, n_obj = 1, 1024, 8, 80, 3
    q, k, v, masks, coef = make_case(B, n, h, d, n_obj, seed=2)
    masks.zero_()
    ref_u = O.attention_core(q[:B].float(), k[:, 0].float(), v[:, 0].float(), h)
    ref_c = O.attention_core(q[B:].float(), k[:, 1].float(), v[:, 1].float(), h)
    out, _ = ops.xattn_fwd(q.cuda(), k.cuda(), v.cuda(), masks.cuda(), coef.cuda(), h)
    torch.cuda.synchronize()
    _close(out[:B], ref_u)
    _close(out[B:], ref_c)


@pytest.mark.gpu
def test_xattn_linearity_in_coef():
    """out_c is affine in coef: f(2c) - f(c) == f(c) - f(0) (up to fp16 rounding of the outputs)."""
    B, n, h, d, n_obj = 1, 256, 8, 160, 2
    q, k, v, masks, coef = make_case(B, n, h, d, n_obj, seed=4)
    dev = lambda t: t.cuda()
    f = lambda c: ops.xattn_fwd(dev(q), dev(k), dev(v), dev(masks), dev(c), h)[0].float()
    f0, f1, f2 = f(torch.zeros_like(coef)), f(coef), f(2 * coef)
    torch.cuda.synchronize()
    assert ((f2 - f1) - (f1 - f0)).abs().max().item() < 2e-2


BWD_SHAPES = [
    (1, 4096, 8, 40, 2),
    (1, 1024, 8, 80, 2),
    (1, 256, 8, 160, 2),
    (1, 64, 8, 160, 3),
    (1, 1024, 8, 80, 0),
    (2, 576, 8, 160, 6),
    (2, 2304, 4, 80, 5),
    (1, 100, 2, 40, 1),
]


@pytest.mark.gpu
@pytest.mark.parametrize("shape", BWD_SHAPES, ids=[str(s) for s in BWD_SHAPES])
def test_xattn_bwd_matches_oracle_autograd(shape):
    """d(q) and d(coef) against torch autograd through the oracle restatement (fp32, same fp16-rounded inputs).

    Tolerances: d_q element-wise 3e-3 + 2e-2*|ref| (fp16 dS operand); d_coef relative 2e-2.
    """
    B, n, h, d, n_obj = shape
    q, k, v, masks, coef = make_case(*shape, seed=7)
    g = torch.Generator().manual_seed(11)
    d_out = (torch.randn(2 * B, n, h * d, generator=g) * 0.1).half()
    qf = q.float().requires_grad_(True)
    cf = coef.clone().requires_grad_(n_obj > 0)
    ref = O.dual_cross_attention_core(qf, k.float(), v.float(), masks, cf, h)
    (ref * d_out.float()).sum().backward()
    dev = lambda t: t.cuda()
    out, lse = ops.xattn_fwd(dev(q), dev(k), dev(v), dev(masks) if n_obj else None, dev(coef) if n_obj else None, h)
    d_q, d_coef = ops.xattn_bwd(dev(q), dev(k), dev(v), dev(masks) if n_obj else None,
                                dev(coef) if n_obj else None, lse, dev(d_out), h, out=out)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(d_q, qf.grad, atol=3e-3, rtol=2e-2)
    if n_obj:
        rel = (d_coef.cpu() - cf.grad).abs() / (cf.grad.abs() + 1e-2 * cf.grad.abs().max() + 1e-6)
        assert rel.max().item() < 2e-2, f"d_coef rel err {rel.max().item():.3e}: {d_coef.cpu()} vs {cf.grad}"


@pytest.mark.gpu
@pytest.mark.parametrize("ctx_len", [80, 64, 17, 1])
@pytest.mark.parametrize("d", [40, 160])
def test_xattn_context_lengths_other_than_77(ctx_len, d):
    """The C ABI takes any ctx_len in [1, 80] (CLIP's 77 is only the common case): the padding columns of the 80-wide score
    tiles must be masked in the forward softmax and in the backward's recomputed P for every length."""
    B, n, h, n_obj = 1, 300, 2, 2
    g = torch.Generator().manual_seed(100 + ctx_len + d)
    C = h * d
    q = torch.randn(2 * B, n, C, generator=g).half()
    k = torch.randn(B, 2 + n_obj, ctx_len, C, generator=g).half()
    v = torch.randn(B, 2 + n_obj, ctx_len, C, generator=g).half()
    masks = (torch.rand(B, n_obj, n, generator=g) < 0.4).to(torch.uint8)
    coef = (torch.rand(B, n_obj, generator=g) * 3 + 0.5).float()
    d_out = (torch.randn(2 * B, n, C, generator=g) * 0.1).half()
    qf, cf = q.float().requires_grad_(True), coef.clone().requires_grad_(True)
    ref = O.dual_cross_attention_core(qf, k.float(), v.float(), masks, cf, h)
    (ref * d_out.float()).sum().backward()
    dev = lambda t: t.cuda()
    out, lse = ops.xattn_fwd(dev(q), dev(k), dev(v), dev(masks), dev(coef), h)
    d_q, d_coef = ops.xattn_bwd(dev(q), dev(k), dev(v), dev(masks), dev(coef), lse, dev(d_out), h, out=out)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(out, ref.detach())
    _close(d_q, qf.grad, atol=3e-3, rtol=2e-2)
    rel = (d_coef.cpu() - cf.grad).abs() / (cf.grad.abs() + 1e-2 * cf.grad.abs().max() + 1e-6)
    assert rel.max().item() < 2e-2


@pytest.mark.gpu
def test_xattn_maximum_object_count_and_limits():
    """n_obj = 8 is the library's maximum (every object live in some tiles, dead in others: ring refills + skips);
    9 objects, 81 context tokens or an unsupported head dim are ERRORS, never a fallback."""
    B, n, h, d, n_obj = 1, 1024, 2, 80, 8
    g = torch.Generator().manual_seed(77)
    C = h * d
    q = torch.randn(2 * B, n, C, generator=g).half()
    k = torch.randn(B, 2 + n_obj, 77, C, generator=g).half()
    v = torch.randn(B, 2 + n_obj, 77, C, generator=g).half()
    centers = [[0.15 + 0.1 * i, 0.2 + 0.08 * i] for i in range(n_obj)]
    masks = O.flat_masks(centers, n).unsqueeze(0)
    coef = (torch.rand(B, n_obj, generator=g) * 2 + 0.5).float()
    d_out = (torch.randn(2 * B, n, C, generator=g) * 0.1).half()
    qf, cf = q.float().requires_grad_(True), coef.clone().requires_grad_(True)
    ref = O.dual_cross_attention_core(qf, k.float(), v.float(), masks, cf, h)
    (ref * d_out.float()).sum().backward()
    dev = lambda t: t.cuda()
    out, lse = ops.xattn_fwd(dev(q), dev(k), dev(v), dev(masks), dev(coef), h)
    d_q, d_coef = ops.xattn_bwd(dev(q), dev(k), dev(v), dev(masks), dev(coef), lse, dev(d_out), h, out=out)
    torch.cuda.synchronize()
    assert native.device_error() == 0
    _close(out, ref.detach())
    _close(d_q, qf.grad, atol=3e-3, rtol=2e-2)
    rel = (d_coef.cpu() - cf.grad).abs() / (cf.grad.abs() + 1e-2 * cf.grad.abs().max() + 1e-6)
    assert rel.max().item() < 2e-2
    # limits
    k9 = torch.randn(B, 2 + 9, 77, C).half().cuda()
    with pytest.raises(RuntimeError, match="n_obj"):
        ops.xattn_fwd(dev(q), k9, k9, torch.zeros(B, 9, n, dtype=torch.uint8).cuda(), torch.ones(B, 9).cuda(), h)
    k81 = torch.randn(B, 2, 81, C).half().cuda()
    with pytest.raises(RuntimeError, match="ctx_len"):
        ops.xattn_fwd(dev(q), k81, k81, None, None, h)
    q56 = torch.randn(2, 128, 2 * 56).half().cuda()
    k56 = torch.randn(1, 2, 77, 2 * 56).half().cuda()
    with pytest.raises(RuntimeError, match="head_dim"):
        ops.xattn_fwd(q56, k56, k56, None, None, 2)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.xattn_fwd(q, dev(k), dev(v), dev(masks), dev(coef), h)
