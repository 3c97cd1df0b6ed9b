"""torch-facing wrappers of the C-ABI kernels (include/sta_b200.h) and their autograd Functions.

PyTorch is plumbing here: it owns device memory and the stream; every attention FLOP runs in libsta_b200.so.
There is no fallback path — a missing library or a non-CUDA tensor raises.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import native

# number of kernels launched by this module since import (bench.py reports it as `gpu_launches`)
LAUNCHES = {"sattn_fwd": 0, "sattn_bwd": 0, "xattn_fwd": 0, "xattn_bwd": 0, "groupnorm_fwd": 0, "groupnorm_bwd": 0,
            "add_layernorm_fwd": 0, "add_layernorm_bwd": 0, "geglu_fwd": 0, "geglu_bwd": 0,
            "upsample2x_fwd": 0, "upsample2x_bwd": 0, "plms_step_fwd": 0, "plms_step_bwd": 0}


def launch_count() -> int:
    return sum(LAUNCHES.values())


# Optional per-launch device timing (bench.py turns it on for the timed region): CUDA events are recorded on the
# launching stream right before / after every C-ABI call; `kernel_time_summary()` reduces them after a sync.
_PROFILE = None
_TRACE = None  # when a list: (kind, geometry) of every launch, no events (usable during CUDA-graph capture)


def trace_start() -> None:
    global _TRACE
    _TRACE = []


def trace_stop():
    global _TRACE
    rec, _TRACE = _TRACE, None
    return rec or []


def profile_start() -> None:
    global _PROFILE
    _PROFILE = []


def profile_stop():
    global _PROFILE
    rec, _PROFILE = _PROFILE, None
    return rec or []


class _timed:
    def __init__(self, kind, key):
        self.kind, self.key = kind, key

    def __enter__(self):
        if _TRACE is not None:
            _TRACE.append((self.kind, self.key))
        if _PROFILE is not None:
            self.a = torch.cuda.Event(enable_timing=True)
            self.b = torch.cuda.Event(enable_timing=True)
            self.a.record()
        return self

    def __exit__(self, *exc):
        if _PROFILE is not None:
            self.b.record()
            _PROFILE.append((self.kind, self.key, self.a, self.b))
        return False


def kernel_time_summary(records):
    """{(kind, key): (launches, total_ms)} from profile_stop() records; call after torch.cuda.synchronize()."""
    out = {}
    for kind, key, a, b in records:
        n, t = out.get((kind, key), (0, 0.0))
        out[(kind, key)] = (n + 1, t + a.elapsed_time(b))
    return out


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _require(t: torch.Tensor, name: str, dtype=torch.float16) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the sta_b200 kernels have no CPU fallback")
    if t.dtype != dtype:
        raise RuntimeError(f"{name} must be {dtype}, got {t.dtype}")


def _token_major(t: torch.Tensor, name: str) -> Tuple[int, int]:
    """(token_stride, batch_stride) in elements of a [b, n, c] tensor whose channel dim is dense."""
    if t.dim() != 3 or t.stride(2) != 1:
        raise RuntimeError(f"{name} must be [batch, n, channels] with unit channel stride")
    return t.stride(1), t.stride(0)


# ------------------------------------------------------------------------------------------------------
# self-attention
# ------------------------------------------------------------------------------------------------------
def sattn_fwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, heads: int, scale: Optional[float] = None,
              need_lse: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """softmax(q k^T * scale) v per head.  q/k/v fp16 [b, n, heads*d] (may be strided views of one projection)."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v")):
        _require(t, nm)
    b, n, c = q.shape
    d = c // heads
    scale = float(d ** -0.5) if scale is None else float(scale)
    out = torch.empty((b, n, c), device=q.device, dtype=torch.float16)
    lse = torch.empty((b, heads, n), device=q.device, dtype=torch.float32) if need_lse else None
    a = native.SattnFwdArgs()
    a.q, a.k, a.v, a.out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr()
    a.lse = lse.data_ptr() if lse is not None else None
    a.batch, a.n, a.heads, a.head_dim = b, n, heads, d
    a.q_token_stride, a.q_batch_stride = _token_major(q, "q")
    a.k_token_stride, a.k_batch_stride = _token_major(k, "k")
    a.v_token_stride, a.v_batch_stride = _token_major(v, "v")
    a.o_token_stride, a.o_batch_stride = _token_major(out, "out")
    a.scale = scale
    with _timed("sattn_fwd", (b, n, heads, d)):
        native.check(native.load().sta_sattn_fwd(C.byref(a), _stream()), "sta_sattn_fwd")
    LAUNCHES["sattn_fwd"] += 1
    return out, lse


# ------------------------------------------------------------------------------------------------------
# fused dual cross-attention + alpha-blend
# ------------------------------------------------------------------------------------------------------
def xattn_fwd(q: torch.Tensor, k_ctx: torch.Tensor, v_ctx: torch.Tensor, mask: Optional[torch.Tensor],
              coef: Optional[torch.Tensor], heads: int, scale: Optional[float] = None,
              need_lse: bool = True) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """q fp16 [2B, n, C]; k_ctx/v_ctx fp16 [B, 2+n_obj, L, C] contiguous; mask u8 [B, n_obj, n]; coef f32 [B, n_obj].

    Returns (out fp16 [2B, n, C], lse f32 [B, heads, 2+n_obj, n] | None).  See include/sta_b200.h.
    """
    _require(q, "q")
    _require(k_ctx, "k_ctx")
    _require(v_ctx, "v_ctx")
    b2, n, c = q.shape
    B = b2 // 2
    if k_ctx.dim() != 4 or k_ctx.shape[0] != B or not k_ctx.is_contiguous() or not v_ctx.is_contiguous():
        raise RuntimeError("k_ctx/v_ctx must be contiguous [B, 2+n_obj, ctx_len, C]")
    n_obj, ctx_len = k_ctx.shape[1] - 2, k_ctx.shape[2]
    d = c // heads
    scale = float(d ** -0.5) if scale is None else float(scale)
    if n_obj > 0:
        _require(mask, "mask", torch.uint8)
        _require(coef, "coef", torch.float32)
        if tuple(mask.shape) != (B, n_obj, n) or not mask.is_contiguous():
            raise RuntimeError(f"mask must be contiguous uint8 [{B}, {n_obj}, {n}], got {tuple(mask.shape)}")
        if tuple(coef.shape) != (B, n_obj) or not coef.is_contiguous():
            raise RuntimeError(f"coef must be contiguous float32 [{B}, {n_obj}], got {tuple(coef.shape)}")
    out = torch.empty((b2, n, c), device=q.device, dtype=torch.float16)
    # only the entries the forward evaluates are ever read by the backward (see include/sta_b200.h): no zero-fill node
    lse = torch.empty((B, heads, 2 + n_obj, n), device=q.device, dtype=torch.float32) if need_lse else None
    a = native.XattnFwdArgs()
    a.q, a.k_ctx, a.v_ctx, a.out = q.data_ptr(), k_ctx.data_ptr(), v_ctx.data_ptr(), out.data_ptr()
    a.mask = mask.data_ptr() if n_obj > 0 else None
    a.coef = coef.data_ptr() if n_obj > 0 else None
    a.lse = lse.data_ptr() if lse is not None else None
    a.prompts, a.n, a.heads, a.head_dim, a.n_obj, a.ctx_len = B, n, heads, d, n_obj, ctx_len
    a.q_token_stride, a.q_batch_stride = _token_major(q, "q")
    a.o_token_stride, a.o_batch_stride = _token_major(out, "out")
    a.scale = scale
    with _timed("xattn_fwd", (B, n, heads, d, n_obj)):
        native.check(native.load().sta_xattn_fwd(C.byref(a), _stream()), "sta_xattn_fwd")
    LAUNCHES["xattn_fwd"] += 1
    return out, lse


def xattn_bwd(q: torch.Tensor, k_ctx: torch.Tensor, v_ctx: torch.Tensor, mask: Optional[torch.Tensor],
              coef: Optional[torch.Tensor], lse: torch.Tensor, d_out: torch.Tensor, heads: int,
              scale: Optional[float] = None, out: Optional[torch.Tensor] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Backward of xattn_fwd: returns (d_q fp16 [2B, n, C], d_coef f32 [B, n_obj] | None).  `out` is the forward's
    output (required when there are objects: its unconditional rows give delta_uc for d_coef)."""
    _require(q, "q")
    _require(d_out, "d_out")
    _require(lse, "lse", torch.float32)
    if k_ctx.shape[1] > 2:
        if out is None:
            raise RuntimeError("xattn_bwd needs the forward output `out` when n_obj > 0")
        _require(out, "out")
    b2, n, c = q.shape
    B = b2 // 2
    n_obj, ctx_len = k_ctx.shape[1] - 2, k_ctx.shape[2]
    d = c // heads
    scale = float(d ** -0.5) if scale is None else float(scale)
    d_q = torch.empty((b2, n, c), device=q.device, dtype=torch.float16)
    d_coef = torch.empty((B, n_obj), device=q.device, dtype=torch.float32) if n_obj > 0 else None
    a = native.XattnBwdArgs()
    a.q, a.k_ctx, a.v_ctx = q.data_ptr(), k_ctx.data_ptr(), v_ctx.data_ptr()
    a.mask = mask.data_ptr() if n_obj > 0 else None
    a.coef = coef.data_ptr() if n_obj > 0 else None
    a.lse, a.d_out, a.d_q = lse.data_ptr(), d_out.data_ptr(), d_q.data_ptr()
    a.d_coef = d_coef.data_ptr() if d_coef is not None else None
    a.prompts, a.n, a.heads, a.head_dim, a.n_obj, a.ctx_len = B, n, heads, d, n_obj, ctx_len
    a.q_token_stride, a.q_batch_stride = _token_major(q, "q")
    a.do_token_stride, a.do_batch_stride = _token_major(d_out, "d_out")
    a.scale = scale
    if out is not None:
        a.out = out.data_ptr()
        a.o_token_stride, a.o_batch_stride = _token_major(out, "out")
    with _timed("xattn_bwd", (B, n, heads, d, n_obj)):
        native.check(native.load().sta_xattn_bwd(C.byref(a), _stream()), "sta_xattn_bwd")
    LAUNCHES["xattn_bwd"] += 1
    return d_q, d_coef


def sattn_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: torch.Tensor,
              d_out: torch.Tensor, heads: int, scale: Optional[float] = None,
              d_fused: Optional[torch.Tensor] = None):
    """Backward of sattn_fwd: returns (d_q, d_k, d_v) fp16 [b, n, C].  With `d_fused` (a contiguous fp16 [b, n, 3C]
    buffer) they are written as its three column slices — the gradient of a fused QKV projection, no concatenation."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (d_out, "d_out")):
        _require(t, nm)
    _require(lse, "lse", torch.float32)
    b, n, c = q.shape
    d = c // heads
    scale = float(d ** -0.5) if scale is None else float(scale)
    if d_fused is not None:
        _require(d_fused, "d_fused")
        if tuple(d_fused.shape) != (b, n, 3 * c) or not d_fused.is_contiguous():
            raise RuntimeError(f"d_fused must be contiguous [{b}, {n}, {3 * c}]")
        d_qkv = d_fused.view(b, n, 3, c).permute(2, 0, 1, 3)  # [3, b, n, c] views, token stride 3c
    else:
        d_qkv = torch.empty((3, b, n, c), device=q.device, dtype=torch.float16)
    # head dim 512 (VAE AttnBlock, sta_sattn_wide.cu): every gradient element is owned by one CTA, no fp32 accumulator
    dq_accum = torch.empty((b, n, c), device=q.device, dtype=torch.float32) if d != 512 else None
    delta = torch.empty((b, heads, n), device=q.device, dtype=torch.float32)
    a = native.SattnBwdArgs()
    a.q, a.k, a.v, a.out, a.d_out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), d_out.data_ptr()
    a.lse = lse.data_ptr()
    a.d_q, a.d_k, a.d_v = d_qkv[0].data_ptr(), d_qkv[1].data_ptr(), d_qkv[2].data_ptr()
    a.dq_accum, a.delta = (dq_accum.data_ptr() if dq_accum is not None else None), delta.data_ptr()
    a.batch, a.n, a.heads, a.head_dim = b, n, heads, d
    a.q_token_stride, a.q_batch_stride = _token_major(q, "q")
    a.k_token_stride, a.k_batch_stride = _token_major(k, "k")
    a.v_token_stride, a.v_batch_stride = _token_major(v, "v")
    a.o_token_stride, a.o_batch_stride = _token_major(out, "out")
    a.do_token_stride, a.do_batch_stride = _token_major(d_out, "d_out")
    a.scale = scale
    a.dqkv_token_stride = 3 * c if d_fused is not None else 0
    with _timed("sattn_bwd", (b, n, heads, d)):
        native.check(native.load().sta_sattn_bwd(C.byref(a), _stream()), "sta_sattn_bwd")
    LAUNCHES["sattn_bwd"] += 3 if d != 512 else 2  # delta, main, dq cast (head dim 512: delta, three-role main)
    return d_qkv[0], d_qkv[1], d_qkv[2]


# ------------------------------------------------------------------------------------------------------
# autograd
# ------------------------------------------------------------------------------------------------------
class SelfAttentionFn(torch.autograd.Function):
    """attn1 core.  Saves q, k, v, out, lse (flash style) — the [heads, N, N] matrix is never stored."""

    @staticmethod
    def forward(ctx, q, k, v, heads):
        need = any(ctx.needs_input_grad[:3])
        out, lse = sattn_fwd(q, k, v, heads, need_lse=need)
        if need:
            ctx.save_for_backward(q, k, v, out, lse)
            ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k, v, out, lse = ctx.saved_tensors
        if d_out.stride(2) != 1:
            d_out = d_out.contiguous()
        d_q, d_k, d_v = sattn_bwd(q, k, v, out, lse, d_out.to(torch.float16), ctx.heads)
        return d_q, d_k, d_v, None


class DualCrossAttentionFn(torch.autograd.Function):
    """Fused (1 + n_obj) x attn2 + alpha-blend, before `to_out`.  Gradients flow to q and coef only: the
    context K/V are projections of frozen text embeddings (reference ddpm.py:519-523)."""

    @staticmethod
    def forward(ctx, q, k_ctx, v_ctx, mask, coef, heads):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[4]
        out, lse = xattn_fwd(q, k_ctx, v_ctx, mask, coef, heads, need_lse=need)
        if need:
            ctx.save_for_backward(q, k_ctx, v_ctx, mask, coef, lse, out)
            ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k_ctx, v_ctx, mask, coef, lse, out = ctx.saved_tensors
        if d_out.stride(2) != 1:
            d_out = d_out.contiguous()
        d_q, d_coef = xattn_bwd(q, k_ctx, v_ctx, mask, coef, lse, d_out.to(torch.float16), ctx.heads, out=out)
        return d_q, None, None, None, d_coef, None


class SelfAttentionQKVFn(torch.autograd.Function):
    """attn1 core on the output of ONE fused [C, 3C] projection: q/k/v are column slices of `qkv` [b, n, 3C], and the
    backward kernel writes d_q/d_k/d_v straight into one [b, n, 3C] gradient (no split-backward concatenation)."""

    @staticmethod
    def forward(ctx, qkv, heads):
        q, k, v = qkv.chunk(3, dim=-1)
        need = ctx.needs_input_grad[0]
        out, lse = sattn_fwd(q, k, v, heads, need_lse=need)
        if need:
            ctx.save_for_backward(qkv, out, lse)
            ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, d_out):
        qkv, out, lse = ctx.saved_tensors
        q, k, v = qkv.chunk(3, dim=-1)
        if d_out.stride(2) != 1:
            d_out = d_out.contiguous()
        d_qkv = torch.empty(qkv.shape, device=qkv.device, dtype=torch.float16)
        sattn_bwd(q, k, v, out, lse, d_out.to(torch.float16), ctx.heads, d_fused=d_qkv)
        return d_qkv, None


def self_attention(q, k, v, heads):
    return SelfAttentionFn.apply(q, k, v, heads)


def self_attention_qkv(qkv, heads):
    return SelfAttentionQKVFn.apply(qkv, heads)


def dual_cross_attention(q, k_ctx, v_ctx, mask, coef, heads):
    return DualCrossAttentionFn.apply(q, k_ctx, v_ctx, mask, coef, heads)


# ------------------------------------------------------------------------------------------------------
# fused GroupNorm(32) [+ SiLU], NHWC fp16
# ------------------------------------------------------------------------------------------------------
def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """[B, C, H, W] tensor whose memory is NHWC-dense (channels_last); copies only if it is not already."""
    return x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)


def _check_x_bias(x_bias, b: int, c: int):
    """(pointer, row stride) of an fp16 [b, c] bias whose rows are dense; the rows themselves may be strided (a column slice
    of a wider table)."""
    if x_bias is None:
        return None, 0
    _require(x_bias, "x_bias")
    if tuple(x_bias.shape) != (b, c) or x_bias.stride(1) != 1 or (b > 1 and (x_bias.stride(0) < c or x_bias.stride(0) % 8)):
        raise RuntimeError(f"x_bias must be fp16 [{b}, {c}] with dense rows and a row stride that is a multiple of 8, got "
                           f"{tuple(x_bias.shape)} / {tuple(x_bias.stride())}")
    if x_bias.data_ptr() % 16:
        raise RuntimeError("x_bias must be 16-byte aligned")
    return x_bias.data_ptr(), (x_bias.stride(0) if b > 1 else c)


def groupnorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool,
                  x_bias: Optional[torch.Tensor] = None, x1: Optional[torch.Tensor] = None):
    """(out, stats, x as NHWC memory).  With `x1` the normalised tensor is torch.cat([x, x1], dim=1): the kernel reads the two
    parts in place and also writes the concatenation, which is returned as the third value."""
    _require(x, "x")
    _require(gamma, "gamma", torch.float32)
    _require(beta, "beta", torch.float32)
    b, c, h, w = x.shape
    x = _nhwc(x)
    a = native.GroupNormArgs()
    if x1 is not None:
        _require(x1, "x1")
        if x1.shape[0] != b or tuple(x1.shape[2:]) != (h, w):
            raise RuntimeError(f"groupnorm_fwd: x1 {tuple(x1.shape)} does not concatenate with x {tuple(x.shape)} along channels")
        x1 = _nhwc(x1)
        c_split, c = c, c + x1.shape[1]
        xcat = torch.empty((b, c, h, w), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        a.x1, a.x_cat, a.c_split = x1.data_ptr(), xcat.data_ptr(), c_split
    out = torch.empty((b, c, h, w), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
    stats = torch.empty((b, 32, 2), device=x.device, dtype=torch.float32)
    a.x, a.gamma, a.beta, a.out, a.stats = x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), stats.data_ptr()
    a.batch, a.hw, a.channels, a.silu, a.eps = b, h * w, c, int(silu), float(eps)
    a.x_bias, a.x_bias_stride = _check_x_bias(x_bias, b, c)
    with _timed("groupnorm_fwd", (b, h * w, c)):
        native.check(native.load().sta_groupnorm_fwd(C.byref(a), _stream()), "sta_groupnorm_fwd")
    LAUNCHES["groupnorm_fwd"] += 2
    return out, stats, (x if x1 is None else xcat)


def groupnorm_bwd(x, d_out, gamma, beta, stats, eps: float, silu: bool, x_bias: Optional[torch.Tensor] = None,
                  d_res: Optional[torch.Tensor] = None, split: Optional[int] = None):
    """d_x of GroupNorm[+SiLU]; `d_res` (same shape as x) is a second gradient of x that the kernel adds on the fly.  With
    `split` the result is written as two dense tensors (d_x[:, :split], d_x[:, split:]) — the gradient of a fused cat."""
    b, c, h, w = x.shape
    d_out = _nhwc(d_out if d_out.dtype == torch.float16 else d_out.to(torch.float16))
    if d_res is not None:
        if d_res.shape != x.shape:
            raise RuntimeError(f"groupnorm_bwd: d_res {tuple(d_res.shape)} must match x {tuple(x.shape)}")
        d_res = d_res if d_res.dtype == torch.float16 else d_res.to(torch.float16)
        rs = d_res.stride(3) if w > 1 else (d_res.stride(2) if h > 1 else c)  # NHWC row stride: a channel slice keeps it
        if not (d_res.stride(1) == 1 and rs >= c and rs % 8 == 0 and d_res.data_ptr() % 16 == 0
                and (h == 1 or w == 1 or d_res.stride(2) == w * rs) and (b == 1 or d_res.stride(0) == h * w * rs)):
            d_res, rs = _nhwc(d_res), c
        res_stride = rs
    a = native.GroupNormArgs()
    if split is None:
        d_x = torch.empty_like(x, memory_format=torch.channels_last)
    else:
        d_x = torch.empty((b, split, h, w), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        d_x1 = torch.empty((b, c - split, h, w), device=x.device, dtype=x.dtype, memory_format=torch.channels_last)
        a.out1, a.c_split = d_x1.data_ptr(), split
    bstats = torch.empty((b, 32, 2), device=x.device, dtype=torch.float32)
    a.x, a.d_out, a.gamma, a.beta, a.out = x.data_ptr(), d_out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), d_x.data_ptr()
    a.stats, a.bwd_stats = stats.data_ptr(), bstats.data_ptr()
    a.batch, a.hw, a.channels, a.silu, a.eps = b, h * w, c, int(silu), float(eps)
    a.x_bias, a.x_bias_stride = _check_x_bias(x_bias, b, c)
    a.d_res, a.d_res_stride = (d_res.data_ptr(), res_stride) if d_res is not None else (None, 0)
    with _timed("groupnorm_bwd", (b, h * w, c)):
        native.check(native.load().sta_groupnorm_bwd(C.byref(a), _stream()), "sta_groupnorm_bwd")
    LAUNCHES["groupnorm_bwd"] += 2
    return d_x if split is None else (d_x, d_x1)


class GroupNormSiLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, silu, x_bias):
        out, stats, x_nhwc = groupnorm_fwd(x, gamma, beta, eps, silu, x_bias)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(x_nhwc, gamma, beta, stats, x_bias)
            ctx.eps, ctx.silu = eps, silu
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, gamma, beta, stats, x_bias = ctx.saved_tensors
        return groupnorm_bwd(x, d_out, gamma, beta, stats, ctx.eps, ctx.silu, x_bias), None, None, None, None, None


class GroupNormSiLUForkFn(torch.autograd.Function):
    """(GroupNorm[+SiLU](x), x): the second output is x itself for the residual branch that every ResBlock /
    SpatialTransformer takes next to its GroupNorm.  Both gradients arrive in ONE backward call, so the kernel adds the
    residual gradient while it writes d_x instead of autograd launching a separate accumulation pass per block."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, silu):
        out, stats, x_nhwc = groupnorm_fwd(x, gamma, beta, eps, silu, None)
        ctx.set_materialize_grads(False)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(x_nhwc, gamma, beta, stats)
            ctx.eps, ctx.silu = eps, silu
        return out, x_nhwc.view_as(x_nhwc)

    @staticmethod
    def backward(ctx, d_out, d_x_res):
        x, gamma, beta, stats = ctx.saved_tensors
        if d_out is None:  # only the pass-through output was used
            return d_x_res, None, None, None, None
        return groupnorm_bwd(x, d_out, gamma, beta, stats, ctx.eps, ctx.silu, None, d_res=d_x_res), None, None, None, None


class CatGroupNormSiLUFn(torch.autograd.Function):
    """(GroupNorm[+SiLU](cat([h, skip], 1)), cat([h, skip], 1)) in one pass over h and skip (openaimodel.py:731 + :258): the
    decoder's torch.cat is fused into the first GroupNorm of the ResBlock that consumes it.  Backward: both gradients (of the
    normalised tensor and of the concatenation, i.e. the ResBlock's skip branch) arrive together, and the kernel writes
    d_h and d_skip as two DENSE tensors — the stock cat hands out strided channel slices that every consumer has to copy."""

    @staticmethod
    def forward(ctx, h, skip, gamma, beta, eps, silu):
        out, stats, xcat = groupnorm_fwd(h, gamma, beta, eps, silu, None, x1=skip)
        ctx.set_materialize_grads(False)
        ctx.c_split = h.shape[1]
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            ctx.save_for_backward(xcat, gamma, beta, stats)
            ctx.eps, ctx.silu = eps, silu
        return out, xcat

    @staticmethod
    def backward(ctx, d_out, d_cat):
        xcat, gamma, beta, stats = ctx.saved_tensors
        c0 = ctx.c_split
        if d_out is None:
            if d_cat is None:
                return None, None, None, None, None, None
            return d_cat[:, :c0], d_cat[:, c0:], None, None, None, None
        d_h, d_skip = groupnorm_bwd(xcat, d_out, gamma, beta, stats, ctx.eps, ctx.silu, None, d_res=d_cat, split=c0)
        return d_h, d_skip, None, None, None, None


def cat_group_norm_silu(h, skip, gamma, beta, eps=1e-5, silu=True):
    """(GroupNorm(32)(cat([h, skip], 1).float()).half() [-> SiLU], cat([h, skip], 1)) without a separate cat kernel."""
    return CatGroupNormSiLUFn.apply(h, skip, gamma, beta, eps, silu)


def group_norm_silu_fork(x, gamma, beta, eps=1e-5, silu=True):
    """(GroupNorm(32)(x.float()).half() [-> SiLU], x as NHWC memory) — use the second value for the residual branch."""
    return GroupNormSiLUForkFn.apply(x, gamma, beta, eps, silu)


def group_norm_silu(x, gamma, beta, eps=1e-5, silu=True, x_bias=None):
    """GroupNorm(32)((x + x_bias[:, :, None, None]).float()).half() [-> SiLU] in one pass; x fp16 [B, C, H, W] (any
    layout, NHWC is free); x_bias: optional fp16 [B, C] without gradient (conv bias + timestep embedding)."""
    if x_bias is not None and x_bias.requires_grad:
        raise RuntimeError("group_norm_silu: x_bias is treated as a constant (frozen UNet); it must not require grad")
    return GroupNormSiLUFn.apply(x, gamma, beta, eps, silu, x_bias)


# ------------------------------------------------------------------------------------------------------
# token-major streaming kernels: (bias +) residual add + LayerNorm, GEGLU   (csrc/sta_tokens.cu)
# ------------------------------------------------------------------------------------------------------
def _rows2d(t: torch.Tensor, name: str) -> torch.Tensor:
    """[..., C] fp16 CUDA tensor with dense rows (copies only if it is not already contiguous)."""
    _require(t, name)
    return t if t.is_contiguous() else t.contiguous()


def add_layernorm_fwd(x, bias, residual, gamma, beta, eps: float, need_stats: bool = True):
    """s = x (+ bias) (+ residual); y = LayerNorm(s).  Returns (s, y, stats); s is x itself when nothing is added,
    y / stats are None when gamma is None (plain fused bias + residual add)."""
    x = _rows2d(x, "x")
    c = x.shape[-1]
    rows = x.numel() // c
    added = bias is not None or residual is not None
    if residual is not None:
        residual = _rows2d(residual, "residual")
        if residual.shape != x.shape:
            raise RuntimeError(f"residual {tuple(residual.shape)} must match x {tuple(x.shape)}")
    for t, nm in ((bias, "bias"), (gamma, "gamma"), (beta, "beta")):
        if t is not None:
            _require(t, nm, torch.float32)
            if t.numel() != c or not t.is_contiguous():
                raise RuntimeError(f"{nm} must be a contiguous float32 [{c}] tensor")
    s = torch.empty_like(x) if added else x
    y = torch.empty_like(x) if gamma is not None else None
    stats = torch.empty((rows, 2), device=x.device, dtype=torch.float32) if (gamma is not None and need_stats) else None
    a = native.AddLayerNormArgs()
    a.x = x.data_ptr()
    a.bias = bias.data_ptr() if bias is not None else None
    a.residual = residual.data_ptr() if residual is not None else None
    a.gamma = gamma.data_ptr() if gamma is not None else None
    a.beta = beta.data_ptr() if gamma is not None else None
    a.sum_out = s.data_ptr() if added else None
    a.y = y.data_ptr() if y is not None else None
    a.stats = stats.data_ptr() if stats is not None else None
    a.rows, a.channels, a.eps = rows, c, float(eps)
    with _timed("add_layernorm_fwd", (rows, c)):
        native.check(native.load().sta_add_layernorm_fwd(C.byref(a), _stream()), "sta_add_layernorm_fwd")
    LAUNCHES["add_layernorm_fwd"] += 1
    return s, y, stats


def add_layernorm_bwd(d_y, d_sum, xs, stats, gamma):
    """d_x = d_sum + LN'(d_y) (d_sum may be None)."""
    d_y = _rows2d(d_y if d_y.dtype == torch.float16 else d_y.to(torch.float16), "d_y")
    if d_sum is not None:
        d_sum = _rows2d(d_sum if d_sum.dtype == torch.float16 else d_sum.to(torch.float16), "d_sum")
    c = xs.shape[-1]
    rows = xs.numel() // c
    d_x = torch.empty_like(xs)
    a = native.AddLayerNormBwdArgs()
    a.d_y, a.xs, a.stats, a.gamma, a.d_x = d_y.data_ptr(), xs.data_ptr(), stats.data_ptr(), gamma.data_ptr(), d_x.data_ptr()
    a.d_sum = d_sum.data_ptr() if d_sum is not None else None
    a.rows, a.channels = rows, c
    with _timed("add_layernorm_bwd", (rows, c)):
        native.check(native.load().sta_add_layernorm_bwd(C.byref(a), _stream()), "sta_add_layernorm_bwd")
    LAUNCHES["add_layernorm_bwd"] += 1
    return d_x


class LayerNormFn(torch.autograd.Function):
    """y = LayerNorm(x) on fp16 tokens, fp32 statistics / affine, fp16 result (frozen gamma / beta)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        need = ctx.needs_input_grad[0]
        xs, y, stats = add_layernorm_fwd(x, None, None, gamma, beta, eps, need_stats=need)
        if need:
            ctx.save_for_backward(xs, stats, gamma)
        return y

    @staticmethod
    def backward(ctx, d_y):
        xs, stats, gamma = ctx.saved_tensors
        return add_layernorm_bwd(d_y, None, xs, stats, gamma), None, None, None


class LayerNormForkFn(torch.autograd.Function):
    """(LayerNorm(x), x): the second output is x itself for the residual stream (attention.py:274 `attn1(norm1(x)) + x`);
    both gradients arrive in one backward call and the LayerNorm kernel adds the residual one (its d_sum input) instead of
    autograd launching an accumulation pass per transformer block."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        need = ctx.needs_input_grad[0]
        xs, y, stats = add_layernorm_fwd(x, None, None, gamma, beta, eps, need_stats=need)
        if need:
            ctx.save_for_backward(xs, stats, gamma)
        ctx.set_materialize_grads(False)
        return y, xs.view_as(xs)

    @staticmethod
    def backward(ctx, d_y, d_x_res):
        xs, stats, gamma = ctx.saved_tensors
        if d_y is None:
            return d_x_res, None, None, None
        return add_layernorm_bwd(d_y, d_x_res, xs, stats, gamma), None, None, None


class AddLayerNormFn(torch.autograd.Function):
    """(s, y) = (x + bias + residual, LayerNorm(s)): the residual stream and the next sub-layer's input in one pass.
    Backward: ONE kernel gives d_s_total = d_s + LN'(d_y), which is the gradient of both x and residual."""

    @staticmethod
    def forward(ctx, x, bias, residual, gamma, beta, eps):
        need = ctx.needs_input_grad[0] or ctx.needs_input_grad[2]
        s, y, stats = add_layernorm_fwd(x, bias, residual, gamma, beta, eps, need_stats=need)
        if need:
            ctx.save_for_backward(s, stats, gamma)
        ctx.set_materialize_grads(False)
        return s, y

    @staticmethod
    def backward(ctx, d_s, d_y):
        s, stats, gamma = ctx.saved_tensors
        if d_y is None:
            d = d_s
        else:
            d = add_layernorm_bwd(d_y, d_s, s, stats, gamma)
        return (d if ctx.needs_input_grad[0] else None), None, (d if ctx.needs_input_grad[2] else None), None, None, None


class BiasResidualAddFn(torch.autograd.Function):
    """s = x + bias + residual (fp32 accumulate, one rounding); the gradient passes through to x and residual."""

    @staticmethod
    def forward(ctx, x, bias, residual):
        s, _, _ = add_layernorm_fwd(x, bias, residual, None, None, 0.0, need_stats=False)
        return s

    @staticmethod
    def backward(ctx, d_s):
        return (d_s if ctx.needs_input_grad[0] else None), None, (d_s if ctx.needs_input_grad[2] else None)


def layer_norm(x, gamma, beta, eps=1e-5):
    return LayerNormFn.apply(x, gamma, beta, eps)


def layer_norm_fork(x, gamma, beta, eps=1e-5):
    """(LayerNorm(x), x) — use the second value for the residual branch (see LayerNormForkFn)."""
    return LayerNormForkFn.apply(x, gamma, beta, eps)


def add_layer_norm(x, bias, residual, gamma, beta, eps=1e-5):
    return AddLayerNormFn.apply(x, bias, residual, gamma, beta, eps)


def bias_residual_add(x, bias, residual):
    return BiasResidualAddFn.apply(x, bias, residual)


def geglu_fwd(proj: torch.Tensor) -> torch.Tensor:
    proj = _rows2d(proj, "proj")
    inner = proj.shape[-1] // 2
    rows = proj.numel() // (2 * inner)
    out = torch.empty(proj.shape[:-1] + (inner,), device=proj.device, dtype=torch.float16)
    a = native.GegluArgs()
    a.proj, a.out, a.rows, a.inner = proj.data_ptr(), out.data_ptr(), rows, inner
    with _timed("geglu_fwd", (rows, inner)):
        native.check(native.load().sta_geglu_fwd(C.byref(a), _stream()), "sta_geglu_fwd")
    LAUNCHES["geglu_fwd"] += 1
    return out


def geglu_bwd(proj: torch.Tensor, d_out: torch.Tensor) -> torch.Tensor:
    d_out = _rows2d(d_out if d_out.dtype == torch.float16 else d_out.to(torch.float16), "d_out")
    inner = proj.shape[-1] // 2
    rows = proj.numel() // (2 * inner)
    d_proj = torch.empty_like(proj)
    a = native.GegluArgs()
    a.proj, a.d_out, a.out, a.rows, a.inner = proj.data_ptr(), d_out.data_ptr(), d_proj.data_ptr(), rows, inner
    with _timed("geglu_bwd", (rows, inner)):
        native.check(native.load().sta_geglu_bwd(C.byref(a), _stream()), "sta_geglu_bwd")
    LAUNCHES["geglu_bwd"] += 1
    return d_proj


class GegluFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, proj):
        proj = _rows2d(proj, "proj")
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(proj)
        return geglu_fwd(proj)

    @staticmethod
    def backward(ctx, d_out):
        (proj,) = ctx.saved_tensors
        return geglu_bwd(proj, d_out)


def geglu(proj):
    """value * gelu(gate) of a [.., 2*inner] projection (value | gate), fp16 CUDA."""
    return GegluFn.apply(proj)


# ------------------------------------------------------------------------------------------------------
# nearest x2 upsampling, NHWC fp16
# ------------------------------------------------------------------------------------------------------
def _upsample2x(x: torch.Tensor, backward: bool) -> torch.Tensor:
    _require(x, "x")
    b, c, h, w = x.shape
    x = _nhwc(x)
    if backward:
        if h % 2 or w % 2:
            raise RuntimeError("upsample2x backward needs even spatial sizes")
        lo_h, lo_w = h // 2, w // 2
        out = torch.empty((b, c, lo_h, lo_w), device=x.device, dtype=torch.float16, memory_format=torch.channels_last)
    else:
        lo_h, lo_w = h, w
        out = torch.empty((b, c, 2 * h, 2 * w), device=x.device, dtype=torch.float16, memory_format=torch.channels_last)
    a = native.Upsample2xArgs()
    a.x, a.out, a.batch, a.height, a.width, a.channels = x.data_ptr(), out.data_ptr(), b, lo_h, lo_w, c
    kind = "upsample2x_bwd" if backward else "upsample2x_fwd"
    with _timed(kind, (b, lo_h * lo_w, c)):
        fn = native.load().sta_upsample2x_bwd if backward else native.load().sta_upsample2x_fwd
        native.check(fn(C.byref(a), _stream()), "sta_" + kind)
    LAUNCHES[kind] += 1
    return out


class Upsample2xFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        return _upsample2x(x, False)

    @staticmethod
    def backward(ctx, d_out):
        return _upsample2x(d_out if d_out.dtype == torch.float16 else d_out.to(torch.float16), True)


def upsample_nearest2x(x):
    """F.interpolate(x, scale_factor=2, mode="nearest") for fp16 CUDA images (NHWC is free, NCHW is converted once)."""
    return Upsample2xFn.apply(x)


# ------------------------------------------------------------------------------------------------------
# elementwise part of a PLMS / DDIM sampler step (csrc/sta_sampler.cu)
# ------------------------------------------------------------------------------------------------------
def plms_step_usable(eps: torch.Tensor, x: torch.Tensor, olds) -> bool:
    """The fused kernel takes dense fp32 CUDA tensors (what the CUDA-graph UNet evaluation returns); anything else (fp16 eps
    of the eager autocast path, CPU stand-in models of the host tests) stays on the sampler's torch expressions."""
    ts = [eps, x, *olds]
    return (all(t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() for t in ts) and eps.shape[0] == 2 * x.shape[0]
            and (x[0].numel() % 4) == 0 and all(t.shape == x.shape for t in olds) and len(olds) <= 3)


class PlmsStepFn(torch.autograd.Function):
    """(x_prev, e_t, pred_x0) of include/sta_b200.h::sta_plms_step_fwd; gradients flow to eps, x and the old estimates."""

    @staticmethod
    def forward(ctx, eps, x, old0, old1, old2, guidance, w_e, w_old, a_x, a_e, p_x, p_e):
        olds = [old0, old1, old2]
        B, elems = x.shape[0], x[0].numel()
        e_t, x_prev, pred = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
        a = native.PlmsStepArgs()
        a.eps, a.x, a.e_t, a.x_prev, a.pred_x0 = eps.data_ptr(), x.data_ptr(), e_t.data_ptr(), x_prev.data_ptr(), pred.data_ptr()
        for k in range(3):
            a.old[k] = olds[k].data_ptr() if olds[k] is not None else None
            a.w_old[k] = float(w_old[k]) if olds[k] is not None else 0.0
        a.prompts, a.elems = B, elems
        a.guidance, a.w_e, a.a_x, a.a_e, a.p_x, a.p_e = float(guidance), float(w_e), float(a_x), float(a_e), float(p_x), float(p_e)
        with _timed("plms_step_fwd", (B, elems)):
            native.check(native.load().sta_plms_step_fwd(C.byref(a), _stream()), "sta_plms_step_fwd")
        LAUNCHES["plms_step_fwd"] += 1
        ctx.consts = (B, elems, float(guidance), float(w_e), [float(w) for w in w_old], float(a_x), float(a_e))
        ctx.has_old = [o is not None for o in olds]
        ctx.mark_non_differentiable(pred)
        ctx.set_materialize_grads(False)
        return x_prev, e_t, pred

    @staticmethod
    def backward(ctx, g_x_prev, g_e_t, _g_pred):
        B, elems, guidance, w_e, w_old, a_x, a_e = ctx.consts
        ref = g_x_prev if g_x_prev is not None else g_e_t
        if ref is None:
            return (None,) * 12
        g_x_prev = g_x_prev.contiguous() if g_x_prev is not None else None
        g_e_t = g_e_t.contiguous() if g_e_t is not None else None
        g_eps = torch.empty((2 * B,) + tuple(ref.shape[1:]), device=ref.device, dtype=torch.float32)
        g_x = torch.empty_like(ref)
        g_old = [torch.empty_like(ref) if h else None for h in ctx.has_old]
        a = native.PlmsStepBwdArgs()
        a.g_x_prev = g_x_prev.data_ptr() if g_x_prev is not None else None
        a.g_e_t = g_e_t.data_ptr() if g_e_t is not None else None
        a.g_eps, a.g_x = g_eps.data_ptr(), g_x.data_ptr()
        for k in range(3):
            a.g_old[k] = g_old[k].data_ptr() if g_old[k] is not None else None
            a.w_old[k] = w_old[k] if g_old[k] is not None else 0.0
        a.prompts, a.elems, a.guidance, a.w_e, a.a_x, a.a_e = B, elems, guidance, w_e, a_x, a_e
        with _timed("plms_step_bwd", (B, elems)):
            native.check(native.load().sta_plms_step_bwd(C.byref(a), _stream()), "sta_plms_step_bwd")
        LAUNCHES["plms_step_bwd"] += 1
        return (g_eps, g_x, g_old[0], g_old[1], g_old[2]) + (None,) * 7


def plms_step(eps, x, olds, guidance, w_e, w_old, a_x, a_e, p_x, p_e):
    """One fused sampler step: eps [2B, ...] (uncond rows first), x [B, ...], olds = [e_{t-1}, e_{t-2}, e_{t-3}][:k]."""
    o = list(olds) + [None] * (3 - len(olds))
    w = list(w_old) + [0.0] * (3 - len(w_old))
    return PlmsStepFn.apply(eps, x, o[0], o[1], o[2], guidance, w_e, w, a_x, a_e, p_x, p_e)
