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
LAUNCHES = {"sattn_fwd": 0, "sattn_bwd": 0, "xattn_fwd": 0, "xattn_bwd": 0, "groupnorm_fwd": 0, "groupnorm_bwd": 0}


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
    lse = torch.zeros((B, heads, 2 + n_obj, n), device=q.device, dtype=torch.float32) if need_lse else None
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
              scale: Optional[float] = None) -> Tuple[torch.Tensor, Optional[torch.Tensor]]:
    """Backward of xattn_fwd: returns (d_q fp16 [2B, n, C], d_coef f32 [B, n_obj] | None)."""
    _require(q, "q")
    _require(d_out, "d_out")
    _require(lse, "lse", torch.float32)
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
    with _timed("xattn_bwd", (B, n, heads, d, n_obj)):
        native.check(native.load().sta_xattn_bwd(C.byref(a), _stream()), "sta_xattn_bwd")
    LAUNCHES["xattn_bwd"] += 1
    return d_q, d_coef


def sattn_bwd(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, out: torch.Tensor, lse: torch.Tensor,
              d_out: torch.Tensor, heads: int, scale: Optional[float] = None):
    """Backward of sattn_fwd: returns (d_q, d_k, d_v) fp16 [b, n, C]."""
    for t, nm in ((q, "q"), (k, "k"), (v, "v"), (out, "out"), (d_out, "d_out")):
        _require(t, nm)
    _require(lse, "lse", torch.float32)
    b, n, c = q.shape
    d = c // heads
    scale = float(d ** -0.5) if scale is None else float(scale)
    d_qkv = torch.empty((3, b, n, c), device=q.device, dtype=torch.float16)
    dq_accum = torch.empty((b, n, c), device=q.device, dtype=torch.float32)
    delta = torch.empty((b, heads, n), device=q.device, dtype=torch.float32)
    a = native.SattnBwdArgs()
    a.q, a.k, a.v, a.out, a.d_out = q.data_ptr(), k.data_ptr(), v.data_ptr(), out.data_ptr(), d_out.data_ptr()
    a.lse = lse.data_ptr()
    a.d_q, a.d_k, a.d_v = d_qkv[0].data_ptr(), d_qkv[1].data_ptr(), d_qkv[2].data_ptr()
    a.dq_accum, a.delta = dq_accum.data_ptr(), delta.data_ptr()
    a.batch, a.n, a.heads, a.head_dim = b, n, heads, d
    a.q_token_stride, a.q_batch_stride = _token_major(q, "q")
    a.k_token_stride, a.k_batch_stride = _token_major(k, "k")
    a.v_token_stride, a.v_batch_stride = _token_major(v, "v")
    a.o_token_stride, a.o_batch_stride = _token_major(out, "out")
    a.do_token_stride, a.do_batch_stride = _token_major(d_out, "d_out")
    a.scale = scale
    with _timed("sattn_bwd", (b, n, heads, d)):
        native.check(native.load().sta_sattn_bwd(C.byref(a), _stream()), "sta_sattn_bwd")
    LAUNCHES["sattn_bwd"] += 3  # delta, main, dq cast
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
            ctx.save_for_backward(q, k_ctx, v_ctx, mask, coef, lse)
            ctx.heads = heads
        return out

    @staticmethod
    def backward(ctx, d_out):
        q, k_ctx, v_ctx, mask, coef, lse = ctx.saved_tensors
        if d_out.stride(2) != 1:
            d_out = d_out.contiguous()
        d_q, d_coef = xattn_bwd(q, k_ctx, v_ctx, mask, coef, lse, d_out.to(torch.float16), ctx.heads)
        return d_q, None, None, None, d_coef, None


def self_attention(q, k, v, heads):
    return SelfAttentionFn.apply(q, k, v, heads)


def dual_cross_attention(q, k_ctx, v_ctx, mask, coef, heads):
    return DualCrossAttentionFn.apply(q, k_ctx, v_ctx, mask, coef, heads)


# ------------------------------------------------------------------------------------------------------
# fused GroupNorm(32) [+ SiLU], NHWC fp16
# ------------------------------------------------------------------------------------------------------
def _nhwc(x: torch.Tensor) -> torch.Tensor:
    """[B, C, H, W] tensor whose memory is NHWC-dense (channels_last); copies only if it is not already."""
    return x if x.is_contiguous(memory_format=torch.channels_last) else x.contiguous(memory_format=torch.channels_last)


def groupnorm_fwd(x: torch.Tensor, gamma: torch.Tensor, beta: torch.Tensor, eps: float, silu: bool):
    _require(x, "x")
    _require(gamma, "gamma", torch.float32)
    _require(beta, "beta", torch.float32)
    b, c, h, w = x.shape
    x = _nhwc(x)
    out = torch.empty_like(x, memory_format=torch.channels_last)
    stats = torch.empty((b, 32, 2), device=x.device, dtype=torch.float32)
    a = native.GroupNormArgs()
    a.x, a.gamma, a.beta, a.out, a.stats = x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), out.data_ptr(), stats.data_ptr()
    a.batch, a.hw, a.channels, a.silu, a.eps = b, h * w, c, int(silu), float(eps)
    with _timed("groupnorm_fwd", (b, h * w, c)):
        native.check(native.load().sta_groupnorm_fwd(C.byref(a), _stream()), "sta_groupnorm_fwd")
    LAUNCHES["groupnorm_fwd"] += 2
    return out, stats, x


def groupnorm_bwd(x, d_out, gamma, beta, stats, eps: float, silu: bool):
    b, c, h, w = x.shape
    d_out = _nhwc(d_out if d_out.dtype == torch.float16 else d_out.to(torch.float16))
    d_x = torch.empty_like(x, memory_format=torch.channels_last)
    bstats = torch.empty((b, 32, 2), device=x.device, dtype=torch.float32)
    a = native.GroupNormArgs()
    a.x, a.d_out, a.gamma, a.beta, a.out = x.data_ptr(), d_out.data_ptr(), gamma.data_ptr(), beta.data_ptr(), d_x.data_ptr()
    a.stats, a.bwd_stats = stats.data_ptr(), bstats.data_ptr()
    a.batch, a.hw, a.channels, a.silu, a.eps = b, h * w, c, int(silu), float(eps)
    with _timed("groupnorm_bwd", (b, h * w, c)):
        native.check(native.load().sta_groupnorm_bwd(C.byref(a), _stream()), "sta_groupnorm_bwd")
    LAUNCHES["groupnorm_bwd"] += 2
    return d_x


class GroupNormSiLUFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps, silu):
        out, stats, x_nhwc = groupnorm_fwd(x, gamma, beta, eps, silu)
        if ctx.needs_input_grad[0]:
            ctx.save_for_backward(x_nhwc, gamma, beta, stats)
            ctx.eps, ctx.silu = eps, silu
        return out

    @staticmethod
    def backward(ctx, d_out):
        x, gamma, beta, stats = ctx.saved_tensors
        return groupnorm_bwd(x, d_out, gamma, beta, stats, ctx.eps, ctx.silu), None, None, None, None


def group_norm_silu(x, gamma, beta, eps=1e-5, silu=True):
    """GroupNorm(32)(x.float()).half() [-> SiLU] in one pass; x fp16 [B, C, H, W] (any layout, NHWC is free)."""
    return GroupNormSiLUFn.apply(x, gamma, beta, eps, silu)
