"""Drop-in for the reference `ldm.modules.attention` hook API, backed by the sm_100a kernels in libsta_b200.so.

Same class names, constructor/forward signatures and parameter names as
/root/reference/attention_optimization/stable-diffusion/ldm/modules/attention.py (CrossAttention :157-215,
BasicTransformerBlock :223-300, SpatialTransformer :303-345, FeedForward :52-69, GEGLU :42-49), so an
sd-v1-4 `state_dict` loads unchanged.  What differs is how a block evaluates:

  reference (attention.py:268-300)                      here
  ------------------------------------------------      ---------------------------------------------------
  attn1: 3 Linear + bmm + softmax + bmm ([H,N,N] in HBM) one fused QKV GEMM + sta_sattn_fwd (flash, tcgen05)
  attn2 called 1 + n_obj times, each re-running          to_q once; K/V of the (frozen) contexts cached per
  to_q(norm2(x)), to_k, to_v, to_out                     prompt; ONE sta_xattn_fwd; ONE to_out after the blend
  ~5 elementwise launches per object for the blend       blend folded into the kernel (row-scaled P operand)
  `time == 981` on a CUDA scalar: a host sync per block  step index is a Python int; no sync
  local embeddings read from c{i}_*.pt at step 0         passed in memory (disk protocol kept as a fallback
                                                         for unmodified callers)

There is no PyTorch fallback for the attention math: without the CUDA library (or on CPU tensors) calls raise.
"""
from __future__ import annotations

import math
import os
from typing import List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F
from torch import nn

from ... import ops

mode = "fix_radius_0p2"  # reference attention.py:14
MASK_RADIUS_SQ = 0.04    # r = 0.2 (attention.py:261)


def exists(val):
    return val is not None


def default(val, d):
    return val if exists(val) else (d() if callable(d) else d)


def zero_module(module):
    for p in module.parameters():
        p.detach().zero_()
    return module


def Normalize(in_channels):
    return torch.nn.GroupNorm(num_groups=32, num_channels=in_channels, eps=1e-6, affine=True)


def build_object_masks(bboxs_curr: Sequence[Sequence[float]], n_tokens: int, device) -> torch.Tensor:
    """uint8 [n_obj, n_tokens]: 1 where (col/dim - x)^2 + (row/dim - y)^2 < 0.04.

    Evaluated with the reference's exact fp32 torch expression (attention.py:254-261) on the host, so boundary
    pixels agree bit for bit; the kernel only ever sees the byte mask.
    """
    dim = int(math.isqrt(n_tokens))
    if dim * dim != n_tokens:
        raise ValueError("the layout masks assume a square latent (reference attention.py:243)")
    out = torch.zeros(len(bboxs_curr), n_tokens, dtype=torch.uint8)
    axis = torch.arange(dim, dtype=torch.float32) / dim
    for i, box in enumerate(bboxs_curr):
        dist1 = (axis - box[0]) ** 2
        dist2 = (axis - box[1]) ** 2
        out[i] = (dist1.unsqueeze(0) + dist2.unsqueeze(1) < MASK_RADIUS_SQ).reshape(-1).to(torch.uint8)
    return out.to(device)


def layout_shape(bboxs_curr) -> Tuple[bool, int]:
    """(per_prompt, n_obj) of a layout argument: `[n_obj][2]` (one layout shared by the batch, the reference's form) or
    `[B][n_obj][2]` (one layout per prompt; the inner lists are EMPTY for prompts without objects)."""
    if not bboxs_curr:
        return False, 0
    first = bboxs_curr[0]
    if isinstance(first, (list, tuple)) and (len(first) == 0 or isinstance(first[0], (list, tuple))):
        return True, len(first)
    return False, len(bboxs_curr)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, dim_out=None, mult=4, glu=False, dropout=0.0):
        super().__init__()
        inner_dim = int(dim * mult)
        dim_out = default(dim_out, dim)
        project_in = nn.Sequential(nn.Linear(dim, inner_dim), nn.GELU()) if not glu else GEGLU(dim, inner_dim)
        self.net = nn.Sequential(project_in, nn.Dropout(dropout), nn.Linear(inner_dim, dim_out))

    def forward(self, x):
        return self.net(x)


_LN_MIXED = None


def _layer_norm(norm: nn.LayerNorm, x: torch.Tensor) -> torch.Tensor:
    """LayerNorm of an fp16 activation with fp32 statistics and fp32 affine parameters WITHOUT autocast's fp32 input
    copy and fp32 output (the next Linear would cast it back to fp16 anyway): same rounding point, two copies fewer.
    Falls back to the module call when this torch build has no mixed-dtype layer_norm kernel."""
    global _LN_MIXED
    if x.is_cuda and x.dtype == torch.float16 and norm.weight.dtype == torch.float32 and _LN_MIXED is not False:
        try:
            with torch.autocast("cuda", enabled=False):
                y = F.layer_norm(x, norm.normalized_shape, norm.weight, norm.bias, norm.eps)
            _LN_MIXED = True
            return y
        except RuntimeError:
            if _LN_MIXED:  # it worked before: a real error
                raise
            _LN_MIXED = False
    return norm(x)


def _fp16(t: torch.Tensor) -> torch.Tensor:
    return t if t.dtype == torch.float16 else t.to(torch.float16)


def cached_sum(owner: nn.Module, name: str, params, dtype) -> torch.Tensor:
    """Sum of (frozen) parameters in `dtype`, cached on `owner` and rebuilt only when one of them changed or moved.
    Used for the fp32 bias / LayerNorm vectors the fused kernels read (no per-evaluation casts inside a CUDA graph)."""
    params = [p for p in params if p is not None]
    key = (dtype,) + tuple((p.data_ptr(), p._version, p.device) for p in params)
    slot = owner.__dict__.setdefault("_sta_cached", {})
    hit = slot.get(name)
    if hit is None or hit[0] != key:
        with torch.no_grad():
            t = params[0].detach().to(dtype)
            for p in params[1:]:
                t = t + p.detach().to(dtype)
            if t.dim() <= 2:  # conv weights keep their (channels_last) memory format
                t = t.contiguous()
            if hit is not None and _same_storage_class(hit[1], t):
                # a weight changed (load_state_dict, broadcast, .to()): refresh IN PLACE so that CUDA graphs captured
                # over the old tensor (graphed.py) keep reading valid, current data
                hit[1].copy_(t)
                t = hit[1]
        slot[name] = hit = (key, t)
    return hit[1]


def _same_storage_class(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a.shape == b.shape and a.dtype == b.dtype and a.device == b.device


def _fused_ok(x: torch.Tensor) -> bool:
    return x.is_cuda and x.dtype == torch.float16


def _frozen(module: nn.Module) -> bool:
    """The fused paths read cached, detached copies of the weights (sampling / alpha optimisation: the UNet is frozen,
    reference ddpm.py:519-523); a module with trainable parameters takes the plain autograd path instead."""
    if not torch.is_grad_enabled():
        return True
    return not any(p.requires_grad for p in module.parameters())


def nhwc_tokens(x: torch.Tensor) -> torch.Tensor:
    """[B, C, H, W] -> [B, H*W, C] ('b c h w -> b (h w) c'); a free view when x is channels_last."""
    b, c, h, w = x.shape
    return x.permute(0, 2, 3, 1).reshape(b, h * w, c)


def tokens_nhwc(t: torch.Tensor, h: int, w: int) -> torch.Tensor:
    """[B, H*W, C] -> channels_last [B, C, H, W] view ('b (h w) c -> b c h w')."""
    b, _, c = t.shape
    return t.reshape(b, h, w, c).permute(0, 3, 1, 2)


class CrossAttention(nn.Module):
    """Same parameters as the reference module (to_q / to_k / to_v bias-free, to_out.0 with bias)."""

    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64, dropout=0.0):
        super().__init__()
        inner_dim = dim_head * heads
        context_dim = default(context_dim, query_dim)
        self.scale = dim_head ** -0.5
        self.heads = heads
        self.to_q = nn.Linear(query_dim, inner_dim, bias=False)
        self.to_k = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_v = nn.Linear(context_dim, inner_dim, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner_dim, query_dim), nn.Dropout(dropout))

    # -- building blocks used by BasicTransformerBlock -------------------------------------------------
    def _fused_qkv_weight(self) -> torch.Tensor:
        """fp16 [3C, C] concatenation of to_q/to_k/to_v, rebuilt only when one of them changed (frozen in sampling)."""
        ws = (self.to_q.weight, self.to_k.weight, self.to_v.weight)
        key = tuple((w.data_ptr(), w._version, w.device) for w in ws)
        if getattr(self, "_wqkv_key", None) != key:
            with torch.no_grad():
                w = torch.cat([w.detach() for w in ws], dim=0).to(torch.float16)
                prev = getattr(self, "_wqkv", None)
                if prev is not None and _same_storage_class(prev, w):
                    prev.copy_(w)  # in place: captured CUDA graphs keep a valid pointer (see cached_sum)
                else:
                    self._wqkv = w
            self._wqkv_key = key
        return self._wqkv

    def self_attention_core(self, x: torch.Tensor) -> torch.Tensor:
        """softmax(q k^T * scale) v for context = x: one [C, 3C] GEMM, then the flash kernel on strided views."""
        qkv = F.linear(_fp16(x), self._fused_qkv_weight())
        return ops.self_attention_qkv(qkv, self.heads)  # d(qkv) comes back as ONE buffer (no cat in backward)

    def project_out(self, a: torch.Tensor) -> torch.Tensor:
        """to_out on a kernel output (fp16); under autocast the Linear casts itself, otherwise match the weights."""
        if not torch.is_autocast_enabled() and a.dtype != self.to_out[0].weight.dtype:
            a = a.to(self.to_out[0].weight.dtype)
        return self.to_out(a)

    @torch.no_grad()
    def project_contexts(self, contexts: torch.Tensor):
        """to_k / to_v of frozen contexts [B, S, L, ctx_dim] -> two fp16 [B, S, L, C] tensors (cached by callers)."""
        w_dtype = self.to_k.weight.dtype
        contexts = contexts if torch.is_autocast_enabled() else contexts.to(w_dtype)
        k = _fp16(F.linear(contexts, self.to_k.weight)).contiguous()
        v = _fp16(F.linear(contexts, self.to_v.weight)).contiguous()
        return k, v

    def forward(self, x, context=None, mask=None, self_attention_region=None):
        """Plain (single-context) attention with the reference signature (attention.py:175).

        `mask` / `self_attention_region` are never passed by the reference's own callers (dead code,
        attention.py:187-191,199-213) and are rejected here rather than silently ignored.
        """
        if mask is not None or self_attention_region is not None:
            raise NotImplementedError("mask / self_attention_region are dead code in the reference and not built")
        if context is None:
            out = self.self_attention_core(x)
        else:
            # generic cross-attention = the fused kernel with no objects: both halves attend to their own context
            b = x.shape[0]
            q = _fp16(self.to_q(x))
            k, v = self.project_contexts(context.unsqueeze(1))  # [b, 1, L, C]
            # treat every row as a "conditional" row of its own prompt: slots (unused uncond, ctx)
            q2 = torch.cat([q, q], dim=0)
            k2 = torch.cat([k, k], dim=1)
            v2 = torch.cat([v, v], dim=1)
            out = ops.dual_cross_attention(q2, k2, v2, None, None, self.heads)[b:]
        return self.project_out(out)


class BasicTransformerBlock(nn.Module):
    """Transformer block with the dual (global + per-object local) cross-attention and alpha-blend."""

    def __init__(self, dim, n_heads, d_head, dropout=0.0, context_dim=None, gated_ff=True, checkpoint=False):
        super().__init__()
        self.attn1 = CrossAttention(query_dim=dim, heads=n_heads, dim_head=d_head, dropout=dropout)
        self.ff = FeedForward(dim, dropout=dropout, glu=gated_ff)
        self.attn2 = CrossAttention(query_dim=dim, context_dim=context_dim, heads=n_heads, dim_head=d_head,
                                    dropout=dropout)
        self.norm1 = nn.LayerNorm(dim)
        self.norm2 = nn.LayerNorm(dim)
        self.norm3 = nn.LayerNorm(dim)
        # The reference checkpoints every block (attention.py:224,266) to fit 48 GB; with flash-style saved state
        # and 180 GB of HBM the default here is to keep activations (no recompute in backward).
        self.checkpoint = checkpoint
        # reference: torch.load("uncond_fix_radius_0p2_g0.pt") at construction (attention.py:234).  Only row 1 of
        # gs[i] is ever used (attention.py:290), so this tensor does not influence the output; it is loaded when
        # present for attribute compatibility.
        self.uncond = None
        path = "uncond_%s_g0.pt" % mode
        if os.path.exists(path):
            try:
                self.uncond = torch.load(path, map_location="cpu")
            except Exception:  # noqa: BLE001 - the shipped file is pickled on cuda:0
                self.uncond = None
        self.defer_ff_bias = False  # set by SpatialTransformer._forward_fused for its last block (see _forward_fused)
        self.first_timestep = 981  # reference hard-codes 981 (attention.py:240); samplers overwrite per schedule
        self.local_contexts: Optional[List[torch.Tensor]] = None  # in-memory c_i [B, 77, ctx_dim] (or [1, ...])
        self._cache = None
        self._caches = {}

    # -- per-prompt state ------------------------------------------------------------------------------
    # One cache per (prompts B, n_obj, tokens N) signature, holding the projected context K/V [B, 2+n_obj, L, C]
    # and the byte masks [B, n_obj, N].  The buffers are allocated once and refreshed IN PLACE, so CUDA graphs
    # captured over this block (graphed.py) keep pointing at valid, current data.
    def set_local_contexts(self, local_contexts: Optional[Sequence[torch.Tensor]]):
        self.local_contexts = list(local_contexts) if local_contexts is not None else None
        self.reset_cache()

    def reset_cache(self):
        for c in self._caches.values():
            c["stale"] = True

    def _load_local_contexts_from_disk(self, n_obj: int, device) -> List[torch.Tensor]:
        """The reference's transport for local embeddings: c{i}_fix_radius_0p2_g{id}.pt in CWD (attention.py:246)."""
        try:
            from process_id import NON_EXISTING_NAME_ID  # type: ignore
        except Exception:  # noqa: BLE001
            NON_EXISTING_NAME_ID = 0
        return [torch.load("c%d_%s_g%d.pt" % (i, mode, NON_EXISTING_NAME_ID), map_location=device) for i in range(n_obj)]

    @torch.no_grad()
    def _fill_cache(self, key, context, bboxs_curr):
        B, n_obj, n = key
        dev = context.device
        locs = self.local_contexts
        if n_obj and locs is None:
            locs = self._load_local_contexts_from_disk(n_obj, dev)
        slots = [context[:B], context[B:]]
        for i in range(n_obj):
            c_i = locs[i].to(device=dev, dtype=context.dtype)
            if c_i.dim() == 2:
                c_i = c_i.unsqueeze(0)
            slots.append(c_i.expand(B, -1, -1))
        k_ctx, v_ctx = self.attn2.project_contexts(torch.stack(slots, dim=1))  # [B, 2 + n_obj, L, C]
        masks = None
        if n_obj:
            if layout_shape(bboxs_curr)[0]:  # per-prompt layouts: [B][n_obj][2]
                masks = torch.stack([build_object_masks(bb, n, dev) for bb in bboxs_curr])
            else:
                masks = build_object_masks(bboxs_curr, n, dev).unsqueeze(0).expand(B, -1, -1).contiguous()
        c = self._caches.get(key)
        if c is None or c["k"].shape != k_ctx.shape:
            c = {"k": k_ctx, "v": v_ctx, "masks": masks, "n_obj": n_obj, "n": n, "B": B}
            self._caches[key] = c
        else:
            c["k"].copy_(k_ctx)
            c["v"].copy_(v_ctx)
            if n_obj:
                c["masks"].copy_(masks)
        c["stale"] = False
        return c

    def refresh_cache(self, context, bboxs_curr, batch2):
        """Rebuild, in place, every existing cache of this (B, n_obj) signature (graphed.py calls this per prompt)."""
        B = batch2 // 2
        n_obj = layout_shape(bboxs_curr)[1]
        for key in list(self._caches.keys()):
            if key[0] == B and key[1] == n_obj:
                self._fill_cache(key, context, bboxs_curr)

    # -- forward ---------------------------------------------------------------------------------------
    def forward(self, x, context=None, time=None, text_index=None, coef=None, bboxs_curr=None):
        if context is None:
            raise ValueError("BasicTransformerBlock needs the text context (SD-v1 always passes one)")
        n_obj = layout_shape(bboxs_curr)[1]
        if torch.is_tensor(time):
            time = int(time.item())  # unmodified callers pass timesteps[0] (a device scalar): one sync, as upstream
        # The reference rebuilds masks / local contexts when `time == 981` (attention.py:240); here the projected
        # K/V are cached as well and refreshed at the schedule's first timestep or after reset_cache() /
        # set_local_contexts() (what the samplers in this package call once per prompt).
        key = (x.shape[0] // 2, n_obj, x.shape[1])
        c = self._caches.get(key)
        if c is None or c["stale"] or time == self.first_timestep:
            c = self._fill_cache(key, context, bboxs_curr)
        self._cache = c
        if self.checkpoint and torch.is_grad_enabled() and x.shape[1] >= getattr(self, "checkpoint_min_tokens", 0):
            from torch.utils.checkpoint import checkpoint as _ckpt

            return _ckpt(self._forward, x, context, coef, use_reentrant=False)
        return self._forward(x, context, coef)

    def _dropout_free(self) -> bool:
        return not self.training or all(m.p == 0.0 for m in self.modules() if isinstance(m, nn.Dropout))

    def _use_fused(self, x) -> bool:
        return _fused_ok(x) and isinstance(self.ff.net[0], GEGLU) and self._dropout_free() and _frozen(self)

    def _forward(self, x, context=None, coef=None, bboxs_curr_input=None):
        if self._use_fused(x):
            return self._forward_fused(x, coef)
        return self._forward_unfused(x, coef)

    def _coef_rows(self, coef, B, n_obj, device):
        if not n_obj:
            return None
        coef_b = coef.reshape(-1, n_obj).to(device=device, dtype=torch.float32)
        if coef_b.shape[0] != B:
            coef_b = coef_b.expand(B, n_obj)
        return coef_b.contiguous()

    def _forward_fused(self, x, coef):
        """attention.py:268-300 with every elementwise / normalisation step between the GEMMs fused (csrc/sta_tokens.cu):
        LN1 | QKV GEMM | flash self-attention | to_out GEMM | bias+residual+LN2 | to_q GEMM | fused dual cross-attention |
        to_out GEMM | bias+residual+LN3 | GEGLU proj GEMM (+bias) | gate | out GEMM | bias+residual.  13 launches."""
        c = self._cache
        B, n_obj = c["B"], c["n_obj"]
        a1, a2, ff = self.attn1, self.attn2, self.ff
        f32 = torch.float32

        def ln(norm):
            return cached_sum(self, "g" + str(id(norm)), [norm.weight], f32), cached_sum(self, "b" + str(id(norm)), [norm.bias], f32)

        def w16(lin):
            return lin.weight if lin.weight.dtype == torch.float16 else cached_sum(self, "w" + str(id(lin)), [lin.weight], torch.float16)

        x = x if x.is_contiguous() else x.contiguous()
        g1, b1 = ln(self.norm1)
        g2, b2 = ln(self.norm2)
        g3, b3 = ln(self.norm3)
        with torch.autocast("cuda", enabled=False):
            n1, x = ops.layer_norm_fork(x, g1, b1, self.norm1.eps)  # x: residual stream; one backward launch for both uses
            sa = ops.self_attention_qkv(F.linear(n1, a1._fused_qkv_weight()), a1.heads)
            t = F.linear(sa, w16(a1.to_out[0]))
            x1, n2 = ops.add_layer_norm(t, cached_sum(self, "bo1", [a1.to_out[0].bias], f32), x, g2, b2, self.norm2.eps)
            q = F.linear(n2, w16(a2.to_q))  # computed ONCE (the reference recomputes it 1 + n_obj times)
            blended = ops.dual_cross_attention(q, c["k"], c["v"], c["masks"], self._coef_rows(coef, B, n_obj, x.device),
                                               a2.heads)
            t = F.linear(blended, w16(a2.to_out[0]))
            x2, n3 = ops.add_layer_norm(t, cached_sum(self, "bo2", [a2.to_out[0].bias], f32), x1, g3, b3, self.norm3.eps)
            proj, out = ff.net[0].proj, ff.net[2]
            pb = proj.bias if proj.bias.dtype == torch.float16 else cached_sum(self, "pb", [proj.bias], torch.float16)
            gated = ops.geglu(F.linear(n3, w16(proj), pb))
            if self.defer_ff_bias:
                # last block of a SpatialTransformer: the residual rides in the GEMM (D = gated W^T + x2, fp32 accumulate) and
                # the bias is folded into proj_out's bias by the caller (proj_out is affine) — one launch fewer per block
                return torch.addmm(x2.reshape(-1, x2.shape[-1]), gated.reshape(-1, gated.shape[-1]), w16(out).t()).view_as(x2)
            t = F.linear(gated, w16(out))
            return ops.bias_residual_add(t, cached_sum(self, "bo3", [out.bias], f32), x2)

    def _forward_unfused(self, x, coef):
        c = self._cache
        B, n_obj = c["B"], c["n_obj"]
        a1 = self.attn1
        x = a1.project_out(a1.self_attention_core(_layer_norm(self.norm1, x))) + x  # attention.py:274
        a2 = self.attn2
        q = _fp16(a2.to_q(_layer_norm(self.norm2, x)))  # computed ONCE (the reference recomputes it 1 + n_obj times)
        coef_b = None
        if n_obj:
            coef_b = coef.reshape(-1, n_obj).to(device=x.device, dtype=torch.float32)
            if coef_b.shape[0] != B:
                coef_b = coef_b.expand(B, n_obj)
            coef_b = coef_b.contiguous()
        blended = ops.dual_cross_attention(q, c["k"], c["v"], c["masks"], coef_b, a2.heads)  # :278-294, pre-to_out
        x = a2.project_out(blended) + x  # :281 + :297 (to_out commutes with the blend)
        return self.ff(_layer_norm(self.norm3, x)) + x  # :299


class SpatialTransformer(nn.Module):
    """GroupNorm -> 1x1 conv -> tokens -> transformer block(s) -> image -> 1x1 conv + residual (attention.py:303-345)."""

    def __init__(self, in_channels, n_heads, d_head, depth=1, dropout=0.0, context_dim=None):
        super().__init__()
        self.in_channels = in_channels
        inner_dim = n_heads * d_head
        self.norm = Normalize(in_channels)
        self.proj_in = nn.Conv2d(in_channels, inner_dim, kernel_size=1, stride=1, padding=0)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner_dim, n_heads, d_head, dropout=dropout, context_dim=context_dim)
             for _ in range(depth)]
        )
        self.proj_out = zero_module(nn.Conv2d(inner_dim, in_channels, kernel_size=1, stride=1, padding=0))

    def _forward_fused(self, x, context, time, text_index, coef, bboxs_curr):
        """Same arithmetic with the two 1x1 convolutions as token GEMMs (NHWC memory IS the token matrix): proj_in's
        bias rides in the GEMM epilogue, proj_out's bias and the `+ x_in` are one fused pass — instead of two
        broadcast bias kernels (ATen adds the cuDNN conv bias separately) and a residual add."""
        b, c, h, w = x.shape
        f32 = torch.float32
        nw, nb = cached_sum(self, "gn_w", [self.norm.weight], f32), cached_sum(self, "gn_b", [self.norm.bias], f32)
        # x feeds the GroupNorm and the `+ x_in` residual: the fork hands both gradients to one GroupNorm backward launch
        xn, x = ops.group_norm_silu_fork(x, nw, nb, self.norm.eps, False)
        xn = nhwc_tokens(xn)
        x_in = nhwc_tokens(x)
        x_in = x_in if x_in.is_contiguous() else x_in.contiguous()
        w_in = cached_sum(self, "w_in", [self.proj_in.weight], torch.float16).reshape(self.proj_in.out_channels, c)
        w_out = cached_sum(self, "w_out", [self.proj_out.weight], torch.float16).reshape(c, -1)
        with torch.autocast("cuda", enabled=False):
            t = F.linear(xn, w_in, cached_sum(self, "b_in", [self.proj_in.bias], torch.float16))
        last = self.transformer_blocks[-1]
        defer = last._use_fused(t)
        for block in self.transformer_blocks:
            block.defer_ff_bias = defer and block is last
            t = block(t, context=context, time=time, text_index=text_index, coef=coef, bboxs_curr=bboxs_curr)
        b_out = self._deferred_out_bias(last, w_out) if defer else cached_sum(self, "b_out", [self.proj_out.bias], f32)
        with torch.autocast("cuda", enabled=False):
            t = F.linear(_fp16(t), w_out)
            out = ops.bias_residual_add(t, b_out, x_in)
        return tokens_nhwc(out, h, w)

    @torch.no_grad()
    def _deferred_out_bias(self, last, w_out) -> torch.Tensor:
        """proj_out.bias + W_out @ b_ff: the last block left out its FeedForward output bias b_ff (defer_ff_bias) and
        proj_out(y + b_ff) = proj_out(y) + W_out b_ff.  fp32, cached, refreshed in place when a weight changes."""
        params = (self.proj_out.bias, self.proj_out.weight, last.ff.net[2].bias)
        key = tuple((p.data_ptr(), p._version, p.device) for p in params)
        hit = self.__dict__.get("_sta_defer_bias")
        if hit is None or hit[0] != key:
            t = self.proj_out.bias.float() + w_out.float() @ last.ff.net[2].bias.float()
            if hit is not None and hit[1].shape == t.shape and hit[1].device == t.device:
                hit[1].copy_(t)
                t = hit[1]
            self.__dict__["_sta_defer_bias"] = hit = (key, t)
        return hit[1]

    def forward(self, x, context=None, time=None, text_index=None, coef=None, bboxs_curr=None):
        b, c, h, w = x.shape
        if _fused_ok(x) and _frozen(self):
            return self._forward_fused(x, context, time, text_index, coef, bboxs_curr)
        x_in = x
        if x.is_cuda and x.dtype == torch.float16:  # fused NHWC GroupNorm (fp32 statistics), no activation here
            nw = self.norm.weight if self.norm.weight.dtype == torch.float32 else self.norm.weight.float()
            nb = self.norm.bias if self.norm.bias.dtype == torch.float32 else self.norm.bias.float()
            x = ops.group_norm_silu(x, nw, nb, self.norm.eps, False)
        else:
            x = self.norm(x)
        x = self.proj_in(x)
        x = x.permute(0, 2, 3, 1).reshape(b, h * w, -1)  # 'b c h w -> b (h w) c'
        for block in self.transformer_blocks:
            block.defer_ff_bias = False
            x = block(x, context=context, time=time, text_index=text_index, coef=coef, bboxs_curr=bboxs_curr)
        x = x.reshape(b, h, w, -1).permute(0, 3, 1, 2)  # 'b (h w) c -> b c h w'
        x = self.proj_out(x)
        return x + x_in
