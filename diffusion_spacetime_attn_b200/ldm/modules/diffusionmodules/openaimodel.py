"""SD-v1 UNet with the reference's module tree and forward signature
(/root/reference/.../ldm/modules/diffusionmodules/openaimodel.py: TimestepEmbedSequential :74-88, ResBlock
:163-275, UNetModel :413-742), so `model.diffusion_model.*` checkpoint keys load unchanged.

Only what `configs/stable-diffusion/v1-inference.yaml:29-44` instantiates is built (spatial transformers at every
attention resolution, conv resampling, no class conditioning, no scale-shift norm).  The convolutions, GroupNorms
and the feed-forward stay library calls (cuDNN / cuBLAS through torch); the attention inside every
SpatialTransformer runs on the sm_100a kernels (ldm/modules/attention.py of this package).

Additions over the reference API (all optional, the reference call shape keeps working):
  * `timesteps[0]` is forwarded to the blocks as a Python int when the caller passes `step_time=` (no device sync);
  * `set_local_contexts()` / `reset_attention_cache()` hand the per-object embeddings to all 16 blocks in memory.
"""
from __future__ import annotations

from typing import Iterable, List, Optional, Sequence

import torch as th
import torch.nn as nn
import torch.nn.functional as F

from .... import ops as _ops
from ..attention import BasicTransformerBlock, SpatialTransformer, _frozen, cached_sum, nhwc_tokens, tokens_nhwc
from .util import checkpoint, conv_nd, linear, normalization, timestep_embedding, zero_module


_ATEN_UPSAMPLE = __import__("os").environ.get("STA_ATEN_UPSAMPLE") == "1"  # A/B timing knob


def _fusable(x) -> bool:
    return x.is_cuda and x.dtype == th.float16


def gn_silu(norm: nn.GroupNorm, x, silu: bool = True):
    """GroupNorm32 (+ SiLU) on an fp16 CUDA activation through the fused NHWC kernel (csrc/sta_groupnorm.cu): fp32
    statistics like the reference's `x.float()` path (util.py:214-216), one rounding to fp16."""
    w = norm.weight if norm.weight.dtype == th.float32 else norm.weight.float()
    b = norm.bias if norm.bias.dtype == th.float32 else norm.bias.float()
    return _ops.group_norm_silu(x, w, b, norm.eps, silu)


class _TimeEmbedding:
    """Stands in for the timestep embedding `emb` when UNetModel's per-timestep table already holds what every ResBlock
    needs from it (`_sta_xb`); `value()` evaluates the time_embed MLP on first use, for blocks that take the plain path."""

    def __init__(self, compute):
        self._compute, self._value = compute, None

    def value(self):
        if self._value is None:
            self._value = self._compute()
        return self._value


def _emb_tensor(emb):
    return emb.value() if isinstance(emb, _TimeEmbedding) else emb


class TimestepBlock(nn.Module):
    """Marker: forward(x, emb)."""


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    def forward(self, x, emb, context=None, time=None, text_index=None, coef=None, bboxs_curr=None, skip=None):
        """`skip`: the encoder activation the decoder concatenates to x first (openaimodel.py:731); a leading ResBlock fuses
        that cat into its first GroupNorm, anything else gets the plain torch.cat."""
        if skip is not None and not (len(self) and isinstance(self[0], ResBlock)):
            x, skip = th.cat([x, skip], dim=1), None
        for layer in self:
            if isinstance(layer, ResBlock) and skip is not None:
                x, skip = layer(x, emb, skip), None
            elif isinstance(layer, TimestepBlock):
                x = layer(x, emb)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, context, time, text_index, coef=coef, bboxs_curr=bboxs_curr)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_conv = use_conv
        if use_conv:
            self.conv = conv_nd(dims, self.channels, self.out_channels, 3, padding=padding)

    def forward(self, x):
        if _fusable(x) and x.shape[1] % 8 == 0 and not _ATEN_UPSAMPLE:
            x = _ops.upsample_nearest2x(x)  # one 16-byte vector per thread (ATen's NHWC kernel moves single elements)
        else:
            x = F.interpolate(x, scale_factor=2, mode="nearest")
        return self.conv(x) if self.use_conv else x


class Downsample(nn.Module):
    def __init__(self, channels, use_conv, dims=2, out_channels=None, padding=1):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        if use_conv:
            self.op = conv_nd(dims, self.channels, self.out_channels, 3, stride=2, padding=padding)
        else:
            self.op = nn.AvgPool2d(kernel_size=2, stride=2)

    def forward(self, x):
        return self.op(x)


class ResBlock(TimestepBlock):
    """GroupNorm-SiLU-conv, + timestep embedding, GroupNorm-SiLU-conv, skip (reference :252-275)."""

    def __init__(self, channels, emb_channels, dropout, out_channels=None, dims=2, use_checkpoint=False):
        super().__init__()
        self.channels = channels
        self.out_channels = out_channels or channels
        self.use_checkpoint = use_checkpoint
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(),
                                       conv_nd(dims, channels, self.out_channels, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(
            normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
            zero_module(conv_nd(dims, self.out_channels, self.out_channels, 3, padding=1)),
        )
        if self.out_channels == channels:
            self.skip_connection = nn.Identity()
        else:
            self.skip_connection = conv_nd(dims, channels, self.out_channels, 1)

    def forward(self, x, emb, skip=None):
        """`skip` (decoder blocks): the block's input is torch.cat([x, skip], dim=1) (openaimodel.py:731)."""
        flag = self.use_checkpoint and x.shape[-1] * x.shape[-2] >= getattr(self, "checkpoint_min_tokens", 0)
        return checkpoint(self._forward, (x, emb, skip), self.parameters(), flag)

    def _forward_fused(self, x, emb, skip=None):
        """openaimodel.py:252-275 for a frozen fp16 NHWC ResBlock.  ATen adds a cuDNN convolution's bias as a separate
        broadcast kernel, and `h + emb_out[:, :, None, None]` is another one; here the convolutions run bias-free and
          * conv1's bias + the projected timestep embedding become ONE fp16 [B, C] vector (the embedding GEMM's epilogue
            adds the conv bias) that the second GroupNorm kernel adds on the fly (sta_groupnorm x_bias),
          * conv2's bias (+ the 1x1 skip convolution's bias) and the residual add are one fused pass; the 1x1 skip
            convolution itself is a token GEMM on the NHWC memory."""
        f32, f16 = th.float32, th.float16
        conv1, conv2, emb_lin = self.in_layers[2], self.out_layers[3], self.emb_layers[1]
        n1, n2 = self.in_layers[0], self.out_layers[0]
        gw, gb = cached_sum(self, "g1", [n1.weight], f32), cached_sum(self, "b1", [n1.bias], f32)
        if skip is None:
            # x feeds the first GroupNorm and the skip branch: the fork hands both gradients to one GroupNorm backward launch
            g1, x = _ops.group_norm_silu_fork(x, gw, gb, n1.eps, True)
        else:
            # decoder block: the cat of (x, encoder skip) is read in place by the GroupNorm, which also writes the concatenation
            # for the 1x1 skip convolution; its backward returns the two gradients as dense tensors
            g1, x = _ops.cat_group_norm_silu(x, skip, gw, gb, n1.eps, True)
        b, c, hh, ww = x.shape
        with th.autocast("cuda", enabled=False):
            h = F.conv2d(g1, cached_sum(self, "w1", [conv1.weight], f16), None, conv1.stride, conv1.padding)
            # x_bias = emb_layers(emb) + conv1.bias.  UNetModel.forward hands over one [B, sum C] gather from its
            # per-timestep table (a strided column slice is this block's vector); standalone blocks project themselves
            slot = getattr(emb, "_sta_xb", None)
            if slot is not None and id(self) in slot[1]:
                off = slot[1][id(self)]
                xb = slot[0][:, off:off + self.out_channels]
            else:
                emb = _emb_tensor(emb)
                act = getattr(emb, "_sta_silu", None)  # SiLU(emb) is shared by all ResBlocks
                if act is None:
                    act = F.silu(emb)
                xb = F.linear(act if act.dtype == f16 else act.to(f16), cached_sum(self, "we", [emb_lin.weight], f16),
                              cached_sum(self, "be", [emb_lin.bias, conv1.bias], f16))
            g2 = _ops.group_norm_silu(h, cached_sum(self, "g2", [n2.weight], f32), cached_sum(self, "b2", [n2.bias], f32),
                                      n2.eps, True, x_bias=xb)
            h = F.conv2d(g2, cached_sum(self, "w2", [conv2.weight], f16), None, conv2.stride, conv2.padding)
            x_tok = nhwc_tokens(x)
            if isinstance(self.skip_connection, nn.Identity):
                skip = x_tok if x_tok.is_contiguous() else x_tok.contiguous()
                bias = cached_sum(self, "bo", [conv2.bias], f32)
            else:
                sk = self.skip_connection
                skip = F.linear(x_tok, cached_sum(self, "ws", [sk.weight], f16).reshape(sk.out_channels, c))
                bias = cached_sum(self, "bo", [conv2.bias, sk.bias], f32)
            out = _ops.bias_residual_add(nhwc_tokens(h), bias, skip)
        return tokens_nhwc(out, hh, ww)

    def _fusable_block(self, x) -> bool:
        sk = self.skip_connection
        return (_fusable(x) and _frozen(self) and (not self.training or self.out_layers[2].p == 0.0)
                and (isinstance(sk, nn.Identity) or (isinstance(sk, nn.Conv2d) and sk.kernel_size == (1, 1))))

    def _forward(self, x, emb, skip=None):
        if skip is not None and not (self._fusable_block(x) and _fusable(skip) and x.shape[1] % 8 == 0 and skip.shape[1] % 8 == 0):
            x, skip = th.cat([x, skip], dim=1), None
        if self._fusable_block(x):
            return self._forward_fused(x, emb, skip)
        fused = _fusable(x)
        h = self.in_layers[2](gn_silu(self.in_layers[0], x)) if fused else self.in_layers(x)
        emb_out = self.emb_layers(_emb_tensor(emb)).type(h.dtype)
        h = h + emb_out[:, :, None, None]
        if fused and _fusable(h):
            h = self.out_layers[3](self.out_layers[2](gn_silu(self.out_layers[0], h)))  # [2] = Dropout(p)
        else:
            h = self.out_layers(h)
        return self.skip_connection(x) + h


class UNetModel(nn.Module):
    XB_TABLE_STEPS = 1000  # DDPM training timesteps (v1-inference.yaml:8); sampler timesteps are always below it

    def __init__(self, image_size=32, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), dropout=0, channel_mult=(1, 2, 4, 4), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=8, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=True, transformer_depth=1, context_dim=768,
                 n_embed=None, legacy=False):
        super().__init__()
        if not use_spatial_transformer or context_dim is None:
            raise ValueError("only the spatial-transformer UNet of SD-v1 is built")
        if num_classes is not None or use_scale_shift_norm or resblock_updown or n_embed is not None:
            raise ValueError("class conditioning / scale-shift norm / resblock resampling are not used by SD-v1")
        if isinstance(context_dim, (list, tuple)) or type(context_dim).__name__ == "ListConfig":
            context_dim = list(context_dim)[0]
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.attention_resolutions = num_res_blocks, tuple(attention_resolutions)
        self.channel_mult, self.num_heads = tuple(channel_mult), num_heads
        self.num_classes = None
        self.dtype = th.float32

        time_embed_dim = model_channels * 4
        self.time_embed = nn.Sequential(linear(model_channels, time_embed_dim), nn.SiLU(),
                                        linear(time_embed_dim, time_embed_dim))

        def res(cin, cout):
            return ResBlock(cin, time_embed_dim, dropout, out_channels=cout, dims=dims, use_checkpoint=use_checkpoint)

        def attn(ch):
            heads = num_heads if num_head_channels == -1 else ch // num_head_channels
            return SpatialTransformer(ch, heads, ch // heads, depth=transformer_depth, context_dim=context_dim)

        self.input_blocks = nn.ModuleList([TimestepEmbedSequential(conv_nd(dims, in_channels, model_channels, 3, padding=1))])
        skip_chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(self.channel_mult):
            for _ in range(num_res_blocks):
                layers: List[nn.Module] = [res(ch, mult * model_channels)]
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    layers.append(attn(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                skip_chans.append(ch)
            if level != len(self.channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, conv_resample, dims=dims, out_channels=ch)))
                skip_chans.append(ch)
                ds *= 2

        self.middle_block = TimestepEmbedSequential(res(ch, ch), attn(ch), res(ch, ch))

        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [res(ch + skip_chans.pop(), model_channels * mult)]
                ch = model_channels * mult
                if ds in self.attention_resolutions:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(Upsample(ch, conv_resample, dims=dims, out_channels=ch))
                    ds //= 2
                self.output_blocks.append(TimestepEmbedSequential(*layers))

        self.out = nn.Sequential(normalization(ch), nn.SiLU(),
                                 zero_module(conv_nd(dims, model_channels, out_channels, 3, padding=1)))
        self.set_checkpointing(use_checkpoint)

    def set_checkpointing(self, flag: bool, min_tokens: int = 0):
        """Gradient checkpointing of ResBlocks and transformer blocks (the reference checkpoints all of them:
        v1-inference.yaml:43, attention.py:224).  `min_tokens` keeps the activations of blocks whose feature map has
        fewer pixels than that (cheap to store, so their recompute can be skipped on a 180 GB part)."""
        self.use_checkpoint = bool(flag)
        self.checkpoint_min_tokens = int(min_tokens)
        for m in self.modules():
            if isinstance(m, ResBlock):
                m.use_checkpoint = bool(flag)
                m.checkpoint_min_tokens = int(min_tokens)
            elif isinstance(m, BasicTransformerBlock):
                m.checkpoint = bool(flag)
                m.checkpoint_min_tokens = int(min_tokens)

    # -- per-timestep ResBlock biases ---------------------------------------------------------------------------
    def _timestep_bias(self, timesteps):
        """([B, sum C] fp16, {id(ResBlock): column offset}): `emb_layers(time_embed(t)) + conv1.bias` of EVERY ResBlock for
        the given integer timesteps, gathered from a table over all `XB_TABLE_STEPS` timesteps.  The UNet is frozen while
        sampling, so the 22 per-block projections of an evaluation (22 tiny GEMMs = 22 graph nodes) collapse into one row
        gather; the table (1000 x 20160 fp16 = 40 MB for SD-v1) is one GEMM, rebuilt only when a weight changes."""
        blocks = [m for m in self.modules() if isinstance(m, ResBlock)]
        params = [p for p in self.time_embed.parameters()]
        for rb in blocks:
            params += [rb.emb_layers[1].weight, rb.emb_layers[1].bias, rb.in_layers[2].bias]
        key = tuple((p.data_ptr(), p._version) for p in params)
        cache = self.__dict__.get("_xb_cache")
        if cache is None or cache[0] != key:
            f16 = th.float16
            dev = timesteps.device
            with th.no_grad(), th.autocast("cuda", enabled=False):
                t = th.arange(self.XB_TABLE_STEPS, device=dev)
                e = timestep_embedding(t, self.model_channels, repeat_only=False).to(f16)
                l0, l2 = self.time_embed[0], self.time_embed[2]
                e = F.linear(F.silu(F.linear(e, l0.weight.to(f16), l0.bias.to(f16))), l2.weight.to(f16), l2.bias.to(f16))
                w_all = th.cat([rb.emb_layers[1].weight.to(f16) for rb in blocks])
                b_all = th.cat([(rb.emb_layers[1].bias.float() + rb.in_layers[2].bias.float()).to(f16) for rb in blocks])
                table = F.linear(F.silu(e), w_all, b_all).contiguous()
            offs, off = {}, 0
            for rb in blocks:
                offs[id(rb)] = off
                off += rb.out_channels
            if cache is not None and cache[1].shape == table.shape and cache[1].device == table.device:
                cache[1].copy_(table)  # in place: CUDA graphs captured over the old table keep a valid pointer
                table = cache[1]
            cache = (key, table, offs)
            self.__dict__["_xb_cache"] = cache
        return cache[1].index_select(0, timesteps), cache[2]

    # -- per-prompt attention state (in-memory replacement of the reference's c{i}_*.pt files) ---------------
    def transformer_blocks(self) -> Iterable[BasicTransformerBlock]:
        for m in self.modules():
            if isinstance(m, BasicTransformerBlock):
                yield m

    def set_local_contexts(self, local_contexts: Optional[Sequence[th.Tensor]], first_timestep: Optional[int] = None):
        for blk in self.transformer_blocks():
            blk.set_local_contexts(local_contexts)
            if first_timestep is not None:
                blk.first_timestep = first_timestep

    def reset_attention_cache(self):
        for blk in self.transformer_blocks():
            blk.reset_cache()

    def forward(self, x, text_index=None, timesteps=None, context=None, y=None, coef=None, bboxs_curr=None,
                step_time: Optional[int] = None, **kwargs):
        """x [2B,4,H,W], timesteps [2B], context [2B,77,768] -> eps [2B,4,H,W]   (reference :710-742)."""
        assert y is None, "SD-v1 is not class-conditional"
        hs = []
        def compute_emb():
            return self.time_embed(timestep_embedding(timesteps, self.model_channels, repeat_only=False))

        if x.is_cuda and timesteps.is_cuda and timesteps.dtype == th.long and _frozen(self):
            # every ResBlock's projected embedding comes out of the per-timestep table: the time_embed MLP itself (sin / cos /
            # cat / two Linear + SiLU = 13 launches per evaluation) only runs if some block asks for the raw embedding
            emb = _TimeEmbedding(compute_emb)
            emb._sta_xb = self._timestep_bias(timesteps)
        else:
            emb = compute_emb()
            if not emb.requires_grad:
                emb._sta_silu = F.silu(emb)  # every ResBlock starts its embedding branch with the same SiLU (:217-223)
        # the reference hands timesteps[0] (a device scalar) to every block, which then syncs on `time == 981`
        # (attention.py:240); callers of this package pass the same value as a host int instead
        time = step_time if step_time is not None else int(timesteps[0].item())  # one sync instead of 16
        h = x.type(self.dtype)
        for module in self.input_blocks:
            h = module(h, emb, context, time, text_index, coef=coef, bboxs_curr=bboxs_curr)
            hs.append(h)
        h = self.middle_block(h, emb, context, time, text_index, coef=coef, bboxs_curr=bboxs_curr)
        for module in self.output_blocks:  # h = cat([h, hs.pop()], dim=1) happens inside the block's first GroupNorm
            h = module(h, emb, context, time, text_index, coef=coef, bboxs_curr=bboxs_curr, skip=hs.pop())
        if _fusable(h):  # GroupNorm32 works in fp32 either way; skipping the cast only skips two copies
            return self.out[2](gn_silu(self.out[0], h))
        h = h.type(x.dtype)
        return self.out(h)
