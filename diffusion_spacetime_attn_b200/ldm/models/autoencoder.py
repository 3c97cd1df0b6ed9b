"""KL-VAE decode half, with the reference's parameter names (`first_stage_model.decoder.*`, `post_quant_conv`).

Reference: ldm/models/autoencoder.py:285-333 (AutoencoderKL.decode) -> ldm/modules/diffusionmodules/model.py:462-568
(Decoder), ResnetBlock :82-142, AttnBlock :150-202, Upsample :42-58.  Only `decode` runs on the path (once per alpha
epoch, differentiably: ddpm.py:705 has @no_grad commented out); the encoder is not built (checkpoint keys for it are
ignored by load_state_dict(strict=False), as the reference does at scripts/txt2img-gpt.py:62).

SURVEY.md §8f rank 3.  The convolutions stay cuDNN; on fp16 CUDA activations (what the pipeline feeds it) everything
between them runs on this package's streaming kernels, NHWC end to end:
  * GroupNorm(32) + SiLU -> sta_groupnorm (one fused pass pair instead of fp32 copy / moments / normalise / fp16 copy /
    SiLU on NCHW tensors — the reference's chain was 65 % of the decode + loss + backward time on B200);
  * conv1's bias rides in the second GroupNorm (x_bias), conv2's bias (+ the 1x1 shortcut's) and the residual add are
    one pass (sta_add_layernorm_fwd without a norm) — ATen adds a cuDNN conv bias as a separate broadcast kernel;
  * the 1x1 convolutions (shortcut, attention q/k/v/proj_out) are token GEMMs on the NHWC memory; q/k/v are ONE GEMM.
  * the single mid-block attention (1 head, d = 512, N = 4096) is the 512-wide flash kernel (csrc/sta_sattn_wide.cu through
    sta_sattn_fwd / _bwd): no [N, N] score matrix, which the reference materialises three times (model.py:176-191);
    STA_VAE_ATTN=materialised selects bmm / softmax / bmm on cuBLAS (the A/B arm of tools/bench_wide.py); torch SDPA on
    the plain path.  On CPU / fp32 tensors the plain torch path below runs
(that is what tests/test_vae_golden.py checks against the reference's output on CPU).
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from ... import ops
from ..modules.attention import _frozen, cached_sum, nhwc_tokens, tokens_nhwc


def _fusable(x: torch.Tensor, module: nn.Module) -> bool:
    return x.is_cuda and x.dtype == torch.float16 and _frozen(module)


def _gn(owner: nn.Module, tag: str, norm: nn.GroupNorm, x, silu: bool, x_bias=None):
    f32 = torch.float32
    return ops.group_norm_silu(x, cached_sum(owner, tag + "w", [norm.weight], f32), cached_sum(owner, tag + "b", [norm.bias], f32),
                               norm.eps, silu, x_bias=x_bias)


def _conv_nobias(owner: nn.Module, tag: str, conv: nn.Conv2d, x):
    return F.conv2d(x, cached_sum(owner, tag, [conv.weight], torch.float16), None, conv.stride, conv.padding)


def Normalize(in_channels, num_groups=32):
    return nn.GroupNorm(num_groups=num_groups, num_channels=in_channels, eps=1e-6, affine=True)


class Upsample(nn.Module):
    def __init__(self, in_channels, with_conv):
        super().__init__()
        self.with_conv = with_conv
        if with_conv:
            self.conv = nn.Conv2d(in_channels, in_channels, kernel_size=3, stride=1, padding=1)

    def forward(self, x):
        if x.is_cuda and x.dtype == torch.float16 and x.shape[1] % 8 == 0:
            x = ops.upsample_nearest2x(x)
        else:
            x = F.interpolate(x, scale_factor=2.0, mode="nearest")
        return self.conv(x) if self.with_conv else x


class ResnetBlock(nn.Module):
    def __init__(self, *, in_channels, out_channels=None, dropout=0.0):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(dropout)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        if in_channels != out_channels:
            self.nin_shortcut = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1, padding=0)

    def _forward_fused(self, x):
        b, c, hh, ww = x.shape
        f16 = torch.float16
        with torch.autocast("cuda", enabled=False):
            h = _conv_nobias(self, "w1", self.conv1, _gn(self, "n1", self.norm1, x, True))
            xb = cached_sum(self, "b1", [self.conv1.bias], f16)
            xb = xb.unsqueeze(0) if b == 1 else xb.unsqueeze(0).expand(b, -1).contiguous()
            h = _conv_nobias(self, "w2", self.conv2, _gn(self, "n2", self.norm2, h, True, x_bias=xb))
            x_tok = nhwc_tokens(x)
            if self.in_channels != self.out_channels:
                sk = self.nin_shortcut
                skip = F.linear(x_tok, cached_sum(self, "ws", [sk.weight], f16).reshape(self.out_channels, c))
                bias = cached_sum(self, "bo", [self.conv2.bias, sk.bias], torch.float32)
            else:
                skip = x_tok if x_tok.is_contiguous() else x_tok.contiguous()
                bias = cached_sum(self, "bo", [self.conv2.bias], torch.float32)
            return tokens_nhwc(ops.bias_residual_add(nhwc_tokens(h), bias, skip), hh, ww)

    def forward(self, x, temb=None):
        if _fusable(x, self) and (not self.training or self.dropout.p == 0.0):
            return self._forward_fused(x)
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(self.dropout(F.silu(self.norm2(h))))
        if self.in_channels != self.out_channels:
            x = self.nin_shortcut(x)
        return x + h


class AttnBlock(nn.Module):
    def __init__(self, in_channels):
        super().__init__()
        self.norm = Normalize(in_channels)
        self.q = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.k = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.v = nn.Conv2d(in_channels, in_channels, kernel_size=1)
        self.proj_out = nn.Conv2d(in_channels, in_channels, kernel_size=1)

    def _forward_fused(self, x):
        b, c, hh, ww = x.shape
        f16 = torch.float16
        with torch.autocast("cuda", enabled=False):
            t = nhwc_tokens(_gn(self, "n", self.norm, x, False))  # [b, hw, c]
            slot = self.__dict__.setdefault("_sta_cached", {})
            key = tuple((m.weight.data_ptr(), m.weight._version, m.bias.data_ptr(), m.bias._version) for m in (self.q, self.k, self.v))
            if slot.get("qkv", (None,))[0] != key:
                with torch.no_grad():
                    wq = torch.cat([m.weight.detach().reshape(c, c) for m in (self.q, self.k, self.v)]).to(f16).contiguous()
                    bq = torch.cat([m.bias.detach() for m in (self.q, self.k, self.v)]).to(f16).contiguous()
                prev = slot.get("qkv")
                if prev is not None and prev[1].shape == wq.shape and prev[1].device == wq.device:
                    prev[1].copy_(wq)  # in place: captured CUDA graphs keep valid pointers
                    prev[2].copy_(bq)
                    wq, bq = prev[1], prev[2]
                slot["qkv"] = (key, wq, bq)
            _, wq, bq = slot["qkv"]
            qkv = F.linear(t, wq, bq)  # [b, hw, 3c]: q / k / v are its column slices
            if c == 512 and os.environ.get("STA_VAE_ATTN", "flash") != "materialised":
                # One head of d = C = 512 (SD-v1's KL-VAE): the 512-wide flash kernel (csrc/sta_sattn_wide.cu) — no [hw, hw]
                # score matrix (model.py:176-191 writes it three times: bmm output, scaled copy, softmax), and d(qkv) comes
                # back as ONE [b, hw, 3c] gradient.  B200, 64 x 64 latent: forward 87 us (bmm / softmax / bmm on cuBLAS:
                # 159 us), forward + backward 397 us (398 us); two 96 x 96 latents: 2.74 ms (3.61 ms).
                o = ops.self_attention_qkv(qkv, 1)
            else:
                # other widths: bmm / softmax / bmm on cuBLAS with the materialised score matrix, fp32 softmax as the
                # reference computes it under autocast (torch's SDPA has only an sm80 kernel for one wide head)
                q, k, v = qkv.chunk(3, dim=-1)
                w_ = torch.bmm(q, k.transpose(1, 2))
                w_ = torch.softmax(w_.float() * (int(c) ** -0.5), dim=2).to(f16)
                o = torch.bmm(w_, v)
            o = F.linear(o, cached_sum(self, "wo", [self.proj_out.weight], f16).reshape(c, c))
            x_tok = nhwc_tokens(x)
            x_tok = x_tok if x_tok.is_contiguous() else x_tok.contiguous()
            out = ops.bias_residual_add(o, cached_sum(self, "bo", [self.proj_out.bias], torch.float32), x_tok)
            return tokens_nhwc(out, hh, ww)

    def forward(self, x):
        if _fusable(x, self):
            return self._forward_fused(x)
        h_ = self.norm(x)
        b, c, h, w = h_.shape
        q, k, v = (f(h_).reshape(b, 1, c, h * w).transpose(2, 3) for f in (self.q, self.k, self.v))
        o = F.scaled_dot_product_attention(q, k, v)  # scale = c ** -0.5, as model.py:183
        o = o.transpose(2, 3).reshape(b, c, h, w)
        return x + self.proj_out(o)


class Decoder(nn.Module):
    def __init__(self, *, ch=128, out_ch=3, ch_mult=(1, 2, 4, 4), num_res_blocks=2, attn_resolutions=(), dropout=0.0,
                 resamp_with_conv=True, in_channels=3, resolution=256, z_channels=4, **ignored):
        super().__init__()
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = nn.Conv2d(z_channels, block_in, kernel_size=3, stride=1, padding=1)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(in_channels=block_in, out_channels=block_in, dropout=dropout)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(in_channels=block_in, out_channels=block_out, dropout=dropout))
                block_in = block_out
            up = nn.Module()
            up.block = block
            up.attn = nn.ModuleList()
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = nn.Conv2d(block_in, out_ch, kernel_size=3, stride=1, padding=1)

    def forward(self, z):
        h = self.conv_in(z)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                h = self.up[i_level].block[i_block](h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        if _fusable(h, self.norm_out):
            return self.conv_out(_gn(self, "no", self.norm_out, h, True))
        return self.conv_out(F.silu(self.norm_out(h)))


class AutoencoderKL(nn.Module):
    def __init__(self, ddconfig=None, lossconfig=None, embed_dim=4, **ignored):
        super().__init__()
        ddconfig = dict(ddconfig or {})
        ddconfig.setdefault("z_channels", 4)
        self.decoder = Decoder(**ddconfig)
        self.post_quant_conv = nn.Conv2d(embed_dim, ddconfig["z_channels"], 1)
        self.embed_dim = embed_dim

    def decode(self, z):
        return self.decoder(self.post_quant_conv(z.to(self.post_quant_conv.weight.dtype)))
