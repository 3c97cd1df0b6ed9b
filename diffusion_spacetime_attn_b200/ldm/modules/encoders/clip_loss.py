"""CLIP-fidelity loss of the alpha optimisation (reference ldm/models/diffusion/plms.py:21-61, DCLIPLoss).

The reference calls `clip.load("ViT-B/32")` (openai/CLIP, unpinned).  That package and its weights are
not available offline, so this file carries a compact CLIP ViT-B/32 with OpenAI's parameter names (a downloaded
`ViT-B-32.pt` state_dict loads with `load_openai_state_dict`); without weights it is a seeded random-init network —
the loss is outside the kernels and only has to be a fixed differentiable image -> scalar map.

Resampling follows the reference exactly: the global loss uses nearest Upsample(x7) -> AvgPool2d(16) (512 -> 224,
plms.py:26-27,41); the per-object crops use a bilinear Resize((224,224)) (plms.py:28,31).  No CLIP mean/std
normalisation is applied (the reference applies none).
"""
from __future__ import annotations

import zlib
from typing import List

import torch
import torch.nn as nn
import torch.nn.functional as F


class _ResBlock(nn.Module):
    def __init__(self, width, heads):
        super().__init__()
        self.attn = nn.MultiheadAttention(width, heads)
        self.ln_1 = nn.LayerNorm(width)
        self.mlp = nn.Sequential()
        self.mlp.add_module("c_fc", nn.Linear(width, width * 4))
        self.mlp.add_module("gelu", _QuickGELU())
        self.mlp.add_module("c_proj", nn.Linear(width * 4, width))
        self.ln_2 = nn.LayerNorm(width)
        self.heads = heads

    def forward(self, x, causal=False):  # x: [B, L, W]
        y = self.ln_1(x)
        b, l, w = y.shape
        qkv = F.linear(y, self.attn.in_proj_weight, self.attn.in_proj_bias)
        q, k, v = (t.reshape(b, l, self.heads, w // self.heads).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        o = F.scaled_dot_product_attention(q, k, v, is_causal=causal).transpose(1, 2).reshape(b, l, w)
        x = x + self.attn.out_proj(o)
        return x + self.mlp(self.ln_2(x))


class _QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.ModuleList([_ResBlock(width, heads) for _ in range(layers)])

    def forward(self, x, causal=False):
        for blk in self.resblocks:
            x = blk(x, causal)
        return x


class _VisionTower(nn.Module):
    def __init__(self, res=224, patch=32, width=768, layers=12, heads=12, out_dim=512):
        super().__init__()
        self.conv1 = nn.Conv2d(3, width, patch, patch, bias=False)
        scale = width ** -0.5
        self.class_embedding = nn.Parameter(scale * torch.randn(width))
        self.positional_embedding = nn.Parameter(scale * torch.randn((res // patch) ** 2 + 1, width))
        self.ln_pre = nn.LayerNorm(width)
        self.transformer = _Transformer(width, layers, heads)
        self.ln_post = nn.LayerNorm(width)
        self.proj = nn.Parameter(scale * torch.randn(width, out_dim))

    def forward(self, x):
        x = self.conv1(x).flatten(2).transpose(1, 2)
        cls = self.class_embedding.to(x.dtype).expand(x.shape[0], 1, -1)
        x = torch.cat([cls, x], dim=1) + self.positional_embedding.to(x.dtype)
        x = self.transformer(self.ln_pre(x))
        return self.ln_post(x[:, 0]) @ self.proj.to(x.dtype)


class CLIPViTB32(nn.Module):
    def __init__(self, embed_dim=512, context_length=77, vocab_size=49408, text_width=512, text_layers=12, text_heads=8):
        super().__init__()
        self.visual = _VisionTower(out_dim=embed_dim)
        self.transformer = _Transformer(text_width, text_layers, text_heads)
        self.token_embedding = nn.Embedding(vocab_size, text_width)
        self.positional_embedding = nn.Parameter(0.01 * torch.randn(context_length, text_width))
        self.ln_final = nn.LayerNorm(text_width)
        self.text_projection = nn.Parameter(text_width ** -0.5 * torch.randn(text_width, embed_dim))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6593)
        self.context_length = context_length

    def encode_image(self, image):
        return self.visual(image)

    def encode_text(self, tokens):
        x = self.token_embedding(tokens) + self.positional_embedding
        x = self.ln_final(self.transformer(x, causal=True))
        # device-side row index: a CPU arange would be copied to the device synchronously (a stream sync inside every call)
        return x[torch.arange(x.shape[0], device=x.device), tokens.argmax(dim=-1)] @ self.text_projection

    def load_openai_state_dict(self, sd):
        return self.load_state_dict({k: v for k, v in sd.items() if k in self.state_dict()}, strict=False)


def upsample_avgpool_matrix(size_in: int, scale: int = 7, pool: int = 16) -> torch.Tensor:
    """The reference's global-loss resampling, nn.Upsample(scale_factor=7) (nearest) followed by nn.AvgPool2d(16)
    (plms.py:26-27,41), is a fixed separable linear map: out = A @ img @ A^T with A [size_in*scale/pool, size_in],
    A[i, r // scale] += 1/pool for r in [pool*i, pool*(i+1)).  Two tiny GEMMs instead of materialising the 3584 x 3584
    fp32 image and pooling it back (and the same again in backward) — SURVEY.md §8f rank 4."""
    up = size_in * scale
    if up % pool:
        raise ValueError(f"{size_in} x {scale} is not a multiple of the {pool}-pixel pooling window")
    A = torch.zeros(up // pool, size_in, dtype=torch.float64)
    src = torch.arange(up) // scale
    rows = torch.arange(up) // pool
    A.index_put_((rows, src), torch.full((up,), 1.0 / pool, dtype=torch.float64), accumulate=True)
    return A.float()


def hash_tokenize(texts: List[str], context_length: int = 77, vocab_size: int = 49408) -> torch.Tensor:
    """Offline stand-in for clip.tokenize (its BPE vocabulary is not in this image): <sot> word-hash ids <eot>."""
    out = torch.zeros(len(texts), context_length, dtype=torch.long)
    for i, t in enumerate(texts):
        ids = [49406] + [1 + zlib.crc32(w.encode()) % (vocab_size - 3) for w in t.lower().split()][: context_length - 2]
        ids.append(49407)  # <eot> is the largest id: encode_text pools at argmax, as OpenAI CLIP does
        out[i, : len(ids)] = torch.tensor(ids)
    return out


class DCLIPLoss(nn.Module):
    def __init__(self, device="cuda", seed=0, tokenizer=None):
        super().__init__()
        g = torch.random.get_rng_state()
        torch.manual_seed(seed)
        self.model = CLIPViTB32()
        torch.random.set_rng_state(g)
        self.model.to(device).eval().requires_grad_(False)
        self.upsample = nn.Upsample(scale_factor=7)
        self.avg_pool = nn.AvgPool2d(kernel_size=16)
        self.tokenizer = tokenizer or hash_tokenize
        self._text_cache = {}
        self._resample = {}
        # CUDA-graph execution (pipeline.py sets it): the image tower's forward-with-grad + backward become one captured graph
        # pair per number of images (graphed.GraphedDifferentiable; ~1200 eager launches per loss evaluation otherwise)
        self.graph_encode = False
        self._graphed_encode = {}

    def _text_feat(self, text):
        if text not in self._text_cache:
            dev = next(self.model.parameters()).device
            tokens = self.tokenizer([text])
            if dev.type == "cuda":
                # pinned + non_blocking: a pageable .to(dev) synchronises the stream, i.e. the host would wait here for the whole
                # forward trajectory it has enqueued ahead of the GPU (measured: 133 ms in the first loss call of every image,
                # after which the GPU waited ~7 ms for the host to catch up)
                tokens = tokens.pin_memory().to(dev, non_blocking=True)
            else:
                tokens = tokens.to(dev)
            with torch.no_grad():
                self._text_cache[text] = self.model.encode_text(tokens).float()
        return self._text_cache[text]

    def _encode_image(self, images_224):
        if not (self.graph_encode and images_224.is_cuda and images_224.requires_grad and torch.is_grad_enabled()):
            return self.model.encode_image(images_224)
        from ....graphed import GraphedDifferentiable, drop_if_weights_changed

        drop_if_weights_changed(self, "_graphed_fp", self.model.visual, self._graphed_encode, "CLIP image tower")
        amp = torch.is_autocast_enabled()  # the graph replays what the caller's autocast state would have computed eagerly
        key = (tuple(images_224.shape), images_224.dtype, amp)
        g = self._graphed_encode.get(key)
        if g is None:

            def fn(x):
                with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                    return self.model.encode_image(x)

            g = self._graphed_encode[key] = GraphedDifferentiable(fn, images_224.detach())
        return g(images_224)

    def _one_minus_cos(self, image_224, text):
        feat = self._encode_image(image_224).float()
        return 1 - F.cosine_similarity(feat, self._text_feat(text))

    def one_minus_cos_batched(self, images_224, texts):
        """[K] losses of K already-resized images [K, 3, 224, 224] against K texts: ONE pass through the image tower instead of
        the reference's K separate batch-1 passes (plms.py:252-273 calls forward_2 / forward_3 per image; the per-image
        arithmetic is the same, the tower is launch-bound at batch 1)."""
        feat = self._encode_image(images_224).float()
        return 1 - F.cosine_similarity(feat, torch.cat([self._text_feat(t) for t in texts], dim=0))

    def resize_global(self, image):
        """forward_2's resampling of a [3, H, W] image to [1, 3, 224, 224] (plms.py:41)."""
        h, w = image.shape[-2:]
        if (h * 7) % 224 or (w * 7) % 224:
            return F.interpolate(image.unsqueeze(0).float(), size=(224, 224), mode="bilinear", antialias=True, align_corners=False)
        mats = []
        for n in (h, w):
            key = (n, image.device)
            if key not in self._resample:
                self._resample[key] = upsample_avgpool_matrix(n, 7, n * 7 // 224).to(image.device)
            mats.append(self._resample[key])
        with torch.autocast(image.device.type, enabled=False):  # exact fp32, like the reference's pooling of an fp32 image
            small = mats[0] @ image.float() @ mats[1].t()
        return small.unsqueeze(0)

    @staticmethod
    def resize_crop(image):
        """forward_3's bilinear resize of a [3, h, w] crop to [1, 3, 224, 224] (plms.py:31)."""
        return F.interpolate(image.unsqueeze(0), size=(224, 224), mode="bilinear", antialias=True, align_corners=False)

    def forward_2(self, image, text):
        """Global loss: image [3, 512, 512] in [0,1] -> 1 - cos(CLIP(img), CLIP(text))   (plms.py:38-45).
        512 px: Upsample(x7) -> AvgPool2d(16) -> 224 px, the reference's literal pipeline (as two small GEMMs).  Other sizes
        (the reference hard-codes 512) keep the x7 upsample and pool with the window that lands on CLIP's
        224 px (768 px -> window 24); sizes where that is not an integer are resized bilinearly instead."""
        return self._one_minus_cos(self.resize_global(image), text)

    def forward_3(self, image, text):
        """Per-object crop loss with a bilinear resize to 224 x 224   (plms.py:29-36)."""
        return self._one_minus_cos(self.resize_crop(image), text)
