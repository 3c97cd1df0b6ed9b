"""Text conditioning stage (reference ldm/modules/encoders/modules.py:137-162, FrozenCLIPEmbedder: HF CLIP-L/14 text
tower -> [B, 77, 768]).  Out of the kernel scope: it runs 2 + n_obj times per prompt before the
denoising loop and only produces the frozen K/V source tensors.

No CLIP-L/14 weights exist offline.  `FrozenCLIPEmbedder` loads them from a local directory when one is given
(`version` = path, or $STA_CLIP_L_PATH); otherwise `SyntheticTextEmbedder` returns deterministic N(0, 1.04^2)
embeddings keyed by the text (the statistics of the shipped unconditional embedding) so that
the data path has the right shapes and scale.  bench.py and the tests say `data: synthetic` for that reason.
"""
from __future__ import annotations

import os
import zlib
from typing import List, Sequence, Union

import torch
import torch.nn as nn


class SyntheticTextEmbedder(nn.Module):
    def __init__(self, max_length: int = 77, dim: int = 768, std: float = 1.04, device="cuda"):
        super().__init__()
        self.max_length, self.dim, self.std, self.device = max_length, dim, std, device

    def encode(self, text: Union[str, Sequence[str]]) -> torch.Tensor:
        texts: List[str] = [text] if isinstance(text, str) else list(text)
        out = []
        for t in texts:
            g = torch.Generator().manual_seed(zlib.crc32(t.encode()) & 0x7FFFFFFF)
            out.append(torch.randn(self.max_length, self.dim, generator=g) * self.std)
        return torch.stack(out).to(self.device)

    forward = encode


class FrozenCLIPEmbedder(nn.Module):
    """HF CLIP text encoder from a LOCAL path (no hub access here)."""

    def __init__(self, version: str = "", device="cuda", max_length=77):
        super().__init__()
        path = version if os.path.isdir(version) else os.environ.get("STA_CLIP_L_PATH", "")
        if not os.path.isdir(path):
            raise FileNotFoundError("CLIP-L/14 weights are not available offline; use SyntheticTextEmbedder")
        from transformers import CLIPTextModel, CLIPTokenizer

        self.tokenizer = CLIPTokenizer.from_pretrained(path)
        self.transformer = CLIPTextModel.from_pretrained(path).to(device).eval().requires_grad_(False)
        self.device, self.max_length = device, max_length

    @torch.no_grad()
    def encode(self, text):
        enc = self.tokenizer(text, truncation=True, max_length=self.max_length, return_length=True,
                             return_overflowing_tokens=False, padding="max_length", return_tensors="pt")
        return self.transformer(input_ids=enc["input_ids"].to(self.device)).last_hidden_state

    forward = encode
