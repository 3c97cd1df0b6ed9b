"""Seeded inputs shared by tools/ref_on_gpu.py (runs the UNMODIFIED reference on the B200 and writes
tests/golden/ref_gpu.npz) and tests/test_ref_gpu_golden.py (compares the sta_* kernels with that fixture).

Everything is regenerated from seeds on both sides, so the fixture only has to hold the reference's outputs."""
from __future__ import annotations

from pathlib import Path

import torch

from oracle import sta_oracle as O

GOLD = Path(__file__).resolve().parent / "golden"
BBOXES = [[0.30, 0.50], [0.70, 0.50]]
# (tag, tokens, channels): the UNet's four transformer geometries at 512x512
BLOCKS = [("L0", 4096, 320), ("L1", 1024, 640), ("L2", 256, 1280), ("mid", 64, 1280)]
FULL = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
            channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768)
TINY = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(1, 2), num_res_blocks=1,
            channel_mult=(1, 2), num_heads=8, context_dim=768)
UNETS = [("unet_tiny", TINY, 16, 5, 501), ("unet_full", FULL, 64, 0, 501)]  # (tag, cfg, latent, weight seed, timestep)


def ctx_tensor(seed, shape=(1, 77, 768)):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * 1.04


def uncond():
    return torch.load(GOLD / "uncond_embedding.pt", map_location="cpu").float()


def token_subsample(n: int, C: int) -> slice:
    """Tokens of a block output kept in the fixture: 64 / 32 / 16 / 16 tokens at C = 320 / 640 / 1280 / 1280 (about 41 k
    values per array), evenly spread over the image."""
    return slice(0, n, max(1, n // max(8, 20480 // C)))


def block_case(n: int, C: int):
    """x [2,n,C], context [2,77,768], two local contexts, coef, upstream gradient G, block weights (seed 2)."""
    g = torch.Generator().manual_seed(7 * n + C)
    x = torch.randn(2, n, C, generator=g)
    G = torch.randn(2, n, C, generator=g) * 0.05
    context = torch.cat([uncond(), ctx_tensor(100)])
    locs = [ctx_tensor(101), ctx_tensor(102)]
    coef = torch.tensor([2.5, 1.5])
    return x, G, context, locs, coef


def block_weights(shapes):
    return O.seeded_state_dict(shapes, seed=2)


def unet_case(latent: int):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, latent, latent, generator=g)
    G = torch.randn(2, 4, latent, latent, generator=g)
    context = torch.cat([uncond(), ctx_tensor(100)])
    locs = [ctx_tensor(101), ctx_tensor(102)]
    coef = torch.tensor([2.5, 2.5])
    return torch.cat([x, x]), G, context, locs, coef
