"""Generate tests/golden/vae_decoder.npz by running the UNMODIFIED reference KL-VAE decoder
(/root/reference/.../ldm/modules/diffusionmodules/model.py:462-568, Decoder, with ResnetBlock :82-142, AttnBlock
:150-202, Upsample :42-58) on CPU in fp32: the image for a seeded latent AND d(sum(image * G))/d(latent) — the
gradient the alpha optimisation pushes back through `decode_first_stage` (ddpm.py:705-763, plms.py:249-277).

TEST INFRASTRUCTURE; runs only in the build container (the reference tree does not travel).  Weights: the v1 decoder
architecture (v1-inference.yaml:45-66: ch 128, ch_mult 1/2/4/4, 2 res blocks, z_channels 4) with
oracle.sta_oracle.seeded_state_dict tensors — no checkpoint exists offline.  Usage: python oracle/make_golden_vae.py
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import sta_oracle as O  # noqa: E402

REF_SD = Path("/root/reference/attention_optimization/stable-diffusion")
GOLD = ROOT / "tests" / "golden"
DDCONFIG = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                num_res_blocks=2, attn_resolutions=[], dropout=0.0)
LATENT, SEED = 16, 11


def main():
    if "omegaconf" not in sys.modules:  # ldm.util imports nothing from it at module level, but be safe
        sys.modules.setdefault("omegaconf", types.ModuleType("omegaconf"))
    sys.path.insert(0, str(REF_SD))
    import os
    import tempfile

    with tempfile.TemporaryDirectory() as td:  # ldm.modules.attention loads uncond_*.pt lazily only in blocks; CWD-safe
        os.chdir(td)
        from ldm.modules.diffusionmodules.model import Decoder  # the reference module, unmodified

        dec = Decoder(**DDCONFIG).eval()
        os.chdir(ROOT)
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    sd = O.seeded_state_dict(shapes, SEED)
    dec.load_state_dict(sd)
    g = torch.Generator().manual_seed(SEED)
    z = torch.randn(1, 4, LATENT, LATENT, generator=g).requires_grad_(True)
    G = torch.randn(1, 3, LATENT * 8, LATENT * 8, generator=g)
    torch.set_num_threads(os.cpu_count())
    img = dec(z)
    (img * G).sum().backward()
    np.savez_compressed(GOLD / "vae_decoder.npz", image=img.detach().numpy().astype(np.float32),
                        d_latent=z.grad.numpy().astype(np.float32), latent=LATENT, seed=SEED)
    print("vae_decoder.npz written; |img| mean %.4f, |dz| mean %.4f, %d parameters tensors" %
          (img.abs().mean(), z.grad.abs().mean(), len(shapes)))


if __name__ == "__main__":
    main()
