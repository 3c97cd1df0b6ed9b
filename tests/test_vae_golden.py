"""KL-VAE decoder against the committed output of the UNMODIFIED reference Decoder
(tests/golden/vae_decoder.npz, written by oracle/make_golden_vae.py from
/root/reference/.../ldm/modules/diffusionmodules/model.py:462-568 on CPU in fp32): image and d(sum(image*G))/d(latent).

CPU test: the plain torch path of this package's Decoder (same parameter names) reproduces the reference to 1e-4.
GPU test: the fused fp16 NHWC path (sta_groupnorm + bias/residual kernels + token GEMMs + SDPA) — relative L2 error
<= 5e-3 on the image and <= 2e-2 on the gradient (fp16 storage, fp32 statistics / accumulation).
"""
from __future__ import annotations

from pathlib import Path

import numpy as np
import pytest
import torch

from diffusion_spacetime_attn_b200.ldm.models.autoencoder import Decoder
from diffusion_spacetime_attn_b200.ldm.models.diffusion.ddpm import V1_VAE
from oracle import sta_oracle as O

GOLD = Path(__file__).resolve().parent / "golden" / "vae_decoder.npz"


def _decoder_and_inputs():
    gold = np.load(GOLD)
    dec = Decoder(**V1_VAE["ddconfig"]).eval()
    shapes = {k: tuple(v.shape) for k, v in dec.state_dict().items()}
    dec.load_state_dict(O.seeded_state_dict(shapes, int(gold["seed"])), strict=True)
    g = torch.Generator().manual_seed(int(gold["seed"]))
    lat = int(gold["latent"])
    z = torch.randn(1, 4, lat, lat, generator=g)
    G = torch.randn(1, 3, lat * 8, lat * 8, generator=g)
    return dec, z, G, torch.from_numpy(gold["image"]), torch.from_numpy(gold["d_latent"])


def _rel(a, b):
    return float((a.detach().float().cpu() - b).norm() / b.norm())


def test_decoder_torch_path_matches_reference_on_cpu():
    dec, z, G, img_ref, dz_ref = _decoder_and_inputs()
    torch.set_num_threads(8)
    z = z.requires_grad_(True)
    img = dec(z)
    (img * G).sum().backward()
    assert _rel(img, img_ref) < 1e-4
    assert _rel(z.grad, dz_ref) < 1e-4


@pytest.mark.gpu
def test_decoder_fused_fp16_path_matches_reference():
    from diffusion_spacetime_attn_b200 import native, ops

    dec, z, G, img_ref, dz_ref = _decoder_and_inputs()
    dec = dec.cuda().half().requires_grad_(False)
    for m in dec.modules():
        if isinstance(m, torch.nn.GroupNorm):
            m.float()
    dec.to(memory_format=torch.channels_last)
    before, attn_before = ops.launch_count(), (ops.LAUNCHES["sattn_fwd"], ops.LAUNCHES["sattn_bwd"])
    zd = z.cuda().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        img = dec(zd)
        (img.float() * G.cuda()).sum().backward()
    torch.cuda.synchronize()
    assert native.device_error() == 0
    assert ops.launch_count() - before > 100, "the fused kernels did not run"
    # the mid-block AttnBlock (model.py:150-202, one head of d = 512) went through the 512-wide flash kernel, both ways
    assert (ops.LAUNCHES["sattn_fwd"], ops.LAUNCHES["sattn_bwd"]) == (attn_before[0] + 1, attn_before[1] + 2)
    e_img, e_dz = _rel(img, img_ref), _rel(zd.grad, dz_ref)
    assert e_img < 5e-3 and e_dz < 2e-2, (e_img, e_dz)


@pytest.mark.gpu
def test_decoder_flash_attention_agrees_with_the_materialised_formulation(monkeypatch):
    """A/B of the mid-block attention inside the whole decoder: 512-wide flash kernel vs bmm / softmax / bmm with the [N, N]
    score matrix in HBM (what model.py:176-191 does) — same image and latent gradient to fp16 rounding."""
    from diffusion_spacetime_attn_b200 import native, ops

    dec, z, G, _, _ = _decoder_and_inputs()
    dec = dec.cuda().half().requires_grad_(False)
    for m in dec.modules():
        if isinstance(m, torch.nn.GroupNorm):
            m.float()
    dec.to(memory_format=torch.channels_last)
    res = {}
    for arm in ("flash", "materialised"):
        monkeypatch.setenv("STA_VAE_ATTN", arm)
        n0 = ops.LAUNCHES["sattn_fwd"]
        zd = z.cuda().requires_grad_(True)
        with torch.autocast("cuda", dtype=torch.float16):
            img = dec(zd)
            (img.float() * G.cuda()).sum().backward()
        torch.cuda.synchronize()
        assert ops.LAUNCHES["sattn_fwd"] - n0 == (1 if arm == "flash" else 0)
        res[arm] = (img.detach().float().cpu(), zd.grad.float().cpu())
    assert native.device_error() == 0
    assert _rel(res["flash"][0], res["materialised"][0]) < 2e-3
    assert _rel(res["flash"][1], res["materialised"][1]) < 1e-2


@pytest.mark.gpu
def test_graphed_decode_replays_the_eager_decode():
    """graphed.GraphedDifferentiable (the captured forward-with-grad / backward pair the sampler uses for the alpha-optimisation
    tail): several replays with fresh inputs reproduce the eager fp16 decode and its latent gradient (same kernels, same
    order; GroupNorm statistics use atomics, so equality is to fp16 rounding, not bitwise), and still match the reference."""
    from diffusion_spacetime_attn_b200 import native
    from diffusion_spacetime_attn_b200.graphed import GraphedDifferentiable

    dec, z, G, img_ref, dz_ref = _decoder_and_inputs()
    dec = dec.cuda().half().requires_grad_(False)
    for m in dec.modules():
        if isinstance(m, torch.nn.GroupNorm):
            m.float()
    dec.to(memory_format=torch.channels_last)

    def fn(zz):
        with torch.autocast("cuda", dtype=torch.float16):
            return torch.clamp((dec(zz) + 1.0) / 2.0, 0.0, 1.0)

    Gd = G.cuda()
    graphed = GraphedDifferentiable(fn, z.cuda())
    gen = torch.Generator().manual_seed(5)
    for k in range(3):
        zk = (z if k == 2 else torch.randn(z.shape, generator=gen)).cuda()
        za, zb = zk.clone().requires_grad_(True), zk.clone().requires_grad_(True)
        ia = fn(za)
        (ia.float() * Gd).sum().backward()
        ib = graphed(zb)
        (ib.float() * Gd).sum().backward()
        torch.cuda.synchronize()
        assert ib.dtype == ia.dtype and ib.shape == ia.shape
        # same kernels, but cuDNN may autotune the eager call and the captured one to different algorithms: fp16 rounding
        e_img, e_dz = _rel(ib, ia.detach().float().cpu()), _rel(zb.grad, za.grad.float().cpu())
        assert e_img < 5e-3 and e_dz < 2e-2, (k, e_img, e_dz)  # the bounds each fp16 evaluation meets against the fp32 reference
    assert native.device_error() == 0
    with torch.no_grad():  # no graph needed: plain call
        assert _rel(graphed(z.cuda()), ia.detach().float().cpu()) < 1e-3
    # a second differentiable call before the first one's backward runs eagerly (the captured activations are still needed)
    za, zb = z.cuda().requires_grad_(True), z.cuda().requires_grad_(True)
    first, second = graphed(za), graphed(zb)
    assert graphed.pending and second.grad_fn.__class__.__name__ != "_GraphedFnBackward"
    ((first.float() + second.float()) * Gd).sum().backward()
    torch.cuda.synchronize()
    assert not graphed.pending and _rel(za.grad, zb.grad.float().cpu()) < 2e-2
    with pytest.raises(RuntimeError, match="backward without a matching forward"):
        zc = z.cuda().requires_grad_(True)
        out = graphed(zc)
        out.sum().backward(retain_graph=True)
        out.sum().backward()
