"""GPU parity of the CLIP-loss front end (reference ldm/models/diffusion/plms.py:26-28, 31, 41 and the
crop arithmetic :256-273): the product's A @ img @ A^T resample and its crop + bilinear resize against the reference's
literal torch pipeline — nn.Upsample(scale_factor=7) -> nn.AvgPool2d(16) (a 3 x 3584 x 3584 fp32 intermediate) and
torchvision Resize((224, 224)) — on the device, values AND gradients w.r.t. the decoded image, and the loss end to end
through the same (seeded) CLIP ViT-B/32.  fp32 in, fp32 out: tolerance 1e-5 absolute on [0, 1] images."""
from __future__ import annotations

import pytest
import torch
import torch.nn as nn

from diffusion_spacetime_attn_b200.ldm.modules.encoders.clip_loss import DCLIPLoss


@pytest.mark.gpu
def test_global_resample_and_loss_match_the_reference_pipeline_on_gpu():
    g = torch.Generator().manual_seed(11)
    loss = DCLIPLoss(device="cuda", seed=3)
    img = torch.rand(3, 512, 512, generator=g).cuda().requires_grad_(True)
    # ---- reference: plms.py:38-45 (forward_2) with the same CLIP weights ----
    ref_small = nn.AvgPool2d(kernel_size=16)(nn.Upsample(scale_factor=7)(img.unsqueeze(0)))
    ref = 1 - torch.nn.CosineSimilarity()(loss.model.encode_image(ref_small).float(), loss._text_feat("a red cube"))
    (g_ref,) = torch.autograd.grad(ref.sum(), img)
    # ---- product ----
    out = loss.forward_2(img, "a red cube")
    (g_out,) = torch.autograd.grad(out.sum(), img)
    assert abs(float(out) - float(ref)) < 1e-5
    assert (g_out - g_ref).abs().max().item() <= 1e-5 * max(1.0, g_ref.abs().max().item()) + 1e-7
    # the resample alone, against the materialised 3584^2 image
    A = loss._resample[(512, img.device)]
    assert (A @ img.detach() @ A.t() - ref_small[0].detach()).abs().max().item() < 1e-5


@pytest.mark.gpu
def test_object_crop_loss_matches_the_reference_pipeline_on_gpu():
    """plms.py:256-273: crop [y - 0.2, y + 0.2] x [x - 0.2, x + 0.2] (clamped, x 512 px), Resize((224, 224)), forward_3."""
    import torchvision.transforms as transforms

    g = torch.Generator().manual_seed(12)
    loss = DCLIPLoss(device="cuda", seed=3)
    img = torch.rand(3, 512, 512, generator=g).cuda().requires_grad_(True)
    for box in ([0.30, 0.50], [0.05, 0.95], [0.70, 0.50]):
        x1, x2 = max(box[0] - 0.2, 0), min(box[0] + 0.2, 1)
        y1, y2 = max(box[1] - 0.2, 0), min(box[1] + 0.2, 1)
        crop = img[:, int(512 * y1):int(512 * y2), int(512 * x1):int(512 * x2)]
        ref_in = transforms.Resize((224, 224))(crop).unsqueeze(0)  # plms.py:28,31 (antialiased bilinear)
        ref = 1 - torch.nn.CosineSimilarity()(loss.model.encode_image(ref_in).float(), loss._text_feat("A photo of cube"))
        (g_ref,) = torch.autograd.grad(ref.sum(), img)
        out = loss.forward_3(crop, "A photo of cube")
        (g_out,) = torch.autograd.grad(out.sum(), img)
        assert abs(float(out) - float(ref)) < 1e-5
        assert (g_out - g_ref).abs().max().item() <= 1e-5 * max(1.0, g_ref.abs().max().item()) + 1e-7


@pytest.mark.gpu
@pytest.mark.parametrize("graphed", [False, True], ids=["eager", "graph"])
def test_batched_loss_equals_the_per_image_loss(graphed):
    """The sampler sends the full frame and every object crop through the CLIP image tower in ONE pass (and, under CUDA graphs,
    through a captured forward/backward pair) instead of the reference's one call per image (plms.py:252-273).  fp32 weights:
    loss and d(loss)/d(image) agree with the literal per-image loop to 1e-4 relative (batch-3 GEMMs round differently from
    batch-1 GEMMs), for two consecutive images through the same captured graph."""
    from diffusion_spacetime_attn_b200.ldm.models.diffusion.plms import PLMSSampler

    class _Model:  # the sampler only needs .device and a scheduler-free loss path here
        device = torch.device("cuda")

    loss = DCLIPLoss(device="cuda", seed=3)
    loss.graph_encode = graphed
    sampler = PLMSSampler.__new__(PLMSSampler)
    sampler.clip_loss_model, sampler.local_loss_weight = loss, 5.0
    g = torch.Generator().manual_seed(13)
    boxes, names = [[[0.30, 0.50], [0.70, 0.45]]], [["the red cube", "The blue sphere"]]
    for _ in range(2):
        img = torch.rand(1, 3, 512, 512, generator=g).cuda()
        a, b = img.clone().requires_grad_(True), img.clone().requires_grad_(True)
        loss.graph_encode = False  # the literal loop, eagerly
        ref, ref_pp = sampler._loss_per_image(a, ["a red cube left of a blue sphere"], boxes, names)
        loss.graph_encode = graphed
        out, out_pp = sampler._loss(b, ["a red cube left of a blue sphere"], boxes, names)
        ref.backward()
        out.backward()
        assert abs(float(out) - float(ref)) <= 1e-4 * abs(float(ref))
        assert abs(float(out_pp[0]) - float(ref_pp[0])) <= 1e-4 * abs(float(ref))
        assert (a.grad - b.grad).abs().max().item() <= 1e-3 * a.grad.abs().max().item()
    assert bool(loss._graphed_encode) == graphed
    if graphed:  # a weight rewritten after the capture: the captured pair is dropped (warning) and re-captured, never replayed stale
        first = next(iter(loss._graphed_encode.values()))
        with torch.no_grad():
            next(loss.model.visual.parameters()).mul_(0.5)
        a, b = img.clone().requires_grad_(True), img.clone().requires_grad_(True)
        loss.graph_encode = False
        ref, _ = sampler._loss_per_image(a, ["a red cube left of a blue sphere"], boxes, names)
        loss.graph_encode = True
        with pytest.warns(UserWarning, match="weights changed after CUDA-graph capture"):
            out, _ = sampler._loss(b, ["a red cube left of a blue sphere"], boxes, names)
        assert abs(float(out) - float(ref)) <= 1e-4 * abs(float(ref))
        assert next(iter(loss._graphed_encode.values())) is not first
