"""CPU: host-side logic of the drop-in (schedules, masks, prompt readers, sharding, sampler arithmetic) against the
oracle and against closed forms.  No CUDA needed."""
from __future__ import annotations

import numpy as np
import pytest
import torch

from diffusion_spacetime_attn_b200 import prompts as P
from diffusion_spacetime_attn_b200.ldm.modules.attention import build_object_masks
from diffusion_spacetime_attn_b200.ldm.modules.diffusionmodules import util as U
from diffusion_spacetime_attn_b200.pipeline import WorkItem, shard_prompts, synthetic_layout
from oracle import sta_oracle as O

from pathlib import Path

GOLD = Path(__file__).resolve().parent / "golden"


@pytest.mark.parametrize("S", [10, 50, 100])
def test_schedule_matches_oracle(S):
    betas = U.make_beta_schedule("linear", 1000, linear_start=0.00085, linear_end=0.012)
    acp = np.cumprod(1.0 - betas).astype(np.float32)
    ts = U.make_ddim_timesteps("uniform", S, 1000, verbose=False)
    sig, a, ap = U.make_ddim_sampling_parameters(acp, ts, 0.0, verbose=False)
    sch = O.make_schedule(S)
    assert np.array_equal(ts, sch.timesteps) and ts[-1] == 1000 // S * (S - 1) + 1
    assert np.array_equal(a, sch.alphas) and np.array_equal(ap, sch.alphas_prev) and not sig.any()
    if S == 50:
        assert ts[-1] == 981  # the constant the reference hard-codes (attention.py:240)


def test_timestep_embedding_matches_oracle():
    t = torch.tensor([1, 501, 981])
    assert torch.equal(U.timestep_embedding(t, 320), O.timestep_embedding(t, 320))


@pytest.mark.parametrize("n", [64, 144, 256, 1024, 4096, 9216])
def test_masks_match_oracle_and_reference_semantics(n):
    boxes = [[0.30, 0.50], [0.70, 0.50], [0.0, 1.0]]
    m = build_object_masks(boxes, n, "cpu")
    assert torch.equal(m, O.flat_masks(boxes, n))
    dim = int(n ** 0.5)
    r, c = int(0.5 * dim), int(0.3 * dim)
    assert m[0, r * dim + c] == 1  # the disc centre: x indexes columns, y rows (attention.py:254-261)
    assert m[0].sum() > 0 and m[0].sum() < 0.2 * n  # pi * 0.2^2 = 12.6 % of the latent


def test_plms_closed_form_for_constant_eps():
    """With eps == const every Adams-Bashforth combination returns that constant, so the trajectory is the DDIM
    closed form x_0 = x_T * prod(sqrt(a_prev/a_t)) + e * sum(...): checks the multistep coefficients sum to 1."""
    S = 10
    sch = O.make_schedule(S)
    e = torch.full((1, 4, 8, 8), 0.3)
    x = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(0))
    z = O.plms_trajectory(lambda xx, t, i: e, x, S, sch)
    ref = x.clone()
    for index in reversed(range(S)):
        a_t, a_p = float(sch.alphas[index]), float(sch.alphas_prev[index])
        pred = (ref - (1 - a_t) ** 0.5 * e) / a_t ** 0.5
        ref = a_p ** 0.5 * pred + (1 - a_p) ** 0.5 * e
    assert (z - ref).abs().max() < 1e-5


def test_product_sampler_arithmetic_matches_oracle_on_cpu():
    """PLMSSampler's update/combination code (no UNet involved) against the oracle's restatement."""
    from diffusion_spacetime_attn_b200.ldm.models.diffusion.plms import PLMSSampler

    class FakeModel:
        num_timesteps = 1000
        device = torch.device("cpu")
        alphas_cumprod = torch.tensor(np.cumprod(1.0 - U.make_beta_schedule("linear", 1000, 0.00085, 0.012)), dtype=torch.float32)

        def apply_model_extra(self, x_in, text_index, t_in, c_in, coef=None, bboxs_curr=None, step_time=None):
            # a linear "UNet" whose output depends on x, t and coef, so every code path is exercised
            s = (t_in.float() / 1000.0).reshape(-1, 1, 1, 1)
            k = 1.0 + (coef.sum() if coef is not None else 0.0) * 0.01
            return 0.1 * k * x_in * s + 0.05 * torch.cat([torch.zeros_like(x_in[:1]), torch.ones_like(x_in[1:])])

    S = 10
    sampler = PLMSSampler(FakeModel(), clip_loss_model=torch.nn.Identity(), save_images=False)
    sampler.make_schedule(S, verbose=False)
    x_T = torch.randn(1, 4, 8, 8, generator=torch.Generator().manual_seed(1))
    W = torch.full((1, 2, S), 2.5)
    z = sampler._trajectory(x_T.clone(), torch.zeros(1, 77, 768), torch.zeros(1, 77, 768), 7.5, W, [[0.3, 0.5], [0.7, 0.5]], 0)

    def eps_model(x, t, i):
        out = FakeModel().apply_model_extra(torch.cat([x, x]), 0, torch.full((2,), t), None, coef=W[0, :, i])
        e_u, e_c = out.chunk(2)
        return e_u + 7.5 * (e_c - e_u)

    z_ref = O.plms_trajectory(eps_model, x_T, S)
    assert (z - z_ref).abs().max() < 1e-4


def _linear_unet_eps(x_in, t_in, c_in, coef):
    """The linear stand-in UNet of oracle/make_golden_sampler.py (restated: the fixture script is not imported)."""
    s = (t_in.float() / 1000.0).reshape(-1, 1, 1, 1)
    k = 1.0 + coef.sum() * 0.01
    row = torch.cat([torch.zeros_like(x_in[:1]), torch.ones_like(x_in[1:])])
    return 0.1 * k * x_in * s + 0.02 * torch.roll(x_in, 1, dims=-1) + 0.05 * row * (1.0 + c_in.mean())


@pytest.mark.parametrize("S", [5, 10, 50])
def test_sampler_pinned_by_execution_of_the_reference_plms(S):
    """tests/golden/plms_sampler.npz holds the output of the UNMODIFIED reference `make_schedule` + `p_sample_plms`
    (ldm/models/diffusion/plms.py:81-112, 296-358; driven as plms.py:227-247 does) on a linear stand-in UNet.  The oracle's
    restatement and the product sampler must reproduce the schedule constants, every e_t and the final latent."""
    from diffusion_spacetime_attn_b200.ldm.models.diffusion.plms import PLMSSampler

    fx = np.load(GOLD / "plms_sampler.npz")
    g = torch.Generator().manual_seed(S)
    x_T = torch.randn(1, 4, 8, 8, generator=g)
    c = torch.randn(1, 77, 768, generator=g)
    uc = torch.randn(1, 77, 768, generator=g)
    W = 2.5 + 0.5 * torch.randn(2, S, generator=g)
    # ---- oracle ----
    sch = O.make_schedule(S)
    assert np.array_equal(sch.timesteps, fx[f"S{S}_timesteps"])
    assert np.allclose(sch.alphas, fx[f"S{S}_alphas"], rtol=1e-6) and np.allclose(sch.alphas_prev, fx[f"S{S}_alphas_prev"], rtol=1e-6)
    seen = []

    def eps_model(x, t, i):
        out = _linear_unet_eps(torch.cat([x, x]), torch.full((2,), t), torch.cat([uc, c]), W[:, i])
        e_u, e_c = out.chunk(2)
        e = e_u + 7.5 * (e_c - e_u)
        seen.append(e)
        return e

    z = O.plms_trajectory(eps_model, x_T, S, sch)
    ref_z = torch.from_numpy(fx[f"S{S}_latent"])
    assert (z - ref_z).abs().max().item() < 2e-5 * ref_z.abs().max().item() + 1e-5
    # e_t of step i: the first step evaluates the model twice (plms.py:341-345), so seen[0], seen[2], seen[3], ...
    e_hist = torch.stack([seen[0]] + seen[2:])
    assert (e_hist - torch.from_numpy(fx[f"S{S}_eps"])).abs().max().item() < 2e-5 * float(np.abs(fx[f"S{S}_eps"]).max()) + 1e-5

    # ---- product sampler (host arithmetic only; the stand-in replaces the UNet) ----
    class FakeModel:
        num_timesteps = 1000
        device = torch.device("cpu")
        alphas_cumprod = torch.tensor(np.cumprod(1.0 - U.make_beta_schedule("linear", 1000, 0.00085, 0.012)), dtype=torch.float32)

        def apply_model_extra(self, x_in, text_index, t_in, c_in, coef=None, bboxs_curr=None, step_time=None):
            return _linear_unet_eps(x_in, t_in, c_in, coef)

    sampler = PLMSSampler(FakeModel(), clip_loss_model=torch.nn.Identity(), save_images=False)
    sampler.make_schedule(S, verbose=False)
    assert np.array_equal(np.asarray(sampler.ddim_timesteps), fx[f"S{S}_timesteps"])
    zp = sampler._trajectory(x_T.clone(), c, uc, 7.5, W.unsqueeze(0), [[0.3, 0.5], [0.7, 0.5]], 0)
    assert (zp - ref_z).abs().max().item() < 2e-5 * ref_z.abs().max().item() + 1e-5


def test_prompt_readers_and_layout():
    recs = P.read_gpt(P.SYNTHETIC_GPT)
    assert recs[0] == ("a red cube left of a blue sphere", ["red cube", "blue sphere"])  # BASELINE.json configs[0]
    assert all(2 <= len(o) <= 3 for _, o in recs)
    items = P.build_work_items(recs, start=10)
    assert items[0].prompt_idx == 10 and items[0].seed == 1
    for it in items:
        assert len(it.bboxes) == len(it.object_names)
        assert all(0.0 <= v <= 1.0 for b in it.bboxes for v in b)
    lay = {recs[0][0]: {"silver bed": [0.574, 0.503], "white couch": [0.269, 0.442]}}  # README.md:56-62 format
    it = P.build_work_items(recs[:1], layouts=lay)[0]
    assert it.object_names == ["silver bed", "white couch"] and it.bboxes[0] == [0.574, 0.503]
    assert synthetic_layout(["a", "b"], "p") == synthetic_layout(["a", "b"], "p")
    assert P.guess_objects("The bed is below the cat.", 2) == ["bed", "cat"]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
def test_sharding_is_a_partition_with_global_indices(world):
    n = 37
    shards = [shard_prompts(n, r, world) for r in range(world)]
    flat = sorted(i for s in shards for i in s)
    assert flat == list(range(n))
    assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_object_crop_box_matches_reference_expression():
    assert O.object_crop_box([0.3, 0.5]) == (int(512 * 0.3), int(512 * 0.7), int(512 * (0.3 - 0.2)), int(512 * 0.5))
    assert O.object_crop_box([0.05, 0.95]) == (int(512 * 0.75), 512, 0, int(512 * 0.25))
    assert float(O.alpha_init(2, 50)[0, 0]) == 2.5 and tuple(O.alpha_init(3, 10).shape) == (3, 10)


def test_clip_global_resample_matrix_equals_upsample_avgpool():
    """plms.py:26-27,41: Upsample(x7, nearest) -> AvgPool2d(16) == A @ img @ A^T, incl. gradients."""
    import torch.nn as nn

    from diffusion_spacetime_attn_b200.ldm.modules.encoders.clip_loss import upsample_avgpool_matrix

    g = torch.Generator().manual_seed(3)
    for h, w in ((512, 512), (64, 96), (16, 16)):
        img = torch.rand(3, h, w, generator=g, requires_grad=True)
        ref = nn.AvgPool2d(16)(nn.Upsample(scale_factor=7)(img.unsqueeze(0)))[0]
        G = torch.randn(ref.shape, generator=g)
        (gref,) = torch.autograd.grad((ref * G).sum(), img)
        A, B = upsample_avgpool_matrix(h), upsample_avgpool_matrix(w)
        out = A @ img @ B.t()
        (gout,) = torch.autograd.grad((out * G).sum(), img)
        assert out.shape == ref.shape == (3, h * 7 // 16, w * 7 // 16)
        assert (out - ref).abs().max().item() < 1e-5
        assert (gout - gref).abs().max().item() < 1e-5
        assert torch.allclose(A.sum(1), torch.ones(A.shape[0]))


def test_clip_global_loss_accepts_non_512_images():
    """configs[4] decodes 768 px images (the reference hard-codes 512): the x7 upsample is kept and the
    pooling window becomes 7 * size / 224 so that CLIP still sees 224 x 224; the 512 px case is the reference's literal one."""
    import torch.nn as nn

    from diffusion_spacetime_attn_b200.ldm.modules.encoders.clip_loss import DCLIPLoss, upsample_avgpool_matrix

    g = torch.Generator().manual_seed(5)
    img = torch.rand(3, 768, 768, generator=g)
    A = upsample_avgpool_matrix(768, 7, 24)
    ref = nn.AvgPool2d(24)(nn.Upsample(scale_factor=7)(img.unsqueeze(0)))[0]
    assert ref.shape == (3, 224, 224) and (A @ img @ A.t() - ref).abs().max().item() < 1e-5
    loss = DCLIPLoss(device="cpu")
    for size in (512, 768, 100):  # 100: no integer window -> bilinear resize
        out = loss.forward_2(torch.rand(3, size, size, generator=g), "a red cube")
        assert out.shape == (1,) and bool(torch.isfinite(out).all())


def test_timestep_bias_table_matches_per_block_projection():
    """UNetModel._timestep_bias: row t of the table == emb_layers(time_embed(timestep_embedding(t))) + conv1.bias of every
    ResBlock (openaimodel.py:217-223,259-268), computed the ordinary way in fp32 (the table is fp16)."""
    import torch.nn.functional as F

    from diffusion_spacetime_attn_b200.ldm.modules.diffusionmodules.openaimodel import ResBlock, UNetModel
    from diffusion_spacetime_attn_b200.ldm.modules.diffusionmodules.util import timestep_embedding

    torch.manual_seed(0)
    m = UNetModel(attention_resolutions=(1,), num_res_blocks=1, channel_mult=(1, 2), model_channels=32, num_heads=4,
                  context_dim=16).eval().requires_grad_(False)
    for p in m.parameters():  # zero-initialised convs / random biases: make every term visible
        if p.dim() == 1:
            p.copy_(torch.randn_like(p) * 0.3)
    m.XB_TABLE_STEPS = 64
    t = torch.tensor([3, 41], dtype=torch.long)
    xb, offs = m._timestep_bias(t)
    blocks = [b for b in m.modules() if isinstance(b, ResBlock)]
    assert xb.shape == (2, sum(b.out_channels for b in blocks)) and xb.dtype == torch.float16
    emb = m.time_embed(timestep_embedding(t, m.model_channels))
    for b in blocks:
        ref = b.emb_layers(emb) + b.in_layers[2].bias
        got = xb[:, offs[id(b)]:offs[id(b)] + b.out_channels].float()
        assert (got - ref).abs().max().item() < 2e-2 * (1 + ref.abs().max().item())
    again, _ = m._timestep_bias(t)
    assert again.data_ptr() != xb.data_ptr() and m.__dict__["_xb_cache"][1].shape[0] == 64  # cached table, fresh gather


def test_layout_shape_handles_shared_per_prompt_and_empty_layouts():
    from diffusion_spacetime_attn_b200.ldm.modules.attention import layout_shape

    assert layout_shape(None) == (False, 0) and layout_shape([]) == (False, 0)
    assert layout_shape([[0.3, 0.5], [0.7, 0.5]]) == (False, 2)            # the reference's form: one layout, 2 objects
    assert layout_shape([[[0.3, 0.5]], [[0.7, 0.5]]]) == (True, 1)         # B = 2 prompts, one object each
    assert layout_shape([[], []]) == (True, 0)                             # B = 2 prompts WITHOUT objects (used to raise)
    assert layout_shape([(0.3, 0.5)]) == (False, 1)


def test_checkpoint_without_real_clip_is_an_error_not_a_silent_fallback(monkeypatch):
    """with --ckpt the text stage / CLIP loss must be real, or the caller must opt in to synthetic stand-ins."""
    from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline

    for var in ("STA_CLIP_L_PATH", "STA_CLIP_B32_PATH", "STA_CLIP_TOKENIZER_PATH"):
        monkeypatch.delenv(var, raising=False)
    pipe = object.__new__(SpaceTimeAttnPipeline)  # no networks needed for the policy itself
    pipe.data = "checkpoint"
    with pytest.raises(RuntimeError, match="allow_synthetic_conditioning"):
        pipe._real_text_stage({}, None)
    with pytest.raises(RuntimeError, match="allow_synthetic_conditioning"):
        pipe._real_clip_loss(False)
    with pytest.warns(UserWarning, match="synthetic"):
        assert pipe._real_text_stage({}, True) is None
    assert "SYNTHETIC" in pipe.data
