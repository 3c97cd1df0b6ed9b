"""PLMS sampler WITH the alpha (blend-weight) optimisation, keeping the reference's public API
(/root/reference/.../ldm/models/diffusion/plms.py: PLMSSampler.sample :114-180, plms_sampling :182-293,
p_sample_plms :296-358, DCLIPLoss :21-61).

Behaviour kept: 3 "epochs", each a full differentiable S-step trajectory from the same x_T (S + 1 UNet evaluations:
the first step evaluates the model twice with the same coefficient column, :341-345), VAE decode, CLIP loss
(global + 5 x per-object crops, :252-273), backward to `weighting_parameter`, one Adam(lr=0.005) step; the image of
the LAST epoch is written to result_outputs/final{epoch}_s{seed}_index_{idx}.png (:280-288).

Generalised (the reference hard-codes them): the number of columns of `weighting_parameter` is S (not
50); the "first step" the attention blocks key on is the schedule's largest timestep (not 981); B > 1 prompts per call
(rows [uc x B, c x B]); the crop size follows the decoded image (not 512).  Host-side differences that do not change
the arithmetic: per-step scalars are Python floats (no per-step device tensors), the timestep reaches the attention
blocks as a host int (no sync), local embeddings are handed over in memory (`local_conditionings=`) instead of
through c{i}_*.pt files.
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence

import numpy as np
import torch

from ...modules.diffusionmodules.util import make_ddim_sampling_parameters, make_ddim_timesteps

mode = "fix_radius_0p2"

# NVTX ranges (prompt / epoch / step / decode+loss / backward) for nsys / ncu --nvtx timelines: STA_NVTX=1.  Off by default:
# a range push/pop per step is host work inside the sampling loop.
_NVTX = os.environ.get("STA_NVTX", "0") == "1"


class _nvtx_range:
    def __init__(self, name):
        self.name = name

    def __enter__(self):
        if _NVTX:
            torch.cuda.nvtx.range_push(self.name)

    def __exit__(self, *exc):
        if _NVTX:
            torch.cuda.nvtx.range_pop()
        return False


def _per_prompt(value, B):
    """bboxs / names may be given once (shared by all prompts) or per prompt."""
    if value is None:
        return [[] for _ in range(B)]
    if len(value) and isinstance(value[0], (list, tuple)) and len(value[0]) and isinstance(value[0][0], (list, tuple)):
        return [list(v) for v in value]
    if len(value) and isinstance(value[0], (list, tuple)) and (len(value[0]) == 0 or isinstance(value[0][0], str)):
        return [list(v) for v in value]
    return [list(value) for _ in range(B)]


class PLMSSampler(object):
    method = "plms"

    def __init__(self, model, schedule="linear", clip_loss_model=None, num_epochs=3, lr=0.005,
                 weight_initialize_coef=5.0, local_loss_weight=5.0, save_images=True, out_dir="result_outputs",
                 verbose=False, loss_fn=None, decode_fn=None, **kwargs):
        super().__init__()
        self.model = model
        self.ddpm_num_timesteps = model.num_timesteps
        self.schedule = schedule
        if clip_loss_model is None:
            from ...modules.encoders.clip_loss import DCLIPLoss

            clip_loss_model = DCLIPLoss(device=model.device)
        self.clip_loss_model = clip_loss_model
        self.clip_loss_model.requires_grad_(False)
        self.num_epochs, self.lr = num_epochs, lr
        self.weight_initialize_coef, self.local_loss_weight = weight_initialize_coef, local_loss_weight
        self.save_images, self.out_dir, self.verbose = save_images, out_dir, verbose
        # test hooks: replace the VAE decode / CLIP loss by any differentiable stand-ins (both are outside the kernels)
        self.loss_fn = loss_fn or self._loss
        self.decode_fn = decode_fn or self._decode
        self._graphed_decode = {}  # latent shape -> graphed.GraphedDifferentiable (CUDA-graph execution only)
        self.last_result = None

    # ------------------------------------------------------------------------------------------------
    # Larger decodes (config 5: 2 x 96 x 96) stay eager: their saved activations would sit in a private graph pool for the
    # life of the pipeline next to 27 activation slots that already fill the GPU.
    GRAPHED_DECODE_MAX_LATENT = 2 * 64 * 64

    def _decode_eager(self, z):
        return torch.clamp((self.model.decode_first_stage(z) + 1.0) / 2.0, 0.0, 1.0)  # plms.py:249-250

    def _decode(self, z):
        """decode_first_stage + the clamp to [0, 1].  When the UNet runs under CUDA graphs (the model carries a graph runner)
        the differentiable decode — fixed shapes, frozen weights, ~1000 eager launches forward + backward — is one captured
        (forward-with-grad, backward) graph pair per latent shape."""
        graphable = (getattr(self.model, "graph_runner", None) is not None and z.is_cuda and torch.is_grad_enabled()
                     and z.requires_grad and z.shape[0] * z.shape[2] * z.shape[3] <= self.GRAPHED_DECODE_MAX_LATENT
                     and os.environ.get("STA_GRAPH_DECODE", "1") != "0")
        if not graphable:
            return self._decode_eager(z)
        from ....graphed import GraphedDifferentiable, drop_if_weights_changed

        vae = getattr(self.model, "first_stage_model", None)
        if vae is not None:
            drop_if_weights_changed(self, "_graphed_decode_fp", vae, self._graphed_decode, "VAE decoder")
        amp = torch.is_autocast_enabled()  # the graph replays what the caller's autocast state would have computed eagerly
        key = (tuple(z.shape), z.dtype, amp)
        g = self._graphed_decode.get(key)
        if g is None:

            def fn(zz):
                with torch.autocast("cuda", dtype=torch.float16, enabled=amp):
                    return self._decode_eager(zz)

            g = self._graphed_decode[key] = GraphedDifferentiable(fn, z.detach())
        return g(z)

    # ------------------------------------------------------------------------------------------------
    def make_schedule(self, ddim_num_steps, ddim_discretize="uniform", ddim_eta=0.0, verbose=True):
        if ddim_eta != 0:
            raise ValueError("ddim_eta must be 0 for PLMS")
        self.ddim_timesteps = make_ddim_timesteps(ddim_discretize, ddim_num_steps, self.ddpm_num_timesteps, verbose=False)
        acp = self.model.alphas_cumprod.detach().float().cpu().numpy()
        sigmas, alphas, alphas_prev = make_ddim_sampling_parameters(acp, self.ddim_timesteps, ddim_eta, verbose=False)
        self.ddim_sigmas, self.ddim_alphas, self.ddim_alphas_prev = sigmas, alphas, alphas_prev
        self.ddim_sqrt_one_minus_alphas = np.sqrt(1.0 - alphas)
        dev = self.model.device
        self._ts = {int(t): torch.full((1,), int(t), device=dev, dtype=torch.long) for t in self.ddim_timesteps}

    def sample(self, S, batch_size, shape, conditioning=None, callback=None, normals_sequence=None, img_callback=None,
               quantize_x0=False, eta=0.0, mask=None, x0=None, temperature=1.0, noise_dropout=0.0,
               score_corrector=None, corrector_kwargs=None, verbose=True, x_T=None, log_every_t=100,
               unconditional_guidance_scale=1.0, unconditional_conditioning=None, text_index=None, curr_text="",
               bboxs_curr=None, seed=None, prompt_idx=None, object_names=None, local_conditionings=None,
               optimize_alpha=True, alpha=None, **kwargs):
        if conditioning is not None and conditioning.shape[0] != batch_size:
            print(f"Warning: Got {conditioning.shape[0]} conditionings but batch-size is {batch_size}")
        if mask is not None or x0 is not None or score_corrector is not None or quantize_x0:
            raise NotImplementedError("inpainting masks / score correctors are not on the txt2img-* path")
        self.make_schedule(ddim_num_steps=S, ddim_eta=eta, verbose=False)
        C, H, W = shape
        self.plms_sampling(conditioning, (batch_size, C, H, W), x_T=x_T,
                           unconditional_guidance_scale=unconditional_guidance_scale,
                           unconditional_conditioning=unconditional_conditioning, text_index=text_index,
                           curr_text=curr_text, bboxs_curr=bboxs_curr, seed=seed, prompt_idx=prompt_idx,
                           object_names=object_names, local_conditionings=local_conditionings,
                           optimize_alpha=optimize_alpha, alpha=alpha)
        return None

    # ------------------------------------------------------------------------------------------------
    def _trajectory(self, img, cond, uc, scale, W, bboxes, text_index):
        time_range = np.flip(self.ddim_timesteps)
        total = len(time_range)
        old_eps: List[torch.Tensor] = []
        for i, step in enumerate(time_range):
            index = total - i - 1
            t_next = int(time_range[min(i + 1, total - 1)])
            coef = W[:, :, i] if W is not None else None
            with _nvtx_range(f"{self.method}_step_{i}_t{int(step)}"):
                img, _, e_t = self.p_sample_plms(img, cond, int(step), index=index, unconditional_guidance_scale=scale,
                                                 unconditional_conditioning=uc, old_eps=old_eps, t_next=t_next,
                                                 text_index=text_index, coef=coef, bboxs_curr=bboxes)
            old_eps.append(e_t)
            if len(old_eps) >= 4:
                old_eps.pop(0)
        return img

    def _loss_weights(self, weights, device):
        key = (tuple(weights), device)
        if getattr(self, "_lw_cache", (None, None))[0] != key:
            self._lw_cache = (key, torch.tensor(weights, dtype=torch.float32, device=device))
        return self._lw_cache[1]

    def _loss_per_image(self, images, texts, bboxes_pp, names_pp):
        """plms.py:252-273 literally: one forward_2 / forward_3 call per image."""
        total = images.new_zeros((), dtype=torch.float32)
        per_prompt = []
        for b in range(images.shape[0]):
            img = images[b].float()
            size_y, size_x = img.shape[1], img.shape[2]
            loss = self.clip_loss_model.forward_2(img, texts[b]).sum()
            for box, name in zip(bboxes_pp[b], names_pp[b]):
                x1, x2 = max(box[0] - 0.2, 0), min(box[0] + 0.2, 1)
                y1, y2 = max(box[1] - 0.2, 0), min(box[1] + 0.2, 1)
                obj = name.lower().replace("the ", "")
                crop = img[:, int(size_y * y1):int(size_y * y2), int(size_x * x1):int(size_x * x2)]
                loss = loss + self.local_loss_weight * self.clip_loss_model.forward_3(crop, "A photo of " + obj).sum()
            per_prompt.append(loss)
            total = total + loss
        return total, per_prompt

    def _loss(self, images, texts, bboxes_pp, names_pp):
        """plms.py:252-273 per prompt; images [B, 3, Hpx, Wpx] in [0, 1]."""
        clip = self.clip_loss_model
        if not hasattr(clip, "one_minus_cos_batched"):  # a drop-in loss model with only forward_2 / forward_3
            return self._loss_per_image(images, texts, bboxes_pp, names_pp)
        # every image of every prompt (the full frame and one crop per object) through the CLIP image tower in ONE pass
        resized, prompts, weights, first = [], [], [], []
        for b in range(images.shape[0]):
            img = images[b].float()
            size_y, size_x = img.shape[1], img.shape[2]
            first.append(len(resized))
            resized.append(clip.resize_global(img))
            prompts.append(texts[b])
            weights.append(1.0)
            for box, name in zip(bboxes_pp[b], names_pp[b]):
                x1, x2 = max(box[0] - 0.2, 0), min(box[0] + 0.2, 1)
                y1, y2 = max(box[1] - 0.2, 0), min(box[1] + 0.2, 1)
                obj = name.lower().replace("the ", "")
                crop = img[:, int(size_y * y1):int(size_y * y2), int(size_x * x1):int(size_x * x2)]
                resized.append(clip.resize_crop(crop))
                prompts.append("A photo of " + obj)
                weights.append(self.local_loss_weight)
        first.append(len(resized))
        terms = clip.one_minus_cos_batched(torch.cat(resized, dim=0), prompts)
        if any(w != 1.0 for w in weights):
            terms = terms * self._loss_weights(weights, terms.device)
        per_prompt = [terms[first[b]:first[b + 1]].sum() for b in range(images.shape[0])]
        total = per_prompt[0] if len(per_prompt) == 1 else terms.sum()
        return total, per_prompt

    def plms_sampling(self, cond, shape, x_T=None, unconditional_guidance_scale=1.0, unconditional_conditioning=None,
                      text_index=None, curr_text="", bboxs_curr=None, seed=None, prompt_idx=None, object_names=None,
                      local_conditionings=None, optimize_alpha=True, alpha=None):
        assert seed is not None
        device = self.model.device
        B = shape[0]
        bboxes_pp = _per_prompt(bboxs_curr, B)
        names_pp = _per_prompt(object_names, B)
        n_obj = len(bboxes_pp[0])
        assert all(len(bb) == n_obj for bb in bboxes_pp), "prompts of one batch must have the same number of objects"
        assert all(len(nm) == len(bb) for nm, bb in zip(names_pp, bboxes_pp))
        texts = [curr_text] * B if isinstance(curr_text, str) else list(curr_text)
        idxs = [prompt_idx] * B if not isinstance(prompt_idx, (list, tuple)) else list(prompt_idx)
        img_input = torch.randn(shape, device=device) if x_T is None else x_T
        S = len(self.ddim_timesteps)

        unet = self.model.model.diffusion_model
        bboxes_arg = bboxes_pp if B > 1 else bboxes_pp[0]
        runner = getattr(self.model, "graph_runner", None)
        if runner is not None:  # CUDA-graph execution: fix the per-prompt state once, capture on first use
            runner.begin_prompt(shape, torch.cat([unconditional_conditioning, cond]), local_conditionings, bboxes_arg,
                                first_timestep=int(self.ddim_timesteps[-1]))
        else:
            unet.set_local_contexts(local_conditionings, first_timestep=int(self.ddim_timesteps[-1]))

        if n_obj:
            W = torch.full((B, n_obj, S), self.weight_initialize_coef / n_obj, device=device, dtype=torch.float32)
            if alpha is not None:
                W = torch.as_tensor(alpha, device=device, dtype=torch.float32).reshape(-1, n_obj, S).expand(B, n_obj, S).clone()
        else:
            W = torch.zeros((B, 0, S), device=device, dtype=torch.float32)
        do_opt = bool(optimize_alpha and n_obj > 0)
        optimizer = None
        if do_opt:
            W.requires_grad_(True)
            optimizer = torch.optim.Adam([W], lr=self.lr)
        epochs = self.num_epochs if do_opt else 1
        losses, img, decoded, alpha_grads = [], None, None, []
        for epoch in range(epochs):
            if runner is not None:
                runner.new_trajectory()
            with torch.set_grad_enabled(do_opt):
                with _nvtx_range(f"alpha_epoch_{epoch}_trajectory"):
                    img = self._trajectory(img_input.clone(), cond, unconditional_conditioning,
                                           unconditional_guidance_scale, W if n_obj else None, bboxes_arg, text_index)
                if do_opt or self.save_images:
                    with _nvtx_range("vae_decode"):
                        decoded = self.decode_fn(img)  # (decode + 1) / 2 clamped to [0, 1]   (plms.py:249-250)
                if do_opt:
                    with _nvtx_range("clip_loss"):
                        loss, per_prompt = self.loss_fn(decoded, texts, bboxes_pp, names_pp)
                    optimizer.zero_grad(set_to_none=True)
                    with _nvtx_range(f"alpha_epoch_{epoch}_backward"):
                        loss.backward()
                    alpha_grads.append(W.grad.detach().clone())  # dL/dalpha of this epoch (diagnostics / parity tests)
                    optimizer.step()
                    losses.append(torch.stack(per_prompt).detach())  # read back after the last epoch: no host sync per epoch
            if epoch == epochs - 1 and self.save_images and decoded is not None:
                self._save(decoded.detach(), epoch if do_opt else 2, seed, idxs)
        losses = [[float(v) for v in t.cpu()] for t in losses]
        if runner is not None:
            runner.active = None
        unet.set_local_contexts(None)
        self.last_result = {"latent": img.detach(), "weighting_parameter": W.detach(), "losses": losses,
                            "image": decoded.detach() if decoded is not None else None, "alpha_grads": alpha_grads}
        return None

    def _save(self, images, epoch, seed, idxs):
        from PIL import Image

        os.makedirs(self.out_dir, exist_ok=True)
        arr = (255.0 * images.float().cpu().numpy()).transpose(0, 2, 3, 1).astype(np.uint8)
        for b, idx in enumerate(idxs):
            Image.fromarray(arr[b]).save(os.path.join(self.out_dir, "final%d_s%d_index_%d.png" % (epoch, seed, idx or 0)))

    # ------------------------------------------------------------------------------------------------
    def _model_eps(self, x, t_int, c, uc, text_index, coef, bboxs_curr):
        """The UNet's two output rows [uncond x B, cond x B] through apply_model_extra (plms.py:304-307)."""
        B = x.shape[0]
        t_in = self._ts[t_int].expand(2 * B)
        x_in = torch.cat([x, x])
        c_in = torch.cat([uc, c])
        return self.model.apply_model_extra(x_in, text_index, t_in, c_in, coef=coef, bboxs_curr=bboxs_curr, step_time=t_int)

    def _model_output(self, x, t_int, c, uc, scale, text_index, coef, bboxs_curr):
        """get_model_output (plms.py:299-308): the two rows, then classifier-free guidance."""
        e_t_uncond, e_t = self._model_eps(x, t_int, c, uc, text_index, coef, bboxs_curr).chunk(2)
        return e_t_uncond + scale * (e_t - e_t_uncond)

    def _step_consts(self, index):
        """(a_x, a_e, p_x, p_e) with x_prev = a_x x + a_e e' and pred_x0 = p_x x + p_e e' (plms.py:321-338, eta = 0)."""
        a_t, a_prev = float(self.ddim_alphas[index]), float(self.ddim_alphas_prev[index])
        sigma_t, sqrt_1m = float(self.ddim_sigmas[index]), float(self.ddim_sqrt_one_minus_alphas[index])
        p_x, p_e = 1.0 / (a_t ** 0.5), -sqrt_1m / (a_t ** 0.5)
        return (a_prev ** 0.5) * p_x, ((1.0 - a_prev - sigma_t ** 2) ** 0.5) + (a_prev ** 0.5) * p_e, p_x, p_e

    _AB = {0: (1.0, ()), 1: (3 / 2, (-1 / 2,)), 2: (23 / 12, (-16 / 12, 5 / 12)), 3: (55 / 24, (-59 / 24, 37 / 24, -9 / 24))}

    def _p_sample_fused(self, x, eps, index, scale, old_eps, model_eps_at):
        """p_sample_plms with every elementwise operation of the step in ONE kernel launch (ops.plms_step) and one more in
        backward; `eps` is the UNet output at (x, t), `model_eps_at(x_prev)` evaluates it at t_next (first PLMS step only)."""
        from .... import ops as _ops

        consts = self._step_consts(index)
        if self.method == "ddim":
            x_prev, e_t, pred = _ops.plms_step(eps, x, [], scale, 1.0, [], *consts)
            return x_prev, pred, e_t
        if len(old_eps) == 0:  # pseudo improved Euler: e' = (e_t + e_t_next) / 2 with the SAME coefficient column
            x_mid, e_t, _ = _ops.plms_step(eps, x, [], scale, 1.0, [], *consts)
            x_prev, _, pred = _ops.plms_step(model_eps_at(x_mid), x, [e_t], scale, 0.5, [0.5], *consts)
            return x_prev, pred, e_t
        k = min(len(old_eps), 3)
        w_e, w_old = self._AB[k]
        x_prev, e_t, pred = _ops.plms_step(eps, x, [old_eps[-1 - j] for j in range(k)], scale, w_e, list(w_old), *consts)
        return x_prev, pred, e_t

    def _x_prev_and_pred_x0(self, x, e_t, index):
        a_t, a_prev = float(self.ddim_alphas[index]), float(self.ddim_alphas_prev[index])
        sigma_t, sqrt_1m = float(self.ddim_sigmas[index]), float(self.ddim_sqrt_one_minus_alphas[index])
        pred_x0 = (x - sqrt_1m * e_t) / (a_t ** 0.5)
        dir_xt = ((1.0 - a_prev - sigma_t ** 2) ** 0.5) * e_t
        return (a_prev ** 0.5) * pred_x0 + dir_xt, pred_x0  # eta = 0: no noise term

    def p_sample_plms(self, x, c, t, index, repeat_noise=False, use_original_steps=False, quantize_denoised=False,
                      temperature=1.0, noise_dropout=0.0, score_corrector=None, corrector_kwargs=None,
                      unconditional_guidance_scale=1.0, unconditional_conditioning=None, old_eps=None, t_next=None,
                      text_index=None, coef=None, bboxs_curr=None):
        t_int = int(t if not torch.is_tensor(t) else t.flatten()[0].item())
        tn_int = int(t_next if not torch.is_tensor(t_next) else t_next.flatten()[0].item())
        from .... import ops as _ops

        eps = self._model_eps(x, t_int, c, unconditional_conditioning, text_index, coef, bboxs_curr)
        if _ops.plms_step_usable(eps, x, old_eps[-3:]):
            return self._p_sample_fused(
                x, eps, index, unconditional_guidance_scale, old_eps,
                lambda xp: self._model_eps(xp, tn_int, c, unconditional_conditioning, text_index, coef, bboxs_curr))
        out = lambda xx, tt: self._model_output(xx, tt, c, unconditional_conditioning, unconditional_guidance_scale,
                                                text_index, coef, bboxs_curr)
        e_u0, e_c0 = eps.chunk(2)
        e_t = e_u0 + unconditional_guidance_scale * (e_c0 - e_u0)
        if self.method == "ddim":
            e_t_prime = e_t
        elif len(old_eps) == 0:  # pseudo improved Euler: second evaluation with the SAME coefficient column
            x_prev, _ = self._x_prev_and_pred_x0(x, e_t, index)
            e_t_prime = (e_t + out(x_prev, tn_int)) / 2
        elif len(old_eps) == 1:
            e_t_prime = (3 * e_t - old_eps[-1]) / 2
        elif len(old_eps) == 2:
            e_t_prime = (23 * e_t - 16 * old_eps[-1] + 5 * old_eps[-2]) / 12
        else:
            e_t_prime = (55 * e_t - 59 * old_eps[-1] + 37 * old_eps[-2] - 9 * old_eps[-3]) / 24
        x_prev, pred_x0 = self._x_prev_and_pred_x0(x, e_t_prime, index)
        return x_prev, pred_x0, e_t
