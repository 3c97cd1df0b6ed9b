"""Pin the SAMPLER arithmetic by execution: import the UNMODIFIED reference ldm/models/diffusion/plms.py (it needs the
`clip` package, absent here -> a 6-line sys.modules shim; nothing of `clip` is touched by p_sample_plms) and drive its
own `make_schedule` + `p_sample_plms` (plms.py:81-112, 296-358) exactly as `plms_sampling` does (plms.py:227-247) with
a linear stand-in for the UNet.  Writes tests/golden/plms_sampler.npz; tests/test_host_cpu.py checks the oracle
(`plms_trajectory`) and the product sampler against it.

TEST INFRASTRUCTURE; runs only in the build container (/root/reference).  The only harness-side adaptation is a
subclass that overrides `register_buffer` (plms.py:74-78 force-moves every buffer to "cuda"; there is no GPU here) —
the sampler arithmetic itself runs unmodified, on CPU, in fp32.

    python oracle/make_golden_sampler.py
"""
from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
REF_SD = Path("/root/reference/attention_optimization/stable-diffusion")
GOLD = ROOT / "tests" / "golden"


class LinearUNet:
    """Stand-in for LatentDiffusion: the schedule buffers of ddpm.py:116-160 (linear in sqrt(beta), 0.00085 -> 0.012, 1000
    steps: configs/stable-diffusion/v1-inference.yaml:5-9) and an `apply_model_extra` that is linear in x and depends on
    t, on the coefficient column and on the row (unconditional / conditional), so every branch of p_sample_plms matters."""

    num_timesteps = 1000
    parameterization = "eps"
    device = torch.device("cpu")

    def __init__(self):
        betas = torch.linspace(0.00085 ** 0.5, 0.012 ** 0.5, 1000, dtype=torch.float64) ** 2
        acp = np.cumprod(1.0 - betas.numpy(), axis=0)
        self.betas = betas.float()
        self.alphas_cumprod = torch.tensor(acp, dtype=torch.float32)
        self.alphas_cumprod_prev = torch.tensor(np.append(1.0, acp[:-1]), dtype=torch.float32)

    @staticmethod
    def apply_model_extra(x_in, text_index, t_in, c_in, coef=None, bboxs_curr=None, **kw):
        s = (t_in.float() / 1000.0).reshape(-1, 1, 1, 1)
        k = 1.0 + (coef.sum() if coef is not None else 0.0) * 0.01
        rolled = torch.roll(x_in, 1, dims=-1)
        row = torch.cat([torch.zeros_like(x_in[:1]), torch.ones_like(x_in[1:])])
        return 0.1 * k * x_in * s + 0.02 * rolled + 0.05 * row * (1.0 + c_in.mean())


def import_reference_plms():
    clip = types.ModuleType("clip")
    clip.load = lambda *a, **k: (torch.nn.Identity(), None)  # DCLIPLoss.__init__ (plms.py:24); never called afterwards
    clip.tokenize = lambda *a, **k: None
    sys.modules["clip"] = clip
    sys.path.insert(0, str(REF_SD))
    import ldm.models.diffusion.plms as ref_plms

    return ref_plms


def run_reference(ref_plms, S, x_T, c, uc, W, scale=7.5):
    class CPUSampler(ref_plms.PLMSSampler):
        def register_buffer(self, name, attr):  # plms.py:74-78 would move every buffer to "cuda"
            setattr(self, name, attr)

    sampler = CPUSampler(LinearUNet())
    sampler.make_schedule(ddim_num_steps=S, ddim_eta=0.0, verbose=False)
    timesteps = sampler.ddim_timesteps
    time_range = np.flip(timesteps)
    total_steps = timesteps.shape[0]
    img, old_eps, eps_hist = x_T.clone(), [], []
    for i, step in enumerate(time_range):  # plms.py:227-247, verbatim control flow
        index = total_steps - i - 1
        ts = torch.full((1,), step, dtype=torch.long)
        ts_next = torch.full((1,), time_range[min(i + 1, len(time_range) - 1)], dtype=torch.long)
        img, pred_x0, e_t = sampler.p_sample_plms(img, c, ts, index=index, unconditional_guidance_scale=scale,
                                                  unconditional_conditioning=uc, old_eps=old_eps, t_next=ts_next,
                                                  text_index=0, coef=W[:, i], bboxs_curr=[[0.3, 0.5], [0.7, 0.5]])
        old_eps.append(e_t)
        eps_hist.append(e_t.clone())
        if len(old_eps) >= 4:
            old_eps.pop(0)
    return img, torch.stack(eps_hist), sampler


def main():
    ref_plms = import_reference_plms()
    out = {}
    for S in (5, 10, 50):
        g = torch.Generator().manual_seed(S)
        x_T = torch.randn(1, 4, 8, 8, generator=g)
        c = torch.randn(1, 77, 768, generator=g)
        uc = torch.randn(1, 77, 768, generator=g)
        W = 2.5 + 0.5 * torch.randn(2, S, generator=g)
        z, eps, sampler = run_reference(ref_plms, S, x_T, c, uc, W)
        out[f"S{S}_latent"] = z.numpy()
        out[f"S{S}_eps"] = eps.numpy()
        out[f"S{S}_timesteps"] = np.asarray(sampler.ddim_timesteps)
        out[f"S{S}_alphas"] = np.asarray(sampler.ddim_alphas, dtype=np.float64)
        out[f"S{S}_alphas_prev"] = np.asarray(sampler.ddim_alphas_prev, dtype=np.float64)
        print(f"S={S}: timesteps {sampler.ddim_timesteps[:3]}..{sampler.ddim_timesteps[-1]}  |z| {float(z.abs().mean()):.4f}")
    np.savez_compressed(GOLD / "plms_sampler.npz", **out)
    print("tests/golden/plms_sampler.npz written")


if __name__ == "__main__":
    main()
