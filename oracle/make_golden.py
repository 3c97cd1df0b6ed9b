"""Generate tests/golden/* by running the UNMODIFIED reference modules (imported from /root/reference) on CPU.

TEST INFRASTRUCTURE.  Runs only in the build container (the reference tree does not travel to the GPU box); the
fixtures it writes are committed.  Usage:  python oracle/make_golden.py [--full] [--traj]

What the reference needs to run here: an `omegaconf.listconfig.ListConfig` shim
(openaimodel.py:476), CWD containing a CPU-resaved `uncond_fix_radius_0p2_g0.pt` and the `c{i}_fix_radius_0p2_g0.pt`
local embeddings (attention.py:234,246), and one throw-away call at timestep 981 to make every block build its
masks and contexts (attention.py:240-263) before any other timestep is evaluated.

Weights: no checkpoint is available offline, so both sides use oracle.sta_oracle.seeded_state_dict (the
reference's zero-initialised layers are randomised, else the UNet output is identically zero).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import tempfile
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from oracle import sta_oracle as O  # noqa: E402

REF_SD = Path("/root/reference/attention_optimization/stable-diffusion")
GOLD = ROOT / "tests" / "golden"


def import_reference():
    if "omegaconf" not in sys.modules:
        oc = types.ModuleType("omegaconf")
        lc = types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        sys.modules["omegaconf"] = oc
        sys.modules["omegaconf.listconfig"] = lc
    sys.path.insert(0, str(REF_SD))
    import ldm.modules.attention as ref_attn  # noqa: E402
    import ldm.modules.diffusionmodules.openaimodel as ref_unet  # noqa: E402

    return ref_attn, ref_unet


def ctx_tensor(seed, shape=(1, 77, 768)):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(shape, generator=g) * 1.04


def prepare_cwd(tmp: Path, local_cs):
    """The reference reads its embeddings from files in CWD."""
    unc = torch.load(REF_SD / "uncond_fix_radius_0p2_g0.pt", map_location="cpu") if torch.cuda.is_available() else None
    if unc is None:
        import pickle

        class CPUUnpickler(pickle.Unpickler):
            def find_class(self, module, name):
                if module == "torch.storage" and name == "_load_from_bytes":
                    import io

                    return lambda b: torch.load(io.BytesIO(b), map_location="cpu")
                return super().find_class(module, name)

        unc = torch.load(REF_SD / "uncond_fix_radius_0p2_g0.pt", map_location=torch.device("cpu"), weights_only=False)
    unc = unc.detach().float().cpu()
    torch.save(unc, tmp / "uncond_fix_radius_0p2_g0.pt")
    for i, c in enumerate(local_cs):
        torch.save(c, tmp / ("c%d_fix_radius_0p2_g0.pt" % i))
    return unc


def golden_blocks(ref_attn, uncond):
    """One BasicTransformerBlock per channel width (d = 40 / 80 / 160)."""
    out = {}
    for tag, (dim_px, C) in {"L0": (8, 320), "L1": (4, 640), "L2": (4, 1280)}.items():
        n = dim_px * dim_px
        bboxes = [[0.30, 0.50], [0.70, 0.50]]
        blk = ref_attn.BasicTransformerBlock(C, 8, C // 8, context_dim=768)
        shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
        sd = O.seeded_state_dict(shapes, seed=1)
        blk.load_state_dict(sd)
        g = torch.Generator().manual_seed(20 + C)
        x = torch.randn(2, n, C, generator=g)
        context = torch.cat([uncond, ctx_tensor(100)])
        coef = torch.tensor([2.5, 1.5])
        with torch.no_grad():
            y = blk(x.clone(), context=context, time=torch.tensor(981), text_index=0, coef=coef, bboxs_curr=bboxes)
        out[tag] = y.numpy()
        out[tag + "_shapes"] = json.dumps({k: list(v) for k, v in shapes.items()})
    np.savez_compressed(GOLD / "blocks.npz", **out)
    print("blocks.npz written")


TINY = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(1, 2), num_res_blocks=1,
            channel_mult=(1, 2), num_heads=8, context_dim=768)
FULL = dict(in_channels=4, out_channels=4, model_channels=320, attention_resolutions=(4, 2, 1), num_res_blocks=2,
            channel_mult=(1, 2, 4, 4), num_heads=8, context_dim=768)


def build_ref_unet(ref_unet, cfg, seed):
    m = ref_unet.UNetModel(image_size=32, use_spatial_transformer=True, transformer_depth=1, use_checkpoint=False,
                           legacy=False, **cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    ours = O.unet_param_shapes(O.UNetConfig(**cfg))
    assert ours == shapes, "oracle.unet_param_shapes disagrees with the reference module tree"
    m.load_state_dict(O.seeded_state_dict(shapes, seed))
    return m.eval(), shapes


def ref_eval(model, x, t, context, coef, bboxes):
    """One reference UNet evaluation at timestep t; primes the blocks with a 981 call first if needed."""
    with torch.no_grad():
        if not getattr(model, "_primed", False):
            model(x, 0, torch.full((x.shape[0],), 981, dtype=torch.long), context=context, coef=coef, bboxs_curr=bboxes)
            model._primed = True
        return model(x, 0, torch.full((x.shape[0],), t, dtype=torch.long), context=context, coef=coef, bboxs_curr=bboxes)


def golden_unet(ref_unet, uncond, cfg, name, latent, seed, t=501):
    model, shapes = build_ref_unet(ref_unet, cfg, seed)
    g = torch.Generator().manual_seed(1)
    x = torch.randn(1, 4, latent, latent, generator=g)
    x_in = torch.cat([x, x])
    context = torch.cat([uncond, ctx_tensor(100)])
    coef = torch.tensor([2.5, 2.5])
    bboxes = [[0.30, 0.50], [0.70, 0.50]]
    y = ref_eval(model, x_in, t, context, coef, bboxes)
    np.savez_compressed(GOLD / f"{name}.npz", eps=y.numpy(), t=t, latent=latent, seed=seed)
    (GOLD / f"{name}_shapes.json").write_text(json.dumps({k: list(v) for k, v in shapes.items()}))
    print(f"{name}.npz written; |eps| mean {y.abs().mean():.4f}")
    return model


def golden_trajectory(ref_unet, uncond, model, steps=10):
    """BASELINE.json configs[0]: 'a red cube left of a blue sphere', 2 objects, 64x64 latent, 10 PLMS steps, fixed alpha.
    The UNet is the unmodified reference module; the PLMS arithmetic is the oracle's restatement (plms.py cannot be
    imported here: it needs the `clip` package)."""
    g = torch.Generator().manual_seed(1)
    x_T = torch.randn(1, 4, 64, 64, generator=g)
    uc, c = uncond, ctx_tensor(100)
    context = torch.cat([uc, c])
    coef = torch.tensor([2.5, 2.5])
    bboxes = [[0.30, 0.50], [0.70, 0.50]]

    def eps_model(x, t, i):
        e_u, e_c = ref_eval(model, torch.cat([x, x]), t, context, coef, bboxes).chunk(2)
        return e_u + 7.5 * (e_c - e_u)

    z = O.plms_trajectory(eps_model, x_T, steps)
    np.savez_compressed(GOLD / "config1_trajectory.npz", latent=z.numpy(), steps=steps)
    print("config1_trajectory.npz written; |z| mean", float(z.abs().mean()))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--full", action="store_true", help="also the full SD-v1 UNet single evaluation (~1 min)")
    ap.add_argument("--traj", action="store_true", help="also the 10-step config-1 trajectory (~3 min, implies --full)")
    args = ap.parse_args()
    GOLD.mkdir(parents=True, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    local_cs = [ctx_tensor(101), ctx_tensor(102)]
    with tempfile.TemporaryDirectory() as td:
        tmp = Path(td)
        uncond = prepare_cwd(tmp, local_cs)
        os.chdir(tmp)
        ref_attn, ref_unet = import_reference()
        torch.save(uncond, GOLD / "uncond_embedding.pt")
        golden_blocks(ref_attn, uncond)
        golden_unet(ref_unet, uncond, TINY, "unet_tiny", latent=8, seed=3)
        if args.full or args.traj:
            model = golden_unet(ref_unet, uncond, FULL, "unet_full", latent=64, seed=0)
            if args.traj:
                golden_trajectory(ref_unet, uncond, model)
        os.chdir(ROOT)


if __name__ == "__main__":
    main()
