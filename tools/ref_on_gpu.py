"""Run the UNMODIFIED reference (vendored by tools/vendor_ref.py into untracked baseline/_ref/) on the B200 under
`torch.autocast("cuda")` — the only mode in which its backward is legal (in-place hazard) — and

  (i)  time `UNetModel` forward and checkpointed forward+backward at BASELINE.json configs[1]'s geometry (batch 2 CFG,
       64x64 latent, 2 objects, use_checkpoint=True as configs/stable-diffusion/v1-inference.yaml:43) and extrapolate one
       alpha-optimised image exactly as bench.py's CPU arm does: 153 t_fwd + 150 (t_fwd+bwd - t_fwd)
       -> gpurun_out/ref_on_gpu.json   (copied to profiles/ and BASELINE.md §5 "Reference on 1xB200 (PyTorch)")
  (ii) dump the reference's outputs and gradients (dL/dx, dL/dalpha) for the seeded cases of tests/ref_cases.py: one
       BasicTransformerBlock per geometry (fp16 autocast = r16, and fp32 forward = r32) and the tiny / full UNet
       -> gpurun_out/ref_gpu.npz       (committed as tests/golden/ref_gpu.npz; tests/test_ref_gpu_golden.py reads it)

TEST INFRASTRUCTURE: imports oracle/ (seeded weights) and the reference; never imported by the product.

    python tools/vendor_ref.py && gpurun -- python tools/ref_on_gpu.py
"""
from __future__ import annotations

import json
import os
import sys
import tempfile
import time
import types
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import ref_cases as RC  # noqa: E402
from oracle import sta_oracle as O  # noqa: E402

REF = ROOT / "baseline" / "_ref"
OUT = ROOT / "gpurun_out"


def import_reference():
    if not (REF / "ldm" / "modules" / "attention.py").exists():
        raise SystemExit("baseline/_ref is empty: run `python tools/vendor_ref.py` in the build container first")
    if "omegaconf" not in sys.modules:  # openaimodel.py:476 only needs the ListConfig type
        oc, lc = types.ModuleType("omegaconf"), types.ModuleType("omegaconf.listconfig")

        class ListConfig(list):
            pass

        lc.ListConfig = ListConfig
        oc.listconfig = lc
        sys.modules["omegaconf"], sys.modules["omegaconf.listconfig"] = oc, lc
    sys.path.insert(0, str(REF))
    import ldm.modules.attention as ref_attn
    import ldm.modules.diffusionmodules.openaimodel as ref_unet

    return ref_attn, ref_unet


def prepare_cwd(tmp: Path, locs, device):
    """attention.py:234,246 read the embeddings from files in CWD with torch.load (device = the device they were saved on)."""
    torch.save(RC.uncond().to(device), tmp / "uncond_fix_radius_0p2_g0.pt")
    for i, c in enumerate(locs):
        torch.save(c.to(device), tmp / ("c%d_fix_radius_0p2_g0.pt" % i))


def rel(a, b):
    return float((a.float() - b.float()).norm() / b.float().norm())


def dump_blocks(ref_attn, out):
    for tag, n, C in RC.BLOCKS:
        x, G, context, locs, coef = RC.block_case(n, C)
        blk = ref_attn.BasicTransformerBlock(C, 8, C // 8, context_dim=768)
        blk.load_state_dict(RC.block_weights({k: tuple(v.shape) for k, v in blk.state_dict().items()}))
        blk = blk.cuda().eval()
        sub = RC.token_subsample(n, C)
        # r32: fp32 forward (the backward is illegal outside CUDA autocast: attention.py:282,294 write in place)
        with torch.no_grad():
            y32 = blk(x.cuda(), context=context.cuda(), time=981, text_index=0, coef=coef.cuda(), bboxs_curr=RC.BBOXES)
        # r16: what scripts/txt2img-gpt.py:310 runs — fp16 activations under autocast, fp32 master weights
        xg = x.cuda().half().requires_grad_(True)
        cg = coef.cuda().requires_grad_(True)
        with torch.autocast("cuda"):
            y16 = blk(xg, context=context.cuda(), time=981, text_index=0, coef=cg, bboxs_curr=RC.BBOXES)
            (y16.float() * G.cuda()).sum().backward()
        out[f"{tag}_y16"] = y16.detach()[:, sub].half().cpu().numpy()
        out[f"{tag}_y32"] = y32[:, sub].float().cpu().numpy()
        out[f"{tag}_dx16"] = xg.grad[:, sub].half().cpu().numpy()
        out[f"{tag}_dcoef16"] = cg.grad.float().cpu().numpy()
        out[f"{tag}_gap_y"] = np.float32(rel(y16, y32))
        print(f"block {tag}: |r16 - r32| / |r32| = {rel(y16, y32):.3e}   d_coef = {cg.grad.tolist()}", flush=True)


def build_unet(ref_unet, cfg, seed, use_checkpoint):
    m = ref_unet.UNetModel(image_size=32, use_spatial_transformer=True, transformer_depth=1, use_checkpoint=use_checkpoint,
                           legacy=False, **cfg)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(O.seeded_state_dict(shapes, seed))
    return m.cuda().eval()


def prime(model, x, context, coef):
    """Every block builds its masks / local contexts when it sees timestep 981 (attention.py:240-263)."""
    with torch.no_grad(), torch.autocast("cuda"):
        model(x, 0, torch.full((x.shape[0],), 981, dtype=torch.long, device="cuda"), context=context, coef=coef,
              bboxs_curr=RC.BBOXES)


def dump_unets(ref_unet, out):
    for tag, cfg, latent, seed, t in RC.UNETS:
        x, G, context, locs, coef = RC.unet_case(latent)
        model = build_unet(ref_unet, cfg, seed, use_checkpoint=True)
        x, G, context = x.cuda(), G.cuda(), context.cuda()
        prime(model, x, context, coef.cuda())
        tt = torch.full((2,), t, dtype=torch.long, device="cuda")
        with torch.no_grad():
            y32 = model(x, 0, tt, context=context, coef=coef.cuda(), bboxs_curr=RC.BBOXES)
        xg, cg = x.clone().requires_grad_(True), coef.cuda().requires_grad_(True)
        with torch.autocast("cuda"):
            y16 = model(xg, 0, tt, context=context, coef=cg, bboxs_curr=RC.BBOXES)
            (y16.float() * G).sum().backward()
        out[f"{tag}_y16"] = y16.detach().half().cpu().numpy()
        out[f"{tag}_y32"] = y32.float().cpu().numpy()
        out[f"{tag}_dx16"] = xg.grad.float().cpu().numpy()
        out[f"{tag}_dcoef16"] = cg.grad.float().cpu().numpy()
        out[f"{tag}_gap_y"] = np.float32(rel(y16, y32))
        print(f"{tag}: |r16 - r32| / |r32| = {rel(y16, y32):.3e}   d_coef = {cg.grad.tolist()}", flush=True)
        del model
        torch.cuda.empty_cache()


def cuda_time(fn, iters, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters / 1000.0


def time_reference(ref_unet, iters=8):
    tag, cfg, latent, seed, t = RC.UNETS[1]
    x, G, context, locs, coef = RC.unet_case(latent)
    x, context = x.cuda(), context.cuda()
    tt = torch.full((2,), t, dtype=torch.long, device="cuda")
    res = {}
    for label, freeze in (("as_shipped", False),):
        # as shipped the UNet parameters require grad (scripts/txt2img-gpt.py:55-72 never freezes them), so
        # loss.backward() (plms.py:276) also computes 859.5 M weight gradients that nobody reads.  Freezing them is not an
        # option for the unmodified code: CheckpointFunction.backward (util.py:139) differentiates w.r.t. every parameter
        # and torch raises "One of the differentiated Tensors does not require grad" (tried on this box).
        model = build_unet(ref_unet, cfg, seed, use_checkpoint=True)
        model.requires_grad_(not freeze)
        prime(model, x, context, coef.cuda())

        def fwd():
            with torch.no_grad(), torch.autocast("cuda"):
                model(x, 0, tt, context=context, coef=coef.cuda(), bboxs_curr=RC.BBOXES)

        def fwd_bwd():
            xg, cg = x.clone().requires_grad_(True), coef.cuda().requires_grad_(True)
            with torch.autocast("cuda"):
                y = model(xg, 0, tt, context=context, coef=cg, bboxs_curr=RC.BBOXES)
                y.float().sum().backward()
            model.zero_grad(set_to_none=True)

        t_f, t_fb = cuda_time(fwd, iters), cuda_time(fwd_bwd, iters)
        t_img = 153 * t_f + 150 * max(t_fb - t_f, 0.0)
        res[label] = {"t_fwd_s": t_f, "t_fwd_bwd_s": t_fb, "s_per_image_extrapolated": t_img, "images_per_s": 1.0 / t_img}
        print(label, json.dumps(res[label]), flush=True)
        del model
        torch.cuda.empty_cache()
    return res


def main():
    torch.backends.cuda.matmul.allow_tf32 = False  # r32 must be true fp32: no TF32 GEMMs ...
    torch.backends.cudnn.allow_tf32 = False        # ... and no TF32 convolutions (cuDNN's default would add ~8e-4)
    OUT.mkdir(exist_ok=True)
    locs = [RC.ctx_tensor(101), RC.ctx_tensor(102)]
    with tempfile.TemporaryDirectory() as td:
        prepare_cwd(Path(td), locs, "cuda")
        os.chdir(td)
        ref_attn, ref_unet = import_reference()
        out = {}
        dump_blocks(ref_attn, out)
        dump_unets(ref_unet, out)
        np.savez_compressed(OUT / "ref_gpu.npz", **out)
        timing = time_reference(ref_unet)
        os.chdir(ROOT)
    info = {"what": "UNMODIFIED reference UNetModel (baseline/_ref, sha256 in MANIFEST.json) under torch.autocast('cuda'), "
                    "batch 2 (CFG), 64x64 latent, 2 objects, use_checkpoint=True; one alpha-optimised image extrapolated "
                    "as 153 t_fwd + 150 (t_fwd+bwd - t_fwd) (no VAE decode / CLIP loss), CUDA events",
            "gpu": torch.cuda.get_device_name(0), "torch": torch.__version__, "extrapolated": True, "timing": timing,
            "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime())}
    (OUT / "ref_on_gpu.json").write_text(json.dumps(info, indent=1))
    print(json.dumps(info))


if __name__ == "__main__":
    main()
