"""Shared body of the three entrypoints (the reference's txt2img-{gpt,mscoco,vsr}.py are identical except for the
dataset parsing at lines 254-261 and a None-guard in the vsr variant).

Kept from the reference CLI (scripts/txt2img-gpt.py:105-247): --prompt --outdir --skip_grid --skip_save --ddim_steps
--plms --laion400m --fixed_code --ddim_eta --n_iter --H --W --C --f --n_samples --n_rows --scale --from-file --config
--ckpt --seed --process_id --precision --dpm_solver.  Flags that select code paths the fork broke or never used
(--laion400m, --dpm_solver, DDIM without injection) are accepted and rejected with a message.  Added: --dataset
(path of the prompt file; default = the reference's relative path, falling back to the bundled synthetic set),
--bboxes-from (layout sidecar JSON), --start/--limit (the reference's hand-edited `start`), --no_alpha_opt,
--eager.  Multi-GPU: launch with torchrun; rank r renders prompts r, r+W, ... and output names keep the global index
(result_outputs/final2_s{seed}_index_{idx}.png, plms.py:288).
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parents[2]
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))


def build_parser(default_dataset: str) -> argparse.ArgumentParser:
    p = argparse.ArgumentParser()
    p.add_argument("--prompt", type=str, nargs="?", default="a painting of a virus monster playing guitar")
    p.add_argument("--outdir", type=str, nargs="?", default="outputs/notuse")
    p.add_argument("--skip_grid", action="store_true")
    p.add_argument("--skip_save", action="store_true")
    p.add_argument("--ddim_steps", type=int, default=50)
    p.add_argument("--plms", action="store_true")
    p.add_argument("--laion400m", action="store_true")
    p.add_argument("--fixed_code", action="store_true")
    p.add_argument("--ddim_eta", type=float, default=0.0)
    p.add_argument("--n_iter", type=int, default=1)
    p.add_argument("--H", type=int, default=512)
    p.add_argument("--W", type=int, default=512)
    p.add_argument("--C", type=int, default=4)
    p.add_argument("--f", type=int, default=8)
    p.add_argument("--n_samples", type=int, default=1)
    p.add_argument("--n_rows", type=int, default=0)
    p.add_argument("--scale", type=float, default=7.5)
    p.add_argument("--from-file", type=str)
    p.add_argument("--config", type=str, default="configs/stable-diffusion/v1-inference.yaml")
    p.add_argument("--ckpt", type=str, default="models/ldm/stable-diffusion-v1/model.ckpt")
    p.add_argument("--seed", type=int, default=42)
    p.add_argument("--process_id", type=int, default=0)
    p.add_argument("--precision", type=str, choices=["full", "autocast"], default="autocast")
    p.add_argument("--dpm_solver", action="store_true")
    # additions
    p.add_argument("--dataset", type=str, default=default_dataset)
    p.add_argument("--bboxes-from", dest="bboxes_from", type=str, default=None)
    p.add_argument("--start", type=int, default=0)
    p.add_argument("--limit", type=int, default=500)
    p.add_argument("--no_alpha_opt", action="store_true")
    p.add_argument("--eager", action="store_true")
    p.add_argument("--force_objects", type=int, default=None)
    p.add_argument("--allow_synthetic_conditioning", action="store_true",
                   help="with --ckpt: accept hashed-noise text embeddings / a random-weight CLIP loss when the real CLIP "
                        "weights are not installed (STA_CLIP_L_PATH, STA_CLIP_B32_PATH); otherwise that is an error")
    return p


def run(kind: str, default_dataset: str, argv=None) -> int:
    opt = build_parser(default_dataset).parse_args(argv)
    if opt.laion400m or opt.dpm_solver:
        raise SystemExit("--laion400m / --dpm_solver select vanilla samplers without the attention injection; the "
                         "reference's own versions of these paths raise a TypeError (SURVEY.md §0)")
    if opt.precision != "autocast":
        raise SystemExit("--precision full: the attention kernels are fp16 (the reference's supported mode is autocast)")
    if opt.H != opt.W:
        raise SystemExit("the layout masks assume a square latent (reference attention.py:243)")
    import torch
    import torch.distributed as dist

    from diffusion_spacetime_attn_b200 import prompts as P
    from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline, broadcast_weights, shard_prompts

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    path = Path(opt.dataset)
    if not path.exists():
        print(f"[txt2img-{kind}] {path} not found: using the bundled synthetic prompt set", file=sys.stderr)
        path, reader = P.SYNTHETIC_GPT, P.read_gpt
    else:
        reader = P.read_gpt if kind == "gpt" else P.read_lines
    records = reader(path, limit=opt.start + opt.limit)[opt.start:]
    items = P.build_work_items(records, layouts=P.load_layouts(opt.bboxes_from), start=opt.start, seed=1,
                               force_objects=opt.force_objects)
    ckpt = opt.ckpt if os.path.exists(opt.ckpt) else None
    if ckpt is None and rank == 0:
        print(f"[txt2img-{kind}] checkpoint {opt.ckpt} not found: seeded random weights", file=sys.stderr)
    pipe = SpaceTimeAttnPipeline(device=f"cuda:{local_rank}", ckpt=ckpt, steps=opt.ddim_steps, scale=opt.scale,
                                 latent_size=opt.H // opt.f, sampler="plms" if opt.plms else "ddim",
                                 save_images=not opt.skip_save, out_dir="result_outputs", cuda_graphs=not opt.eager,
                                 half_weights=not opt.eager, allow_synthetic_conditioning=opt.allow_synthetic_conditioning)
    if rank == 0:
        print(f"[txt2img-{kind}] weights: {pipe.weights}; data: {pipe.data}", file=sys.stderr)
    if world > 1:
        broadcast_weights(pipe.model)
        broadcast_weights(pipe.clip_loss)
    log = open(f"result_outputs_rank{rank}.jsonl", "a") if not opt.skip_save else None
    done = 0
    for i in shard_prompts(len(items), rank, world):
        it = items[i]
        out_png = Path("result_outputs") / ("final2_s%d_index_%d.png" % (it.seed, it.prompt_idx))
        if out_png.exists() and not opt.skip_save:
            continue  # sweep resume: skip prompts whose output exists
        t0 = time.perf_counter()
        try:
            pipe.generate([it], pipe.encode([it]), optimize_alpha=not opt.no_alpha_opt)
        except Exception as ex:  # noqa: BLE001 - a failed prompt is logged and skipped (independent work items)
            print(f"[rank {rank}] prompt {it.prompt_idx} failed: {ex}", file=sys.stderr)
            continue
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        done += 1
        if log:
            log.write(json.dumps({"prompt_idx": it.prompt_idx, "prompt": it.prompt, "n_obj": len(it.object_names), "data": pipe.data,
                                  "seconds": dt, "losses": pipe.sampler.last_result["losses"]}) + "\n")
            log.flush()
        print(f"[rank {rank}] prompt {it.prompt_idx}: {dt:.2f} s  ({it.prompt})")
    if world > 1:
        counts = torch.tensor([done], device="cuda")
        gathered = [torch.zeros_like(counts) for _ in range(world)]
        dist.all_gather(gathered, counts)
        if rank == 0:
            print("images per rank:", [int(c) for c in gathered])
        dist.destroy_process_group()
    return 0
