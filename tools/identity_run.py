"""Per-prompt outputs of a small prompt sweep, written to --out as prompt_<idx>.pt, from 1 process or from W torchrun
ranks (prompts sharded round-robin, weights broadcast from rank 0).  tests/test_multigpu_identity.py runs it both ways and
compares the files: a prompt's result must not depend on how many GPUs the sweep was spread over (SURVEY.md §4 item 6;
the reference shards by hand-edited `start` ranges and process_id.py, scripts/txt2img-gpt.py:303-305).

Reduced UNet (2 levels) and 4 steps keep it to seconds; the code path (pipeline, sampler, CUDA graphs, kernels) is the product's.
"""
import argparse
import os
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
os.environ.setdefault("STA_CUDNN_BENCHMARK", "0")  # the autotuner may pick different (differently rounding) algorithms per process


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", required=True)
    ap.add_argument("--prompts", type=int, default=4)
    ap.add_argument("--optimize", type=int, default=0)
    args = ap.parse_args()
    import torch.distributed as dist

    from diffusion_spacetime_attn_b200 import native, prompts as P
    from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline, broadcast_weights, shard_prompts

    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    tiny = dict(image_size=32, in_channels=4, out_channels=4, model_channels=320, attention_resolutions=[1, 2],
                num_res_blocks=1, channel_mult=[1, 2], num_heads=8, use_spatial_transformer=True, transformer_depth=1,
                context_dim=768, legacy=False)
    # rank r > 0 starts from DIFFERENT weights on purpose: only the broadcast makes the sweep consistent
    pipe = SpaceTimeAttnPipeline(device=f"cuda:{local_rank}", seed=rank, steps=4, num_epochs=2, latent_size=16,
                                 with_vae=False, unet_config=tiny, save_images=False)
    broadcast_weights(pipe.model)
    if args.optimize:
        G = torch.randn(1, 4, 16, 16, generator=torch.Generator().manual_seed(5)).to(pipe.device)
        pipe.sampler.decode_fn = lambda z: z
        pipe.sampler.loss_fn = lambda imgs, *a: ((imgs.float() * G).sum(), [(imgs.float() * G).sum()])
    items = P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT))[:args.prompts]
    os.makedirs(args.out, exist_ok=True)
    for i in shard_prompts(len(items), rank, world):
        it = items[i]
        pipe.generate([it], pipe.encode([it]), optimize_alpha=bool(args.optimize))
        r = pipe.sampler.last_result
        torch.save({"latent": r["latent"].cpu(), "weighting_parameter": r["weighting_parameter"].cpu(), "rank": rank},
                   os.path.join(args.out, f"prompt_{it.prompt_idx}.pt"))
    torch.cuda.synchronize()
    assert native.device_error() == 0
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
