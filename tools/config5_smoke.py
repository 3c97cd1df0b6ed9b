"""BASELINE.json configs[4] at FULL geometry, shortened: 768x768 (96x96 latent), DDIM + injection, 6 local descriptions,
2 prompts per GPU — S steps / E alpha epochs instead of 100 / 3.  Checks that the N = 9216 self-attention, the 6-object fused
cross-attention, the B = 2 activation slots and the memory budget logic hold at the stress geometry.

  python tools/config5_smoke.py [S] [E]
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import native, prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 10
E = int(sys.argv[2]) if len(sys.argv) > 2 else 2
pipe = SpaceTimeAttnPipeline(steps=S, num_epochs=E, latent_size=96, sampler="ddim", save_images=False)
items = P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT), force_objects=6)[:4]
for rep in range(2):
    batch = items[2 * rep:2 * rep + 2]
    cond = pipe.to_device(pipe.encode(batch))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    img = pipe.generate(batch, cond)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    res = pipe.sampler.last_result
    print(f"pass {rep}: {dt:.2f} s for 2 images of {tuple(img.shape[1:])}, {S} DDIM steps x {E} epochs; losses {res['losses']}")
    assert torch.isfinite(img).all() and torch.isfinite(res["weighting_parameter"]).all()
print("slots", pipe.model.graph_runner.slot_summary(), "reserved GiB", round(torch.cuda.memory_reserved() / 2 ** 30, 1))
print("alpha moved by", float((res["weighting_parameter"] - 5.0 / 6).abs().max()))
print("device_error", native.device_error())
