"""How much of an image is NOT inside the captured UNet-evaluation graphs?  Times (CUDA events) one complete configs[1] image
through the pipeline, then replays the very same slot graphs back to back (51 forward + 50 backward per epoch, 3 epochs)
without the sampler's Python / elementwise work in between.  The difference is the host + small-kernel overhead per step."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

pipe = SpaceTimeAttnPipeline(steps=50, num_epochs=3, save_images=False)
items = [it for it in P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT)) if len(it.object_names) == 2][:3]
conds = [pipe.to_device(pipe.encode([it])) for it in items]
pipe.generate([items[0]], conds[0])
torch.cuda.synchronize()


def timed(fn, reps=2):
    best = 1e9
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b))
    return best


t_image = timed(lambda: pipe.generate([items[1]], conds[1], check_device_error=False))
g = next(iter(pipe.model.graph_runner.graphs.values()))
slots = g.slots


def replay_only():
    for _ in range(3):
        for s in slots[:51]:
            s.g_fwd.replay()
        for s in slots[:50]:
            s.g_bwd.replay()


t_graphs = timed(replay_only)
print(f"image through the pipeline: {t_image:.1f} ms; the same {3 * 51} forward + {3 * 50} backward graph replays alone: "
      f"{t_graphs:.1f} ms; outside the graphs (sampler arithmetic, copies, VAE + CLIP tail 3 x ~21.7 ms, Adam): {t_image - t_graphs:.1f} ms "
      f"= {100 * (t_image - t_graphs) / t_image:.1f} %")
