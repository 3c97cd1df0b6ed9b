"""Per-kernel device time inside the captured CUDA graphs of one UNet evaluation (forward graph, recompute+backward
graph), warm and un-serialised (CUPTI activity records through torch.profiler — not ncu's cold-cache replay).

  python tools/profile_graph.py [replays] > gpurun_out/graph_profile.txt
"""
import collections
import os
import re
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

R = int(sys.argv[1]) if len(sys.argv) > 1 else 5
pipe = SpaceTimeAttnPipeline(steps=4, num_epochs=1, save_images=False)
items = [it for it in P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT)) if len(it.object_names) == 2][:1]
cond = pipe.to_device(pipe.encode([items[0]]))
pipe.generate([items[0]], cond)
torch.cuda.synchronize()
g = next(iter(pipe.model.graph_runner.graphs.values()))
from torch.profiler import ProfilerActivity, profile  # noqa: E402


def short(name: str) -> str:
    name = re.sub(r"std::array<[^>]*>", "arr", name)
    name = re.sub(r"\[lambda[^\]]*\]", "λ", name)
    return name[:150]


graphs = [("forward graph (no grad)", g.g_fwd), ("recompute+backward graph", g.g_bwd)]
if g.slots:  # what a differentiable evaluation actually replays: forward-with-grad + backward-only of one slot
    graphs = [("slot forward graph (activations kept)", g.slots[0].g_fwd), ("slot backward graph", g.slots[0].g_bwd)]
for label, graph in graphs:
    for _ in range(2):
        graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(R):
        graph.replay()
    b.record()
    torch.cuda.synchronize()
    wall = a.elapsed_time(b) / R * 1e3
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(R):
            graph.replay()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            k = agg[short(ev.name)]
            k[0] += 1
            k[1] += ev.device_time
    tot = sum(v[1] for v in agg.values()) / R
    print(f"==== {label}: {wall:.0f} us per replay (events), kernel-time sum {tot:.0f} us, "
          f"{sum(v[0] for v in agg.values()) // R} launches")
    for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:70]:
        print(f"{t / R:9.1f} us {t / R / tot * 100:5.1f}% {n // R:5d}x {t / n:8.2f} us  {name}")
