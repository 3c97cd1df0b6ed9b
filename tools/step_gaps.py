"""GPU idle time inside one differentiable trajectory: runs a short image (S steps, 1 epoch) under torch.profiler (CUPTI,
kernels inside CUDA-graph replays included), sorts every device activity by start time and reports where the GPU sat idle
(gaps between consecutive kernels) and which kernels ran OUTSIDE the captured evaluation graphs."""
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

S = int(sys.argv[1]) if len(sys.argv) > 1 else 6
pipe = SpaceTimeAttnPipeline(steps=S, num_epochs=1, save_images=False, with_vae=False)
G = torch.randn(1, 4, 64, 64, device="cuda")
pipe.sampler.decode_fn = lambda z: z
pipe.sampler.loss_fn = lambda imgs, *a: ((imgs.float() * G).sum(), [(imgs.float() * G).sum()])
items = [it for it in P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT)) if len(it.object_names) == 2][:2]
conds = [pipe.to_device(pipe.encode([it])) for it in items]
pipe.generate([items[0]], conds[0])
pipe.generate([items[0]], conds[0])
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    pipe.generate([items[1]], conds[1], check_device_error=False)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range.end > e.time_range.start]
evs.sort(key=lambda e: e.time_range.start)
t0, t1 = evs[0].time_range.start, max(e.time_range.end for e in evs)
busy, idle, cur_end = 0.0, 0.0, evs[0].time_range.start
gaps = []
for e in evs:
    s, en = e.time_range.start, e.time_range.end
    if s > cur_end:
        idle += s - cur_end
        gaps.append((s - cur_end, e.name[:70]))
    busy += max(0.0, en - max(s, cur_end))
    cur_end = max(cur_end, en)
print(f"S={S}: span {(t1 - t0) / 1e3:.2f} ms, busy {busy / 1e3:.2f} ms, idle {idle / 1e3:.2f} ms ({100 * idle / (t1 - t0):.1f} %), {len(evs)} device activities")
by = defaultdict(lambda: [0, 0.0])
for g, nm in gaps:
    by[nm][0] += 1
    by[nm][1] += g
print("idle time by the kernel that FOLLOWS the gap (top 12):")
for nm, (n, t) in sorted(by.items(), key=lambda kv: -kv[1][1])[:12]:
    print(f"  {t / 1e3:8.3f} ms  {n:5d}x  {t / n:7.1f} us each   {nm}")
big = sorted(gaps, reverse=True)[:8]
print("largest single gaps:", [(round(g, 1), nm[:40]) for g, nm in big])
