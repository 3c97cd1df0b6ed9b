"""How much of an image is NOT the UNet trajectory?  Times generate() with the real VAE decode + CLIP loss and with
trivial stand-ins, then profiles one decode + loss + backward (the once-per-epoch tail, reference plms.py:249-277).

  python tools/profile_tail.py > gpurun_out/tail_profile.txt
"""
import collections
import re
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

pipe = SpaceTimeAttnPipeline(steps=50, num_epochs=3, save_images=False)
items = [it for it in P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT)) if len(it.object_names) == 2][:4]
items = (items * 3)[:8]
conds = [pipe.to_device(pipe.encode([it])) for it in items]
pipe.generate([items[0]], conds[0])
torch.cuda.synchronize()


def timed(i):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe.generate([items[i]], conds[i])
    torch.cuda.synchronize()
    return time.perf_counter() - t0


# medians of three images each way (a single pair of images differs by +-1 % = +-5 ms per epoch on its own)
full = sorted(timed(i) for i in (1, 2, 3))[1]
s = pipe.sampler
real_decode, real_loss = s.decode_fn, s.loss_fn
s.decode_fn = lambda z: z
s.loss_fn = lambda imgs, *a: (imgs.float().sum(), [imgs.float().sum()])
timed(4)
bare = sorted(timed(i) for i in (5, 6, 7))[1]
s.decode_fn, s.loss_fn = real_decode, real_loss
print(f"image with VAE+CLIP tail: {full * 1e3:.0f} ms; trajectory only: {bare * 1e3:.0f} ms; tail = {(full - bare) * 1e3 / 3:.1f} ms per epoch")

z = torch.randn(1, 4, 64, 64, device="cuda", requires_grad=True)
it = items[0]


def tail():
    with torch.autocast("cuda", dtype=torch.float16):
        img = real_decode(z)
        loss, _ = real_loss(img, [it.prompt], [it.bboxes], [it.object_names])
    (g,) = torch.autograd.grad(loss, z)
    return g


for _ in range(2):
    tail()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(3):
    tail()
torch.cuda.synchronize()
print(f"tail alone (eager, wall): {(time.perf_counter() - t0) / 3 * 1e3:.1f} ms")
from torch.profiler import ProfilerActivity, profile  # noqa: E402

with profile(activities=[ProfilerActivity.CUDA]) as prof:
    tail()
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"\[lambda[^\]]*\]", "λ", re.sub(r"std::array<[^>]*>", "arr", ev.name))[:150]
        agg[name][0] += 1
        agg[name][1] += ev.device_time
tot = sum(v[1] for v in agg.values())
print(f"kernel-time sum {tot:.0f} us, {sum(v[0] for v in agg.values())} launches")
for name, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
    print(f"{t:9.1f} us {t / tot * 100:5.1f}% {n:5d}x {t / n:8.2f} us  {name}")
