"""Where does the once-per-epoch tail (VAE decode + CLIP loss + their backward, reference plms.py:249-277) spend its time INSIDE
an image?  CUDA events on the sampling stream around the decode, the loss and the arrival of d(loss)/d(latent), plus the
host's wall-clock time in the same calls, for every epoch of a few images.

  python tools/tail_in_situ.py > gpurun_out/tail_in_situ.txt
"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import prompts as P  # noqa: E402
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline  # noqa: E402

pipe = SpaceTimeAttnPipeline(steps=50, num_epochs=3, save_images=False)
items = [it for it in P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT)) if len(it.object_names) == 2][:4]
conds = [pipe.to_device(pipe.encode([it])) for it in items]
pipe.generate([items[0]], conds[0])
torch.cuda.synchronize()

s = pipe.sampler
real_decode, real_loss = s.decode_fn, s.loss_fn
rec = []  # per epoch: dict of events / host times


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def decode(z):
    r = {"a": ev(), "h0": time.perf_counter()}
    rec.append(r)
    z.register_hook(lambda g: r.__setitem__("d", ev()))  # d(loss)/d(latent) is ready: the tail's backward has been enqueued
    out = real_decode(z)
    r["b"] = ev()
    r["h1"] = time.perf_counter()
    return out


def loss(*a):
    r = rec[-1]
    out = real_loss(*a)
    r["c"] = ev()
    r["h2"] = time.perf_counter()
    return out


s.decode_fn, s.loss_fn = decode, loss
traj_start = []
real_traj = s._trajectory


def traj(*a, **k):
    traj_start.append(ev())  # also marks the end of the previous epoch's backward + Adam step
    return real_traj(*a, **k)


s._trajectory = traj
# host time per piece of the loss call (where does the host block?)
clip = s.clip_loss_model
host = {}


def wrap(obj, name):
    f = getattr(obj, name)

    def g(*a, **k):
        t = time.perf_counter()
        out = f(*a, **k)
        host.setdefault(name, []).append(1e3 * (time.perf_counter() - t))
        return out

    setattr(obj, name, g)


for nm in ("_text_feat", "resize_global", "resize_crop", "_encode_image", "tokenizer"):
    wrap(clip, nm)
wrap(clip.model, "encode_text")
for i in (1, 2, 3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pipe.generate([items[i]], conds[i])
    torch.cuda.synchronize()
    print(f"image {i}: {1e3 * (time.perf_counter() - t0):.0f} ms")
for k, r in enumerate(rec):
    print(f"epoch {k}: GPU decode fwd {r['a'].elapsed_time(r['b']):6.2f} ms, loss fwd {r['b'].elapsed_time(r['c']):6.2f} ms, "
          f"tail backward {r['c'].elapsed_time(r['d']):6.2f} ms, total {r['a'].elapsed_time(r['d']):6.2f} ms | host: decode call "
          f"{1e3 * (r['h1'] - r['h0']):6.2f} ms, loss call {1e3 * (r['h2'] - r['h1']):6.2f} ms")
traj_start.append(ev())
torch.cuda.synchronize()
for k, r in enumerate(rec):
    fwd = traj_start[k].elapsed_time(r["a"])
    line = f"epoch {k}: forward trajectory {fwd:7.2f} ms = 51 x {fwd / 51:.3f}"
    if k % 3 != 2:  # the next trajectory of the same image starts right behind this epoch's backward
        bwd = r["d"].elapsed_time(traj_start[k + 1])
        line += f"; trajectory backward + Adam {bwd:7.2f} ms = 50 x {bwd / 50:.3f}"
    print(line)
for nm, v in host.items():
    print(f"host {nm}: " + " ".join(f"{x:.2f}" for x in v[:24]))
