"""Which CUDA kernels dominate one UNet evaluation (forward and forward+backward)?  torch.profiler over eager mode."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200.pipeline import SpaceTimeAttnPipeline
from diffusion_spacetime_attn_b200 import prompts as P

pipe = SpaceTimeAttnPipeline(steps=50, with_vae=False, cuda_graphs=False, half_weights=True, use_checkpoint=False)
unet = pipe.model.model.diffusion_model
unet.set_checkpointing(False)
if len(sys.argv) > 1 and sys.argv[1] == "cl":
    unet.to(memory_format=torch.channels_last)
items = P.build_work_items(P.read_gpt(P.SYNTHETIC_GPT))[:1]
cond = pipe.to_device(pipe.encode(items))
unet.set_local_contexts([cond["locals"][i] for i in range(2)], first_timestep=981)
x = torch.randn(2, 4, 64, 64, device="cuda")
t = torch.full((2,), 501, device="cuda", dtype=torch.long)
ctx = torch.cat([cond["uc"], cond["c"]])
coef = torch.full((1, 2), 2.5, device="cuda", requires_grad=True)
def run(bwd):
    xg = x.clone().requires_grad_(True)
    with torch.autocast("cuda", dtype=torch.float16):
        y = unet(xg, 0, t, context=ctx, coef=coef, bboxs_curr=items[0].bboxes, step_time=501).float()
    if bwd:
        torch.autograd.grad(y, [xg, coef], torch.ones_like(y))
for _ in range(3): run(True)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
for bwd in (False, True):
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        run(bwd); torch.cuda.synchronize()
    print("==== backward" if bwd else "==== forward only")
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))
