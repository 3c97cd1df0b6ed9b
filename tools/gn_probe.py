"""Run each GroupNorm shape a few times (for `ncu --metrics gpu__time_duration.sum`: which kernels run, how long)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import ops
for b, c, h in [(2, 320, 64), (2, 640, 64), (2, 1920, 32), (2, 1280, 32), (2, 1280, 8)]:
    x = torch.randn(b, c, h, h, device="cuda").half().contiguous(memory_format=torch.channels_last)
    gam, bet = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    for _ in range(3):
        y, stats, xn = ops.groupnorm_fwd(x, gam, bet, 1e-5, True, None)
        ops.groupnorm_bwd(xn, y, gam, bet, stats, 1e-5, True, None)
torch.cuda.synchronize()
