"""Launch the streaming kernels at their dominant UNet shapes (for `ncu --set full -k regex:...`)."""
import sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import ops
x = torch.randn(2, 320, 64, 64, device="cuda").half().contiguous(memory_format=torch.channels_last)
gam, bet = torch.ones(320, device="cuda"), torch.zeros(320, device="cuda")
t, r = torch.randn(8192, 320, device="cuda").half(), torch.randn(8192, 320, device="cuda").half()
bias = torch.zeros(320, device="cuda")
proj, do = torch.randn(8192, 2560, device="cuda").half(), torch.randn(8192, 1280, device="cuda").half()
for _ in range(3):
    y, stats, xn = ops.groupnorm_fwd(x, gam, bet, 1e-5, True, None)
    ops.groupnorm_bwd(xn, y, gam, bet, stats, 1e-5, True, None)
    s, yy, st = ops.add_layernorm_fwd(t, bias, r, gam, bet, 1e-5)
    ops.add_layernorm_bwd(yy, r, s, st, gam)
    ops.geglu_fwd(proj)
    ops.geglu_bwd(proj, do)
torch.cuda.synchronize()
