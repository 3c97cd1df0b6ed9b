"""Launch each sta_* kernel a few times at the SD-v1 512^2 geometries (for `ncu --set full -k regex:...`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

which = sys.argv[1:] or ["sattn_fwd", "sattn_bwd", "xattn_fwd", "xattn_bwd"]
geoms = {"sattn": [(2, 4096, 8, 40), (2, 1024, 8, 80)], "xattn": [(1, 4096, 8, 40, 2), (1, 1024, 8, 80, 2)]}
for kind in list(which):
    if kind.startswith("sattn_wide"):  # the VAE mid-block geometry: one head of 512 over a 64 x 64 latent
        from diffusion_spacetime_attn_b200 import ops

        qkv = torch.randn(1, 4096, 1536, device="cuda").half()
        q, k, v = qkv.chunk(3, dim=-1)
        for _ in range(5):
            out, lse = ops.sattn_fwd(q, k, v, 1)
            ops.sattn_bwd(q, k, v, out, lse, (out * 0.1).half(), 1)
        torch.cuda.synchronize()
        print(kind, "done")
        which.remove(kind)
for kind in which:
    for key in geoms[kind[:5]]:
        ms = bench.standalone_kernel_ms(kind, key, iters=3)
        print(kind, key, f"{ms * 1000:.1f} us")
