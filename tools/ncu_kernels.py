"""Launch each sta_* kernel a few times at the SD-v1 512^2 geometries (for `ncu --set full -k regex:...`)."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

which = sys.argv[1:] or ["sattn_fwd", "sattn_bwd", "xattn_fwd", "xattn_bwd"]
geoms = {"sattn": [(2, 4096, 8, 40), (2, 1024, 8, 80)], "xattn": [(1, 4096, 8, 40, 2), (1, 1024, 8, 80, 2)]}
for kind in which:
    for key in geoms[kind[:5]]:
        ms = bench.standalone_kernel_ms(kind, key, iters=3)
        print(kind, key, f"{ms * 1000:.1f} us")
