"""Per-launch time of sta_sattn_bwd (delta + main + dq-cast launches) at the four UNet geometries of a 512^2 image
(bench.standalone_kernel_ms: 10 launches per CUDA graph, CUDA events, L2 flushed)."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402

for key in [(2, 4096, 8, 40), (2, 1024, 8, 80), (2, 256, 8, 160), (2, 64, 8, 160)]:
    ms = sorted(bench.standalone_kernel_ms("sattn_bwd", key, iters=10) for _ in range(3))[1]
    print("sattn_bwd", key, f"{ms * 1000:.1f} us", flush=True)
