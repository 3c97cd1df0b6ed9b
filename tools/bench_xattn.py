"""Per-launch device time of sta_xattn_fwd / sta_xattn_bwd at the UNet's four geometries (and the 768x768 ones):
`iters` launches captured in one CUDA graph between two events (bench.standalone_kernel_ms), L2 flushed before.
Prints one JSON line per (kernel, geometry): mean us, algorithmic GB/s and TFLOP/s, fraction of the measured HBM peak."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from diffusion_spacetime_attn_b200 import native  # noqa: E402

GEOMS = [(1, 4096, 8, 40, 2), (1, 4096, 8, 40, 3), (1, 1024, 8, 80, 2), (1, 256, 8, 160, 2), (1, 64, 8, 160, 2),
         (1, 4096, 8, 40, 0), (2, 9216, 8, 40, 6), (2, 2304, 8, 80, 6), (2, 576, 8, 160, 6)]

if __name__ == "__main__":
    peaks = bench.measured_peaks()
    print(torch.cuda.get_device_name(0))
    for key in GEOMS:
        for kind in ("xattn_fwd", "xattn_bwd"):
            ms = min(bench.standalone_kernel_ms(kind, key, iters=20) for _ in range(3))
            by, fl = bench.launch_bytes(kind, key), bench.launch_flops(kind, key)
            print(json.dumps({"kernel": kind, "geometry": list(key), "us": round(1000 * ms, 2),
                              "gbs": round(by / ms / 1e6, 1), "frac_hbm": round(by / ms / 1e6 / peaks["hbm_gbs"], 3),
                              "tflops": round(fl / ms / 1e9, 1)}), flush=True)
    print("device_error", native.device_error())
