"""Per-shape device time of the streaming kernels (GroupNorm, LayerNorm, GEGLU, bias/residual add) at the SD-v1 UNet's
shapes: 20 launches captured in one CUDA graph, timed with CUDA events (no host launch overhead), L2 flushed before.

  python tools/bench_stream_kernels.py > gpurun_out/stream_kernels.txt
"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import native, ops  # noqa: E402

HBM = 6550.0  # GB/s, MEASURED_PEAKS.json


def timed(fn, iters=20):
    for _ in range(3):
        fn()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    best = 1e9
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        best = min(best, a.elapsed_time(b) / iters * 1e3)
    return best


def row(name, shape, us, nbytes):
    print(f"{name:22s} {str(shape):22s} {us:8.2f} us  {nbytes / 1e6:8.2f} MB  {nbytes / us / 1e3:8.0f} GB/s  {nbytes / us / 1e3 / HBM:5.2f} of HBM peak")


print(torch.cuda.get_device_name(0))
for b, c, h in [(2, 320, 64), (2, 640, 64), (2, 960, 64), (2, 640, 32), (2, 1280, 32), (2, 1920, 32), (2, 1280, 16),
                (2, 2560, 16), (2, 1280, 8), (2, 2560, 8)]:
    x = torch.randn(b, c, h, h, device="cuda").half().contiguous(memory_format=torch.channels_last)
    gam, bet = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda")
    xb = torch.randn(b, c, device="cuda").half()
    y, stats, xn = ops.groupnorm_fwd(x, gam, bet, 1e-5, True, xb)
    n = x.numel() * 2
    row("groupnorm_fwd", (b, h * h, c), timed(lambda: ops.groupnorm_fwd(x, gam, bet, 1e-5, True, xb)), 3 * n)
    row("groupnorm_bwd", (b, h * h, c), timed(lambda: ops.groupnorm_bwd(xn, y, gam, bet, stats, 1e-5, True, xb)), 5 * n)
for rows, c in [(8192, 320), (2048, 640), (512, 1280), (128, 1280)]:
    x, r = torch.randn(rows, c, device="cuda").half(), torch.randn(rows, c, device="cuda").half()
    gam, bet, bias = torch.ones(c, device="cuda"), torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    n = x.numel() * 2
    row("layernorm_fwd", (rows, c), timed(lambda: ops.add_layernorm_fwd(x, None, None, gam, bet, 1e-5)), 2 * n)
    row("add_layernorm_fwd", (rows, c), timed(lambda: ops.add_layernorm_fwd(x, bias, r, gam, bet, 1e-5)), 4 * n)
    row("bias_residual_add", (rows, c), timed(lambda: ops.add_layernorm_fwd(x, bias, r, None, None, 0.0)), 3 * n)
    s, y, st = ops.add_layernorm_fwd(x, bias, r, gam, bet, 1e-5)
    row("add_layernorm_bwd", (rows, c), timed(lambda: ops.add_layernorm_bwd(y, r, s, st, gam)), 4 * n)
    proj = torch.randn(rows, 8 * c, device="cuda").half()
    do = torch.randn(rows, 4 * c, device="cuda").half()
    row("geglu_fwd", (rows, 4 * c), timed(lambda: ops.geglu_fwd(proj)), 3 * rows * 4 * c * 2)
    row("geglu_bwd", (rows, 4 * c), timed(lambda: ops.geglu_bwd(proj, do)), 5 * rows * 4 * c * 2)
print("device_error", native.device_error())
