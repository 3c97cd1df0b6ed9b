"""Micro-benchmarks of the individual kernels (CUDA events, L2 flushed between iterations)."""
import sys, json
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import ops, native

PEAK_TFLOPS = 1645.3

def time_fn(fn, iters=20, warmup=5, flush=True):
    fl = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device="cuda") if flush else None
    for _ in range(warmup): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush: fl.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]

def bench_sattn():
    for (b, n, h, d) in [(2, 4096, 8, 40), (2, 1024, 8, 80), (2, 256, 8, 160), (2, 64, 8, 160), (2, 9216, 8, 40), (8, 4096, 8, 40)]:
        q, k, v = (torch.randn(b, n, h * d, device="cuda").half() for _ in range(3))
        ms = time_fn(lambda: ops.sattn_fwd(q, k, v, h))
        fl = 4.0 * n * n * h * d * b
        print(json.dumps({"kernel": "sattn_fwd", "shape": [b, n, h, d], "ms": round(ms, 4), "tflops": round(fl / ms / 1e9, 1), "frac_of_measured_peak": round(fl / ms / 1e9 / PEAK_TFLOPS, 3)}))

if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    which = sys.argv[1:] or ["sattn"]
    if "sattn" in which: bench_sattn()
    if "xattn" in which and hasattr(sys.modules[__name__], "bench_xattn"): bench_xattn()
    print("device_error", native.device_error())
