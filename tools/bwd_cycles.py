"""Cycle breakdown of sta_sattn_bwd (CTA 0: one math warp and the MMA-issue thread), STA_DEBUG_FLAGS=8.

  STA_DEBUG_FLAGS=8 python tools/bwd_cycles.py [b n h d]
"""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ.setdefault("STA_DEBUG_FLAGS", "8")
import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import native, ops  # noqa: E402

b, n, h, d = (int(v) for v in sys.argv[1:5]) if len(sys.argv) >= 5 else (2, 4096, 8, 40)
q, k, v = (torch.randn(b, n, h * d, device="cuda").half() for _ in range(3))
do = torch.randn(b, n, h * d, device="cuda").half() * 0.1
out, lse = ops.sattn_fwd(q, k, v, h)
for _ in range(3):
    ops.sattn_bwd(q, k, v, out, lse, do, h)
torch.cuda.synchronize()
a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    ops.sattn_bwd(q, k, v, out, lse, do, h)
e.record()
torch.cuda.synchronize()
buf = (C.c_longlong * 16)()
lib = native.load()
lib.sta_debug_read.argtypes = [C.POINTER(C.c_longlong), C.c_int]
lib.sta_debug_read(buf, 16)
vals = list(buf)
tiles = (n + 127) // 128
print(f"sattn_bwd {b}x{n}x{h}x{d}: {a.elapsed_time(e) / 10 * 1e3:.1f} us per call (3 launches), {tiles} query tiles per CTA")
names = ["math: wait S/dP", "math: P/dS compute", "math: wait dQ", "math: drain dQ", "math: total"]
for nm, val in zip(names, vals[:5]):
    print(f"  {nm:22s} {val:9d} cycles  {val / max(tiles, 1):8.0f} per tile")
for nm, val in zip(["mma: wait P/dS", "mma: issue", "mma: total"], vals[8:11]):
    print(f"  {nm:22s} {val:9d} cycles  {val / max(tiles, 1):8.0f} per tile")
print("device_error", native.device_error())
