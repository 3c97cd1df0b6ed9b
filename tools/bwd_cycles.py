"""Cycle breakdown of sta_sattn_bwd (CTA 0: one math warp and the MMA-issue thread), STA_DEBUG_FLAGS=8, from a separate
-DSTA_BWD_DEBUG build of the library (the product has these paths compiled out).

  STA_DEBUG_FLAGS=8 python tools/bwd_cycles.py [b n h d]
"""
import ctypes as C
import os
import sys
from pathlib import Path

os.environ.setdefault("STA_DEBUG_FLAGS", "8")
import subprocess  # noqa: E402

import torch  # noqa: E402

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import build as B  # noqa: E402


def build_debug():
    """The counters / ablations are compiled out of the product: build csrc/build/libsta_b200_bwddbg.so with -DSTA_BWD_DEBUG."""
    B.build_native()
    out, obj = B.BUILD / "libsta_b200_bwddbg.so", B.BUILD / "sta_sattn_bwd_dbg.o"
    subprocess.run([B._nvcc(), *B.NVCC_FLAGS, "-DSTA_BWD_DEBUG", "-I", str(B.INCLUDE), "-c", str(B.CSRC / "sta_sattn_bwd.cu"),
                    "-o", str(obj)], check=True, capture_output=True)
    objs = [str(B.BUILD / (s.stem + ".o")) for s in sorted(B.CSRC.glob("*.cu")) if s.stem != "sta_sattn_bwd"] + [str(obj)]
    subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs], check=True)
    return out


os.environ["STA_B200_LIB"] = str(build_debug())
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
for nm, val in zip(["mma: wait P/dS", "mma: wait Q/dO", "mma: total"], vals[8:11]):
    print(f"  {nm:22s} {val:9d} cycles  {val / max(tiles, 1):8.0f} per tile")
print("device_error", native.device_error())
