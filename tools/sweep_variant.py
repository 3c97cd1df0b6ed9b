"""Build a variant of ONE kernel translation unit with extra -D flags into csrc/build/libsta_b200_<tag>.so (the product library
is untouched) and print its path:   python tools/sweep_variant.py sta_sattn_fwd pe3 -DSTA_POLY_EVERY=3
Used by tuning sweeps on the GPU box:  STA_B200_LIB=$(python tools/sweep_variant.py ...) python tools/bench_kernels.py sattn"""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from diffusion_spacetime_attn_b200 import build as B  # noqa: E402

unit, tag, flags = sys.argv[1], sys.argv[2], sys.argv[3:]
B.build_native()
obj, out = B.BUILD / f"{unit}_{tag}.o", B.BUILD / f"libsta_b200_{tag}.so"
subprocess.run([B._nvcc(), *B.NVCC_FLAGS, *flags, "-I", str(B.INCLUDE), "-c", str(B.CSRC / f"{unit}.cu"), "-o", str(obj)],
               check=True, capture_output=True)
objs = [str(B.BUILD / (s.stem + ".o")) for s in sorted(B.CSRC.glob("*.cu")) if s.stem != unit] + [str(obj)]
subprocess.run([B._nvcc(), "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(out), *objs], check=True)
print(out)
