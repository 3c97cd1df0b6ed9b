"""Copy the handful of UNMODIFIED reference files the hot path needs into the untracked `baseline/_ref/` so that the
reference itself can run on the GPU box (`/root/reference` does not exist there; `baseline/_ref/` is git-ignored but NOT
gpurun-ignored, so it travels with the snapshot).  Nothing under baseline/_ref is ever committed or imported by the product.

    python tools/vendor_ref.py            # in the build container, before tools/ref_on_gpu.py is sent to the GPU

Writes baseline/_ref/MANIFEST.json with the sha256 of every file (source and copy are identical byte for byte); the shipped
`uncond_fix_radius_0p2_g0.pt` was pickled on cuda:0 and is re-saved from tests/golden/uncond_embedding.pt (same values, CPU).
"""
from __future__ import annotations

import hashlib
import json
import shutil
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference/attention_optimization/stable-diffusion")
DST = ROOT / "baseline" / "_ref"
FILES = [
    "process_id.py",
    "ldm/util.py",
    "ldm/modules/attention.py",                       # BasicTransformerBlock._forward :268-300 (the hot path)
    "ldm/modules/diffusionmodules/__init__.py",
    "ldm/modules/diffusionmodules/util.py",           # checkpoint :102-148, schedule helpers
    "ldm/modules/diffusionmodules/openaimodel.py",    # UNetModel :443-742
    "ldm/modules/diffusionmodules/model.py",          # KL-VAE Decoder :150-202, :535-568
    "ldm/models/diffusion/plms.py",                   # p_sample_plms :296-358 (needs a `clip` shim to import)
]


def sha(p: Path) -> str:
    return hashlib.sha256(p.read_bytes()).hexdigest()


def main() -> int:
    if not SRC.exists():
        print(f"{SRC} not found: vendor_ref.py only runs in the build container", file=sys.stderr)
        return 1
    manifest = {}
    for rel in FILES:
        d = DST / rel
        d.parent.mkdir(parents=True, exist_ok=True)
        shutil.copyfile(SRC / rel, d)
        assert sha(d) == sha(SRC / rel)
        manifest[rel] = sha(d)
    (DST / "MANIFEST.json").write_text(json.dumps({"source": str(SRC), "sha256": manifest}, indent=1))
    print(f"vendored {len(FILES)} reference files into {DST}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
