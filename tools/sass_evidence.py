"""Per-kernel SASS evidence that the attention kernels are Blackwell-native: counts of the tcgen05 / TMEM / TMA mnemonics
in every `sta::*` kernel of the in-tree libsta_b200.so (`cuobjdump -sass`), written to profiles/r2_sass_summary.txt.

  UTCHMMA = tcgen05.mma (kind::f16)   LDTM / STTM = tcgen05.ld / st   UTMALDG / UTMASTG / UTMAREDG = TMA load / store /
  reduce-add   UTCBAR = tcgen05.commit   HMMA would be the legacy mma.sync path (expected: 0)
"""
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "diffusion_spacetime_attn_b200" / "libsta_b200.so"
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "SYNCS", "MUFU.EX2", "HMMA", "LDL", "STL"]

if __name__ == "__main__":
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    chunks = re.split(r"\n\s*Function : \S+", sass)[1:]
    lines = ["# cuobjdump -sass diffusion_spacetime_attn_b200/libsta_b200.so  (sm_100a), instruction counts per kernel",
             "%-58s %s" % ("kernel", " ".join("%9s" % o for o in OPS))]
    for name, body in zip(filt, chunks):
        short = re.sub(r"\(.*", "", name).replace("void ", "")
        if "sta::" not in short:
            continue
        counts = [len(re.findall(r"\b" + re.escape(o) + r"\b", body)) for o in OPS]
        lines.append("%-58s %s" % (short[:58], " ".join("%9d" % c for c in counts)))
    out = ROOT / "profiles" / "r2_sass_summary.txt"
    out.write_text("\n".join(lines) + "\n")
    sys.stdout.write(out.read_text())
