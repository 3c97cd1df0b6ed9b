"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (count, total, share)."""
import csv
import re
import sys
from collections import defaultdict

rows = defaultdict(lambda: [0, 0.0])
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    if "sta::" in name:
        name = re.sub(r"\(.*", "", name)  # keep the template arguments of our own kernels (head dim, rows per thread)
    else:
        name = re.sub(r"\[lambda[^\]]*\]", "L", re.sub(r"std::array<[^>]*>", "arr", name))
        m = re.match(r"(void )?(at::native::)?(\(anonymous namespace\)::)?([A-Za-z0-9_:]+)(<[^(]{0,120})?", name)
        name = (m.group(4) + (m.group(5) or "")) if m else name[:120]
    name = name.replace(",", ";")[:140]
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    v_us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows[name][0] += 1
    rows[name][1] += v_us
tot = sum(v[1] for v in rows.values())
print(f"# {sum(v[0] for v in rows.values())} launches, {tot / 1000.0:.2f} ms total (cold-cache, serialised: compare SHARES)")
print("kernel,launches,total_us,share")
for k, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{k},{n},{t:.1f},{t / tot:.4f}")
